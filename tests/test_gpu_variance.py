"""GPU parity test for PictureAnalysis' per-LCU mean / variance pyramid (row N1, through the C ABI) against the reference's own static
ComputeBlockMeanComputeVariance when oracle/_ref travelled to this box, else the plain-C port."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_pa import PA_CASES, padded_luma, port, refpa, run_variance

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", PA_CASES + [("uhd_3840x2160", 3840, 2160, "synth"), ("hd_1920x1080", 1920, 1080, "random")], ids=lambda c: c[0])
def test_variance_cuda_vs_oracle(case):
    name, w, h, kind = case
    pl, _keep = padded_luma(w, h, seed=w + h, kind=kind)
    if refpa is not None:
        wm, wv = run_variance(refpa, "svtref_variance_picture", pl, C.c_int(1 if common.ref_asm_simd_ok() else 0))
    else:
        wm, wv = run_variance(port, "port_variance_picture", pl)
    gm, gv = S.variance_picture(pl)
    assert np.array_equal(gm, wm), (name, np.argwhere(gm != wm)[:5])
    assert np.array_equal(gv, wv), (name, np.argwhere(gv != wv)[:5])


def test_variance_rejects_a_plane_without_padding():
    y = np.zeros((72, 200), np.uint8)
    pl = S.Plane()
    pl.buf, pl.stride, pl.origin_x, pl.origin_y, pl.width, pl.height = y.ctypes.data, 200, 0, 0, 200, 72
    with pytest.raises(S.Svtb200Error):
        S.variance_picture(pl)
