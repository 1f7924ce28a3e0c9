"""GPU parity tests for the boundary-strength derivation (row N3, through the C ABI) against the reference's own SetBSArrayBasedOn{PU,TU}Boundary when
oracle/_ref travelled to this box, else the plain-C port; then the derived strengths drive the deblocking of a picture."""
import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_bs import BS_CASES, make_md_maps, run_bs

pytestmark = pytest.mark.gpu
port = common.load_port()
ref = common.load_ref()


@pytest.mark.parametrize("case", BS_CASES + [("B_4k", 3840, 2160, 0, 0, 16), ("P_1080p", 1920, 1080, 1, 0, 0)], ids=lambda c: c[0])
def test_bs_cuda_vs_oracle(case):
    name, w, h, st, p0, p1 = case
    cu, cbf = make_md_maps(w, h, seed=w + h + st, slice_type=st)
    if ref is not None:
        wv, wh = run_bs(ref, "svtref_dlf_bs_picture", cu, cbf, w, h, st, p0, p1)
    else:
        wv, wh = run_bs(port, "port_dlf_bs_picture", cu, cbf, w, h, st, p0, p1)
    gv, gh = S.dlf_bs_picture(cu, cbf, w, h, st, p0, p1)
    assert np.array_equal(gv, wv), (name, np.argwhere(gv != wv)[:5])
    assert np.array_equal(gh, wh), (name, np.argwhere(gh != wh)[:5])


def test_bs_feeds_the_deblocking_filter():
    """derive on the device, filter with the result: equal to filtering with the oracle's strengths"""
    from test_oracle_dlf import make_dlf_inputs, run_dlf
    w, h = 256, 192
    pic, _, _, qp, prm = make_dlf_inputs(w, h, seed=5, density=0.5, kind="synth")
    cu, cbf = make_md_maps(w, h, seed=8, slice_type=0)
    wv, wh = run_bs(port, "port_dlf_bs_picture", cu, cbf, w, h, 0, 0, 16)
    gv, gh = S.dlf_bs_picture(cu, cbf, w, h, 0, 0, 16)
    want = run_dlf(port, "port_dlf_picture", pic, wv, wh, qp, prm)
    got = S.HostYuvPicture(*(a.copy() for a in pic.interior()))
    S.dlf_picture(got, gv, gh, qp, prm)
    for a, b in zip(got.interior(), want.interior()):
        assert np.array_equal(a, b)


def test_bs_rejects_bad_maps():
    cu, cbf = make_md_maps(128, 64, seed=1)
    bad = cu.copy(); bad[0, 0]["log2_cu_size"] = 7
    with pytest.raises(S.Svtb200Error):
        S.dlf_bs_picture(bad, cbf, 128, 64, 0)
    with pytest.raises(S.Svtb200Error):
        S.dlf_bs_picture(cu, cbf, 132, 64, 0)      # width not a multiple of 8
