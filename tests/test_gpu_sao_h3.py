"""GPU parity of the H3 SAO statistics (device pictures: deblocked + unfiltered reconstruction) against the oracle rule and, when present, the reference's
own interleaving of deblocking and gathering (tests/test_oracle_sao_h3.py)."""
import numpy as np
import pytest

import svtb200 as S
from test_oracle_dlf import make_dlf_inputs
from test_oracle_sao import assert_stats_equal
from test_oracle_sao_h3 import CASES, noisy_source, port_h3, ref, reference_interleaved

pytestmark = pytest.mark.gpu


def dev_picture(host):
    import torch
    planes = [torch.from_numpy(np.ascontiguousarray(p)).cuda() for p in (host.y, host.cb, host.cr)]
    st = S.make_yuv_struct(tuple(t.data_ptr() for t in planes), tuple(t.shape[1] for t in planes), 80, host.width, host.height)
    return planes, st


@pytest.mark.parametrize("case", CASES + [("4k_full_frame", 3840, 2160, "synth", 0.5)], ids=lambda c: c[0])
def test_sao_stats_h3_cuda_vs_oracle(case):
    import torch
    name, w, h, kind, density = case
    pic, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    src = noisy_source(pic, w * 7 + h)
    final, _, want = port_h3(pic, src, vbs, hbs, qp, prm)
    keep = [dev_picture(p) for p in (src, final, pic)]
    n = ((w + 63) // 64) * ((h + 63) // 64)
    d_out = torch.zeros(n * 3 * S.SAO_STATS_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    S.sao_stats_picture_h3_dev(keep[0][1], keep[1][1], keep[2][1], d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(S.SAO_STATS_DTYPE).reshape(n, 3)
    assert_stats_equal(got, want, name + " H3 CUDA vs port")
    if ref is not None and w < 2000:
        _, _, wr = reference_interleaved(pic, src, vbs, hbs, qp, prm, asm_mode=0)
        assert_stats_equal(got, wr, name + " H3 CUDA vs the reference's interleaved run")


@pytest.mark.parametrize("case", CASES[:5] + CASES[-1:] + [("4k_full_frame16", 3840, 2160, "synth", 0.5)], ids=lambda c: c[0])
def test_sao_stats_h3_16bit_cuda_vs_oracle(case):
    """16-bit sample planes (10-bit video, configs 3 / 4): the host entry point against the port's rule (itself pinned against the reference's interleaved
    LCU*DLFCore16bit + GatherSaoStatisticsLcu_62x62_16bit run by tests/test_oracle_sao_h3.py)"""
    import ctypes as C
    from test_oracle_dlf import run_dlf16, to16
    from test_oracle_sao_h3 import port, vp
    name, w, h, kind, density = case
    pic8, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    pic = to16(pic8, seed=w, extremes=(kind == "extremes"))
    rng = np.random.default_rng(w * 7 + h)
    src = S.HostYuvPicture(*[np.clip(a.astype(np.int32) + rng.integers(-30, 31, a.shape), 0, 1023).astype(np.uint16) for a in pic.interior()], dtype=np.uint16)
    n = ((w + 63) // 64) * ((h + 63) // 64)
    final = run_dlf16(port, "port_dlf_picture16", pic, vbs, hbs, qp, prm)
    scratch = S.HostYuvPicture(*(np.zeros_like(a) for a in pic.interior()), dtype=np.uint16)
    want = np.zeros((n, 3), dtype=S.SAO_STATS_DTYPE)
    pf = port.port_sao_stats_picture16_h3
    pf.restype = C.c_int
    assert pf(C.byref(src.struct), C.byref(final.struct), C.byref(pic.struct), C.byref(scratch.struct), vp(want)) == 0
    got = S.sao_stats_picture16_h3(src, final, pic)
    assert_stats_equal(got, want, name + " 16-bit H3 CUDA vs port")
