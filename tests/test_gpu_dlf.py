"""GPU parity tests for deblocking (through the C ABI) against the plain-C oracle and the reference's per-LCU drivers."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_dlf import DLF_CASES, edge_block, make_dlf_inputs, run_dlf

pytestmark = pytest.mark.gpu
port = common.load_port()
ref = common.load_ref()
vp = lambda a, off=0: C.c_void_p(a.ctypes.data + off)


@pytest.mark.parametrize("kind", ["random", "step", "ramp"])
def test_edge_filter_leafs(kind):
    rng = np.random.default_rng(3)
    port.port_luma_edge.restype = port.port_chroma_edge.restype = None
    for trial in range(60):
        blk = edge_block(kind, rng)
        tc, beta, vert = int(rng.integers(0, 25)), int(rng.integers(0, 65)), int(rng.integers(0, 2))
        want, got = blk.copy(), blk.copy()
        port.port_luma_edge(vp(want, 8 * 16 + 8), C.c_uint32(16), C.c_uint8(vert), C.c_int32(tc), C.c_int32(beta))
        S.luma_edge_dlf(got, 8 * 16 + 8, 16, vert, tc, beta)
        assert np.array_equal(got, want)
        cb_tc, cr_tc = int(rng.integers(0, 25)), int(rng.integers(0, 25))
        wa, wb, ga, gb = blk.copy(), blk.T.copy(), blk.copy(), blk.T.copy()
        port.port_chroma_edge(vp(wa, 8 * 16 + 8), vp(wb, 8 * 16 + 8), C.c_uint32(16), C.c_uint8(vert), C.c_uint8(cb_tc), C.c_uint8(cr_tc))
        S.chroma_edge_dlf(ga, gb, 8 * 16 + 8, 16, vert, cb_tc, cr_tc)
        assert np.array_equal(ga, wa) and np.array_equal(gb, wb)


@pytest.mark.parametrize("case", DLF_CASES + [("4k_full_frame", 3840, 2160, "synth", 0.5)], ids=lambda c: c[0])
def test_dlf_picture_cuda_vs_oracle(case):
    name, w, h, kind, density = case
    pic, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    if ref is not None:
        want = run_dlf(ref, "svtref_dlf_picture", pic, vbs, hbs, qp, prm, asm_mode=1 if common.ref_asm_simd_ok() else 0)
    else:
        want = run_dlf(port, "port_dlf_picture", pic, vbs, hbs, qp, prm)
    y, cb, cr = (a.copy() for a in pic.interior())
    got = S.HostYuvPicture(y, cb, cr)
    S.dlf_picture(got, vbs, hbs, qp, prm)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs: {np.argwhere(a != b)[:5]}"
    # padding untouched
    fresh = S.HostYuvPicture(*(a.copy() for a in pic.interior()))
    assert np.array_equal(got.y[:80], fresh.y[:80]) and np.array_equal(got.cb[:, :40], fresh.cb[:, :40])


@pytest.mark.parametrize("case", DLF_CASES, ids=lambda c: c[0])
def test_dlf_picture16_cuda_vs_oracle(case):
    """16-bit sample planes (10-bit video)"""
    from test_oracle_dlf import run_dlf16, to16
    name, w, h, kind, density = case
    pic8, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    pic = to16(pic8, seed=w, extremes=kind == "extremes")
    want = run_dlf16(port, "port_dlf_picture16", pic, vbs, hbs, qp, prm)
    y, cb, cr = (a.copy() for a in pic.interior())
    got = S.HostYuvPicture(y, cb, cr, dtype=np.uint16)
    S.dlf_picture(got, vbs, hbs, qp, prm, hbd=True)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs (16-bit CUDA vs port): {np.argwhere(a != b)[:5]}"
    if ref is not None:
        wr = run_dlf16(ref, "svtref_dlf_picture16", pic, vbs, hbs, qp, prm, asm_mode=1 if common.ref_asm_simd_ok() else 0)
        for a, b in zip(got.interior(), wr.interior()):
            assert np.array_equal(a, b), name
