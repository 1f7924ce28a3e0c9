"""GPU parity tests for SAO statistics and application (through the C ABI) against the plain-C oracle and the reference."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_sao import (SAO_APPLY_CASES, SAO_CASES, assert_stats_equal, make_sao_inputs, port, refsao, run_apply_port, run_apply_ref,
                             run_stats)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", SAO_CASES + [("4k_full_frame", 3840, 2160, "synth")], ids=lambda c: c[0])
def test_sao_stats_cuda_vs_oracle(case):
    name, w, h, kind = case
    src, rec, _ = make_sao_inputs(w, h, seed=w * 3 + h, kind=kind)
    got = S.sao_stats_picture(src, rec)
    want = run_stats(port, "port_sao_stats_picture", src, rec)
    assert_stats_equal(got, want, name + " CUDA vs port")
    assert not got["eo_diff"][..., 4].any() and not got["eo_count"][..., 4].any()
    if refsao is not None:
        assert_stats_equal(got, run_stats(refsao, "svtref_sao_stats_picture", src, rec, asm_mode=1 if common.ref_asm_simd_ok() else 0), name + " CUDA vs reference")


@pytest.mark.parametrize("case", SAO_APPLY_CASES + [("4k_full_frame", 3840, 2160, "synth")], ids=lambda c: c[0])
@pytest.mark.parametrize("flags", [(1, 1), (1, 0), (0, 1)])
def test_sao_apply_cuda_vs_oracle(case, flags):
    name, w, h, kind = case
    if w > 2000 and flags != (1, 1):
        pytest.skip("4K once")
    _, rec, prm = make_sao_inputs(w, h, seed=w * 3 + h, kind=kind)
    out = S.HostYuvPicture(*(np.zeros_like(a) for a in rec.interior()))
    S.sao_apply_picture(rec, out, prm, flags)
    want = run_apply_ref(rec, prm, flags, asm_mode=1 if common.ref_asm_simd_ok() else 0) if refsao is not None else run_apply_port(rec, prm, flags)
    for a, b, nm in zip(out.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs: {np.argwhere(a != b)[:5]}"


@pytest.mark.parametrize("case", SAO_CASES, ids=lambda c: c[0])
def test_sao_stats16_cuda_vs_oracle(case):
    """16-bit sample planes (10-bit video)"""
    from test_oracle_sao import make_sao_inputs16
    name, w, h, kind = case
    src, rec, _ = make_sao_inputs16(w, h, seed=w * 3 + h, kind=kind)
    want = run_stats(port, "port_sao_stats_picture16", src, rec)
    got = S.sao_stats_picture(src, rec, hbd=True)
    assert_stats_equal(got, want, name + " 16-bit CUDA vs port")
    assert np.all(got["eo_diff"][..., 4] == 0) and np.all(got["eo_count"][..., 4] == 0)


@pytest.mark.parametrize("case", SAO_APPLY_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("flags", [(1, 1), (0, 1)])
def test_sao_apply16_cuda_vs_oracle(case, flags):
    from test_oracle_sao import make_sao_inputs16
    name, w, h, kind = case
    _, rec, prm = make_sao_inputs16(w, h, seed=w * 3 + h, kind=kind)
    want = S.HostYuvPicture(*(np.zeros_like(a) for a in rec.interior()), dtype=np.uint16)
    fn = port.port_sao_apply_picture16; fn.restype = C.c_int
    assert fn(C.byref(rec.struct), C.byref(want.struct), prm.ctypes.data_as(C.c_void_p), (C.c_uint8 * 2)(*flags)) == 0
    got = S.HostYuvPicture(*(np.zeros_like(a) for a in rec.interior()), dtype=np.uint16)
    S.sao_apply_picture(rec, got, prm, flags, hbd=True)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs (16-bit CUDA vs port)"
