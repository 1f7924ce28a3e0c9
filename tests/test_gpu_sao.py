"""GPU parity tests for SAO statistics and application (through the C ABI) against the plain-C oracle and the reference."""
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
