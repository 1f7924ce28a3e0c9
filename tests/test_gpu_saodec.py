"""GPU parity of the SAO rate-distortion decision (row f2 / next-row N3) through the C ABI against the plain-C oracle and, when present,
the reference's own static functions (same inputs as tests/test_oracle_saodec.py)."""
import numpy as np
import pytest

import svtb200 as S
from test_oracle_sao import SAO_CASES
from test_oracle_saodec import LAMBDAS, assert_decisions_equal, picture_stats, port, rate_tables, refdec, run_decision

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", SAO_CASES + [("4k_full_frame", 3840, 2160, "synth")], ids=lambda c: c[0])
@pytest.mark.parametrize("rates", ["typical", "free", "expensive"])
def test_sao_decision_cuda_vs_oracle(case, rates):
    stats, cols, rows = picture_stats(case)
    t = rate_tables(rates)
    for lam, lam_c in LAMBDAS:
        got = S.sao_decision_picture(stats, cols, rows, t, lam, lam_c)
        want = run_decision(port, "port_sao_decision_picture", stats, cols, rows, t, lam, lam_c)
        assert_decisions_equal(got, want, f"{case[0]} {rates} lambda {lam}: CUDA vs port")
        if refdec is not None:
            assert_decisions_equal(got, run_decision(refdec, "svtref_sao_decision_picture", stats, cols, rows, t, lam, lam_c), f"{case[0]}: CUDA vs reference")


@pytest.mark.parametrize("is10bit", [0, 1])
def test_sao_decision_cuda_random_statistics(is10bit):
    rng = np.random.default_rng(12 + is10bit)
    cols, rows = 61, 35          # more LCUs per anti-diagonal than one warp, more diagonals than a 4K picture
    st = np.zeros((cols * rows, 3), dtype=S.SAO_STATS_DTYPE)
    scale = 4 if is10bit else 1
    st["bo_count"] = rng.integers(0, 400, st["bo_count"].shape) * (rng.random(st["bo_count"].shape) < 0.6)
    st["bo_diff"] = (rng.normal(0, 3 * scale, st["bo_diff"].shape) * st["bo_count"]).astype(np.int32)
    st["eo_count"] = rng.integers(0, 1500, st["eo_count"].shape) * (rng.random(st["eo_count"].shape) < 0.8)
    st["eo_diff"] = (rng.normal(0, 2.5 * scale, st["eo_diff"].shape) * st["eo_count"]).astype(np.int32)
    st["eo_count"][3::7] = st["eo_count"][2::7][:len(st["eo_count"][3::7])]
    st["eo_diff"][3::7] = st["eo_diff"][2::7][:len(st["eo_diff"][3::7])]
    for rates in ("typical", "expensive"):
        t = rate_tables(rates)
        for lam, lam_c in LAMBDAS:
            got = S.sao_decision_picture(st, cols, rows, t, lam, lam_c, is10bit)
            want = run_decision(port, "port_sao_decision_picture", st, cols, rows, t, lam, lam_c, is10bit)
            assert_decisions_equal(got, want, f"random {rates} lambda {lam} 10bit={is10bit}")
