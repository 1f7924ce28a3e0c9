"""SAO rate-distortion decision: the plain-C restatement (oracle/port/sao_decision_port.c) against the
reference's own static functions (oracle/refinc_saodec_harness.c #includes EbSampleAdaptiveOffsetGenerationDecision.c in place).
CPU only: the decision still runs on the host; these tests pin the oracle a device version will be checked against."""
import ctypes as C
import os

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_sao import SAO_CASES, make_sao_inputs, run_stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
port = common.load_port()


class RateTables(C.Structure):
    _fields_ = [("merge_flag_bits", C.c_uint32 * 2), ("type_index_bits", C.c_uint32 * 6), ("offset_trun_unary_bits", C.c_uint32 * 8)]


DECISION_DTYPE = np.dtype([("merge_left", np.uint8), ("merge_up", np.uint8), ("type", np.uint8, 2), ("band_position", np.uint8, 3), ("reserved", np.uint8),
                           ("offset", np.int32, (3, 4)), ("luma_cost", np.int64), ("chroma_cost", np.int64)], align=True)
assert DECISION_DTYPE.itemsize == 72


def load_ref_saodec():
    variant = "avx512" if common.cpu_has("avx512bw") else "avx2"
    p = os.path.join(ROOT, "oracle", "_ref", f"libsvtref_saodec_{variant}.so")
    return C.CDLL(p) if os.path.exists(p) else None


refdec = load_ref_saodec()
needs_ref = pytest.mark.skipif(refdec is None, reason="oracle/_ref/libsvtref_saodec_*.so not built (no /root/reference here)")


def rate_tables(kind):
    """plausible CABAC estimates (1 bit = 32768) and two degenerate sets"""
    t = RateTables()
    if kind == "typical":
        t.merge_flag_bits[:] = [9000, 70000]
        t.type_index_bits[:] = [14000, 80000, 80000, 80000, 80000, 81000]
        t.offset_trun_unary_bits[:] = [32768 * (k + 1) if k < 7 else 32768 * 7 for k in range(8)]
    elif kind == "free":
        t.merge_flag_bits[:] = [0, 0]; t.type_index_bits[:] = [0] * 6; t.offset_trun_unary_bits[:] = [0] * 8
    else:   # "expensive": off and merge are cheap, everything else costs a lot
        t.merge_flag_bits[:] = [40000, 20000]
        t.type_index_bits[:] = [1000, 300000, 280000, 290000, 310000, 250000]
        t.offset_trun_unary_bits[:] = [20000 + 50000 * k for k in range(8)]
    return t


def run_decision(lib, name, stats, cols, rows, rates, lam, lam_c, is10bit=0):
    out = np.zeros(cols * rows, dtype=DECISION_DTYPE)
    fn = getattr(lib, name)
    fn.restype = C.c_int
    assert fn(stats.ctypes.data_as(C.c_void_p), C.c_int(cols), C.c_int(rows), C.byref(rates), C.c_uint64(lam), C.c_uint64(lam_c), C.c_int(is10bit),
              out.ctypes.data_as(C.c_void_p)) == 0
    return out


def assert_decisions_equal(a, b, what):
    for f in ("merge_left", "merge_up", "type", "band_position", "offset", "luma_cost", "chroma_cost"):
        if not np.array_equal(a[f], b[f]):
            bad = np.argwhere(a[f] != b[f])[0]
            raise AssertionError(f"{what}: {f} differs first at {tuple(bad)}: {a[f][tuple(bad)]} vs {b[f][tuple(bad)]}")


def picture_stats(case):
    name, w, h, kind = case
    src, rec, _ = make_sao_inputs(w, h, seed=w * 3 + h, kind=kind)
    return run_stats(port, "port_sao_stats_picture", src, rec), (w + 63) // 64, (h + 63) // 64


# lambdas: fullLambda / fullChromaLambdaSao of a QP 20 / 32 / 45 picture are of these magnitudes (lambdaModeDecision* tables, EbLambdaRateTables.h)
LAMBDAS = [(30000, 25000), (439162, 350000), (9000000, 7000000)]


@needs_ref
@pytest.mark.parametrize("case", SAO_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("rates", ["typical", "free", "expensive"])
def test_decision_port_vs_reference_on_picture_statistics(case, rates):
    stats, cols, rows = picture_stats(case)
    t = rate_tables(rates)
    for lam, lam_c in LAMBDAS:
        want = run_decision(refdec, "svtref_sao_decision_picture", stats, cols, rows, t, lam, lam_c)
        got = run_decision(port, "port_sao_decision_picture", stats, cols, rows, t, lam, lam_c)
        assert_decisions_equal(got, want, f"{case[0]} {rates} lambda {lam}")


@needs_ref
@pytest.mark.parametrize("is10bit", [0, 1])
def test_decision_port_vs_reference_on_random_statistics(is10bit):
    """statistics no picture would produce: large and sign-mixed sums, empty categories, counts of one -- every clip, tie and merge path"""
    rng = np.random.default_rng(12 + is10bit)
    cols, rows = 9, 7
    st = np.zeros((cols * rows, 3), dtype=S.SAO_STATS_DTYPE)
    scale = 4 if is10bit else 1
    st["bo_count"] = rng.integers(0, 400, st["bo_count"].shape) * (rng.random(st["bo_count"].shape) < 0.6)
    st["bo_diff"] = (rng.normal(0, 3 * scale, st["bo_diff"].shape) * st["bo_count"]).astype(np.int32)
    st["eo_count"] = rng.integers(0, 1500, st["eo_count"].shape) * (rng.random(st["eo_count"].shape) < 0.8)
    st["eo_diff"] = (rng.normal(0, 2.5 * scale, st["eo_diff"].shape) * st["eo_count"]).astype(np.int32)
    st["eo_count"][::5, :, :, 1] = 1
    st["eo_diff"][::5, :, :, 1] = rng.integers(-40, 40, st["eo_diff"][::5, :, :, 1].shape)
    st["eo_count"][3::7] = st["eo_count"][2::7][:len(st["eo_count"][3::7])]          # neighbours with identical statistics: merge candidates tie
    st["eo_diff"][3::7] = st["eo_diff"][2::7][:len(st["eo_diff"][3::7])]
    for rates in ("typical", "free", "expensive"):
        t = rate_tables(rates)
        for lam, lam_c in LAMBDAS:
            want = run_decision(refdec, "svtref_sao_decision_picture", st, cols, rows, t, lam, lam_c, is10bit)
            got = run_decision(port, "port_sao_decision_picture", st, cols, rows, t, lam, lam_c, is10bit)
            assert_decisions_equal(got, want, f"random {rates} lambda {lam} 10bit={is10bit}")
        assert got["merge_left"].any() or got["merge_up"].any() or rates != "expensive"


def test_decision_port_basic_properties():
    """runs without the reference: an LCU with empty statistics is off, merge flags only with neighbours, offsets inside the HEVC ranges"""
    rng = np.random.default_rng(5)
    cols, rows = 5, 4
    st = np.zeros((cols * rows, 3), dtype=S.SAO_STATS_DTYPE)
    st["eo_count"][1:] = rng.integers(0, 900, st["eo_count"][1:].shape)
    st["eo_diff"][1:] = (rng.normal(0, 2, st["eo_diff"][1:].shape) * st["eo_count"][1:]).astype(np.int32)
    st["bo_count"][1:] = rng.integers(0, 300, st["bo_count"][1:].shape)
    st["bo_diff"][1:] = (rng.normal(0, 2, st["bo_diff"][1:].shape) * st["bo_count"][1:]).astype(np.int32)
    d = run_decision(port, "port_sao_decision_picture", st, cols, rows, rate_tables("typical"), 439162, 350000)
    assert d["type"][0].tolist() == [0, 0] and not d["offset"][0].any()
    assert not d["merge_left"][::cols].any() and not d["merge_up"][:cols].any()
    assert np.abs(d["offset"]).max() <= 7 and (d["type"] <= 5).all()
    eo = (d["type"][:, 0] >= 1) & (d["type"][:, 0] <= 4) & (d["merge_left"] == 0) & (d["merge_up"] == 0)
    assert (d["offset"][eo, 0, :2] >= 0).all() and (d["offset"][eo, 0, 2:] <= 0).all()      # EO categories 1, 2 raise / 3, 4 lower
