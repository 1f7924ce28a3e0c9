"""CPU tests: pin the plain-C SAO restatement (oracle/port/sao_port.c) against the UNMODIFIED reference: statistics through
its own gather dispatch table and application through its own (static) ApplySaoOffsetsPicture, reached by compiling
Codec/EbEncDecProcess.c in place into oracle/_ref/libsvtref_sao_*.so (oracle/refinc_sao_harness.c)."""
import ctypes as C
import os

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_mc import make_pictures

port = common.load_port()


def load_ref_sao():
    variant = "avx512" if common.cpu_has("avx512bw") else "avx2"
    common.load_ref(variant)
    so = os.path.join(common.ORACLE, "_ref", f"libsvtref_sao_{variant}.so")
    return C.CDLL(so) if os.path.exists(so) else None


refsao = load_ref_sao()
needs_ref = pytest.mark.skipif(refsao is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a: a.ctypes.data_as(C.c_void_p)


def make_sao_inputs(w, h, seed, kind="synth"):
    rng = np.random.default_rng(seed)
    src, rec = make_pictures(w, h, seed=seed, n=2, kind=kind)
    if kind == "synth":   # recon = source + coding noise
        for a, b in zip(rec.interior(), src.interior()):
            a[:] = np.clip(b.astype(int) + rng.integers(-6, 7, b.shape), 0, 255)
    n_lcu = ((w + 63) // 64) * ((h + 63) // 64)
    prm = np.zeros(n_lcu, dtype=S.SAO_PARAMS_DTYPE)
    prm["type"] = rng.integers(0, 6, (n_lcu, 2))
    prm["band_position"] = rng.integers(0, 29, (n_lcu, 3))
    prm["offset"] = rng.integers(-7, 8, (n_lcu, 3, 4))
    return src, rec, prm


SAO_CASES = [("synth_256x192", 256, 192, "synth"), ("synth_partial_200x136", 200, 136, "synth"), ("random_192x128", 192, 128, "random"),
             ("extremes_128x64", 128, 64, "extremes"), ("tiny_64x64", 64, 64, "synth")]


def run_stats(lib, name, src, rec, asm_mode=None):
    n_lcu = ((rec.width + 63) // 64) * ((rec.height + 63) // 64)
    out = np.zeros((n_lcu, 3), dtype=S.SAO_STATS_DTYPE)
    fn = getattr(lib, name); fn.restype = C.c_int
    args = [C.byref(src.struct), C.byref(rec.struct), vp(out)] + ([C.c_int(asm_mode)] if asm_mode is not None else [])
    assert fn(*args) == 0
    return out


def assert_stats_equal(a, b, what):
    """Only the 4 real EO categories are compared: slot [4] is scratch (the C row leaves a copy of [3] there, the SSE2 row 0)."""
    for f in ("bo_diff", "bo_count"):
        assert np.array_equal(a[f], b[f]), f"{what}: {f} differs"
    for f in ("eo_diff", "eo_count"):
        assert np.array_equal(a[f][..., :4], b[f][..., :4]), f"{what}: {f} differs"


@needs_ref
@pytest.mark.parametrize("case", SAO_CASES, ids=lambda c: c[0])
def test_sao_stats_port_vs_reference(case):
    name, w, h, kind = case
    src, rec, _ = make_sao_inputs(w, h, seed=w * 3 + h, kind=kind)
    want = run_stats(refsao, "svtref_sao_stats_picture", src, rec, asm_mode=0)
    got = run_stats(port, "port_sao_stats_picture", src, rec)
    assert_stats_equal(got, want, name + " port vs reference C row")
    if common.ref_asm_simd_ok():
        simd = run_stats(refsao, "svtref_sao_stats_picture", src, rec, asm_mode=1)
        assert_stats_equal(simd, want, name + " reference SIMD vs C row")


def run_apply_ref(rec, prm, flags, asm_mode):
    y, cb, cr = (a.copy() for a in rec.interior())
    out = S.HostYuvPicture(y, cb, cr)
    fn = refsao.svtref_sao_apply_picture; fn.restype = C.c_int
    assert fn(C.byref(out.struct), vp(prm), (C.c_uint8 * 2)(*flags), C.c_int(asm_mode)) == 0
    return out


def run_apply_port(rec, prm, flags):
    out = S.HostYuvPicture(*(np.zeros_like(a) for a in rec.interior()))
    fn = port.port_sao_apply_picture; fn.restype = C.c_int
    assert fn(C.byref(rec.struct), C.byref(out.struct), vp(prm), (C.c_uint8 * 2)(*flags)) == 0
    return out


# Pictures with a single LCU row or column are left out of the application tests: ApplySaoOffsetsLcu saves the first AND the
# last row (column) of such an LCU into the same scratch array (tmpY / tmp, Codec/EbEncDecProcess.c:291-330, :440-500), so the
# reference restores the wrong samples there -- an artefact no >= 2-LCU picture can hit.
SAO_APPLY_CASES = [c for c in SAO_CASES if c[1] > 64 and c[2] > 64] + [("extremes_192x128", 192, 128, "extremes")]


@needs_ref
@pytest.mark.parametrize("case", SAO_APPLY_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("flags", [(1, 1), (1, 0), (0, 1)])
def test_sao_apply_port_vs_reference(case, flags):
    name, w, h, kind = case
    _, rec, prm = make_sao_inputs(w, h, seed=w * 3 + h, kind=kind)
    want = run_apply_ref(rec, prm, flags, asm_mode=0)
    got = run_apply_port(rec, prm, flags)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs (port vs reference C rows): {np.argwhere(a != b)[:5]}"
    if common.ref_asm_simd_ok():
        simd = run_apply_ref(rec, prm, flags, asm_mode=1)
        for a, b, nm in zip(simd.interior(), want.interior(), "Y Cb Cr".split()):
            assert np.array_equal(a, b), f"{name}: {nm} differs (reference SIMD vs C rows)"


# ---- 16-bit samples (10-bit video) ---------------------------------------------------------------------------------------------
def make_sao_inputs16(w, h, seed, kind="synth"):
    rng = np.random.default_rng(seed + 5)
    src8, rec8, prm = make_sao_inputs(w, h, seed, kind)
    def up(p, noise):
        planes = [(a.astype(np.uint16) // 255) * 1023 if kind == "extremes" else np.clip(a.astype(int) * 4 + rng.integers(-noise, noise + 1, a.shape), 0, 1023).astype(np.uint16)
                  for a in p.interior()]
        return S.HostYuvPicture(*planes, dtype=np.uint16)
    return up(src8, 2), up(rec8, 9), prm


@needs_ref
@pytest.mark.parametrize("case", SAO_CASES, ids=lambda c: c[0])
def test_sao_stats16_port_vs_reference(case):
    name, w, h, kind = case
    src, rec, _ = make_sao_inputs16(w, h, seed=w * 3 + h, kind=kind)
    want = run_stats(refsao, "svtref_sao_stats_picture16", src, rec, asm_mode=0)
    got = run_stats(port, "port_sao_stats_picture16", src, rec)
    assert_stats_equal(got, want, name + " 16-bit port vs reference C row")
    if common.ref_asm_simd_ok():
        simd = run_stats(refsao, "svtref_sao_stats_picture16", src, rec, asm_mode=1)
        assert_stats_equal(simd, want, name + " 16-bit reference SIMD vs C row")


@needs_ref
@pytest.mark.parametrize("case", SAO_APPLY_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("flags", [(1, 1), (1, 0), (0, 1)])
def test_sao_apply16_port_vs_reference(case, flags):
    name, w, h, kind = case
    _, rec, prm = make_sao_inputs16(w, h, seed=w * 3 + h, kind=kind)
    outs = []
    for asm in ([0, 1] if common.ref_asm_simd_ok() else [0]):
        y, cb, cr = (a.copy() for a in rec.interior())
        o = S.HostYuvPicture(y, cb, cr, dtype=np.uint16)
        fn = refsao.svtref_sao_apply_picture16; fn.restype = C.c_int
        assert fn(C.byref(o.struct), vp(prm), (C.c_uint8 * 2)(*flags), C.c_int(asm)) == 0
        outs.append(o)
    got = S.HostYuvPicture(*(np.zeros_like(a) for a in rec.interior()), dtype=np.uint16)
    fn = port.port_sao_apply_picture16; fn.restype = C.c_int
    assert fn(C.byref(rec.struct), C.byref(got.struct), vp(prm), (C.c_uint8 * 2)(*flags)) == 0
    for o in outs:
        for a, b, nm in zip(got.interior(), o.interior(), "Y Cb Cr".split()):
            assert np.array_equal(a, b), f"{name}: {nm} differs (16-bit port vs reference): {np.argwhere(a != b)[:5]}"
