"""which deblocking state do the SAO statistics of an LCU see?  The reference gathers them right after that LCU's own deblocking step,
before the LCUs to its right and below exist (Codec/EbCodingLoop.c:4600-4760).  oracle/ref_dlf_harness.c (svtref_dlf_sao_stats_picture) runs exactly
that interleaving with the reference's own LCU drivers and gather kernels; the plain-C oracle restates the state as a rule (final deblocked samples,
except the last 4 columns / rows of an LCU that has a right / bottom neighbour, which are still unfiltered) -- these tests pin the rule."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_dlf import DLF_CASES, make_dlf_inputs, run_dlf
from test_oracle_sao import assert_stats_equal, run_stats

port = common.load_port()
ref = common.load_ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a: a.ctypes.data_as(C.c_void_p)
CASES = DLF_CASES + [("big_synth_448x320", 448, 320, "synth", 1.0), ("extremes_320x256", 320, 256, "extremes", 1.0), ("ragged_456x328", 456, 328, "synth", 1.0)]


def noisy_source(pic, seed):
    rng = np.random.default_rng(seed)
    return S.HostYuvPicture(*[np.clip(a.astype(np.int16) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8) for a in pic.interior()])


def reference_interleaved(pic, src, vbs, hbs, qp, prm, asm_mode):
    work = S.HostYuvPicture(*(a.copy() for a in pic.interior()))
    snap = S.HostYuvPicture(*(np.zeros_like(a) for a in pic.interior()))
    n = ((prm.pic_width + 63) // 64) * ((prm.pic_height + 63) // 64)
    st = np.zeros((n, 3), dtype=S.SAO_STATS_DTYPE)
    fn = ref.svtref_dlf_sao_stats_picture
    fn.restype = C.c_int
    assert fn(C.byref(work.struct), C.byref(src.struct), vp(vbs), vp(hbs), vp(qp), C.byref(prm), vp(st), C.byref(snap.struct), C.c_int(asm_mode)) == 0
    return work, snap, st


def port_h3(pic, src, vbs, hbs, qp, prm):
    final = run_dlf(port, "port_dlf_picture", pic, vbs, hbs, qp, prm)
    scratch = S.HostYuvPicture(*(np.zeros_like(a) for a in pic.interior()))
    n = ((prm.pic_width + 63) // 64) * ((prm.pic_height + 63) // 64)
    st = np.zeros((n, 3), dtype=S.SAO_STATS_DTYPE)
    fn = port.port_sao_stats_picture_h3
    fn.restype = C.c_int
    assert fn(C.byref(src.struct), C.byref(final.struct), C.byref(pic.struct), C.byref(scratch.struct), vp(st)) == 0
    return final, scratch, st


@needs_ref
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_gather_time_state_and_statistics_port_vs_reference(case):
    name, w, h, kind, density = case
    pic, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    src = noisy_source(pic, w * 7 + h)
    work, snap, want = reference_interleaved(pic, src, vbs, hbs, qp, prm, asm_mode=0)
    final, composite, got = port_h3(pic, src, vbs, hbs, qp, prm)
    for a, b, nm in zip(work.interior(), final.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: final {nm} of the interleaved run differs from the picture-order deblocking"
    for a, b, nm in zip(snap.interior(), composite.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: gather-time state of {nm} differs from the rule at {np.argwhere(a != b)[:4]}"
    assert_stats_equal(got, want, name + " H3 statistics port vs reference")
    if common.ref_asm_simd_ok():
        _, _, simd = reference_interleaved(pic, src, vbs, hbs, qp, prm, asm_mode=1)
        assert_stats_equal(simd, want, name + " reference SIMD vs C rows")


@needs_ref
def test_h3_matters():
    """the statistics on the fully deblocked picture are NOT what the encoder gathers (so a picture-order pipeline needs the composite)"""
    name, w, h, kind, density = "big_synth_448x320", 448, 320, "synth", 1.0
    pic, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    src = noisy_source(pic, 5)
    work, _, want = reference_interleaved(pic, src, vbs, hbs, qp, prm, asm_mode=0)
    naive = run_stats(port, "port_sao_stats_picture", src, work)
    assert not np.array_equal(naive["eo_count"][..., :4], want["eo_count"][..., :4])


def test_composite_touches_only_the_strips():
    """reference-free: the composite equals the deblocked picture outside, the unfiltered one inside, the right / bottom 4-sample strips"""
    w, h = 200, 136
    pic, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=3, density=1.0, kind="synth")
    final, comp, _ = port_h3(pic, noisy_source(pic, 1), vbs, hbs, qp, prm)
    for c, (f, u, o) in enumerate(zip(final.interior(), pic.interior(), comp.interior())):
        L = 64 >> (1 if c else 0)
        ph, pw = f.shape
        m = np.zeros_like(f, dtype=bool)
        for y0 in range(0, ph, L):
            for x0 in range(0, pw, L):
                x1, y1 = min(x0 + L, pw), min(y0 + L, ph)
                if x1 < pw:
                    m[y0:y1, x1 - 4:x1] = True
                if y1 < ph:
                    m[y1 - 4:y1, x0:x1] = True
        assert np.array_equal(o, np.where(m, u, f))


# ---- 16-bit sample planes (10-bit video): LCU*DLFCore16bit + GatherSaoStatisticsLcu_62x62_16bit ----------------------------------------------
@needs_ref
@pytest.mark.parametrize("case", CASES[:5] + CASES[-1:], ids=lambda c: c[0])
def test_gather_time_state_and_statistics16_port_vs_reference(case):
    from test_oracle_dlf import run_dlf16, to16
    name, w, h, kind, density = case
    pic8, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    pic = to16(pic8, seed=w, extremes=(kind == "extremes"))
    rng = np.random.default_rng(w * 7 + h)
    src = S.HostYuvPicture(*[np.clip(a.astype(np.int32) + rng.integers(-30, 31, a.shape), 0, 1023).astype(np.uint16) for a in pic.interior()], dtype=np.uint16)
    n = ((w + 63) // 64) * ((h + 63) // 64)
    # the reference's interleaved run
    work = S.HostYuvPicture(*(a.copy() for a in pic.interior()), dtype=np.uint16)
    snap = S.HostYuvPicture(*(np.zeros_like(a) for a in pic.interior()), dtype=np.uint16)
    want = np.zeros((n, 3), dtype=S.SAO_STATS_DTYPE)
    fn = ref.svtref_dlf_sao_stats_picture16
    fn.restype = C.c_int
    assert fn(C.byref(work.struct), C.byref(src.struct), vp(vbs), vp(hbs), vp(qp), C.byref(prm), vp(want), C.byref(snap.struct), C.c_int(0)) == 0
    # the oracle: picture-order deblocking, composite, statistics
    final = run_dlf16(port, "port_dlf_picture16", pic, vbs, hbs, qp, prm)
    scratch = S.HostYuvPicture(*(np.zeros_like(a) for a in pic.interior()), dtype=np.uint16)
    got = np.zeros((n, 3), dtype=S.SAO_STATS_DTYPE)
    pf = port.port_sao_stats_picture16_h3
    pf.restype = C.c_int
    assert pf(C.byref(src.struct), C.byref(final.struct), C.byref(pic.struct), C.byref(scratch.struct), vp(got)) == 0
    for a, b, nm in zip(work.interior(), final.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: final {nm} (16-bit) differs"
    for a, b, nm in zip(snap.interior(), scratch.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: 16-bit gather-time state of {nm} differs from the rule at {np.argwhere(a != b)[:4]}"
    assert_stats_equal(got, want, name + " 16-bit H3 statistics port vs reference")
