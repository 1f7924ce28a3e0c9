"""CPU tests: pin the plain-C ME restatement (oracle/port/me_port.c) against the UNMODIFIED reference
compiled in place (oracle/_ref): leaf kernels against both the C_DEFAULT row and the SIMD rows
(the reference's own asm_test policy, Tests/SVT-HEVC_FunctionalTests.py:830-853, applied per kernel), and the
picture driver against the reference's own MotionEstimateLcu (oracle/ref_me_harness.c)."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S

ref = common.load_ref()
port = common.load_port()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")


def _sad_loop(lib, name, src, sstride, refbuf, roff, rstride, h, w, raw, saw, sah):
    best = C.c_uint64(0)
    xs, ys = C.c_int16(-7), C.c_int16(-9)
    f = getattr(lib, name)
    f.restype = None
    f(src.ctypes.data_as(C.c_void_p), C.c_uint32(sstride), C.c_void_p(refbuf.ctypes.data + roff), C.c_uint32(rstride),
      C.c_uint32(h), C.c_uint32(w), C.byref(best), C.byref(xs), C.byref(ys), C.c_uint32(raw), C.c_int16(saw), C.c_int16(sah))
    return best.value, xs.value, ys.value


SADLOOP_CASES = [  # (w, h, saw, sah)  -- HME L0 (16x8), L1 (32x16), L2 (64x32) block shapes + odd ones
    (16, 8, 64, 32), (16, 8, 128, 64), (16, 8, 80, 44), (16, 8, 32, 22), (16, 8, 16, 1), (16, 8, 37, 5),
    (32, 16, 8, 4), (32, 16, 16, 8), (32, 16, 13, 3), (64, 32, 8, 4), (64, 32, 8, 2), (64, 32, 24, 7),
    (8, 4, 24, 9), (4, 2, 9, 9), (24, 12, 8, 8), (48, 24, 16, 3),
]


@needs_ref
@pytest.mark.parametrize("case", SADLOOP_CASES)
@pytest.mark.parametrize("kind", ["random", "constant", "plant", "smooth"])
def test_sad_loop_port_vs_reference(case, kind):
    w, h, saw, sah = case
    rng = np.random.default_rng(hash((case, kind)) & 0xffff)
    raw = 256
    rows = sah + 2 * h + 4
    if kind == "constant":
        refbuf = np.full((rows, raw), 93, np.uint8)
        src = np.full((h, w), 93, np.uint8)
    elif kind == "smooth":
        refbuf = (np.add.outer(np.arange(rows) * 2, np.arange(raw)) & 255).astype(np.uint8)
        src = np.ascontiguousarray(refbuf[3:3 + 2 * h:2, 5:5 + w])
    else:
        refbuf = rng.integers(0, 256, (rows, raw), dtype=np.uint8)
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if kind == "plant":
            py, px = int(rng.integers(0, sah)), int(rng.integers(0, saw))
            src = np.ascontiguousarray(refbuf[py:py + 2 * h:2, 3 + px:3 + px + w])
    roff = 3   # deliberately unaligned start
    want = _sad_loop(ref, "SadLoopKernel", src, w, refbuf, roff, 2 * raw, h, w, raw, saw, sah)
    got = _sad_loop(port, "port_sad_loop", src, w, refbuf, roff, 2 * raw, h, w, raw, saw, sah)
    assert got == want
    if common.ref_asm_simd_ok() and w in (4, 8, 16, 24, 32, 48, 64):
        simd = _sad_loop(ref, "SadLoopKernel_AVX2_INTRIN", src, w, refbuf, roff, 2 * raw, h, w, raw, saw, sah)
        assert simd == want, "reference C row != AVX2 row"


@needs_ref
def test_nxm_and_avg_sad_port_vs_reference():
    rng = np.random.default_rng(5)
    for (w, h) in [(8, 8), (16, 16), (32, 32), (64, 64), (64, 32), (24, 8), (48, 16)]:
        a = rng.integers(0, 256, (h, 96), dtype=np.uint8)
        b = rng.integers(0, 256, (h, 80), dtype=np.uint8)
        c = rng.integers(0, 256, (h, 72), dtype=np.uint8)
        for lib, n1, n2 in ((ref, "FastLoop_NxMSadKernel", "CombinedAveragingSAD"), (port, "port_sad_nxm", "port_sad_avg")):
            f1, f2 = getattr(lib, n1), getattr(lib, n2)
            f1.restype = f2.restype = C.c_uint32
        vp = lambda x: x.ctypes.data_as(C.c_void_p)
        u = C.c_uint32
        r1 = ref.FastLoop_NxMSadKernel(vp(a), u(96), vp(b), u(80), u(h), u(w))
        p1 = port.port_sad_nxm(vp(a), u(96), vp(b), u(80), u(h), u(w))
        r2 = ref.CombinedAveragingSAD(vp(a), u(96), vp(b), u(80), vp(c), u(72), u(h), u(w))
        p2 = port.port_sad_avg(vp(a), u(96), vp(b), u(80), vp(c), u(72), u(h), u(w))
        assert (r1, r2) == (p1, p2)
        assert r1 == int(np.abs(a[:, :w].astype(int) - b[:, :w].astype(int)).sum())


def _me_cases():
    """(name, width, height, cfg kwargs, content)"""
    return [
        ("m10_4k_style_P_tl0", 448, 256, dict(enc_mode=10, temporal_layer=0, num_lists=1, force_res="4k"), "synth"),
        ("m10_4k_style_B_tl1", 448, 256, dict(enc_mode=10, temporal_layer=1, num_lists=2, force_res="4k"), "synth"),
        ("m10_4k_style_B_tl3", 384, 240, dict(enc_mode=10, temporal_layer=3, num_lists=2, force_res="4k"), "synth"),
        ("m10_4k_style_B_tl0", 320, 192, dict(enc_mode=10, temporal_layer=0, num_lists=2, force_res="4k"), "synth"),
        ("m7_4k_style_B_tl2", 384, 256, dict(enc_mode=7, temporal_layer=2, num_lists=2, force_res="4k"), "synth"),
        ("m7_1080_style_B_tl1_samepoc", 320, 192, dict(enc_mode=7, temporal_layer=1, num_lists=2, force_res="1080p", ref_poc=(8, 8)), "synth"),
        ("m10_1080_style_P_rc", 328, 200, dict(enc_mode=10, temporal_layer=2, num_lists=1, force_res="1080p", rate_control_mode=1), "synth"),
        ("m10_constant", 256, 192, dict(enc_mode=10, temporal_layer=1, num_lists=2, force_res="4k"), "constant"),
        ("m10_random", 256, 192, dict(enc_mode=10, temporal_layer=2, num_lists=2, force_res="4k"), "random"),
        ("m10_checker", 256, 128, dict(enc_mode=10, temporal_layer=1, num_lists=2, force_res="4k"), "checker"),
        ("m7_gradient", 256, 192, dict(enc_mode=7, temporal_layer=0, num_lists=2, force_res="4k"), "gradient"),
        ("m10_bigmotion_B_tl0", 384, 256, dict(enc_mode=10, temporal_layer=0, num_lists=2, force_res="4k"), "bigmotion"),
        ("m10_bigmotion_B_tl2_1080", 320, 192, dict(enc_mode=10, temporal_layer=2, num_lists=2, force_res="1080p"), "bigmotion"),
        ("m7_bigmotion_B_tl0_samepoc", 320, 256, dict(enc_mode=7, temporal_layer=0, num_lists=2, force_res="1080p", ref_poc=(4, 4)), "bigmotion"),
        ("m10_tiny_64x64", 64, 64, dict(enc_mode=10, temporal_layer=1, num_lists=2, force_res="4k"), "synth"),
        ("m10_partial_lcus_72x40", 72, 40, dict(enc_mode=10, temporal_layer=1, num_lists=2, force_res="4k"), "synth"),
        ("m7_partial_lcus_200x136", 200, 136, dict(enc_mode=7, temporal_layer=1, num_lists=2, force_res="1080p", lambda_=1000), "bigmotion"),
        # sub-pel refinement (half + quarter pel, fractional bi-prediction): every PU / selective per temporal layer, both SAD methods, 8x8 PUs on
        ("m7_subpel_all_fullsad_8x8", 256, 192, dict(enc_mode=7, temporal_layer=0, num_lists=2, force_res="4k", frac_model=0, frac_method=1, cu8x8=0), "synth"),
        ("m10_subpel_sel_8x8_tl0", 320, 192, dict(enc_mode=10, temporal_layer=0, num_lists=2, force_res="1080p", cu8x8=0), "bigmotion"),
        ("m10_subpel_P_tl0", 256, 128, dict(enc_mode=10, temporal_layer=0, num_lists=1, force_res="4k"), "synth"),
        ("m7_subpel_sel_tl1", 256, 192, dict(enc_mode=7, temporal_layer=1, num_lists=2, force_res="1080p", use_subpel=True, cu8x8=0), "synth"),
        ("m7_subpel_sel_tl2", 256, 192, dict(enc_mode=7, temporal_layer=2, num_lists=2, force_res="1080p", use_subpel=True, cu8x8=0), "random"),
        ("m7_subpel_sel_tl3_noise", 192, 192, dict(enc_mode=7, temporal_layer=3, num_lists=2, force_res="4k", use_subpel=True, cu8x8=0), "noisy"),
        ("m10_subpel_all_partial_192x136", 192, 136, dict(enc_mode=10, temporal_layer=0, num_lists=2, force_res="4k", frac_model=0, cu16x16=1), "bigmotion"),
        ("m10_subpel_flat", 128, 128, dict(enc_mode=10, temporal_layer=0, num_lists=2, force_res="4k", frac_model=0, cu8x8=0), "constant"),
    ]


def build_me_case(width, height, kw, content):
    kw = dict(kw)
    res = kw.pop("force_res")
    over = {k: kw.pop(k) for k in ("frac_model", "frac_method", "cu8x8", "cu16x16") if k in kw}
    cfg = S.me_config(3840 if res == "4k" else 1920, 2160 if res == "4k" else 1080, **kw)
    cfg.pic_width, cfg.pic_height = width, height
    if "frac_model" in over:
        cfg.fractional_search_model = over["frac_model"]
    if "frac_method" in over:
        cfg.fractional_search_method = over["frac_method"]
    if "cu8x8" in over:
        cfg.cu8x8_mode = over["cu8x8"]
    if "cu16x16" in over:
        cfg.cu16x16_mode = over["cu16x16"]
    if content == "synth":
        v = common.SynthVideo(width, height, seed=1234 + width)
        frames = [v.luma(4), v.luma(0), v.luma(8)]
    elif content == "bigmotion":   # global translations that push search windows across the picture edges
        v = common.SynthVideo(width + 256, height + 256, seed=77 + width)
        big = v.luma(0)
        cut = lambda dx, dy: np.ascontiguousarray(big[128 + dy:128 + dy + height, 128 + dx:128 + dx + width])
        frames = [cut(0, 0), cut(-61, -29), cut(45, 38)]
    elif content == "noisy":       # sub-pel shifted smooth texture + noise: half / quarter positions win often
        v = common.SynthVideo(2 * width + 64, 2 * height + 64, seed=5 + width, texture_box=6)
        big = v.luma(0).astype(np.float64)
        rng = np.random.default_rng(3)
        def cut(dx, dy):       # dx, dy in half samples of the 2x oversampled picture = quarter-ish shifts after decimation
            a = big[16 + dy:16 + dy + 2 * height:2, 16 + dx:16 + dx + 2 * width:2] + big[17 + dy:17 + dy + 2 * height:2, 17 + dx:17 + dx + 2 * width:2]
            return np.clip(np.rint(a / 2 + rng.normal(0, 1.0, (height, width))), 0, 255).astype(np.uint8)
        frames = [cut(0, 0), cut(5, 3), cut(-3, 7)]
    else:
        frames = [common.adversarial_luma(content, width, height, seed=s) for s in (1, 2, 3)]
        if content == "random":   # make it a (noisy) translation so that ME finds structure
            base = np.pad(frames[0], 16, mode="wrap")
            frames[1] = np.ascontiguousarray(base[16 + 3:16 + 3 + height, 16 - 5:16 - 5 + width])
            frames[2] = np.ascontiguousarray(base[16 - 2:16 - 2 + height, 16 + 7:16 + 7 + width])
    pics = [S.HostPaPicture(f) for f in frames]
    return cfg, pics


@needs_ref
@pytest.mark.parametrize("case", _me_cases(), ids=lambda c: c[0])
def test_me_picture_port_vs_reference(case):
    name, w, h, kw, content = case
    cfg, (src, r0, r1) = build_me_case(w, h, kw, content)
    r1 = r1 if cfg.num_lists == 2 else None
    want_c = common.oracle_me_picture(ref, "svtref_me_picture", cfg, src, r0, r1, asm_mode=0)
    got = common.oracle_me_picture(port, "port_me_picture", cfg, src, r0, r1)
    common.assert_me_equal(got, want_c, cfg.num_lists, name + " port vs reference C rows")
    if common.ref_asm_simd_ok():
        want_simd = common.oracle_me_picture(ref, "svtref_me_picture", cfg, src, r0, r1, asm_mode=1)
        common.assert_me_equal(want_simd, want_c, cfg.num_lists, name + " reference SIMD vs C rows")
