"""GPU parity tests for motion estimation: the CUDA path (through the C ABI) against the plain-C oracle and,
when oracle/_ref travelled to this box, against the reference's own MotionEstimateLcu / SadLoopKernel."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_me import SADLOOP_CASES, _me_cases, _sad_loop, build_me_case

pytestmark = pytest.mark.gpu
port = common.load_port()
ref = common.load_ref()


def test_library_sees_a_device():
    assert S.device_count() >= 1, "no CUDA device: the product has no CPU fallback"


@pytest.mark.parametrize("case", SADLOOP_CASES)
@pytest.mark.parametrize("kind", ["random", "constant", "plant"])
def test_sad_loop_leaf(case, kind):
    w, h, saw, sah = case
    rng = np.random.default_rng(hash((case, kind)) & 0xffff)
    raw = 256
    rows = sah + 2 * h + 4
    if kind == "constant":
        refbuf = np.full((rows, raw), 93, np.uint8)
        src = np.full((h, w), 93, np.uint8)
    else:
        refbuf = rng.integers(0, 256, (rows, raw), dtype=np.uint8)
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if kind == "plant":
            py, px = int(rng.integers(0, sah)), int(rng.integers(0, saw))
            src = np.ascontiguousarray(refbuf[py:py + 2 * h:2, 3 + px:3 + px + w])
    for roff in (0, 1, 2, 3):
        want = _sad_loop(port, "port_sad_loop", src, w, refbuf, roff, 2 * raw, h, w, raw, saw, sah)
        got = S.sad_loop_kernel(src, w, refbuf, roff, 2 * raw, h, w, raw, saw, sah, x0=-7, y0=-9)
        assert got == want


def test_nxm_and_avg_sad_leaf():
    rng = np.random.default_rng(5)
    u = C.c_uint32
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    port.port_sad_nxm.restype = port.port_sad_avg.restype = C.c_uint32
    for (w, h) in [(8, 8), (16, 16), (32, 32), (64, 64), (64, 32), (24, 8), (48, 16), (4, 4)]:
        a = rng.integers(0, 256, (h, 96), dtype=np.uint8)
        b = rng.integers(0, 256, (h, 80), dtype=np.uint8)
        c = rng.integers(0, 256, (h, 72), dtype=np.uint8)
        assert S.nxm_sad(a, 96, b, 80, h, w) == port.port_sad_nxm(vp(a), u(96), vp(b), u(80), u(h), u(w))
        assert S.combined_averaging_sad(a, 96, b, 80, c, 72, h, w) == port.port_sad_avg(vp(a), u(96), vp(b), u(80), vp(c), u(72), u(h), u(w))


@pytest.mark.parametrize("sa", [(8, 7), (16, 9), (5, 3), (13, 4), (64, 64), (24, 1), (1, 1), (127, 2)])
@pytest.mark.parametrize("kind", ["random", "constant", "plant", "saturate"])
def test_fullpel_search_leaf(sa, kind):
    saw, sah = sa
    rng = np.random.default_rng(saw * 131 + sah)
    stride = 224
    refbuf = rng.integers(0, 256, (sah + 70, stride), dtype=np.uint8)
    src = rng.integers(0, 256, (64, 64), dtype=np.uint8)
    if kind == "constant":
        refbuf[:] = 17
        src[:] = 17
    elif kind == "plant":
        py, px = int(rng.integers(0, sah)), int(rng.integers(0, saw))
        src = np.ascontiguousarray(refbuf[py:py + 64, 2 + px:2 + px + 64])
    elif kind == "saturate":     # every pixel differs by 255: 2*SAD64 == MAX_SAD_VALUE, the '<=' vs '<' corner
        refbuf[:] = 255
        src[:] = 0
    for roff in (2, 5):
        bs_w = np.full(85, 64 * 64 * 255, np.uint32)
        bm_w = np.arange(85, dtype=np.uint32)
        bs_g, bm_g = bs_w.copy(), bm_w.copy()
        port.port_fullpel_search_lcu(src.ctypes.data_as(C.c_void_p), C.c_uint32(64), C.c_void_p(refbuf.ctypes.data + roff), C.c_uint32(stride),
                                     C.c_int16(-3), C.c_int16(5), C.c_uint32(saw), C.c_uint32(sah),
                                     bs_w.ctypes.data_as(C.c_void_p), bm_w.ctypes.data_as(C.c_void_p))
        S.fullpel_search_lcu(src, 64, refbuf, roff, stride, -3, 5, saw, sah, bs_g, bm_g)
        assert np.array_equal(bs_g, bs_w)
        assert np.array_equal(bm_g, bm_w)


@pytest.mark.parametrize("case", _me_cases(), ids=lambda c: c[0])
def test_me_picture_cuda_vs_oracle(case):
    name, w, h, kw, content = case
    cfg, (src, r0, r1) = build_me_case(w, h, kw, content)
    r1 = r1 if cfg.num_lists == 2 else None
    want = common.oracle_me_picture(port, "port_me_picture", cfg, src, r0, r1)
    got = S.me_picture(cfg, src, r0, r1, want_raw=True, want_rc=True)
    common.assert_me_equal(got, want, cfg.num_lists, name + " CUDA vs port")
    if ref is not None:
        want_ref = common.oracle_me_picture(ref, "svtref_me_picture", cfg, src, r0, r1, asm_mode=1 if common.ref_asm_simd_ok() else 0)
        common.assert_me_equal(got, want_ref, cfg.num_lists, name + " CUDA vs reference MotionEstimateLcu")


@pytest.mark.parametrize("tl,nl", [(0, 2), (3, 2), (1, 1)])
def test_me_picture_4k_m10_full_size(tl, nl):
    """BASELINE configs[1] geometry: 3840x2160, preset M10 parameters, every LCU compared bit-exactly."""
    v = common.SynthVideo(3840, 2160, seed=0x5F7C0001)
    src, r0, r1 = (S.HostPaPicture(v.luma(n)) for n in (4, 0, 8))
    cfg = S.me_config(3840, 2160, enc_mode=10, temporal_layer=tl, num_lists=nl)
    r1 = r1 if nl == 2 else None
    got = S.me_picture(cfg, src, r0, r1, want_raw=True, want_rc=True)
    if ref is not None:
        want = common.oracle_me_picture(ref, "svtref_me_picture", cfg, src, r0, r1, asm_mode=1 if common.ref_asm_simd_ok() else 0)
    else:
        want = common.oracle_me_picture(port, "port_me_picture", cfg, src, r0, r1)
    common.assert_me_equal(got, want, nl, "4K M10 CUDA vs oracle")


@pytest.mark.parametrize("wh", [(448, 256), (328, 200), (72, 40)])
def test_device_picture_build_matches_host_pad_and_decimation(wh):
    w, h = wh
    luma = common.SynthVideo(w, h, seed=3).luma(2)
    host = S.HostPaPicture(luma)
    dev = S.DevPaPicture(w, h)
    dev.upload(luma)
    got = dev.download()
    for (a, pad, pw, ph), (b, _, _, _) in zip(host.planes, got.planes):
        assert np.array_equal(a, b[:, :a.shape[1]])
    dev.destroy()


@pytest.mark.parametrize("wh,pad", [((448, 256), 80), ((328, 200), 16), ((72, 40), 0), ((3840, 2160), 80)])
def test_device_picture_build_from_device_luma(wh, pad):
    """svtb200_pa_picture_build_dev: all three planes straight from a (padded) luma plane already on the device"""
    import torch
    w, h = wh
    luma = common.SynthVideo(w, h, seed=5).luma(1)
    host = S.HostPaPicture(luma)
    t = torch.zeros((h + 2 * pad, w + 2 * pad + (-(w + 2 * pad)) % 4), dtype=torch.uint8, device="cuda")
    t[pad:pad + h, pad:pad + w] = torch.from_numpy(luma).cuda()
    dev = S.DevPaPicture(w, h)
    dev.build_from_device(t[pad:, pad:].data_ptr(), t.shape[1], torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = dev.download()
    for (a, _, _, _), (b, _, _, _) in zip(host.planes, got.planes):
        assert np.array_equal(a, b[:, :a.shape[1]])
    dev.destroy()


def test_me_picture_dev_resident_matches_host_api():
    import torch
    w, h = 448, 256
    v = common.SynthVideo(w, h, seed=9)
    lum = [v.luma(n) for n in (4, 0, 8)]
    hp = [S.HostPaPicture(l) for l in lum]
    dp = [S.DevPaPicture(w, h) for _ in lum]
    for d, l in zip(dp, lum):
        d.upload(l)
    cfg = S.me_config(3840, 2160, enc_mode=10, temporal_layer=2, num_lists=2)
    cfg.pic_width, cfg.pic_height = w, h
    n_lcu = 7 * 4
    out = torch.zeros(n_lcu * 85 * 24, dtype=torch.uint8, device="cuda")
    S.me_picture_dev(cfg, dp[0], dp[1], dp[2], out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(S.ME_PU_RESULT_DTYPE).reshape(n_lcu, 85)
    want = S.me_picture(cfg, hp[0], hp[1], hp[2])[0]
    assert np.array_equal(got, want)
