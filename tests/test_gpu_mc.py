"""GPU parity tests for motion-compensated prediction (through the C ABI) against the plain-C oracle and the
reference's EncodePassInterPrediction."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_mc import MC_CASES, make_jobs, make_pictures, run_oracle

pytestmark = pytest.mark.gpu
port = common.load_port()
ref = common.load_ref()
vp = lambda a, off=0: C.c_void_p(a.ctypes.data + off)
u = C.c_uint32


@pytest.mark.parametrize("size", [(8, 8), (16, 16), (32, 32), (64, 64), (64, 32), (8, 32), (24, 16)])
def test_luma_leaf_all_16_positions(size):
    w, h = size
    rng = np.random.default_rng(w * 100 + h)
    stride = 96
    refbuf = rng.integers(0, 256, (h + 16, stride), dtype=np.uint8)
    refbuf[::3, ::5] = 255
    refbuf[1::4, 2::3] = 0
    for frac in range(16):
        for off in (4 * stride + 5, 4 * stride + 8):
            want = np.zeros((h, w), np.uint8)
            port.port_luma_uni(u(frac), vp(refbuf, off), u(stride), vp(want), u(w), u(w), u(h))
            got = S.luma_interp(frac, refbuf, off, stride, w, h)
            assert np.array_equal(got, want), f"luma frac {frac}"


@pytest.mark.parametrize("size", [(4, 4), (8, 8), (16, 16), (32, 32), (32, 16), (4, 16)])
def test_chroma_leaf_all_64_positions(size):
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    stride = 80
    refbuf = rng.integers(0, 256, (h + 8, stride), dtype=np.uint8)
    for fy in range(8):
        for fx in range(8):
            off = 2 * stride + 3 + (fx & 1)
            want = np.zeros((h, w), np.uint8)
            port.port_chroma_uni(u(fx), u(fy), vp(refbuf, off), u(stride), vp(want), u(w), u(w), u(h))
            got = S.chroma_interp(fx, fy, refbuf, off, stride, w, h)
            assert np.array_equal(got, want), f"chroma frac {fx},{fy}"


def test_bipred_leafs_all_position_pairs():
    rng = np.random.default_rng(3)
    stride = 96
    a = rng.integers(0, 256, (48, stride), dtype=np.uint8)
    b = rng.integers(0, 256, (48, stride), dtype=np.uint8)
    a[::2, ::2] = 255
    b[::2, ::2] = 255      # saturating sums exercise the final clip
    for f0 in range(16):
        for f1 in range(16):
            want = np.zeros((16, 16), np.uint8)
            port.port_luma_bi(u(f0), vp(a, 5 * stride + 7), u(stride), u(f1), vp(b, 4 * stride + 6), u(stride), vp(want), u(16), u(16), u(16))
            got = S.luma_bipred(f0, a, 5 * stride + 7, stride, f1, b, 4 * stride + 6, stride, 16, 16)
            assert np.array_equal(got, want), f"luma bi {f0},{f1}"
    for k in range(0, 64 * 64, 37):
        f0x, f0y, f1x, f1y = k & 7, (k >> 3) & 7, (k >> 6) & 7, (k >> 9) & 7
        want = np.zeros((8, 8), np.uint8)
        port.port_chroma_bi(u(f0x), u(f0y), vp(a, 3 * stride + 5), u(stride), u(f1x), u(f1y), vp(b, 2 * stride + 4), u(stride), vp(want), u(8), u(8), u(8))
        got = S.chroma_bipred(f0x, f0y, a, 3 * stride + 5, stride, f1x, f1y, b, 2 * stride + 4, stride, 8, 8)
        assert np.array_equal(got, want), f"chroma bi {k}"


@pytest.mark.parametrize("case", MC_CASES, ids=lambda c: c[0])
def test_mc_batch_cuda_vs_oracle(case):
    name, w, h, kind, all_fracs, big = case
    refs = make_pictures(w, h, seed=11, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=5, all_fracs=all_fracs, big_mv=big)
    want = run_oracle(port, "port_mc_batch", refs, jobs, w, h)
    dst = S.HostYuvPicture(np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8))
    S.mc_batch(refs, jobs, dst)
    for a, b, nm in zip(dst.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs (CUDA vs port)"
    if ref is not None:
        wr = run_oracle(ref, "svtref_mc_batch", refs, jobs, w, h, asm_mode=1 if common.ref_asm_simd_ok() else 0)
        for a, b, nm in zip(dst.interior(), wr.interior(), "Y Cb Cr".split()):
            assert np.array_equal(a, b), f"{name}: {nm} differs (CUDA vs reference EncodePassInterPrediction)"
    # nothing outside the PU areas may be touched (padding stays zero)
    assert dst.y[:80].sum() == 0 and dst.cb[:40].sum() == 0


def test_mc_batch_4k_full_frame():
    """BASELINE configs[1] geometry: every CTU of a 3840x2160 frame predicted (mixed PU sizes / directions / phases)."""
    w, h = 3840, 2160
    refs = make_pictures(w, h, seed=21, kind="synth", n=2)
    jobs = make_jobs(w, h, len(refs), seed=9, all_fracs=True)
    dst = S.HostYuvPicture(np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8))
    S.mc_batch(refs, jobs, dst)
    if ref is not None:
        want = run_oracle(ref, "svtref_mc_batch", refs, jobs, w, h, asm_mode=1 if common.ref_asm_simd_ok() else 0)
    else:
        want = run_oracle(port, "port_mc_batch", refs, jobs, w, h)
    for a, b, nm in zip(dst.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"4K: {nm} differs"


# ---- 16-bit samples (10-bit video), SURVEY §8 row b5 ---------------------------------------------------------------------
@pytest.mark.parametrize("case", MC_CASES, ids=lambda c: c[0])
def test_mc_batch16_cuda_vs_oracle(case):
    from test_oracle_mc import make_pictures16, run_oracle16
    name, w, h, kind, all_fracs, big_mv = case
    refs = make_pictures16(w, h, seed=w + h, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=w, all_fracs=all_fracs, big_mv=big_mv)
    dst = S.HostYuvPicture(np.zeros((h, w), np.uint16), np.zeros((h // 2, w // 2), np.uint16), np.zeros((h // 2, w // 2), np.uint16), dtype=np.uint16)
    S.mc_batch16(refs, jobs, dst)
    want = run_oracle16(port, "port_mc_batch16", refs, jobs, w, h)
    for a, b, pl in zip(dst.interior(), want.interior(), "y cb cr".split()):
        assert np.array_equal(a, b), (name, pl, np.argwhere(a != b)[:4])
    if ref is not None:
        wr = run_oracle16(ref, "svtref_mc_batch16", refs, jobs, w, h, asm_mode=1 if common.ref_asm_simd_ok() else 0)
        for a, b in zip(dst.interior(), wr.interior()):
            assert np.array_equal(a, b), name


def test_mc_batch16_integer_mvs_and_alignment():
    """integer-pel jobs (copy / average shortcut) and pictures whose planes sit at odd sample offsets inside their arrays"""
    from test_oracle_mc import make_pictures16, run_oracle16
    w, h = 256, 128
    refs = make_pictures16(w, h, seed=5, kind="random")
    jobs = make_jobs(w, h, len(refs), seed=9)
    jobs["mv"] &= ~3
    dst = S.HostYuvPicture(np.zeros((h, w), np.uint16), np.zeros((h // 2, w // 2), np.uint16), np.zeros((h // 2, w // 2), np.uint16), dtype=np.uint16)
    S.mc_batch16(refs, jobs, dst)
    want = run_oracle16(port, "port_mc_batch16", refs, jobs, w, h)
    for a, b in zip(dst.interior(), want.interior()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("case", MC_CASES + [("uhd_3840x2160", 3840, 2160, "synth", True, False)], ids=lambda c: c[0])
def test_mc_ifree_cuda_vs_oracle(case):
    """row b6: the interpolation-free prediction of the mode-decision fast loop (UniPredIFreeRef8Bit / BiPredIFreeRef8Bit)"""
    name, w, h, kind, all_fracs, big = case
    refs = make_pictures(w, h, seed=13, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=6, all_fracs=all_fracs, big_mv=big)
    if ref is not None:
        want = run_oracle(ref, "svtref_mc_ifree_batch", refs, jobs, w, h, asm_mode=1 if common.ref_asm_simd_ok() else 0)
    else:
        want = run_oracle(port, "port_mc_ifree_batch", refs, jobs, w, h)
    got = S.HostYuvPicture(np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8))
    S.mc_ifree_batch(refs, jobs, got)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs: {np.argwhere(a != b)[:5]}"
