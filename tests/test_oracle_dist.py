"""CPU tests: pin the plain-C restatement of mode decision's distortion leaves (oracle/port/dist_port.c) against the UNMODIFIED reference functions --
the C row and the SSE rows the encoder dispatches (Codec/EbPictureOperators.h:520-660) -- called by name in oracle/_ref/libsvtref_*.so."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_mc import make_pictures

port = common.load_port()
ref = common.load_ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a, off=0: C.c_void_p(a.ctypes.data + off)

port.port_spatial_distortion.restype = C.c_uint64
port.port_satd8x8.restype = port.port_satd4x4.restype = C.c_uint64

SIZES = [4, 8, 16, 32, 64]


def make_dist_jobs(w, h, seed, n=300, kinds=(0,), unit_of=lambda k: 4):
    """blocks of every size on every plane, inside the picture, origins multiples of the unit"""
    rng = np.random.default_rng(seed)
    jobs = np.zeros(n, S.DIST_JOB_DTYPE)
    for j in range(n):
        pl = int(rng.integers(0, 3))
        pw, ph = (w, h) if pl == 0 else (w // 2, h // 2)
        kind = int(rng.choice(kinds))
        u = unit_of(kind)
        bw = int(rng.choice([s for s in SIZES if u <= s <= pw]))
        bh = min(bw, max(s for s in SIZES if s <= ph)) if rng.random() < 0.7 else int(rng.choice([s for s in SIZES if u <= s <= ph]))
        jobs[j] = (int(rng.integers(0, (pw - bw) // u + 1)) * u, int(rng.integers(0, (ph - bh) // u + 1)) * u, pl, bw, bh, kind)
    return jobs


def make_coeff_pictures(w, h, seed, extreme=False):
    rng = np.random.default_rng(seed)
    c, r = S.HostCoeffPicture(w, h), S.HostCoeffPicture(w, h)
    for a, b in zip(c.planes, r.planes):
        lim = 32767 if extreme else 1500
        a[:] = rng.integers(-lim - (1 if extreme else 0), lim + 1, a.shape)
        b[:] = np.clip(a.astype(int) + rng.integers(-40, 41, a.shape), -32768, 32767) if not extreme else rng.integers(-32768, 32768, a.shape)
    return c, r


def run_port_batch(name, a, b, jobs, per_job=1):
    out = np.zeros((len(jobs), per_job), np.uint64)
    fn = getattr(port, name); fn.restype = C.c_int
    assert fn(C.byref(a.struct), C.byref(b.struct), vp(jobs), C.c_int(len(jobs)), vp(out)) == 0
    return out if per_job > 1 else out[:, 0]


@needs_ref
@pytest.mark.parametrize("kind", ["synth", "random", "extremes"])
def test_spatial_distortion_port_vs_reference(kind):
    w, h = 192, 128
    a, b = make_pictures(w, h, seed=3, n=2, kind=kind)
    ya, yb = a.interior()[0], b.interior()[0]
    ref.SpatialFullDistortionKernel.restype = C.c_uint64
    simd = {4: "SpatialFullDistortionKernel4x4_SSSE3_INTRIN", 8: "SpatialFullDistortionKernel8x8_SSSE3_INTRIN", 16: "SpatialFullDistortionKernel16MxN_SSSE3_INTRIN",
            32: "SpatialFullDistortionKernel16MxN_SSSE3_INTRIN", 64: "SpatialFullDistortionKernel16MxN_SSSE3_INTRIN"}
    for n in SIZES:
        for (x, y) in [(0, 0), (64, 32), (w - n, h - n)]:
            sa, sb = ya[y:, x:], yb[y:, x:]
            args = (vp(ya, y * ya.strides[0] + x), C.c_uint32(ya.strides[0]), vp(yb, y * yb.strides[0] + x), C.c_uint32(yb.strides[0]), C.c_uint32(n), C.c_uint32(n))
            want = ref.SpatialFullDistortionKernel(*args)
            assert port.port_spatial_distortion(*args) == want
            assert want == int(((sa[:n, :n].astype(np.int64) - sb[:n, :n]) ** 2).sum())
            if common.ref_asm_simd_ok():
                f = getattr(ref, simd[n]); f.restype = C.c_uint64
                assert f(*args) == want, (n, x, y)


@needs_ref
@pytest.mark.parametrize("extreme", [False, True])
def test_full_distortion_port_vs_reference(extreme):
    w, h = 128, 64
    c, r = make_coeff_pictures(w, h, seed=5, extreme=extreme)
    a, b = c.planes[0], r.planes[0]
    names = {S.FULLDIST_DEFAULT: ("FullDistortionKernel_32bit", "FullDistortionKernel%s_32bit_BT_SSE2"), S.FULLDIST_INTRA: ("FullDistortionKernelIntra_32bit", "FullDistortionKernelIntra%s_32bit_BT_SSE2"),
             S.FULLDIST_CBF_ZERO: ("FullDistortionKernelCbfZero_32bit", "FullDistortionKernelCbfZero%s_32bit_BT_SSE2")}
    port.port_full_distortion.restype = None
    for kind, (cname, sname) in names.items():
        for n in [4, 8, 16, 32]:
            for (x, y) in [(0, 0), (32, 16), (w - n, h - n)]:
                want, got, gs = (C.c_uint64 * 2)(), (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
                args = (vp(a, (y * w + x) * 2), C.c_uint32(w), vp(b, (y * w + x) * 2), C.c_uint32(w))
                getattr(ref, cname)(*args, want, C.c_uint32(n), C.c_uint32(n))
                port.port_full_distortion(*args, got, C.c_uint32(n), C.c_uint32(n), C.c_int(kind))
                assert list(got) == list(want), (cname, n, x, y)
                if common.ref_asm_simd_ok():
                    getattr(ref, sname % {4: "4x4", 8: "8x8", 16: "16MxN", 32: "16MxN"}[n])(*args, gs, C.c_uint32(n), C.c_uint32(n))
                    assert list(gs) == list(want), (sname, n, x, y)


@needs_ref
@pytest.mark.parametrize("amp", [255, 1023])
def test_satd_port_vs_reference(amp):
    rng = np.random.default_rng(amp)
    ref.Compute8x8Satd.restype = ref.Compute4x4Satd.restype = C.c_uint64
    for _ in range(200):
        d8 = rng.integers(-amp, amp + 1, 64).astype(np.int16)
        d4 = rng.integers(-amp, amp + 1, 16).astype(np.int16)
        assert port.port_satd8x8(vp(d8)) == ref.Compute8x8Satd(vp(d8))
        assert port.port_satd4x4(vp(d4)) == ref.Compute4x4Satd(vp(d4))
    if common.ref_asm_simd_ok():
        ref.Compute8x8Satd_SSE4.restype = C.c_uint64
        for _ in range(100):
            d8 = rng.integers(-255, 256, 64).astype(np.int16)
            assert ref.Compute8x8Satd_SSE4(vp(d8)) == port.port_satd8x8(vp(d8))


def test_batch_drivers_match_numpy():
    """the batch drivers address the right block of the right plane"""
    w, h = 192, 128
    a, b = make_pictures(w, h, seed=9, n=2, kind="synth")
    jobs = make_dist_jobs(w, h, seed=1, n=60)
    got = run_port_batch("port_spatial_distortion_batch", a, b, jobs)
    for j, t in enumerate(jobs):
        pa, pb = a.interior()[t["plane"]], b.interior()[t["plane"]]
        blk = lambda p: p[t["y"]:t["y"] + t["height"], t["x"]:t["x"] + t["width"]].astype(np.int64)
        assert got[j] == ((blk(pa) - blk(pb)) ** 2).sum()
