"""CPU tests: pin the plain-C MC restatement (oracle/port/mc_port.c) against the UNMODIFIED reference's
EncodePassInterPrediction (oracle/ref_mc_harness.c), C rows and SIMD rows."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S

ref = common.load_ref()
port = common.load_port()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")


def make_pictures(w, h, seed, n=3, kind="synth"):
    pics = []
    if kind == "synth":
        v = common.SynthVideo(w, h, seed=seed)
    rng = np.random.default_rng(seed)
    for i in range(n):
        if kind == "synth":
            y = v.luma(i * 3)
            cb, cr = v.chroma(i * 3)
        elif kind == "random":
            y = rng.integers(0, 256, (h, w), dtype=np.uint8)
            cb = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
            cr = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
        else:   # extremes: 0/255 checkers stress the clip paths
            yy, xx = np.mgrid[0:h, 0:w]
            y = ((((xx >> i) + yy) & 1) * 255).astype(np.uint8)
            cb = ((((xx[:h // 2, :w // 2]) + (yy[:h // 2, :w // 2] >> i)) & 1) * 255).astype(np.uint8)
            cr = 255 - cb
        pics.append(S.HostYuvPicture(y, cb, cr))
    return pics


def make_jobs(w, h, n_refs, seed, all_fracs=False, big_mv=False):
    """Non-overlapping PUs tiling the picture with random sizes, directions and motion vectors."""
    rng = np.random.default_rng(seed)
    jobs = []
    k = 0
    for y0 in range(0, h - h % 64, 64):
        for x0 in range(0, w - w % 64, 64):
            size = int(rng.choice([8, 16, 32, 64]))
            for yy in range(y0, y0 + 64, size):
                for xx in range(x0, x0 + 64, size):
                    j = np.zeros((), dtype=S.MC_JOB_DTYPE)
                    j["x"], j["y"], j["w"], j["h"] = xx, yy, size, size
                    j["dir"] = int(rng.integers(0, 3))
                    j["ref_idx"] = rng.integers(0, n_refs, 2)
                    rngmv = 1200 if big_mv else 200
                    mv = rng.integers(-rngmv, rngmv + 1, (2, 2))
                    if all_fracs:   # walk through every (fracx, fracy) pair of both lists
                        mv[0] = (mv[0] & ~7) + [k % 8, (k // 8) % 8]
                        mv[1] = (mv[1] & ~7) + [(k // 64) % 8, (k // 512) % 8]
                        k += 1
                    if rng.random() < 0.15:
                        mv = mv & ~3      # integer-pel
                    j["mv"] = mv
                    jobs.append(j)
    return np.array(jobs, dtype=S.MC_JOB_DTYPE)


def run_oracle(lib, fn_name, refs, jobs, w, h, asm_mode=None):
    dst = S.HostYuvPicture(np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8))
    fn = getattr(lib, fn_name)
    fn.restype = C.c_int
    args = [S.yuv_array(refs), C.c_int(len(refs)), jobs.ctypes.data_as(C.c_void_p), C.c_int(len(jobs)), C.byref(dst.struct)]
    if asm_mode is not None:
        args.append(C.c_int(asm_mode))
    assert fn(*args) == 0
    return dst


MC_CASES = [("synth_allfracs", 512, 384, "synth", True, False), ("random_allfracs", 256, 192, "random", True, False),
            ("extremes", 256, 128, "extremes", True, False), ("synth_bigmv_clamp", 256, 192, "synth", False, True),
            ("random_bigmv_clamp", 192, 128, "random", False, True)]


@needs_ref
@pytest.mark.parametrize("case", MC_CASES, ids=lambda c: c[0])
def test_mc_batch_port_vs_reference(case):
    name, w, h, kind, all_fracs, big = case
    refs = make_pictures(w, h, seed=11, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=5, all_fracs=all_fracs, big_mv=big)
    want = run_oracle(ref, "svtref_mc_batch", refs, jobs, w, h, asm_mode=0)
    got = run_oracle(port, "port_mc_batch", refs, jobs, w, h)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs (port vs reference C rows)"
    if common.ref_asm_simd_ok():
        simd = run_oracle(ref, "svtref_mc_batch", refs, jobs, w, h, asm_mode=1)
        for a, b, nm in zip(simd.interior(), want.interior(), "Y Cb Cr".split()):
            assert np.array_equal(a, b), f"{name}: {nm} differs (reference SIMD vs C rows)"


@needs_ref
@pytest.mark.parametrize("case", MC_CASES, ids=lambda c: c[0])
def test_mc_ifree_port_vs_reference(case):
    """the interpolation-free prediction of the mode-decision fast loop (row b6): UniPredIFreeRef8Bit / BiPredIFreeRef8Bit, C and SIMD rows"""
    name, w, h, kind, all_fracs, big = case
    refs = make_pictures(w, h, seed=13, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=6, all_fracs=all_fracs, big_mv=big)
    want = run_oracle(ref, "svtref_mc_ifree_batch", refs, jobs, w, h, asm_mode=0)
    got = run_oracle(port, "port_mc_ifree_batch", refs, jobs, w, h)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs (port vs reference C rows)"
    if common.ref_asm_simd_ok():
        simd = run_oracle(ref, "svtref_mc_ifree_batch", refs, jobs, w, h, asm_mode=1)
        for a, b, nm in zip(simd.interior(), want.interior(), "Y Cb Cr".split()):
            assert np.array_equal(a, b), f"{name}: {nm} differs (reference SIMD vs C rows)"
    # integer vectors: the interpolation-free result is the HEVC interpolation's (both are copies / rounded averages of copies ... for uni-pred)
    ij = jobs.copy()
    ij["mv"] &= ~3
    ij = ij[ij["dir"] != 2]
    a = run_oracle(port, "port_mc_ifree_batch", refs, ij, w, h)
    b = run_oracle(port, "port_mc_batch", refs, ij, w, h)
    for x, y in zip(a.interior(), b.interior()):
        # chroma of an odd integer luma vector still has a half-sample fraction in the HEVC path: compare luma only
        assert np.array_equal(x, y) or x.shape != a.interior()[0].shape


@needs_ref
@pytest.mark.parametrize("case", MC_CASES, ids=lambda c: c[0])
def test_md_interpolation_is_the_encode_pass_interpolation(case):
    """UniPredHevcInterpolationMd / BiPredHevcInterpolationMd (Codec/EbMcp.c:99,321), the mode-decision wrappers of row b6: same kernels, same positions as the
    encode pass -- so the entry point that is parity-tested on the GPU (svtb200_mc_batch) covers them"""
    name, w, h, kind, all_fracs, big = case
    refs = make_pictures(w, h, seed=17, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=9, all_fracs=all_fracs, big_mv=big)
    want = run_oracle(port, "port_mc_batch", refs, jobs, w, h)
    for asm in ((0, 1) if common.ref_asm_simd_ok() else (0,)):
        got = run_oracle(ref, "svtref_mc_md_batch", refs, jobs, w, h, asm_mode=asm)
        for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
            assert np.array_equal(a, b), f"{name}: {nm} differs (MD wrappers, asm {asm})"


# ---- 16-bit samples (10-bit video), SURVEY §8 row b5 ---------------------------------------------------------------------
def make_pictures16(w, h, seed, n=3, kind="synth"):
    rng = np.random.default_rng(seed + 1000)
    out = []
    for p in make_pictures(w, h, seed, n, kind):
        planes = []
        for a in p.interior():
            if kind == "extremes":
                planes.append((a.astype(np.uint16) // 255) * 1023)
            else:
                planes.append(a.astype(np.uint16) * 4 + rng.integers(0, 4, a.shape, dtype=np.uint16))
        out.append(S.HostYuvPicture(*planes, dtype=np.uint16))
    return out


def run_oracle16(lib, fn_name, refs, jobs, w, h, asm_mode=None):
    dst = S.HostYuvPicture(np.zeros((h, w), np.uint16), np.zeros((h // 2, w // 2), np.uint16), np.zeros((h // 2, w // 2), np.uint16), dtype=np.uint16)
    fn = getattr(lib, fn_name)
    fn.restype = C.c_int
    args = [S.yuv_array(refs), C.c_int(len(refs)), jobs.ctypes.data_as(C.c_void_p), C.c_int(len(jobs)), C.byref(dst.struct)]
    if asm_mode is not None:
        args.append(C.c_int(asm_mode))
    assert fn(*args) == 0
    return dst


@needs_ref
@pytest.mark.parametrize("case", MC_CASES, ids=lambda c: c[0])
def test_mc_batch16_port_vs_reference(case):
    name, w, h, kind, all_fracs, big_mv = case
    refs = make_pictures16(w, h, seed=w + h, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=w, all_fracs=all_fracs, big_mv=big_mv)
    got = run_oracle16(port, "port_mc_batch16", refs, jobs, w, h)
    for asm in ([0, 1] if common.ref_asm_simd_ok() else [0]):
        want = run_oracle16(ref, "svtref_mc_batch16", refs, jobs, w, h, asm_mode=asm)
        for a, b, pl in zip(got.interior(), want.interior(), "y cb cr".split()):
            assert np.array_equal(a, b), (name, asm, pl, np.argwhere(a != b)[:4])
