"""CPU tests: pin the plain-C transform/quant restatement (oracle/port/tq_port.c) against the UNMODIFIED reference:
leaf transforms against the C and SIMD rows of the dispatch tables, and the TU chain against the reference's own
EstimateTransform / UnifiedQuantizeInvQuantize / EncodeInvTransform / AdditionKernel sequence (oracle/ref_tq_harness.c)."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_mc import make_pictures

ref = common.load_ref()
port = common.load_port()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a: a.ctypes.data_as(C.c_void_p)
u = C.c_uint32


def residual_block(N, kind, rng, amp=255):
    if kind == "random":
        return rng.integers(-amp, amp + 1, (N, N)).astype(np.int16)
    if kind == "flat":
        return np.full((N, N), int(rng.integers(-amp, amp + 1)), np.int16)
    if kind == "extreme":
        return (rng.integers(0, 2, (N, N)) * 2 * amp - amp).astype(np.int16)
    if kind == "smooth":
        y, x = np.mgrid[0:N, 0:N]
        return np.clip(np.rint(40 * np.sin(x / 5.0) + 30 * np.cos(y / 7.0) + rng.normal(0, 2, (N, N))), -amp, amp).astype(np.int16)
    raise ValueError(kind)


FWD_FUNCS = {  # size, dst -> (C row, SIMD row) of transformFunctionTableEstimate, Codec/EbTransforms.h:339
    (32, 0): ("Transform32x32Estimate", "lowPrecisionTransform32x32_AVX2_INTRIN"), (16, 0): ("Transform16x16Estimate", "lowPrecisionTransform16x16_AVX2_INTRIN"),
    (8, 0): ("Transform8x8", "Transform8x8_SSE4_1_INTRIN"), (4, 0): ("Transform4x4", "Transform4x4_SSE2_INTRIN"), (4, 1): ("DstTransform4x4", "DstTransform4x4_SSE2_INTRIN")}
INV_FUNCS = {  # invTransformFunctionTableEncode, Codec/EbTransforms.h:492
    (32, 0): ("InvTransform32x32", "PFinvTransform32x32_SSSE3"), (16, 0): ("InvTransform16x16", "PFinvTransform16x16_SSSE3"),
    (8, 0): ("InvTransform8x8", "InvTransform8x8_SSE2_INTRIN"), (4, 0): ("InvTransform4x4", "InvTransform4x4_SSE2_INTRIN"), (4, 1): ("InvDstTransform4x4", "InvDstTransform4x4_SSE2_INTRIN")}


def call_tr(lib, name, src, N, bit_inc):
    out = np.zeros((N, N), np.int16)
    scratch = np.zeros(64 * 64, np.int16)
    f = getattr(lib, name)
    f.restype = None
    f(vp(src), u(N), vp(out), u(N), vp(scratch), u(bit_inc))
    return out


@needs_ref
@pytest.mark.parametrize("key", list(FWD_FUNCS), ids=lambda k: f"N{k[0]}dst{k[1]}")
@pytest.mark.parametrize("kind", ["random", "flat", "extreme", "smooth"])
def test_forward_transform_port_vs_reference(key, kind):
    N, dst = key
    rng = np.random.default_rng(N * 10 + dst)
    for bit_inc, amp in ((0, 255), (2, 1023)):
        for _ in range(6):
            r = residual_block(N, kind, rng, amp)
            want = call_tr(ref, FWD_FUNCS[key][0], r, N, bit_inc)
            got = np.zeros((N, N), np.int16)
            port.port_estimate_transform(vp(r), u(N), vp(got), u(N), u(N), u(bit_inc), C.c_uint8(dst), C.c_uint8(0))
            assert np.array_equal(got, want)
            # the SIMD rows are only dispatched with bitIncrement 0: EncodeTransform forces row 0 for 10-bit
            # (Codec/EbTransforms.c:3360,3370); Transform8x8_SSE4_1_INTRIN ignores a non-zero bitIncrement
            if common.ref_asm_simd_ok() and bit_inc == 0:
                assert np.array_equal(call_tr(ref, FWD_FUNCS[key][1], r, N, bit_inc), want), "reference SIMD row != C row"


@needs_ref
@pytest.mark.parametrize("key", list(INV_FUNCS), ids=lambda k: f"N{k[0]}dst{k[1]}")
def test_inverse_transform_port_vs_reference(key):
    N, dst = key
    rng = np.random.default_rng(N + dst)
    for trial in range(12):
        c = np.zeros((N, N), np.int16)
        k = [N, max(1, N // 2), max(1, N // 4), 1][trial % 4]          # dense, low-frequency corner, DC only
        c[:k, :k] = rng.integers(-3000, 3001, (k, k))
        if trial >= 8:
            c[:k, :k] = rng.integers(-32768, 32768, (k, k))              # clip paths between the two passes
        for bit_inc in (0, 2):
            want = call_tr(ref, INV_FUNCS[key][0], c, N, bit_inc)
            got = np.zeros((N, N), np.int16)
            port.port_inv_transform(vp(c), u(N), vp(got), u(N), u(N), u(bit_inc), C.c_uint8(dst))
            assert np.array_equal(got, want)
            if common.ref_asm_simd_ok():
                assert np.array_equal(call_tr(ref, INV_FUNCS[key][1], c, N, bit_inc), want), "reference SIMD row != C row"


def make_tq_jobs(w, h, seed, shapes=(0,), flags=(0,), slice_types=(0, 1, 2)):
    """Tile the three planes with TUs of random size; each TU gets random qp / slice type / shape / flags."""
    rng = np.random.default_rng(seed)
    jobs = []
    for plane, (pw, ph) in enumerate(((w, h), (w // 2, h // 2), (w // 2, h // 2))):
        for y0 in range(0, ph - ph % 32, 32):
            for x0 in range(0, pw - pw % 32, 32):
                N = int(rng.choice([4, 8, 16, 32]))
                for yy in range(y0, y0 + 32, N):
                    for xx in range(x0, x0 + 32, N):
                        j = np.zeros((), dtype=S.TQ_JOB_DTYPE)
                        j["x"], j["y"], j["plane"], j["size"] = xx, yy, plane, N
                        j["dst4x4"] = 1 if (N == 4 and plane == 0 and rng.random() < 0.5) else 0
                        j["shape"] = int(rng.choice(shapes)) if N >= 8 else 0
                        if N == 8 and j["shape"] == S.SHAPE_N4:
                            # 8x8 + N4 is undefined in the reference's SIMD row: its 2x2 active area is quantised by the 4x4
                            # kernel (QiQ_funcPtrArray[2 >> 3], Codec/EbTransforms.c:3148), which also counts 12 stale coefficients
                            j["shape"] = S.SHAPE_N2
                        j["qp"] = int(rng.integers(10, 52))
                        j["slice_type"] = int(rng.choice(slice_types))
                        j["temporal_layer"] = int(rng.integers(0, 4)) if j["slice_type"] != 2 else 0
                        j["flags"] = int(rng.choice(flags))
                        j["dz_offset"] = int(rng.choice([0, 0, 0, 5, 9]))
                        jobs.append(j)
    return np.array(jobs, dtype=S.TQ_JOB_DTYPE)


def make_src_pred(w, h, seed, kind="synth"):
    pics = make_pictures(w, h, seed=seed, n=2, kind=kind)
    return pics[0], pics[1]


def run_tq_oracle(lib, name, src, pred, jobs, w, h, asm_mode=None):
    y, cb, cr = (a.copy() for a in pred.interior())
    recon = S.HostYuvPicture(y, cb, cr)
    coeff = S.HostCoeffPicture(w, h)
    nz = np.zeros(len(jobs), np.uint32)
    fn = getattr(lib, name)
    fn.restype = C.c_int
    args = [C.byref(src.struct), C.byref(pred.struct), C.byref(recon.struct), C.byref(coeff.struct), vp(jobs), C.c_int(len(jobs)), vp(nz)]
    if asm_mode is not None:
        args += [C.c_int(asm_mode), C.c_int(0), C.c_int(len(jobs))]
    assert fn(*args) == 0
    return recon, coeff, nz


def assert_tq_equal(a, b, what):
    for pa, pb, nm in zip(a[0].interior(), b[0].interior(), "Y Cb Cr".split()):
        assert np.array_equal(pa, pb), f"{what}: recon {nm} differs"
    for pa, pb, nm in zip(a[1].planes, b[1].planes, "Y Cb Cr".split()):
        assert np.array_equal(pa, pb), f"{what}: coefficients {nm} differ"
    assert np.array_equal(a[2], b[2]), f"{what}: nz counts differ"


TQ_CASES = [("default_synth", 256, 192, "synth", (0,), (0,)), ("default_random", 128, 128, "random", (0,), (0,)),
            ("extremes", 128, 64, "extremes", (0,), (0,)), ("shapes", 256, 128, "synth", (0, 1, 2, 3), (0,)),
            ("flags", 256, 128, "synth", (0, 1), (0, 1, 2, 3, 4, 5, 6, 7)), ("skip_recon", 128, 64, "synth", (0,), (8,))]


@needs_ref
@pytest.mark.parametrize("kind,shapes", [("synth", (0,)), ("random", (0, 1, 2)), ("extremes", (0, 1))], ids=["synth", "random_shapes", "extremes"])
def test_md_quantiser_is_the_skip_recon_chain(kind, shapes):
    """ProductUnifiedQuantizeInvQuantizeMd (Codec/EbFullLoop.c:77), the quantiser of the mode-decision full loop, at enc modes >= 5 (no RDOQ / PM core):
    its coefficients and non-zero counts are those of the encode-pass chain run with SKIP_RECON, no dead-zone offset and no cbf / sparse flags -- so the
    entry point that is parity-tested on the GPU (svtb200_tq_batch, golden case tq/skip_recon) covers it"""
    w, h = 256, 128
    src, pred = make_src_pred(w, h, 23, kind)
    jobs = make_tq_jobs(w, h, seed=12, shapes=shapes, flags=(8, 8 | 4))      # SKIP_RECON, with and without the contouring update
    jobs["dz_offset"] = 0
    want_c = S.HostCoeffPicture(w, h)
    want_nz = np.zeros(len(jobs), np.uint32)
    fn = ref.svtref_tq_md_batch; fn.restype = C.c_int
    for asm in ((0, 1) if common.ref_asm_simd_ok() else (0,)):
        assert fn(C.byref(src.struct), C.byref(pred.struct), C.byref(want_c.struct), vp(jobs), C.c_int(len(jobs)), vp(want_nz), C.c_int(asm)) == 0
        _, got_c, got_nz = run_tq_oracle(port, "port_tq_batch", src, pred, jobs, w, h)
        assert np.array_equal(got_nz, want_nz), (asm, np.argwhere(got_nz != want_nz)[:5])
        for a, b in zip(got_c.planes, want_c.planes):
            assert np.array_equal(a, b), asm


@needs_ref
@pytest.mark.parametrize("case", TQ_CASES, ids=lambda c: c[0])
def test_tq_batch_port_vs_reference(case):
    name, w, h, kind, shapes, flags = case
    src, pred = make_src_pred(w, h, 31, kind)
    jobs = make_tq_jobs(w, h, seed=8, shapes=shapes, flags=flags)
    want = run_tq_oracle(ref, "svtref_tq_batch", src, pred, jobs, w, h, asm_mode=0)
    got = run_tq_oracle(port, "port_tq_batch", src, pred, jobs, w, h)
    assert_tq_equal(got, want, name + " port vs reference C rows")
    if common.ref_asm_simd_ok():
        simd = run_tq_oracle(ref, "svtref_tq_batch", src, pred, jobs, w, h, asm_mode=1)
        assert_tq_equal(simd, want, name + " reference SIMD vs C rows")


# ---- 16-bit samples (10-bit video): EncodeLoop16bit / EncodeGenerateRecon16bit ------------------------------------------------
def make_src_pred16(w, h, seed, kind="synth"):
    from test_oracle_mc import make_pictures16
    pics = make_pictures16(w, h, seed=seed, n=2, kind=kind)
    return pics[0], pics[1]


def run_tq_oracle16(lib, name, src, pred, jobs, w, h, asm_mode=None):
    y, cb, cr = (a.copy() for a in pred.interior())
    recon = S.HostYuvPicture(y, cb, cr, dtype=np.uint16)
    coeff = S.HostCoeffPicture(w, h)
    nz = np.zeros(len(jobs), np.uint32)
    fn = getattr(lib, name)
    fn.restype = C.c_int
    args = [C.byref(src.struct), C.byref(pred.struct), C.byref(recon.struct), C.byref(coeff.struct), vp(jobs), C.c_int(len(jobs)), vp(nz)]
    if asm_mode is not None:
        args += [C.c_int(asm_mode), C.c_int(0), C.c_int(len(jobs))]
    assert fn(*args) == 0
    return recon, coeff, nz


@needs_ref
@pytest.mark.parametrize("case", TQ_CASES, ids=lambda c: c[0])
def test_tq_batch16_port_vs_reference(case):
    name, w, h, kind, shapes, flags = case
    src, pred = make_src_pred16(w, h, 31, kind)
    jobs = make_tq_jobs(w, h, seed=8, shapes=shapes, flags=flags)
    want = run_tq_oracle16(ref, "svtref_tq_batch16", src, pred, jobs, w, h, asm_mode=0)
    got = run_tq_oracle16(port, "port_tq_batch16", src, pred, jobs, w, h)
    assert_tq_equal(got, want, name + " 16-bit port vs reference C rows")
    if common.ref_asm_simd_ok():
        simd = run_tq_oracle16(ref, "svtref_tq_batch16", src, pred, jobs, w, h, asm_mode=1)
        assert_tq_equal(simd, want, name + " 16-bit reference SIMD vs C rows")
