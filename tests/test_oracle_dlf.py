"""CPU tests: pin the plain-C deblocking restatement (oracle/port/dlf_port.c) against the UNMODIFIED reference: edge
filters against the C and SSSE3 rows, and the whole-picture traversal against the reference's own per-LCU drivers
(oracle/ref_dlf_harness.c)."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_mc import make_pictures

ref = common.load_ref()
port = common.load_port()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a, off=0: C.c_void_p(a.ctypes.data + off)


def edge_block(kind, rng):
    if kind == "random":
        return rng.integers(0, 256, (16, 16), dtype=np.uint8)
    if kind == "step":      # a blocking artefact: two flat sides with a small step -> strong filter candidates
        b = np.full((16, 16), int(rng.integers(20, 200)), np.uint8)
        b[:, 8:] += np.uint8(rng.integers(1, 12))
        b[8:, :] += np.uint8(rng.integers(0, 6))
        return (b + rng.integers(0, 2, b.shape).astype(np.uint8)).astype(np.uint8)
    if kind == "ramp":
        y, x = np.mgrid[0:16, 0:16]
        return np.clip(x * rng.integers(1, 9) + y * rng.integers(0, 5) + rng.integers(0, 3, (16, 16)), 0, 255).astype(np.uint8)
    raise ValueError(kind)


@needs_ref
@pytest.mark.parametrize("kind", ["random", "step", "ramp"])
def test_edge_filters_port_vs_reference(kind):
    rng = np.random.default_rng(3)
    for trial in range(300):
        blk = edge_block(kind, rng)
        tc, beta, vert = int(rng.integers(0, 25)), int(rng.integers(0, 65)), int(rng.integers(0, 2))
        outs = []
        for lib, name in ((ref, "Luma4SampleEdgeDLFCore"), (port, "port_luma_edge")) + (((ref, "Luma4SampleEdgeDLFCore_SSSE3"),) if common.ref_asm_simd_ok() else ()):
            b = blk.copy()
            f = getattr(lib, name); f.restype = None
            f(vp(b, 8 * 16 + 8), C.c_uint32(16), C.c_uint8(vert), C.c_int32(tc), C.c_int32(beta))
            outs.append(b)
        for o in outs[1:]:
            assert np.array_equal(o, outs[0]), f"luma edge trial {trial}"
        cb_tc, cr_tc = int(rng.integers(0, 25)), int(rng.integers(0, 25))
        outs = []
        for lib, name in ((ref, "Chroma2SampleEdgeDLFCore"), (port, "port_chroma_edge")) + (((ref, "Chroma2SampleEdgeDLFCore_SSSE3"),) if common.ref_asm_simd_ok() else ()):
            a, b = blk.copy(), blk.T.copy()
            f = getattr(lib, name); f.restype = None
            f(vp(a, 8 * 16 + 8), vp(b, 8 * 16 + 8), C.c_uint32(16), C.c_uint8(vert), C.c_uint8(cb_tc), C.c_uint8(cr_tc))
            outs.append((a, b))
        for o in outs[1:]:
            assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]), f"chroma edge trial {trial}"


def make_dlf_inputs(w, h, seed, density=0.6, kind="synth"):
    rng = np.random.default_rng(seed)
    pic = make_pictures(w, h, seed=seed, n=1, kind=kind)[0]
    if kind == "synth":   # add blocking artefacts so the filters actually engage
        y, cb, cr = pic.interior()
        by = rng.integers(-6, 7, (h // 8 + 1, w // 8 + 1))
        y[:] = np.clip(y.astype(int) + np.kron(by, np.ones((8, 8), int))[:h, :w], 0, 255)
        bc = rng.integers(-5, 6, (h // 16 + 1, w // 16 + 1))
        cb[:] = np.clip(cb.astype(int) + np.kron(bc, np.ones((8, 8), int))[:h // 2, :w // 2], 0, 255)
        cr[:] = np.clip(cr.astype(int) - np.kron(bc, np.ones((8, 8), int))[:h // 2, :w // 2], 0, 255)
    n_lcu = ((w + 63) // 64) * ((h + 63) // 64)
    vbs = rng.integers(0, 3, (n_lcu, 256)).astype(np.uint8)
    hbs = rng.integers(0, 3, (n_lcu, 256)).astype(np.uint8)
    vbs[rng.random(vbs.shape) > density] = 0
    hbs[rng.random(hbs.shape) > density] = 0
    # precondition of the reference drivers (SetBSArrayBasedOn{PU,TU}Boundary never mark them): no strength on the picture's
    # left / top boundary
    cols = (w + 63) // 64
    v3 = vbs.reshape(-1, cols, 16, 16)   # [lcu_row][lcu_col][y4][x4]
    h3 = hbs.reshape(-1, cols, 16, 16)
    v3[:, 0, :, 0] = 0
    h3[0, :, 0, :] = 0
    qs = (w + 7) // 8
    qp = rng.integers(18, 46, ((h + 7) // 8, qs)).astype(np.uint8)
    prm = S.DlfParams(w, h, qs, int(rng.integers(-6, 7)) * 2 // 2, int(rng.integers(-6, 7)), int(rng.integers(-4, 5)), int(rng.integers(-4, 5)))
    return pic, vbs, hbs, qp, prm


def run_dlf(lib, name, pic, vbs, hbs, qp, prm, asm_mode=None):
    y, cb, cr = (a.copy() for a in pic.interior())
    out = S.HostYuvPicture(y, cb, cr)
    fn = getattr(lib, name); fn.restype = C.c_int
    args = [C.byref(out.struct), vp(vbs), vp(hbs), vp(qp), C.byref(prm)]
    if asm_mode is not None:
        args.append(C.c_int(asm_mode))
    assert fn(*args) == 0
    return out


DLF_CASES = [("synth_256x192", 256, 192, "synth", 0.6), ("synth_partial_lcu_200x136", 200, 136, "synth", 0.7), ("random_192x128", 192, 128, "random", 0.5),
             ("extremes_128x64", 128, 64, "extremes", 0.9), ("synth_dense_320x256", 320, 256, "synth", 1.0), ("tiny_64x64", 64, 64, "synth", 0.8)]


@needs_ref
@pytest.mark.parametrize("case", DLF_CASES, ids=lambda c: c[0])
def test_dlf_picture_port_vs_reference(case):
    name, w, h, kind, density = case
    pic, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    want = run_dlf(ref, "svtref_dlf_picture", pic, vbs, hbs, qp, prm, asm_mode=0)
    got = run_dlf(port, "port_dlf_picture", pic, vbs, hbs, qp, prm)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs (port vs reference C rows): {np.argwhere(a != b)[:5]}"
    changed = sum(int((a != b).sum()) for a, b in zip(want.interior(), pic.interior()))
    assert changed > 0 or kind == "extremes"
    if common.ref_asm_simd_ok():
        simd = run_dlf(ref, "svtref_dlf_picture", pic, vbs, hbs, qp, prm, asm_mode=1)
        for a, b, nm in zip(simd.interior(), want.interior(), "Y Cb Cr".split()):
            assert np.array_equal(a, b), f"{name}: {nm} differs (reference SIMD vs C rows)"


# ---- 16-bit samples (10-bit video) ---------------------------------------------------------------------------------------------
def to16(pic, seed, extremes=False):
    rng = np.random.default_rng(seed)
    planes = [(a.astype(np.uint16) // 255) * 1023 if extremes else a.astype(np.uint16) * 4 + rng.integers(0, 4, a.shape, dtype=np.uint16) for a in pic.interior()]
    return S.HostYuvPicture(*planes, dtype=np.uint16)


def run_dlf16(lib, name, pic, vbs, hbs, qp, prm, asm_mode=None):
    y, cb, cr = (a.copy() for a in pic.interior())
    out = S.HostYuvPicture(y, cb, cr, dtype=np.uint16)
    fn = getattr(lib, name); fn.restype = C.c_int
    args = [C.byref(out.struct), vp(vbs), vp(hbs), vp(qp), C.byref(prm)]
    if asm_mode is not None:
        args.append(C.c_int(asm_mode))
    assert fn(*args) == 0
    return out


@needs_ref
@pytest.mark.parametrize("case", DLF_CASES, ids=lambda c: c[0])
def test_dlf_picture16_port_vs_reference(case):
    name, w, h, kind, density = case
    pic8, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    pic = to16(pic8, seed=w, extremes=kind == "extremes")
    want = run_dlf16(ref, "svtref_dlf_picture16", pic, vbs, hbs, qp, prm, asm_mode=0)
    got = run_dlf16(port, "port_dlf_picture16", pic, vbs, hbs, qp, prm)
    for a, b, nm in zip(got.interior(), want.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"{name}: {nm} differs (16-bit port vs reference C rows): {np.argwhere(a != b)[:5]}"
    changed = sum(int((a != b).sum()) for a, b in zip(want.interior(), pic.interior()))
    assert changed > 0 or kind == "extremes"
    if common.ref_asm_simd_ok():
        simd = run_dlf16(ref, "svtref_dlf_picture16", pic, vbs, hbs, qp, prm, asm_mode=1)
        for a, b, nm in zip(simd.interior(), want.interior(), "Y Cb Cr".split()):
            assert np.array_equal(a, b), f"{name}: {nm} differs (16-bit reference SIMD vs C rows)"
