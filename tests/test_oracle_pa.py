"""CPU tests: pin the plain-C restatement of PictureAnalysis' per-LCU mean / variance pyramid (oracle/port/pa_port.c) against the UNMODIFIED reference's static
ComputeBlockMeanComputeVariance, reached by compiling Codec/EbPictureAnalysisProcess.c in place (oracle/refinc_pa_harness.c -> oracle/_ref/libsvtref_pa_*.so),
on its C / SSE2 rows and on its AVX2 row."""
import ctypes as C
import os

import numpy as np
import pytest

import common
import svtb200 as S

port = common.load_port()


def load_ref_pa():
    variant = "avx512" if common.cpu_has("avx512bw") else "avx2"
    common.load_ref(variant)
    so = os.path.join(common.ORACLE, "_ref", f"libsvtref_pa_{variant}.so")
    return C.CDLL(so) if os.path.exists(so) else None


refpa = load_ref_pa()
needs_ref = pytest.mark.skipif(refpa is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a: a.ctypes.data_as(C.c_void_p)

PA_CASES = [("synth_448x256", 448, 256, "synth"), ("random_192x128", 192, 128, "random"), ("partial_328x200", 328, 200, "synth"), ("flat_128x64", 128, 64, "flat"),
            ("extremes_256x128", 256, 128, "extremes")]


def padded_luma(w, h, seed, kind):
    """the padded input picture PictureAnalysis reads: 68 samples of edge replication around the picture (EbEncHandle.c:1947-1950) plus enough below / right
    of a partial last LCU"""
    if kind == "synth":
        y = common.SynthVideo(w, h, seed=seed).luma(1)
    elif kind == "flat":
        y = np.full((h, w), 93, np.uint8)
    elif kind == "extremes":
        yy, xx = np.mgrid[0:h, 0:w]
        y = (((xx // 3 + yy // 5) & 1) * 255).astype(np.uint8)
    else:
        y = np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    buf = np.ascontiguousarray(np.pad(y, 68, mode="edge"))
    pl = S.Plane()
    pl.buf, pl.stride, pl.origin_x, pl.origin_y, pl.width, pl.height = buf.ctypes.data, buf.shape[1], 68, 68, w, h
    return pl, buf


def run_variance(lib, name, pl, *extra):
    n = ((pl.width + 63) // 64) * ((pl.height + 63) // 64)
    ym, vr = np.zeros((n, 85), np.uint8), np.zeros((n, 85), np.uint16)
    fn = getattr(lib, name); fn.restype = C.c_int
    assert fn(C.byref(pl), vp(ym), vp(vr), *extra) == 0
    return ym, vr


@needs_ref
@pytest.mark.parametrize("case", PA_CASES, ids=lambda c: c[0])
def test_variance_port_vs_reference(case):
    name, w, h, kind = case
    pl, _keep = padded_luma(w, h, seed=w + h, kind=kind)
    wm, wv = run_variance(refpa, "svtref_variance_picture", pl, C.c_int(0))
    gm, gv = run_variance(port, "port_variance_picture", pl)
    assert np.array_equal(gm, wm), (name, np.argwhere(gm != wm)[:5])
    assert np.array_equal(gv, wv), (name, np.argwhere(gv != wv)[:5])
    if common.ref_asm_simd_ok():
        sm, sv = run_variance(refpa, "svtref_variance_picture", pl, C.c_int(1))
        assert np.array_equal(sm, wm) and np.array_equal(sv, wv), name + ": AVX2 row differs from the SSE2 row"
    if kind == "flat":
        assert np.all(gm == 93) and np.all(gv == 0)


def test_variance_of_a_half_and_half_block():
    """left half 0, right half 200 in every 8x8 block: mean 100, variance 100^2 at every level"""
    w, h = 128, 64
    y = np.zeros((h, w), np.uint8); y[:, 4::8] = 200; y[:, 5::8] = 200; y[:, 6::8] = 200; y[:, 7::8] = 200
    buf = np.ascontiguousarray(np.pad(y, 68, mode="edge"))
    pl = S.Plane()
    pl.buf, pl.stride, pl.origin_x, pl.origin_y, pl.width, pl.height = buf.ctypes.data, buf.shape[1], 68, 68, w, h
    gm, gv = run_variance(port, "port_variance_picture", pl)
    assert np.all(gm == 100) and np.all(gv == 10000)
