"""CPU tests: pin the plain-C packing restatement (oracle/port/pack_port.c) against the UNMODIFIED reference's C rows and
its AVX2 / AVX512 / SSE2 rows (Codec/EbPackUnPack.h:27-169)."""
import ctypes as C

import numpy as np
import pytest

import common

ref = common.load_ref()
port = common.load_port()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a: C.c_void_p(a.ctypes.data)
u32 = C.c_uint32

# (width, height): SIMD rows need widths the encoder uses (multiples of 8/16/32/64); the C rows take anything (ragged cases)
SIMD_SIZES = [(64, 64), (32, 32), (16, 16), (8, 8), (64, 48), (128, 16)]
RAGGED_SIZES = [(4, 4), (12, 5), (20, 3), (68, 7), (1, 1), (7, 2)]


def aligned(shape, dtype):
    """zeroed array whose data pointer is 64-byte aligned (the reference's _AL rows use aligned 256-bit stores)"""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + 64, np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + n].view(dtype).reshape(shape)


def make_planes(w, h, seed, kind="random"):
    rng = np.random.default_rng(seed)
    if kind == "random":
        s16 = rng.integers(0, 1024, (h, w), dtype=np.uint16)
    elif kind == "max":
        s16 = np.full((h, w), 1023, np.uint16)
    elif kind == "dirty":      # bits above bit 9 set: the (EB_U8) casts of the reference drop them
        s16 = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    else:
        s16 = np.zeros((h, w), np.uint16)
    in8 = (s16 >> 2).astype(np.uint8)
    # unpacked 2-bit plane: the two LSBs in bits 7:6, garbage allowed below (the reference masks it)
    inn = (((s16 & 3) << 6).astype(np.uint8)) | (rng.integers(0, 64, (h, w), dtype=np.uint8) if kind == "dirty" else np.uint8(0))
    return s16, in8, inn


def run_pack(lib, name, in8, inn, w, h, sn):
    out = aligned((h, w), np.uint16)
    f = getattr(lib, name); f.restype = None
    f(vp(in8), u32(in8.strides[0]), vp(inn), vp(out), u32(sn), u32(w), u32(w), u32(h))
    return out


def port_cpack(inn, w, h):
    out = np.zeros((h, max(w // 4, 1)), np.uint8)
    port.port_cpack.restype = None
    port.port_cpack(vp(inn), u32(inn.strides[0]), vp(out), u32(out.strides[0]), u32(w), u32(h))
    return out


@needs_ref
@pytest.mark.parametrize("kind", ["random", "max", "zero", "dirty"])
def test_pack_port_vs_reference(kind):
    for (w, h) in SIMD_SIZES + RAGGED_SIZES:
        s16, in8, inn = make_planes(w, h, w * 131 + h, kind)
        want = run_pack(ref, "EB_ENC_msbPack2D", in8, inn, w, h, inn.strides[0])
        assert np.array_equal(run_pack(port, "port_msb_pack2d", in8, inn, w, h, inn.strides[0]), want)
        if kind != "dirty":
            assert np.array_equal(want, s16 & 1023)
        if (w, h) in SIMD_SIZES and common.ref_asm_simd_ok() and w >= 8:
            assert np.array_equal(run_pack(ref, "EB_ENC_msbPack2D_AVX2_INTRIN_AL", in8, inn, w, h, inn.strides[0]), want), (w, h)
        if w % 4 == 0:
            comp = port_cpack(inn, w, h)
            rc = np.zeros_like(comp)
            ref.CPack_C.restype = None
            ref.CPack_C(vp(inn), u32(inn.strides[0]), vp(rc), u32(rc.strides[0]), None, u32(w), u32(h))
            assert np.array_equal(comp, rc)
            wantc = run_pack(ref, "CompressedPackmsb", in8, comp, w, h, comp.strides[0])
            assert np.array_equal(wantc, want)
            assert np.array_equal(run_pack(port, "port_compressed_packmsb", in8, comp, w, h, comp.strides[0]), wantc)
            if (w, h) in SIMD_SIZES and common.ref_asm_simd_ok() and w in (32, 64):
                # the AVX2 rows exist for the 32- and 64-wide LCU blocks only (CompressedPackLcu, EbPictureOperators.c:567)
                assert np.array_equal(run_pack(ref, "CompressedPackmsb_AVX2_INTRIN", in8, comp, w, h, comp.strides[0]), wantc), (w, h)
                rc2 = aligned(comp.shape, np.uint8)   # the 64x64 branch streams 64 aligned bytes = 4 rows of the stride-16 LCU buffer
                cache = np.zeros(64 * 64, np.uint8)
                ref.CPack_AVX2_INTRIN.restype = None
                ref.CPack_AVX2_INTRIN(vp(inn), u32(inn.strides[0]), vp(rc2), u32(rc2.strides[0]), vp(cache), u32(w), u32(h))
                assert np.array_equal(rc2, comp), (w, h)


@needs_ref
@pytest.mark.parametrize("kind", ["random", "max", "zero", "dirty"])
def test_unpack_port_vs_reference(kind):
    port.port_msb_unpack2d.restype = port.port_unpack_avg.restype = None
    for (w, h) in SIMD_SIZES + RAGGED_SIZES:
        s16, _, _ = make_planes(w, h, w * 17 + h, kind)
        s16b, _, _ = make_planes(w, h, w * 19 + h + 1, kind)

        def unpack(lib, name, nbit=True):
            o8, on = aligned((h, w), np.uint8), aligned((h, w), np.uint8)
            f = getattr(lib, name); f.restype = None
            if nbit:
                f(vp(s16), u32(w), vp(o8), vp(on), u32(w), u32(w), u32(w), u32(h))
            else:
                f(vp(s16), u32(w), vp(o8), u32(w), u32(w), u32(h))
            return o8, on

        w8, wn = unpack(ref, "EB_ENC_msbUnPack2D")
        p8, pn = unpack(port, "port_msb_unpack2d")
        assert np.array_equal(p8, w8) and np.array_equal(pn, wn)
        assert np.array_equal(unpack(ref, "UnPack8BitData", False)[0], w8)
        assert np.array_equal(unpack(ref, "UnPack8BitDataSafeSub", False)[0], w8)
        # SIMD rows saturate (packus) where the C rows truncate: identical on 10-bit data, which is the contract
        simd = (w, h) in SIMD_SIZES and common.ref_asm_simd_ok() and kind != "dirty"
        if simd:
            assert np.array_equal(unpack(ref, "EB_ENC_UnPack8BitData_SSE2_INTRIN", False)[0], w8), (w, h)
            if common.cpu_has("avx512bw") and hasattr(ref, "EB_ENC_msbUnPack2D_AVX512_INTRIN") and w >= 64:   # picture-row kernel: 64-sample chunks
                a8, an = unpack(ref, "EB_ENC_msbUnPack2D_AVX512_INTRIN")
                assert np.array_equal(a8, w8) and np.array_equal(an, wn), (w, h)

        def avg(lib, name):
            o = aligned((h, w), np.uint8)
            f = getattr(lib, name); f.restype = None
            f(vp(s16), u32(w), vp(s16b), u32(w), vp(o), u32(w), u32(w), u32(h))
            return o

        wa = avg(ref, "UnpackAvg")
        assert np.array_equal(avg(port, "port_unpack_avg"), wa)
        assert np.array_equal(avg(ref, "UnpackAvgSafeSub"), wa)
        if simd and w in (8, 16, 32, 64) and h % 2 == 0:   # block kernel: PU widths only
            assert np.array_equal(avg(ref, "UnpackAvg_AVX2_INTRIN"), wa), (w, h)


def test_pack_roundtrip_port():
    """pack(unpack(x)) == x for 10-bit x, through the compressed 2-bit plane as well."""
    port.port_msb_unpack2d.restype = None
    for (w, h) in [(64, 64), (12, 5), (256, 3)]:
        s16, _, _ = make_planes(w, h, 77 + w)
        o8, on = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint8)
        port.port_msb_unpack2d(vp(s16), u32(w), vp(o8), vp(on), u32(w), u32(w), u32(w), u32(h))
        assert np.array_equal(run_pack(port, "port_msb_pack2d", o8, on, w, h, w), s16)
        comp = port_cpack(on, w, h)
        assert np.array_equal(run_pack(port, "port_compressed_packmsb", o8, comp, w, h, comp.strides[0]), s16)
