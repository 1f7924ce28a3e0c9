"""GPU parity tests for 10-bit pack / unpack (through the C ABI) against the plain-C oracle; full-size planes are additionally
checked through the round trip pack(unpack(x)) == x."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_pack import RAGGED_SIZES, SIMD_SIZES, make_planes, port_cpack, run_pack, u32, vp

pytestmark = pytest.mark.gpu
port = common.load_port()

SIZES = SIMD_SIZES + RAGGED_SIZES + [(1920, 1080), (3840 + 8, 9), (1000, 3)]


@pytest.mark.parametrize("kind", ["random", "max", "zero", "dirty"])
def test_pack_cuda_vs_oracle(kind):
    for (w, h) in SIZES:
        s16, in8, inn = make_planes(w, h, w * 131 + h, kind)
        want = run_pack(port, "port_msb_pack2d", in8, inn, w, h, inn.strides[0])
        assert np.array_equal(S.msb_pack2d(in8, inn), want), (w, h)
        if w % 4 == 0:
            comp = port_cpack(inn, w, h)
            assert np.array_equal(S.cpack(inn), comp), (w, h)
            assert np.array_equal(S.msb_pack2d(in8, comp, compressed=True), want), (w, h)


@pytest.mark.parametrize("kind", ["random", "max", "zero", "dirty"])
def test_unpack_cuda_vs_oracle(kind):
    port.port_msb_unpack2d.restype = port.port_unpack_avg.restype = None
    for (w, h) in SIZES:
        s16, _, _ = make_planes(w, h, w * 17 + h, kind)
        s16b, _, _ = make_planes(w, h, w * 19 + h + 1, kind)
        w8, wn = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint8)
        port.port_msb_unpack2d(vp(s16), u32(w), vp(w8), vp(wn), u32(w), u32(w), u32(w), u32(h))
        g8, gn = S.msb_unpack2d(s16)
        assert np.array_equal(g8, w8) and np.array_equal(gn, wn), (w, h)
        assert np.array_equal(S.msb_unpack2d(s16, want_nbit=False), w8), (w, h)
        wa = np.zeros((h, w), np.uint8)
        port.port_unpack_avg(vp(s16), u32(w), vp(s16b), u32(w), vp(wa), u32(w), u32(w), u32(h))
        assert np.array_equal(S.unpack_avg(s16, s16b), wa), (w, h)


def test_strided_views():
    """sub-planes of a larger padded picture: strides larger than the width, unaligned origins (scalar path)"""
    rng = np.random.default_rng(5)
    big16 = rng.integers(0, 1024, (70, 200), dtype=np.uint16)
    v = big16[3:67, 5:133]
    g8, gn = S.msb_unpack2d(v)
    assert np.array_equal(g8, (v >> 2).astype(np.uint8)) and np.array_equal(gn, ((v & 3) << 6).astype(np.uint8))
    big8 = np.zeros((70, 300), np.uint8); bign = np.zeros((70, 300), np.uint8)
    big8[3:67, 7:135] = g8; bign[3:67, 7:135] = gn
    assert np.array_equal(S.msb_pack2d(big8[3:67, 7:135], bign[3:67, 7:135]), v)


def test_4k_roundtrip_and_bad_args():
    s16, _, _ = make_planes(3840, 2160, 99)
    o8, on = S.msb_unpack2d(s16)
    assert np.array_equal(S.msb_pack2d(o8, on), s16)
    assert np.array_equal(S.msb_pack2d(o8, S.cpack(on), compressed=True), s16)
    with pytest.raises(S.Svtb200Error):
        S.cpack(np.zeros((4, 6), np.uint8))
    # empty planes are a no-op
    assert S.msb_pack2d(np.zeros((0, 16), np.uint8), np.zeros((0, 16), np.uint8)).shape == (0, 16)
