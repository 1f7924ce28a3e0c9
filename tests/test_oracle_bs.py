"""CPU tests: pin the plain-C restatement of the deblocking boundary-strength derivation (oracle/port/dlf_port.c: port_dlf_bs_picture) against the
UNMODIFIED reference's SetBSArrayBasedOnPUBoundary / SetBSArrayBasedOnTUBoundary driven CU by CU (oracle/ref_dlf_harness.c: svtref_dlf_bs_picture)."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S

port = common.load_port()
ref = common.load_ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a: a.ctypes.data_as(C.c_void_p)


def make_md_maps(w, h, seed, intra_share=0.15, still_share=0.3, slice_type=0):
    """A random but legal encoder state: a quadtree partition of every LCU into CUs that lie inside the picture, each with a mode, a prediction direction
    and vectors (neighbouring CUs often share vectors, so the 'difference < 4' branches fire), and a cbf map constant over each TU."""
    rng = np.random.default_rng(seed)
    cu = np.zeros((h // 8, w // 8), S.CU_INFO_DTYPE)
    cbf = np.zeros((h // 4, w // 4), np.uint8)
    pool = rng.integers(-40, 41, (6, 2, 2)).astype(np.int16)

    def leaf(x, y, size):
        blk = cu[y // 8:(y + size) // 8, x // 8:(x + size) // 8]
        blk["log2_cu_size"] = int(np.log2(size))
        intra = slice_type == 2 or rng.random() < intra_share
        blk["intra"] = 1 if intra else 0
        if not intra:
            blk["pred_dir"] = 0 if slice_type == 1 else rng.integers(0, 3)
            mv = pool[rng.integers(0, len(pool))].copy()
            if rng.random() > still_share:
                mv += rng.integers(-5, 6, (2, 2)).astype(np.int16)
            blk["mv"] = mv
        tu = 32 if size == 64 else size
        for ty in range(y, y + size, tu):
            for tx in range(x, x + size, tu):
                cbf[ty // 4:(ty + tu) // 4, tx // 4:(tx + tu) // 4] = rng.random() < 0.5

    def split(x, y, size):
        if x >= w or y >= h:
            return
        inside = x + size <= w and y + size <= h
        if size > 8 and (not inside or rng.random() < (0.75 if size == 64 else 0.45)):
            for q in range(4):
                split(x + (q & 1) * size // 2, y + (q >> 1) * size // 2, size // 2)
        elif inside:
            leaf(x, y, size)

    for y in range(0, h, 64):
        for x in range(0, w, 64):
            split(x, y, 64)
    return cu, cbf


BS_CASES = [("B_256x192", 256, 192, 0, 0, 16), ("B_same_poc", 192, 128, 0, 8, 8), ("P_200x136", 200, 136, 1, 0, 0), ("I_128x64", 128, 64, 2, 0, 0),
            ("B_partial_328x200", 328, 200, 0, 4, 12)]


def run_bs(lib, name, cu, cbf, w, h, st, p0, p1):
    n_lcu = ((w + 63) // 64) * ((h + 63) // 64)
    v, hh = np.full((n_lcu, 256), 9, np.uint8), np.full((n_lcu, 256), 9, np.uint8)
    fn = getattr(lib, name); fn.restype = C.c_int
    assert fn(vp(cu), C.c_uint32(cu.shape[1]), vp(cbf), C.c_uint32(cbf.shape[1]), C.c_uint16(w), C.c_uint16(h), C.c_uint8(st), C.c_uint64(p0), C.c_uint64(p1), vp(v), vp(hh)) == 0
    return v, hh


@needs_ref
@pytest.mark.parametrize("case", BS_CASES, ids=lambda c: c[0])
def test_bs_port_vs_reference(case):
    name, w, h, st, p0, p1 = case
    cu, cbf = make_md_maps(w, h, seed=w + h + st, slice_type=st)
    wv, wh = run_bs(ref, "svtref_dlf_bs_picture", cu, cbf, w, h, st, p0, p1)
    gv, gh = run_bs(port, "port_dlf_bs_picture", cu, cbf, w, h, st, p0, p1)
    assert np.array_equal(gv, wv), (name, np.argwhere(gv != wv)[:5])
    assert np.array_equal(gh, wh), (name, np.argwhere(gh != wh)[:5])
    assert set(np.unique(gv)) <= {0, 1, 2} and (gv != 0).any() and (gh != 0).any()
    if st == 0:
        assert (gv == 1).any() and (gv == 2).any() and (gv[:, 0::2] == 0).any()      # every class of the B rule occurs


def test_bs_of_a_uniform_inter_picture_is_the_cbf_edges():
    """one vector everywhere, no intra: a strength is 1 exactly where either side of a CU / TU edge has a coded luma TU"""
    w, h = 128, 128
    cu, cbf = make_md_maps(w, h, seed=3, intra_share=0.0, slice_type=1)
    cu["mv"] = 0
    v, hh = run_bs(port, "port_dlf_bs_picture", cu, cbf, w, h, 1, 0, 0)
    for ly in range(2):
        for lx in range(2):
            for by in range(16):
                for bx in range(16):
                    x, y = lx * 64 + 4 * bx, ly * 64 + 4 * by
                    size = 1 << int(cu[y // 8, x // 8]["log2_cu_size"])
                    edge = x % 8 == 0 and x != 0 and (x % size == 0 or (size == 64 and x % 32 == 0))
                    want = int(edge and bool(cbf[y // 4, x // 4] or cbf[y // 4, x // 4 - 1]))
                    assert v[ly * 2 + lx, by * 16 + bx] == want
