"""CPU tests: pin the plain-C intra restatement (oracle/port/intra_port.c) against the UNMODIFIED reference:
  * every mode x size x luma/chroma on explicit reference lines, through EncodePassIntraPrediction (C rows and SIMD rows);
  * encode-pass batches (neighbour substitution, smoothing, mode dispatch) through GenerateIntraReferenceSamplesEncodePass +
    EncodePassIntraPrediction with hand-built neighbour arrays (oracle/ref_intra_harness.c);
  * the open-loop intra search of whole pictures through OpenLoopIntraSearchLcu, all four control-flow variants."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S

ref = common.load_ref()
port = common.load_port()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference here)")
vp = lambda a: C.c_void_p(a.ctypes.data)
ASM_MODES = [0, 1] if common.ref_asm_simd_ok() else [0]


def line_for(kind, n, rng):
    if kind == "random":
        return rng.integers(0, 256, 4 * n + 1, dtype=np.uint8)
    if kind == "flat":
        return np.full(4 * n + 1, int(rng.integers(0, 256)), np.uint8)
    if kind == "ramp":
        return np.clip(np.arange(4 * n + 1) * int(rng.integers(1, 4)) + int(rng.integers(0, 60)), 0, 255).astype(np.uint8)
    if kind == "extreme":
        return rng.choice(np.array([0, 255], np.uint8), 4 * n + 1)
    raise ValueError(kind)


def port_predict(mode, n, chroma, line):
    out = np.zeros((n, n), np.uint8)
    port.port_intra_predict.restype = None
    port.port_intra_predict(C.c_uint32(mode), C.c_uint32(n), C.c_int(chroma), vp(line), vp(out), C.c_uint32(n))
    return out


def ref_predict(mode, n, chroma, line, asm):
    out = np.zeros((n, n), np.uint8)
    rc = ref.svtref_intra_predict(C.c_uint32(mode), C.c_uint32(n), C.c_int(chroma), vp(line), vp(out), C.c_uint32(n), C.c_int(asm))
    assert rc == 0
    return out


@needs_ref
@pytest.mark.parametrize("kind", ["random", "flat", "ramp", "extreme"])
def test_modes_port_vs_reference(kind):
    rng = np.random.default_rng(11)
    for n in (4, 8, 16, 32):
        for chroma in (0, 1):
            if chroma and n == 32:
                continue      # 4:2:0 chroma blocks are at most 16x16 with 32x32 CUs
            for mode in range(35):
                line = line_for(kind, n, rng)
                want = port_predict(mode, n, chroma, line)
                for asm in ASM_MODES:
                    got = ref_predict(mode, n, chroma, line, asm)
                    assert np.array_equal(got, want), (n, chroma, mode, asm)


# ---- encode pass ---------------------------------------------------------------------------------------------------------
def zscan_avail(x, y, n):
    """(bottom-left available, top-right available) of the n x n block at (x, y) inside its 64x64 LCU: the HEVC z-scan rule"""
    def zidx(bx, by):      # z-order index of the 4x4 unit (bx, by)
        z = 0
        for b in range(4):
            z |= ((bx >> b) & 1) << (2 * b) | ((by >> b) & 1) << (2 * b + 1)
        return z
    lx, ly = (x & 63) >> 2, (y & 63) >> 2
    u = n >> 2
    cur = zidx(lx, ly)
    bl = lx == 0 and ly + u < 16
    if lx > 0 and ly + u < 16:
        bl = zidx(lx - 1, ly + u) < cur
    tr = ly == 0 and True
    if ly > 0:
        tr = lx + u < 16 and zidx(lx + u, ly - 1) < cur
    elif lx + u >= 16:
        tr = True      # the LCU to the right of the one above: decided by picture geometry, not by z-scan
    return bl, tr


def make_intra_jobs(w, h, rng, n_jobs, density, sizes=(4, 8, 16, 32)):
    jobs = np.zeros(n_jobs, S.INTRA_JOB_DTYPE)
    for j in jobs:
        n = int(rng.choice(sizes))
        x = int(rng.integers(0, w // n)) * n
        y = int(rng.integers(0, h // n)) * n
        nu = 2 * n // 4
        la = int(rng.integers(0, 1 << nu)) if rng.random() > density else (1 << nu) - 1
        ta = int(rng.integers(0, 1 << nu)) if rng.random() > density else (1 << nu) - 1
        if rng.random() < 0.08:
            la = ta = 0
        bl, tr = zscan_avail(x, y, n)
        if not bl or y + n >= h:
            la &= ~((1 << (nu // 2)) - 1)                       # bottom-left units are the low bits
        if not tr or x + n >= w:
            ta &= (1 << (nu // 2)) - 1
        if y + 2 * n > h:                                        # units below the picture
            la &= ~((1 << min(nu, (y + 2 * n - h + 3) // 4)) - 1)
        if x + 2 * n > w:
            ta &= (1 << max(0, (w - x) // 4)) - 1
        tl = rng.random() < density
        if x == 0:
            la, tl = 0, False
        if y == 0:
            ta, tl = 0, False
        j["x"], j["y"], j["size"] = x, y, n
        j["luma_mode"] = int(rng.integers(0, 35))
        j["chroma_mode"] = S.INTRA_SKIP if n == 4 else int(rng.integers(0, 5))
        if n != 4 and rng.random() < 0.1:
            j["luma_mode"] = S.INTRA_SKIP
            j["chroma_mode"] = int(rng.integers(0, 4))          # DM needs a luma mode
        j["flags"] = (S.INTRA_TOPLEFT_AVAIL if tl else 0) | (S.INTRA_STRONG_SMOOTHING if rng.random() < 0.7 else 0)
        j["left_avail"], j["top_avail"] = la, ta
    return jobs


def make_recon(w, h, seed, kind="synth"):
    rng = np.random.default_rng(seed)
    if kind == "synth":
        v = common.SynthVideo(w, h, seed=seed)
        y = v.luma(0)
        cb, cr = v.chroma(0)
    elif kind == "smooth":       # near-linear: triggers the strong (bilinear) smoothing of 32x32 blocks
        yy, xx = np.mgrid[0:h, 0:w]
        y = np.clip(40 + xx // 8 + yy // 6 + rng.integers(0, 2, (h, w)), 0, 255).astype(np.uint8)
        cb = np.full((h // 2, w // 2), 120, np.uint8); cr = np.full((h // 2, w // 2), 130, np.uint8)
    else:
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        cb = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8); cr = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    return S.HostYuvPicture(y, cb, cr, pad=80, replicate=False)


def run_intra_batch(lib, name, recon, jobs, *extra):
    pred = S.HostYuvPicture(*[np.zeros_like(a) for a in recon.interior()], pad=80, replicate=False)
    f = getattr(lib, name)
    rc = f(C.byref(recon.struct), C.byref(pred.struct), vp(jobs), C.c_uint32(len(jobs)), *extra)
    assert rc == 0
    return pred


INTRA_CASES = [("synth_256x192", 256, 192, "synth", 0.8), ("random_128x128", 128, 128, "random", 0.5), ("smooth_256x128", 256, 128, "smooth", 0.9),
               ("ragged_200x120", 200, 120, "synth", 0.7)]


@needs_ref
@pytest.mark.parametrize("case", INTRA_CASES, ids=lambda c: c[0])
def test_intra_batch_port_vs_reference(case):
    name, w, h, kind, density = case
    recon = make_recon(w, h, seed=w + h, kind=kind)
    rng = np.random.default_rng(w * 3 + h)
    # jobs overlap in the prediction picture, so compare job by job
    jobs = make_intra_jobs(w & ~31, h & ~31, rng, 400, density)
    for asm in ASM_MODES:
        for j in range(len(jobs)):
            one = jobs[j:j + 1]
            want = run_intra_batch(ref, "svtref_intra_batch", recon, one, C.c_int(asm))
            got = run_intra_batch(port, "port_intra_batch", recon, one)
            for a, b, pl in zip(got.interior(), want.interior(), "y cb cr".split()):
                assert np.array_equal(a, b), (name, asm, j, pl, one)


def test_zscan_rule_matches_reference_tables():
    if ref is None:
        pytest.skip("no reference")
    ref.isBottomLeftAvailable.restype = ref.isUpperRightAvailable.restype = C.c_uint8
    for n in (4, 8, 16, 32):
        depth = {32: 1, 16: 2, 8: 3, 4: 4}[n]
        per = 64 // n
        for by in range(per):
            for bx in range(per):
                bl, tr = zscan_avail(bx * n, by * n, n)
                rbl = bool(ref.isBottomLeftAvailable(C.c_uint32(depth), C.c_uint32(by * per + bx)))
                rtr = bool(ref.isUpperRightAvailable(C.c_uint32(depth), C.c_uint32(by * per + bx)))
                # the reference tables answer for positions INSIDE the LCU; on the LCU's top row / right column the
                # neighbour lies in another LCU and the tables say "available" (the picture geometry decides)
                assert bl == rbl, (n, bx, by, "bl")
                assert tr == rtr, (n, bx, by, "tr")


# ---- open-loop search ------------------------------------------------------------------------------------------------------
def ois_cfg(w, h, **kw):
    c = S.OisConfig()
    c.pic_width, c.pic_height = w, h
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def luma_plane(w, h, seed, kind="synth"):
    if kind == "synth":
        y = common.SynthVideo(w, h, seed=seed).luma(0)
    elif kind == "flat":
        y = np.full((h, w), 77, np.uint8)
    elif kind == "checker":
        yy, xx = np.mgrid[0:h, 0:w]
        y = (((xx // 32 + yy // 32) & 1) * 255).astype(np.uint8)
    else:
        y = np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    buf = np.ascontiguousarray(np.pad(y, 68, mode="edge"))
    pl = S.Plane()
    pl.buf, pl.stride, pl.origin_x, pl.origin_y, pl.width, pl.height = buf.ctypes.data, buf.shape[1], 68, 68, w, h
    return pl, buf


def fake_me(n_lcu, seed, scale):
    rng = np.random.default_rng(seed)
    me = np.zeros((n_lcu, 85), S.ME_PU_RESULT_DTYPE)
    me["distortion"][:, :, 0] = (rng.integers(0, 4000, (n_lcu, 85)) * scale).astype(np.uint32)
    me["distortion"][:, ::7, 0] = 0
    return me


def run_ois(lib, name, cfg, pl, me, cap, *extra):
    n_lcu = ((cfg.pic_width + 63) // 64) * ((cfg.pic_height + 63) // 64)
    total = np.zeros((n_lcu, 85), np.uint8)
    cand = np.zeros((n_lcu, 85, cap), np.uint32)
    rc = getattr(lib, name)(C.byref(cfg), C.byref(pl), vp(me) if me is not None else None, vp(total), vp(cand), C.c_uint32(cap), *extra)
    assert rc == 0
    return total, cand


OIS_VARIANTS = [
    ("i_slice", dict(slice_type=2), 9),
    ("dc_only_m10", dict(slice_type=0, temporal_layer=1, limit_ois_to_dc_mode=1, skip_ois_8x8=1), 1),
    ("dc_only_8x8", dict(slice_type=1, temporal_layer=0, limit_ois_to_dc_mode=1), 9),
    ("kernel_level", dict(slice_type=0, temporal_layer=0, ois_kernel_level=1, skip_ois_8x8=1), 18),
    ("points_th1", dict(slice_type=0, temporal_layer=2, ois_th_set=1, skip_ois_8x8=1), 9),
    ("points_th2_valid", dict(slice_type=1, temporal_layer=0, ois_th_set=2, set_best_ois_distortion_to_valid=1, cu8x8_mode=1), 9),
    ("points_th0_8x8", dict(slice_type=0, temporal_layer=3, ois_th_set=0), 9),
]


@needs_ref
@pytest.mark.parametrize("variant", OIS_VARIANTS, ids=lambda v: v[0])
@pytest.mark.parametrize("pic", [(192, 136, "synth"), (128, 128, "random"), (128, 64, "flat"), (160, 96, "checker")], ids=lambda p: f"{p[0]}x{p[1]}_{p[2]}")
def test_ois_port_vs_reference(variant, pic):
    vname, kw, cap = variant
    w, h, kind = pic
    cfg = ois_cfg(w, h, **kw)
    pl, _keep = luma_plane(w, h, seed=w + h, kind=kind)
    n_lcu = ((w + 63) // 64) * ((h + 63) // 64)
    me = fake_me(n_lcu, seed=5, scale=3 if kind != "flat" else 0)
    for asm in ASM_MODES:
        wt, wc = run_ois(ref, "svtref_ois_picture", cfg, pl, me, cap, C.c_int(asm))
        gt, gc = run_ois(port, "port_ois_picture", cfg, pl, me, cap)
        assert np.array_equal(gt, wt), (vname, asm)
        assert np.array_equal(gc, wc), (vname, asm, np.argwhere(gc != wc)[:5])


# ---- 16-bit samples (10-bit video) ---------------------------------------------------------------------------------------------
def make_recon16(w, h, seed, kind="synth"):
    rng = np.random.default_rng(seed + 77)
    base = make_recon(w, h, seed, kind)
    planes = []
    for a in base.interior():
        planes.append(a.astype(np.uint16) * 4 + (rng.integers(0, 4, a.shape, dtype=np.uint16) if kind != "smooth" else np.uint16(1)))
    return S.HostYuvPicture(*planes, pad=80, replicate=False, dtype=np.uint16)


def run_intra_batch16(lib, name, recon, jobs, *extra):
    pred = S.HostYuvPicture(*[np.zeros_like(a) for a in recon.interior()], pad=80, replicate=False, dtype=np.uint16)
    rc = getattr(lib, name)(C.byref(recon.struct), C.byref(pred.struct), vp(jobs), C.c_uint32(len(jobs)), *extra)
    assert rc == 0
    return pred


@needs_ref
@pytest.mark.parametrize("case", INTRA_CASES, ids=lambda c: c[0])
def test_intra_batch16_port_vs_reference(case):
    name, w, h, kind, density = case
    recon = make_recon16(w, h, seed=w + h, kind=kind)
    rng = np.random.default_rng(w * 3 + h)
    jobs = make_intra_jobs(w & ~31, h & ~31, rng, 250, density)
    for asm in ASM_MODES:
        for j in range(len(jobs)):
            one = jobs[j:j + 1]
            want = run_intra_batch16(ref, "svtref_intra_batch16", recon, one, C.c_int(asm))
            got = run_intra_batch16(port, "port_intra_batch16", recon, one)
            for a, b, pl in zip(got.interior(), want.interior(), "y cb cr".split()):
                assert np.array_equal(a, b), (name, asm, j, pl, one)
