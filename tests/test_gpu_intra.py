"""GPU parity tests for intra prediction (through the C ABI): mode kernels, encode-pass batches and the open-loop intra search,
against the plain-C oracle and -- where oracle/_ref is present -- the reference's own functions."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_intra import (INTRA_CASES, OIS_VARIANTS, fake_me, line_for, luma_plane, make_intra_jobs, make_recon, ois_cfg, port_predict, run_intra_batch,
                               run_ois)

pytestmark = pytest.mark.gpu
port = common.load_port()
ref = common.load_ref()
vp = lambda a: C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("kind", ["random", "flat", "ramp", "extreme"])
def test_mode_leaf_cuda_vs_oracle(kind):
    rng = np.random.default_rng(11)
    for n in (4, 8, 16, 32):
        for chroma in (0, 1):
            for mode in range(35):
                line = line_for(kind, n, rng)
                assert np.array_equal(S.intra_prediction(mode, n, chroma, line), port_predict(mode, n, chroma, line)), (n, chroma, mode)


def disjoint_jobs(w, h, rng, density):
    """one job per 32x32 cell (random size inside it), so that a whole batch can be predicted in place and compared at once"""
    cells = [(cx, cy) for cy in range(0, h - 31, 32) for cx in range(0, w - 31, 32)]
    jobs = make_intra_jobs(w, h, rng, len(cells), density)
    for j, (cx, cy) in zip(jobs, cells):
        n = int(j["size"])
        j["x"], j["y"] = cx + (int(j["x"]) % 32) // n * n, cy + (int(j["y"]) % 32) // n * n
    return jobs


@pytest.mark.parametrize("case", INTRA_CASES + [("uhd_strip_3840x256", 3840, 256, "synth", 0.85)], ids=lambda c: c[0])
def test_intra_batch_cuda_vs_oracle(case):
    name, w, h, kind, density = case
    recon = make_recon(w, h, seed=w + h, kind=kind)
    rng = np.random.default_rng(w * 3 + h)
    jobs = disjoint_jobs(w & ~31, h & ~31, rng, density)
    # masks were drawn for other positions: re-derive nothing, availability words are arbitrary inputs for port and CUDA alike;
    # only picture-border reads must stay inside the padded planes, which the 80-sample padding guarantees
    want = run_intra_batch(port, "port_intra_batch", recon, jobs)
    got = S.HostYuvPicture(*[np.zeros_like(a) for a in recon.interior()], pad=80, replicate=False)
    S.intra_batch(recon, got, jobs)
    for a, b, pl in zip(got.interior(), want.interior(), "y cb cr".split()):
        assert np.array_equal(a, b), (name, pl, np.argwhere(a != b)[:4])


@pytest.mark.skipif(ref is None, reason="oracle/_ref not present")
def test_intra_batch_cuda_vs_reference_in_place():
    """jobs generated with consistent z-scan availability, predicted IN PLACE into the recon picture like the reference does"""
    w, h = 256, 192
    rng = np.random.default_rng(9)
    jobs = make_intra_jobs(w, h, rng, 300, 0.8)
    for j in range(0, len(jobs), 10):
        one = jobs[j:j + 1]
        a = make_recon(w, h, seed=3)
        b = make_recon(w, h, seed=3)
        rc = ref.svtref_intra_batch(C.byref(a.struct), C.byref(a.struct), vp(one), C.c_uint32(1), C.c_int(1 if common.ref_asm_simd_ok() else 0))
        assert rc == 0
        S.intra_batch(b, b, one)
        for p, q in zip(a.interior(), b.interior()):
            assert np.array_equal(p, q), (j, one)


def test_intra_batch_rejects_bad_jobs():
    recon = make_recon(64, 64, seed=1)
    jobs = np.zeros(1, S.INTRA_JOB_DTYPE)
    jobs[0]["size"], jobs[0]["luma_mode"], jobs[0]["chroma_mode"] = 12, 3, S.INTRA_SKIP
    with pytest.raises(S.Svtb200Error):
        S.intra_batch(recon, recon, jobs)
    jobs[0]["size"], jobs[0]["luma_mode"] = 8, 35
    with pytest.raises(S.Svtb200Error):
        S.intra_batch(recon, recon, jobs)
    jobs[0]["size"], jobs[0]["luma_mode"], jobs[0]["x"] = 32, 1, 48
    with pytest.raises(S.Svtb200Error):
        S.intra_batch(recon, recon, jobs)
    S.intra_batch(recon, recon, jobs[:0])      # empty batch is a no-op


@pytest.mark.parametrize("variant", OIS_VARIANTS, ids=lambda v: v[0])
@pytest.mark.parametrize("pic", [(192, 136, "synth"), (128, 128, "random"), (128, 64, "flat"), (160, 96, "checker")], ids=lambda p: f"{p[0]}x{p[1]}_{p[2]}")
def test_ois_cuda_vs_oracle(variant, pic):
    vname, kw, cap = variant
    w, h, kind = pic
    cfg = ois_cfg(w, h, **kw)
    pl, _keep = luma_plane(w, h, seed=w + h, kind=kind)
    n_lcu = ((w + 63) // 64) * ((h + 63) // 64)
    me = fake_me(n_lcu, seed=5, scale=3 if kind != "flat" else 0)
    wt, wc = run_ois(port, "port_ois_picture", cfg, pl, me, cap)
    gt, gc = S.ois_picture(cfg, pl, me, cap)
    assert np.array_equal(gt, wt), vname
    assert np.array_equal(gc, wc), (vname, np.argwhere(gc != wc)[:5])


@pytest.mark.parametrize("variant", [OIS_VARIANTS[1], OIS_VARIANTS[0], OIS_VARIANTS[4]], ids=lambda v: v[0])
def test_ois_4k_frame(variant):
    vname, kw, cap = variant
    w, h = 3840, 2160
    cfg = ois_cfg(w, h, **kw)
    pl, _keep = luma_plane(w, h, seed=7, kind="synth")
    me = fake_me(2040, seed=6, scale=2)
    if ref is not None:
        wt, wc = run_ois(ref, "svtref_ois_picture", cfg, pl, me, cap, C.c_int(1 if common.ref_asm_simd_ok() else 0))
    else:
        wt, wc = run_ois(port, "port_ois_picture", cfg, pl, me, cap)
    gt, gc = S.ois_picture(cfg, pl, me, cap)
    assert np.array_equal(gt, wt) and np.array_equal(gc, wc), vname
    # a checksum of checksums that does not depend on the oracle: every valid CU reports at least one candidate
    assert int(gt.sum()) > 0


def test_ois_rejects_bad_config():
    pl, _keep = luma_plane(128, 64, seed=1)
    cfg = ois_cfg(128, 64, slice_type=0, ois_kernel_level=1)
    with pytest.raises(S.Svtb200Error):
        S.ois_picture(cfg, pl, None, 9)          # kernel level needs 18 slots
    cfg = ois_cfg(128, 64, slice_type=0)
    with pytest.raises(S.Svtb200Error):
        S.ois_picture(cfg, pl, None, 9)          # the OIS-point path reads ME distortions


@pytest.mark.parametrize("case", INTRA_CASES, ids=lambda c: c[0])
def test_intra_batch16_cuda_vs_oracle(case):
    """16-bit sample planes (10-bit video)"""
    from test_oracle_intra import make_recon16, run_intra_batch16
    name, w, h, kind, density = case
    recon = make_recon16(w, h, seed=w + h, kind=kind)
    rng = np.random.default_rng(w * 3 + h)
    jobs = disjoint_jobs(w & ~31, h & ~31, rng, density)
    want = run_intra_batch16(port, "port_intra_batch16", recon, jobs)
    got = S.HostYuvPicture(*[np.zeros_like(a) for a in recon.interior()], pad=80, replicate=False, dtype=np.uint16)
    S.intra_batch(recon, got, jobs, hbd=True)
    for a, b, pl in zip(got.interior(), want.interior(), "y cb cr".split()):
        assert np.array_equal(a, b), (name, pl, np.argwhere(a != b)[:4])
    if ref is not None:
        one = jobs[:1]
        wr = run_intra_batch16(ref, "svtref_intra_batch16", recon, one, C.c_int(1 if common.ref_asm_simd_ok() else 0))
        g1 = S.HostYuvPicture(*[np.zeros_like(a) for a in recon.interior()], pad=80, replicate=False, dtype=np.uint16)
        S.intra_batch(recon, g1, one, hbd=True)
        for a, b in zip(g1.interior(), wr.interior()):
            assert np.array_equal(a, b)
