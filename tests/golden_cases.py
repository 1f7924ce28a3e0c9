"""Golden fixtures of the hot path (the reference stores no vectors for this path, so they are generated HERE from the compiled,
unmodified reference -- oracle/_ref -- by tests/golden/make_golden.py and committed as SHA-256 digests of canonical output bytes).

One case = (stage, name, build() -> inputs, digest(outputs)).  `run_case(case, backend)` produces the digest with
  backend "reference": the reference's own functions (oracle/_ref harness entry points, C rows),
  backend "port"     : oracle/port (the plain-C restatement),
  backend "cuda"     : libsvthevc_b200.so through its C ABI.
All three must hash to the committed value; the inputs are seeded and built by the same helpers the parity tests use."""
import ctypes as C
import hashlib

import numpy as np

import common
import svtb200 as S


def _h(*arrays):
    m = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        m.update(str(a.dtype).encode()); m.update(str(a.shape).encode()); m.update(a.tobytes())
    return m.hexdigest()


def _planes(pic):
    return [np.ascontiguousarray(p) for p in pic.interior()]


# ---- motion estimation ------------------------------------------------------------------------------------------------------
def _me_digest(out, nl):
    res, bs, bm, rc = out
    fields = [res[f] for f in res.dtype.names if not (nl == 1 and f in ("x_mv_l1", "y_mv_l1"))]
    return _h(*fields, bs[:, :nl], bm[:, :nl], rc)


def _me(case, backend):
    from test_oracle_me import build_me_case
    name, w, h, kw, content = case
    cfg, (src, r0, r1) = build_me_case(w, h, kw, content)
    r1 = r1 if cfg.num_lists == 2 else None
    if backend == "reference":
        out = common.oracle_me_picture(common.load_ref(), "svtref_me_picture", cfg, src, r0, r1, asm_mode=0)
    elif backend == "port":
        out = common.oracle_me_picture(common.load_port(), "port_me_picture", cfg, src, r0, r1)
    else:
        out = S.me_picture(cfg, src, r0, r1, want_raw=True, want_rc=True)
    return _me_digest(out, cfg.num_lists)


# ---- motion compensation ------------------------------------------------------------------------------------------------------
def _mc(case, backend):
    from test_oracle_mc import make_jobs, make_pictures, run_oracle
    name, w, h, kind, all_fracs, big = case
    refs = make_pictures(w, h, seed=11, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=5, all_fracs=all_fracs, big_mv=big)
    if backend == "reference":
        out = run_oracle(common.load_ref(), "svtref_mc_batch", refs, jobs, w, h, asm_mode=0)
    elif backend == "port":
        out = run_oracle(common.load_port(), "port_mc_batch", refs, jobs, w, h)
    else:
        out = S.HostYuvPicture(np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8))
        S.mc_batch(refs, jobs, out)
    return _h(*_planes(out))


# ---- transform / quantisation chain ---------------------------------------------------------------------------------------------
def _tq(case, backend):
    from test_oracle_tq import make_src_pred, make_tq_jobs, run_tq_oracle
    name, w, h, kind, shapes, flags = case
    src, pred = make_src_pred(w, h, 31, kind)
    jobs = make_tq_jobs(w, h, seed=8, shapes=shapes, flags=flags)
    if backend == "reference":
        recon, coeff, nz = run_tq_oracle(common.load_ref(), "svtref_tq_batch", src, pred, jobs, w, h, asm_mode=0)
    elif backend == "port":
        recon, coeff, nz = run_tq_oracle(common.load_port(), "port_tq_batch", src, pred, jobs, w, h)
    else:
        y, cb, cr = (a.copy() for a in pred.interior())
        recon = S.HostYuvPicture(y, cb, cr)
        coeff = S.HostCoeffPicture(w, h)
        nz = S.tq_batch(src, pred, recon, coeff, jobs)
    return _h(*_planes(recon), *[np.ascontiguousarray(p) for p in coeff.planes], np.asarray(nz, np.uint32))


# ---- deblocking ---------------------------------------------------------------------------------------------------------------
def _dlf(case, backend):
    from test_oracle_dlf import make_dlf_inputs, run_dlf
    name, w, h, kind, density = case
    pic, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    if backend == "reference":
        out = run_dlf(common.load_ref(), "svtref_dlf_picture", pic, vbs, hbs, qp, prm, asm_mode=0)
    elif backend == "port":
        out = run_dlf(common.load_port(), "port_dlf_picture", pic, vbs, hbs, qp, prm)
    else:
        out = S.HostYuvPicture(*(a.copy() for a in pic.interior()))
        S.dlf_picture(out, vbs, hbs, qp, prm)
    return _h(*_planes(out))


# ---- SAO ------------------------------------------------------------------------------------------------------------------------
def _sao_stats(case, backend):
    from test_oracle_sao import load_ref_sao, make_sao_inputs, run_stats
    name, w, h, kind = case
    src, rec, _ = make_sao_inputs(w, h, seed=w * 3 + h, kind=kind)
    if backend == "reference":
        st = run_stats(load_ref_sao(), "svtref_sao_stats_picture", src, rec, asm_mode=0)
    elif backend == "port":
        st = run_stats(common.load_port(), "port_sao_stats_picture", src, rec)
    else:
        st = S.sao_stats_picture(src, rec)
    # slot [4] of the EO arrays is scratch in the reference (tests/test_oracle_sao.py)
    return _h(st["bo_diff"], st["bo_count"], st["eo_diff"][..., :4], st["eo_count"][..., :4])


def _sao_apply(case, backend):
    from test_oracle_sao import make_sao_inputs, run_apply_port, run_apply_ref
    name, w, h, kind = case
    _, rec, prm = make_sao_inputs(w, h, seed=w * 3 + h, kind=kind)
    if backend == "reference":
        out = run_apply_ref(rec, prm, (1, 1), asm_mode=0)
    elif backend == "port":
        out = run_apply_port(rec, prm, (1, 1))
    else:
        out = S.HostYuvPicture(*(np.zeros_like(a) for a in rec.interior()))
        S.sao_apply_picture(rec, out, prm, (1, 1))
    return _h(*_planes(out))


# ---- open-loop intra search -----------------------------------------------------------------------------------------------------
def _ois(case, backend):
    from test_oracle_intra import fake_me, luma_plane, ois_cfg, run_ois
    name, w, h, kw, cap = case
    cfg = ois_cfg(w, h, **kw)
    pl, _keep = luma_plane(w, h, seed=w + h, kind="synth")
    n_lcu = ((w + 63) // 64) * ((h + 63) // 64)
    me = fake_me(n_lcu, seed=5, scale=3)
    if backend == "reference":
        t, c = run_ois(common.load_ref(), "svtref_ois_picture", cfg, pl, me, cap, C.c_int(0))
    elif backend == "port":
        t, c = run_ois(common.load_port(), "port_ois_picture", cfg, pl, me, cap)
    else:
        t, c = S.ois_picture(cfg, pl, me, cap)
    return _h(t, c)


# ---- 16-bit samples (10-bit video): MC and the T/Q chain ---------------------------------------------------------------------------
def _mc16(case, backend):
    from test_oracle_mc import make_jobs, make_pictures16, run_oracle16
    name, w, h, kind, all_fracs, big = case
    refs = make_pictures16(w, h, seed=w + h, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=w, all_fracs=all_fracs, big_mv=big)
    if backend == "reference":
        out = run_oracle16(common.load_ref(), "svtref_mc_batch16", refs, jobs, w, h, asm_mode=0)
    elif backend == "port":
        out = run_oracle16(common.load_port(), "port_mc_batch16", refs, jobs, w, h)
    else:
        out = S.HostYuvPicture(np.zeros((h, w), np.uint16), np.zeros((h // 2, w // 2), np.uint16), np.zeros((h // 2, w // 2), np.uint16), dtype=np.uint16)
        S.mc_batch16(refs, jobs, out)
    return _h(*_planes(out))


def _tq16(case, backend):
    from test_oracle_tq import make_src_pred16, make_tq_jobs, run_tq_oracle16
    name, w, h, kind, shapes, flags = case
    src, pred = make_src_pred16(w, h, 31, kind)
    jobs = make_tq_jobs(w, h, seed=8, shapes=shapes, flags=flags)
    if backend == "reference":
        recon, coeff, nz = run_tq_oracle16(common.load_ref(), "svtref_tq_batch16", src, pred, jobs, w, h, asm_mode=0)
    elif backend == "port":
        recon, coeff, nz = run_tq_oracle16(common.load_port(), "port_tq_batch16", src, pred, jobs, w, h)
    else:
        y, cb, cr = (a.copy() for a in pred.interior())
        recon = S.HostYuvPicture(y, cb, cr, dtype=np.uint16)
        coeff = S.HostCoeffPicture(w, h)
        nz = S.tq_batch(src, pred, recon, coeff, jobs, hbd=True)
    return _h(*_planes(recon), *[np.ascontiguousarray(p) for p in coeff.planes], np.asarray(nz, np.uint32))


# ---- encode-pass intra prediction ---------------------------------------------------------------------------------------------------
def _intra(case, backend):
    from test_oracle_intra import make_intra_jobs, make_recon, run_intra_batch
    name, w, h, seed, n_jobs = case
    jobs = make_intra_jobs(w, h, np.random.default_rng(seed), n_jobs, 0.8)
    digests = []
    # a sample of the jobs, one at a time, each predicted from the same reconstructed picture into an empty prediction picture
    for j in range(0, len(jobs), max(1, len(jobs) // 24)):
        recon = make_recon(w, h, seed=3)
        one = jobs[j:j + 1]
        if backend == "reference":
            pred = run_intra_batch(common.load_ref(), "svtref_intra_batch", recon, one, C.c_int(0))
        elif backend == "port":
            pred = run_intra_batch(common.load_port(), "port_intra_batch", recon, one)
        else:
            pred = S.HostYuvPicture(*[np.zeros_like(a) for a in recon.interior()], pad=80, replicate=False)
            S.intra_batch(recon, pred, one)
        digests.append(_h(*_planes(pred)))
    return _h(np.frombuffer("".join(digests).encode(), np.uint8))


# ---- SAO decision and H3 statistics ----------------------------------------------------------------------------------------------------------
NO_CUDA_YET = ("none/",)      # (every golden case has a CUDA runner)


def _saodec(case, backend):
    import test_oracle_saodec as T
    name, sao_case, rates, lam, lam_c = case
    stats, cols, rows = T.picture_stats(sao_case)
    if backend == "cuda":
        d = S.sao_decision_picture(stats, cols, rows, T.rate_tables(rates), lam, lam_c)
    else:
        lib, fn = (T.refdec, "svtref_sao_decision_picture") if backend == "reference" else (T.port, "port_sao_decision_picture")
        assert lib is not None
        d = T.run_decision(lib, fn, stats, cols, rows, T.rate_tables(rates), lam, lam_c)
    return _h(*[d[f] for f in ("merge_left", "merge_up", "type", "band_position", "offset", "luma_cost", "chroma_cost")])


def _sao_h3(case, backend):
    import test_oracle_sao_h3 as T
    from test_oracle_dlf import make_dlf_inputs
    name, w, h, kind, density = case
    pic, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=w + h, density=density, kind=kind)
    src = T.noisy_source(pic, w * 7 + h)
    if backend == "cuda":      # picture-order deblocking of a copy on the device, then the H3 statistics from (deblocked, unfiltered)
        final = S.HostYuvPicture(*[a.copy() for a in pic.interior()], pad=pic.pad)
        S.dlf_picture(final, vbs, hbs, qp, prm)
        st = S.sao_stats_picture_h3(src, final, pic)
    else:
        st = T.reference_interleaved(pic, src, vbs, hbs, qp, prm, asm_mode=0)[2] if backend == "reference" else T.port_h3(pic, src, vbs, hbs, qp, prm)[2]
    return _h(st["bo_diff"], st["bo_count"], st["eo_diff"][..., :4], st["eo_count"][..., :4])


# ---- round-2 rows: boundary strengths, distortion leaves, interpolation-free prediction, mean / variance pyramid --------------------------------------
def _bs(case, backend):
    import test_oracle_bs as T
    name, w, h, st, p0, p1 = case
    cu, cbf = T.make_md_maps(w, h, seed=w + h + st, slice_type=st)
    if backend == "cuda":
        v, hh = S.dlf_bs_picture(cu, cbf, w, h, st, p0, p1)
    else:
        lib, fn = (T.ref, "svtref_dlf_bs_picture") if backend == "reference" else (T.port, "port_dlf_bs_picture")
        assert lib is not None
        v, hh = T.run_bs(lib, fn, cu, cbf, w, h, st, p0, p1)
    return _h(v, hh)


def _dist(case, backend):
    """spatial SSD + SATD of one picture pair and the coefficient-domain distortion of one coefficient pair; the reference side calls the leaf kernels block by block"""
    import test_oracle_dist as T
    from test_oracle_mc import make_pictures
    name, w, h, kind = case
    a, b = make_pictures(w, h, seed=w + 1, n=2, kind=kind)
    c, r = T.make_coeff_pictures(w, h, seed=h + 2)
    js = T.make_dist_jobs(w, h, seed=3, n=120)
    jt = T.make_dist_jobs(w, h, seed=4, n=120, kinds=(S.SATD_8X8, S.SATD_4X4), unit_of=lambda k: 8 if k == S.SATD_8X8 else 4)
    jf = T.make_dist_jobs(w, h, seed=5, n=120, kinds=(S.FULLDIST_DEFAULT, S.FULLDIST_INTRA, S.FULLDIST_CBF_ZERO))
    jf = jf[jf["width"] <= 32]
    if backend == "cuda":
        return _h(S.spatial_distortion_batch(a, b, js), S.satd_batch(a, b, jt), S.full_distortion_batch(c, r, w, h, jf))
    if backend == "port":
        return _h(T.run_port_batch("port_spatial_distortion_batch", a, b, js), T.run_port_batch("port_satd_batch", a, b, jt),
                  T.run_port_batch("port_full_distortion_batch", c, r, jf, per_job=2))
    ref = T.ref
    assert ref is not None
    vp = lambda arr, off=0: C.c_void_p(arr.ctypes.data + off)
    ref.SpatialFullDistortionKernel.restype = ref.Compute8x8Satd.restype = ref.Compute4x4Satd.restype = C.c_uint64
    ssd = np.zeros(len(js), np.uint64)
    for j, t in enumerate(js):
        pa, pb = a.interior()[t["plane"]], b.interior()[t["plane"]]
        ssd[j] = ref.SpatialFullDistortionKernel(vp(pa, int(t["y"]) * pa.strides[0] + int(t["x"])), C.c_uint32(pa.strides[0]), vp(pb, int(t["y"]) * pb.strides[0] + int(t["x"])),
                                                 C.c_uint32(pb.strides[0]), C.c_uint32(int(t["width"])), C.c_uint32(int(t["height"])))
    satd = np.zeros(len(jt), np.uint64)
    for j, t in enumerate(jt):
        pa, pb = a.interior()[t["plane"]], b.interior()[t["plane"]]
        u = 8 if t["kind"] == S.SATD_8X8 else 4
        fn = ref.Compute8x8Satd if u == 8 else ref.Compute4x4Satd
        tot = 0
        for uy in range(int(t["height"]) // u):
            for ux in range(int(t["width"]) // u):
                y0, x0 = int(t["y"]) + uy * u, int(t["x"]) + ux * u
                d = np.ascontiguousarray(pa[y0:y0 + u, x0:x0 + u].astype(np.int16) - pb[y0:y0 + u, x0:x0 + u].astype(np.int16))
                tot += fn(vp(d))
        satd[j] = tot
    names = {S.FULLDIST_DEFAULT: "FullDistortionKernel_32bit", S.FULLDIST_INTRA: "FullDistortionKernelIntra_32bit", S.FULLDIST_CBF_ZERO: "FullDistortionKernelCbfZero_32bit"}
    full = np.zeros((len(jf), 2), np.uint64)
    for j, t in enumerate(jf):
        pc, pr = c.planes[t["plane"]], r.planes[t["plane"]]
        out = (C.c_uint64 * 2)()
        off = (int(t["y"]) * pc.shape[1] + int(t["x"])) * 2
        getattr(ref, names[int(t["kind"])])(vp(pc, off), C.c_uint32(pc.shape[1]), vp(pr, off), C.c_uint32(pr.shape[1]), out, C.c_uint32(int(t["width"])), C.c_uint32(int(t["height"])))
        full[j] = (out[0], out[1])
    return _h(ssd, satd, full)


def _mc_ifree(case, backend):
    from test_oracle_mc import make_jobs, make_pictures, run_oracle
    name, w, h, kind, all_fracs, big = case
    refs = make_pictures(w, h, seed=13, kind=kind)
    jobs = make_jobs(w, h, len(refs), seed=6, all_fracs=all_fracs, big_mv=big)
    if backend == "reference":
        out = run_oracle(common.load_ref(), "svtref_mc_ifree_batch", refs, jobs, w, h, asm_mode=0)
    elif backend == "port":
        out = run_oracle(common.load_port(), "port_mc_ifree_batch", refs, jobs, w, h)
    else:
        out = S.HostYuvPicture(np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.zeros((h // 2, w // 2), np.uint8))
        S.mc_ifree_batch(refs, jobs, out)
    return _h(*_planes(out))


def _variance(case, backend):
    import test_oracle_pa as T
    name, w, h, kind = case
    pl, _keep = T.padded_luma(w, h, seed=w + h, kind=kind)
    if backend == "reference":
        assert T.refpa is not None
        ym, vr = T.run_variance(T.refpa, "svtref_variance_picture", pl, C.c_int(0))
    elif backend == "port":
        ym, vr = T.run_variance(T.port, "port_variance_picture", pl)
    else:
        ym, vr = S.variance_picture(pl)
    return _h(ym, vr)


def cases():
    """[(key, runner, case)] -- small, seeded, a few seconds for all of them on any backend"""
    from test_oracle_dlf import DLF_CASES
    from test_oracle_mc import MC_CASES
    from test_oracle_me import _me_cases
    from test_oracle_sao import SAO_APPLY_CASES, SAO_CASES
    from test_oracle_tq import TQ_CASES
    out = []
    me = _me_cases()
    for c in me[:: max(1, len(me) // 6)][:6]:
        out.append((f"me/{c[0]}", _me, c))
    for c in MC_CASES:
        out.append((f"mc/{c[0]}", _mc, c))
    for c in TQ_CASES:
        out.append((f"tq/{c[0]}", _tq, c))
    for c in DLF_CASES[:4]:
        out.append((f"dlf/{c[0]}", _dlf, c))
    for c in SAO_CASES[:3]:
        out.append((f"sao_stats/{c[0]}", _sao_stats, c))
    for c in SAO_APPLY_CASES[:2]:
        out.append((f"sao_apply/{c[0]}", _sao_apply, c))
    for c in MC_CASES[:2]:
        out.append((f"mc16/{c[0]}", _mc16, c))
    for c in TQ_CASES[:2] + TQ_CASES[3:4]:
        out.append((f"tq16/{c[0]}", _tq16, c))
    out.append(("intra/single_jobs_256x192", _intra, ("single_jobs", 256, 192, 9, 300)))
    out.append(("saodec/synth_typical", _saodec, ("synth_typical", SAO_CASES[0], "typical", 439162, 350000)))
    out.append(("saodec/partial_expensive", _saodec, ("partial_expensive", SAO_CASES[1], "expensive", 9000000, 7000000)))
    out.append(("sao_h3/synth_dense_320x256", _sao_h3, DLF_CASES[4]))
    out.append(("sao_h3/synth_partial_lcu_200x136", _sao_h3, DLF_CASES[1]))
    out.append(("ois/islice_256x192", _ois, ("islice", 256, 192, dict(slice_type=2), 9)))
    out.append(("ois/dc_only_448x256", _ois, ("dc_only", 448, 256, dict(slice_type=0, limit_ois_to_dc_mode=1, skip_ois_8x8=1), 1)))
    from test_oracle_bs import BS_CASES
    from test_oracle_pa import PA_CASES
    for c in BS_CASES[:3] + BS_CASES[4:5]:
        out.append((f"bs/{c[0]}", _bs, c))
    out.append(("dist/synth_192x128", _dist, ("synth", 192, 128, "synth")))
    out.append(("dist/random_128x64", _dist, ("random", 128, 64, "random")))
    for c in MC_CASES[:1] + MC_CASES[3:4]:
        out.append((f"mc_ifree/{c[0]}", _mc_ifree, c))
    for c in PA_CASES[:1] + PA_CASES[2:3]:
        out.append((f"variance/{c[0]}", _variance, c))
    return out
