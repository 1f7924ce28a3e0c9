"""Full-size GPU parity tests for the BASELINE configurations beyond configs[1]: a 4K M7 B-frame through ME (HME level 1, sub-pel, two lists), an 8K M10
ME frame, and one 4K frame of 10-bit video through the 16-bit T/Q chain, deblocking and SAO statistics -- every LCU compared bit-exactly with the
reference's own functions when oracle/_ref travelled to this box, else with the plain-C port."""
import numpy as np
import pytest

import common
import svtb200 as S

pytestmark = pytest.mark.gpu
port = common.load_port()
ref = common.load_ref()
ASM = lambda: 1 if common.ref_asm_simd_ok() else 0


def _me_frame(w, h, enc_mode, tl, nl, seed, **kw):
    v = common.SynthVideo(w, h, seed=seed)
    src, r0, r1 = (S.HostPaPicture(v.luma(n)) for n in (4, 0, 8))
    cfg = S.me_config(w, h, enc_mode=enc_mode, temporal_layer=tl, num_lists=nl, **kw)
    r1 = r1 if nl == 2 else None
    got = S.me_picture(cfg, src, r0, r1, want_raw=True, want_rc=True)
    if ref is not None:
        want = common.oracle_me_picture(ref, "svtref_me_picture", cfg, src, r0, r1, asm_mode=ASM())
    else:
        want = common.oracle_me_picture(port, "port_me_picture", cfg, src, r0, r1)
    common.assert_me_equal(got, want, nl, f"{w}x{h} M{enc_mode} tl{tl} CUDA vs oracle")
    return cfg


def test_me_picture_4k_m7_bframe():
    """BASELINE configs[2] geometry (4K, M7): a non-reference B picture of the top layer -- HME level 0 + 1, sub-pel refinement, both lists, bi-prediction"""
    cfg = _me_frame(3840, 2160, 7, 3, 2, 0x5F7C0003, is_reference=False, slice_type="B")
    assert cfg.enable_hme_l1 == 1 and cfg.fractional_search_model == 1


def test_me_picture_4k_m7_reference_p():
    cfg = _me_frame(3840, 2160, 7, 0, 1, 0x5F7C0004, is_reference=True, slice_type="P")
    assert cfg.fractional_search_model == 1


def test_me_picture_8k_m10():
    """BASELINE configs[3] geometry (7680x4320, M10): 8160 LCUs, integer-pel, one-quadrant HME"""
    _me_frame(7680, 4320, 10, 2, 2, 0x5F7C0005)


def test_4k_10bit_frame_tq_dlf_sao():
    """One 3840x2160 frame of 10-bit video: every TU through the 16-bit T/Q chain in place, the reconstruction deblocked, SAO statistics of the result"""
    from test_oracle_dlf import make_dlf_inputs, run_dlf16
    from test_oracle_sao import assert_stats_equal, run_stats
    from test_oracle_tq import assert_tq_equal, make_src_pred16, make_tq_jobs, run_tq_oracle16
    w, h = 3840, 2160
    src, pred = make_src_pred16(w, h, 41, "synth")
    jobs = make_tq_jobs(w, h, seed=6, shapes=(0, 0, 0, 1, 2), flags=(0, 1))
    if ref is not None:
        want = run_tq_oracle16(ref, "svtref_tq_batch16", src, pred, jobs, w, h, asm_mode=ASM())
    else:
        want = run_tq_oracle16(port, "port_tq_batch16", src, pred, jobs, w, h)
    coeff = S.HostCoeffPicture(w, h)
    nz = S.tq_batch(src, pred, pred, coeff, jobs, hbd=True)      # in place, as the encode pass runs it
    assert_tq_equal((pred, coeff, nz), want, "4K 16-bit T/Q CUDA vs oracle")
    assert nz.sum() > 0
    # deblock the reconstruction with a synthetic BS map
    _, vbs, hbs, qp, prm = make_dlf_inputs(w, h, seed=9, density=0.5, kind="flat")
    recon = want[0]
    if ref is not None:
        dl = run_dlf16(ref, "svtref_dlf_picture16", recon, vbs, hbs, qp, prm, asm_mode=ASM())
    else:
        dl = run_dlf16(port, "port_dlf_picture16", recon, vbs, hbs, qp, prm)
    S.dlf_picture(pred, vbs, hbs, qp, prm, hbd=True)
    for a, b, nm in zip(pred.interior(), dl.interior(), "Y Cb Cr".split()):
        assert np.array_equal(a, b), f"4K 16-bit DLF: {nm} differs: {np.argwhere(a != b)[:5]}"
    # SAO statistics of the deblocked frame against the source
    ws = run_stats(port, "port_sao_stats_picture16", src, dl)
    gs = S.sao_stats_picture(src, pred, hbd=True)
    assert_stats_equal(gs, ws, "4K 16-bit SAO statistics CUDA vs port")
    assert int(np.abs(gs["eo_count"]).sum()) > 0
