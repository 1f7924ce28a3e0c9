"""GPU parity tests for mode decision's distortion leaves (through the C ABI) against the plain-C oracle (itself pinned on the reference's C and SSE rows
by tests/test_oracle_dist.py)."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_dist import make_coeff_pictures, make_dist_jobs, run_port_batch
from test_oracle_mc import make_pictures

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wh,kind", [((192, 128), "synth"), ((200, 136), "random"), ((128, 64), "extremes"), ((3840, 2160), "synth")], ids=lambda v: str(v))
def test_spatial_distortion_cuda_vs_oracle(wh, kind):
    w, h = wh
    a, b = make_pictures(w, h, seed=w + h, n=2, kind=kind)
    jobs = make_dist_jobs(w, h, seed=2, n=4000 if w > 1000 else 400)
    want = run_port_batch("port_spatial_distortion_batch", a, b, jobs)
    got = S.spatial_distortion_batch(a, b, jobs)
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]
    assert got.sum() > 0 or kind == "flat"


@pytest.mark.parametrize("extreme", [False, True])
def test_full_distortion_cuda_vs_oracle(extreme):
    w, h = 256, 192
    c, r = make_coeff_pictures(w, h, seed=7, extreme=extreme)
    jobs = make_dist_jobs(w, h, seed=4, n=600, kinds=(S.FULLDIST_DEFAULT, S.FULLDIST_INTRA, S.FULLDIST_CBF_ZERO))
    jobs = jobs[jobs["width"] <= 32]
    want = run_port_batch("port_full_distortion_batch", c, r, jobs, per_job=2)
    got = S.full_distortion_batch(c, r, w, h, jobs)
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]
    if extreme:      # the 32-bit sums of the reference wrap: make sure the case is really exercised
        exact = [(c.planes[t["plane"]][t["y"]:t["y"] + t["height"], t["x"]:t["x"] + t["width"]].astype(np.int64) ** 2).sum() for t in jobs]
        assert max(exact) >= 1 << 32


@pytest.mark.parametrize("wh,kind", [((192, 128), "synth"), ((200, 136), "random"), ((128, 64), "extremes")], ids=lambda v: str(v))
def test_satd_cuda_vs_oracle(wh, kind):
    w, h = wh
    a, b = make_pictures(w, h, seed=w, n=2, kind=kind)
    jobs = make_dist_jobs(w, h, seed=6, n=400, kinds=(S.SATD_8X8, S.SATD_4X4), unit_of=lambda k: 8 if k == S.SATD_8X8 else 4)
    want = run_port_batch("port_satd_batch", a, b, jobs)
    got = S.satd_batch(a, b, jobs)
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]


def test_distortion_rejects_bad_jobs():
    w, h = 128, 64
    a, b = make_pictures(w, h, seed=1, n=2, kind="synth")
    jobs = make_dist_jobs(w, h, seed=1, n=4)
    jobs[1]["x"], jobs[1]["plane"], jobs[1]["width"] = w - 8, 0, 16      # block leaves the picture
    with pytest.raises(S.Svtb200Error):
        S.spatial_distortion_batch(a, b, jobs)
    jobs = make_dist_jobs(w, h, seed=1, n=4)
    jobs[0]["kind"], jobs[0]["width"], jobs[0]["height"], jobs[0]["x"], jobs[0]["y"] = S.SATD_8X8, 12, 8, 0, 0      # not a multiple of the 8x8 unit
    with pytest.raises(S.Svtb200Error):
        S.satd_batch(a, b, jobs)
    assert len(S.spatial_distortion_batch(a, b, jobs[:0])) == 0      # empty batch is a no-op
