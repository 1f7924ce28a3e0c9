"""GPU parity tests for the residual/transform/quant/inverse/recon chain (through the C ABI) against the plain-C oracle
and the reference's own EstimateTransform / UnifiedQuantizeInvQuantize / EncodeInvTransform sequence."""
import ctypes as C

import numpy as np
import pytest

import common
import svtb200 as S
from test_oracle_tq import (FWD_FUNCS, INV_FUNCS, TQ_CASES, assert_tq_equal, call_tr, make_src_pred, make_tq_jobs, residual_block,
                            run_tq_oracle)

pytestmark = pytest.mark.gpu
port = common.load_port()
ref = common.load_ref()
vp = lambda a: a.ctypes.data_as(C.c_void_p)
u = C.c_uint32


@pytest.mark.parametrize("key", list(FWD_FUNCS), ids=lambda k: f"N{k[0]}dst{k[1]}")
@pytest.mark.parametrize("kind", ["random", "flat", "extreme", "smooth"])
def test_forward_transform_leaf(key, kind):
    N, dst = key
    rng = np.random.default_rng(N * 10 + dst)
    shapes = (0,) if N == 4 else (0, 1, 3) if N == 8 else (0, 1, 2, 3)
    for bit_inc, amp in ((0, 255), (2, 1023)):
        for shape in shapes:
            r = residual_block(N, kind, rng, amp)
            want = np.zeros((N, N), np.int16)
            port.port_estimate_transform(vp(r), u(N), vp(want), u(N), u(N), u(bit_inc), C.c_uint8(dst), C.c_uint8(shape))
            got = S.estimate_transform(r, N, bit_inc, dst, shape)
            assert np.array_equal(got, want), f"shape {shape} bit_inc {bit_inc}"
            if ref is not None and shape == 0:
                assert np.array_equal(got, call_tr(ref, FWD_FUNCS[key][0], r, N, bit_inc))


@pytest.mark.parametrize("key", list(INV_FUNCS), ids=lambda k: f"N{k[0]}dst{k[1]}")
def test_inverse_transform_leaf(key):
    N, dst = key
    rng = np.random.default_rng(N + dst)
    for trial in range(12):
        c = np.zeros((N, N), np.int16)
        k = [N, max(1, N // 2), max(1, N // 4), 1][trial % 4]
        c[:k, :k] = rng.integers(-3000, 3001, (k, k)) if trial < 8 else rng.integers(-32768, 32768, (k, k))
        for bit_inc in (0, 2):
            want = np.zeros((N, N), np.int16)
            port.port_inv_transform(vp(c), u(N), vp(want), u(N), u(N), u(bit_inc), C.c_uint8(dst))
            got = S.encode_inv_transform(c, N, bit_inc, dst)
            assert np.array_equal(got, want)
            if ref is not None:
                assert np.array_equal(got, call_tr(ref, INV_FUNCS[key][0], c, N, bit_inc))


def test_qiq_leaf():
    rng = np.random.default_rng(2)
    QF, FF = [26214, 23302, 20560, 18396, 16384, 14564], [40, 45, 51, 57, 64, 72]
    port.port_qiq.restype = None
    for N in (4, 8, 16, 32):
        for qp in (0, 7, 22, 32, 37, 51):
            for intra in (0, 1):
                c = rng.integers(-32768, 32768, (N, N)).astype(np.int16)
                c[rng.random((N, N)) < 0.5] //= 64
                rem, per = qp % 6, qp // 6
                ts = 15 - 8 - int(np.log2(N))
                sq = 14 + per + ts
                qoff = (171 if intra else 85) << (sq - 9)
                sff = FF[rem] << (per - 2) if per > 8 else FF[rem] << per
                sn = 20 - 14 - ts - 2 if per > 8 else 20 - 14 - ts
                wq, wr, wn = np.zeros_like(c), np.zeros_like(c), C.c_uint32(0)
                port.port_qiq(vp(c), u(N), vp(wq), vp(wr), u(QF[rem]), u(qoff), C.c_int32(sq), C.c_int32(sff), C.c_int32(1 << (sn - 1)), C.c_int32(sn), u(N), C.byref(wn))
                gq, gr, gn = S.quantize_inv_quantize(c, QF[rem], qoff, sq, sff, 1 << (sn - 1), sn, N)
                assert np.array_equal(gq, wq) and np.array_equal(gr, wr) and gn == wn.value


@pytest.mark.parametrize("case", TQ_CASES, ids=lambda c: c[0])
def test_tq_batch_cuda_vs_oracle(case):
    name, w, h, kind, shapes, flags = case
    src, pred = make_src_pred(w, h, 31, kind)
    jobs = make_tq_jobs(w, h, seed=8, shapes=shapes, flags=flags)
    want = run_tq_oracle(port, "port_tq_batch", src, pred, jobs, w, h)
    y, cb, cr = (a.copy() for a in pred.interior())
    recon = S.HostYuvPicture(y, cb, cr)
    coeff = S.HostCoeffPicture(w, h)
    nz = S.tq_batch(src, pred, recon, coeff, jobs)
    assert_tq_equal((recon, coeff, nz), want, name + " CUDA vs port")
    if ref is not None:
        wr = run_tq_oracle(ref, "svtref_tq_batch", src, pred, jobs, w, h, asm_mode=1 if common.ref_asm_simd_ok() else 0)
        assert_tq_equal((recon, coeff, nz), wr, name + " CUDA vs reference chain")


def test_tq_batch_in_place_and_4k_full_frame():
    """BASELINE configs[1] geometry: every TU of a 3840x2160 frame, reconstruction in place over the prediction."""
    w, h = 3840, 2160
    src, pred = make_src_pred(w, h, 77, "synth")
    jobs = make_tq_jobs(w, h, seed=4, shapes=(0, 0, 0, 1, 2), flags=(0, 1))
    if ref is not None:
        want = run_tq_oracle(ref, "svtref_tq_batch", src, pred, jobs, w, h, asm_mode=1 if common.ref_asm_simd_ok() else 0)
    else:
        want = run_tq_oracle(port, "port_tq_batch", src, pred, jobs, w, h)
    coeff = S.HostCoeffPicture(w, h)
    nz = S.tq_batch(src, pred, pred, coeff, jobs)      # in place
    assert_tq_equal((pred, coeff, nz), want, "4K in-place CUDA vs oracle")
    assert nz.sum() > 0


@pytest.mark.parametrize("case", TQ_CASES, ids=lambda c: c[0])
def test_tq_batch16_cuda_vs_oracle(case):
    """16-bit sample planes (10-bit video): EncodeLoop16bit / EncodeGenerateRecon16bit"""
    from test_oracle_tq import make_src_pred16, run_tq_oracle16
    name, w, h, kind, shapes, flags = case
    src, pred = make_src_pred16(w, h, 31, kind)
    jobs = make_tq_jobs(w, h, seed=8, shapes=shapes, flags=flags)
    want = run_tq_oracle16(port, "port_tq_batch16", src, pred, jobs, w, h)
    y, cb, cr = (a.copy() for a in pred.interior())
    recon = S.HostYuvPicture(y, cb, cr, dtype=np.uint16)
    coeff = S.HostCoeffPicture(w, h)
    nz = S.tq_batch(src, pred, recon, coeff, jobs, hbd=True)
    assert_tq_equal((recon, coeff, nz), want, name + " 16-bit CUDA vs port")
    if ref is not None:
        wr = run_tq_oracle16(ref, "svtref_tq_batch16", src, pred, jobs, w, h, asm_mode=1 if common.ref_asm_simd_ok() else 0)
        assert_tq_equal((recon, coeff, nz), wr, name + " 16-bit CUDA vs reference chain")
    # in place (recon == pred), as the encode pass runs it
    nz2 = S.tq_batch(src, pred, pred, S.HostCoeffPicture(w, h), jobs, hbd=True)
    assert np.array_equal(nz2, nz)
    for a, b in zip(pred.interior(), recon.interior()):
        assert np.array_equal(a, b)


def test_pack_coeffs_stream_matches_planes():
    """the dense coefficient stream holds exactly the TUs with nz != 0, in job order; offsets are an exclusive scan"""
    import torch
    w, h = 256, 192
    jobs = make_tq_jobs(w, h, seed=3, shapes=(0,), flags=(0,))
    rng = np.random.default_rng(4)
    planes = [rng.integers(-300, 300, (h, w)).astype(np.int16), rng.integers(-300, 300, (h // 2, w // 2)).astype(np.int16),
              rng.integers(-300, 300, (h // 2, w // 2)).astype(np.int16)]
    nz = (rng.random(len(jobs)) < 0.35).astype(np.uint32) * rng.integers(1, 50, len(jobs)).astype(np.uint32)
    dev = torch.device("cuda", 0)
    d_planes = [torch.from_numpy(p).to(dev) for p in planes]
    coeff = S.CoeffPicture()
    for k in range(3):
        coeff.buf[k] = d_planes[k].data_ptr(); coeff.stride[k] = planes[k].shape[1]
    d_jobs = torch.from_numpy(jobs.view(np.uint8)).to(dev)
    d_nz = torch.from_numpy(nz.view(np.int32)).to(dev)
    cap = int(sum(int(j["size"]) ** 2 for j in jobs))
    d_stream = torch.full((cap,), -1, dtype=torch.int16, device=dev)
    d_off = torch.zeros(len(jobs) + 1, dtype=torch.int32, device=dev)
    S.tq_pack_coeffs_dev(coeff, d_jobs.data_ptr(), d_nz.data_ptr(), len(jobs), d_stream.data_ptr(), cap, d_off.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    off = d_off.cpu().numpy().view(np.uint32)
    stream = d_stream.cpu().numpy()
    sizes = np.array([int(j["size"]) ** 2 if n else 0 for j, n in zip(jobs, nz)], np.uint32)
    assert np.array_equal(off[:-1], np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint32)) and off[-1] == sizes.sum()
    for j, n in zip(range(len(jobs)), nz):
        if not n:
            continue
        t = jobs[j]
        N = int(t["size"])
        tu = planes[int(t["plane"])][int(t["y"]):int(t["y"]) + N, int(t["x"]):int(t["x"]) + N]
        assert np.array_equal(stream[off[j]:off[j] + N * N].reshape(N, N), tu), j
    assert np.all(stream[off[-1]:] == -1)
    # a stream that is too small: the total still tells the caller, nothing is written past the capacity
    d_small = torch.full((1000,), -1, dtype=torch.int16, device=dev)
    S.tq_pack_coeffs_dev(coeff, d_jobs.data_ptr(), d_nz.data_ptr(), len(jobs), d_small.data_ptr(), 1000, d_off.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(d_off.cpu().numpy().view(np.uint32)[-1]) == int(sizes.sum())


@pytest.mark.gpu
def test_tq_pack_coeffs_many_jobs():
    """more jobs than one pass of the scan kernel holds (PK_TILE = 32768), several jobs per scan thread, mixed sizes"""
    import torch
    w, h = 2560, 1536
    rng = np.random.default_rng(11)
    tiles = []
    for ty in range(0, h, 32):
        for tx in range(0, w, 32):
            n = int(rng.choice([8, 8, 8, 8, 16, 32]))
            tiles += [(tx + x, ty + y, n) for y in range(0, 32, n) for x in range(0, 32, n)]
    jobs = np.zeros(len(tiles), dtype=S.TQ_JOB_DTYPE)
    arr = np.array(tiles)
    jobs["x"], jobs["y"], jobs["size"], jobs["plane"] = arr[:, 0], arr[:, 1], arr[:, 2], 0
    assert len(jobs) > 32768
    plane = rng.integers(-300, 300, (h, w)).astype(np.int16)
    nz = (rng.random(len(jobs)) < 0.4).astype(np.uint32)
    dev = torch.device("cuda", 0)
    d_plane = torch.from_numpy(plane).to(dev)
    coeff = S.CoeffPicture()
    for k in range(3):
        coeff.buf[k] = d_plane.data_ptr(); coeff.stride[k] = w
    d_jobs = torch.from_numpy(jobs.view(np.uint8)).to(dev)
    d_nz = torch.from_numpy(nz.view(np.int32)).to(dev)
    sizes = np.where(nz != 0, jobs["size"].astype(np.uint32) ** 2, 0).astype(np.uint32)
    cap = int(sizes.sum()) + 64
    d_stream = torch.full((cap,), -1, dtype=torch.int16, device=dev)
    d_off = torch.zeros(len(jobs) + 1, dtype=torch.int32, device=dev)
    S.tq_pack_coeffs_dev(coeff, d_jobs.data_ptr(), d_nz.data_ptr(), len(jobs), d_stream.data_ptr(), cap, d_off.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    off = d_off.cpu().numpy().view(np.uint32)
    stream = d_stream.cpu().numpy()
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(sizes, dtype=np.uint64)]).astype(np.uint32))
    for j in rng.choice(np.nonzero(nz)[0], 400, replace=False):
        N, x, y = int(jobs["size"][j]), int(jobs["x"][j]), int(jobs["y"][j])
        assert np.array_equal(stream[off[j]:off[j] + N * N].reshape(N, N), plane[y:y + N, x:x + N]), j
    assert np.all(stream[off[-1]:] == -1)
