"""GPU tests of the picture hand-off helpers: host -> device upload of Y/Cb/Cr into a padded device picture (linear staging + row placement
for contiguous host planes, pitched copies otherwise) -- byte-exact, interior only, padding untouched."""
import numpy as np
import pytest

import svtb200 as S

pytestmark = pytest.mark.gpu


def _upload(w, h, pad, bps, contiguous, repeat=1):
    import torch
    dev = torch.device("cuda", 0)
    dt_np, dt_t = (np.uint8, torch.uint8) if bps == 1 else (np.uint16, torch.int16)
    rng = np.random.default_rng(w * 131 + h + bps)
    planes_t, want = [], []
    for (pw, ph, pp) in ((w, h, pad), (w // 2, h // 2, pad // 2), (w // 2, h // 2, pad // 2)):
        planes_t.append(torch.full((ph + 2 * pp, pw + 2 * pp), 7, dtype=dt_t, device=dev))
    st = S.make_yuv_struct(tuple(t.data_ptr() for t in planes_t), tuple(t.shape[1] for t in planes_t), pad, w, h)
    for _ in range(repeat):
        hosts, strides = [], []
        for (pw, ph) in ((w, h), (w // 2, h // 2), (w // 2, h // 2)):
            hs = pw if contiguous else pw + 24
            a = rng.integers(0, 256 if bps == 1 else 1024, (ph, hs)).astype(dt_np)
            hosts.append(torch.from_numpy(a).pin_memory()); strides.append(hs)
        S.yuv_picture_upload(st, hosts[0].data_ptr(), strides[0], hosts[1].data_ptr(), strides[1], hosts[2].data_ptr(), strides[2], bps,
                             torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = [hh.numpy()[:, :pw] for hh, pw in zip(hosts, (w, w // 2, w // 2))]
    for t, wnt, pp in zip(planes_t, want, (pad, pad // 2, pad // 2)):
        got = t.cpu().numpy().view(dt_np)
        ph, pw = wnt.shape
        assert np.array_equal(got[pp:pp + ph, pp:pp + pw], wnt)
        border = got.copy(); border[pp:pp + ph, pp:pp + pw] = 7
        assert (border == 7).all(), "padding must stay untouched"


@pytest.mark.parametrize("w,h", [(3840, 2160), (448, 256), (200, 72), (136, 40)])
@pytest.mark.parametrize("contiguous", [True, False])
def test_yuv_picture_upload_8bit(w, h, contiguous):
    _upload(w, h, 80, 1, contiguous, repeat=2)


@pytest.mark.parametrize("contiguous", [True, False])
def test_yuv_picture_upload_16bit(contiguous):
    _upload(448, 256, 80, 2, contiguous)


@pytest.mark.parametrize("w,h,pad", [(3840, 2160, 80), (448, 256, 80), (200, 72, 16), (136, 40, 8)])
def test_yuv_picture_pad_dev_replicates_edges(w, h, pad):
    """GeneratePadding on a device picture == numpy edge padding of the interior"""
    import ctypes as C
    import torch
    rng = np.random.default_rng(w + h)
    planes, want = [], []
    for (pw, ph, pp) in ((w, h, pad), (w // 2, h // 2, pad // 2), (w // 2, h // 2, pad // 2)):
        a = rng.integers(0, 256, (ph, pw), dtype=np.uint8)
        t = torch.full((ph + 2 * pp, pw + 2 * pp + (-(pw + 2 * pp)) % 4), 7, dtype=torch.uint8, device="cuda")
        t[pp:pp + ph, pp:pp + pw] = torch.from_numpy(a).cuda()
        planes.append(t); want.append(np.pad(a, pp, mode="edge"))
    st = S.make_yuv_struct(tuple(t.data_ptr() for t in planes), tuple(t.shape[1] for t in planes), pad, w, h)
    S.check(S.lib().svtb200_yuv_picture_pad_dev(C.byref(st), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    for t, wnt in zip(planes, want):
        assert np.array_equal(t.cpu().numpy()[:, :wnt.shape[1]], wnt)
