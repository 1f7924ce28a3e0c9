"""Reference-side shim, CPU part: the patched encoder (shim/b200.patch + shim/EbB200Shim.c in a scratch copy of the reference) with SVT_HEVC_B200 unset is
the reference's own path (deterministic bitstream), and with SVT_HEVC_B200=1 on a box WITHOUT a usable GPU library EbInitEncoder fails loudly -- there is no
silent CPU fallback.  The encoder-in-the-loop comparison itself needs a GPU: tests/test_gpu_shim.py."""
import os
import subprocess

import pytest

import shim_common as SC

pytestmark = pytest.mark.skipif(not os.path.exists(SC.APP), reason="oracle/_ref/SvtHevcEncApp_b200 not built (shim/build_shim.sh where /root/reference exists)")


def test_patch_touches_no_api_file():
    patch = open(os.path.join(SC.ROOT, "shim", "b200.patch")).read()
    files = [l.split()[1] for l in patch.splitlines() if l.startswith("+++ ")]
    assert files and all("/Source/Lib/Codec/" in f for f in files), files
    assert "Source/API" not in patch


def test_backend_off_is_deterministic_reference_path(tmp_path):
    clip = str(tmp_path / "in.yuv")
    SC.write_clip(clip, 1920, 1080, 6)
    a = SC.encode(str(tmp_path), clip, 1920, 1080, 6, None, tag="a")
    b = SC.encode(str(tmp_path), clip, 1920, 1080, 6, "0", tag="b")
    assert a[0] == 0 and b[0] == 0, a[3][-400:] + b[3][-400:]
    assert a[1] is not None and a[1] == b[1] and a[2] == b[2]
    assert "SVT [B200]" not in a[3]


def test_backend_on_without_gpu_fails_loudly(tmp_path):
    import svtb200 as S
    if S.device_count() > 0:
        pytest.skip("a GPU is visible: covered by tests/test_gpu_shim.py")
    clip = str(tmp_path / "in.yuv")
    SC.write_clip(clip, 1920, 1080, 2)
    rc, bits, rec, log = SC.encode(str(tmp_path), clip, 1920, 1080, 2, "1")
    assert "SVT [B200]" in log and ("svtb200_init failed" in log or "cannot load" in log), log[-600:]
    assert bits is None, "the encoder must not produce a stream on a silent CPU fallback"
