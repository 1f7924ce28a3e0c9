"""Encoder in the loop: the reference's own CLI with the B200 shim encodes the same clip twice -- motion estimation by the reference's MotionEstimateLcu
(SVT_HEVC_B200 unset) and by svtb200_me_picture on the GPU (SVT_HEVC_B200=1) -- and the bitstream and the reconstruction must be IDENTICAL byte for byte
(CQP, closed GOPs: the reference's run-to-run policy).  Every later stage (mode decision, EncDec, entropy coding) is the reference's own code, so one
differing motion vector or distortion anywhere in any picture would change the stream."""
import os

import pytest

import shim_common as SC

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(SC.APP), reason="oracle/_ref/SvtHevcEncApp_b200 not built")]


@pytest.mark.parametrize("w,h,n,enc_mode", [(1920, 1080, 16, 10), (1920, 1080, 9, 7), (3840, 2160, 9, 10)], ids=["1080p_m10", "1080p_m7", "4k_m10"])
def test_bitstream_identical_with_gpu_motion_estimation(tmp_path, w, h, n, enc_mode):
    clip = str(tmp_path / "in.yuv")
    SC.write_clip(clip, w, h, n)
    ref = SC.encode(str(tmp_path), clip, w, h, n, None, enc_mode, tag="_ref")
    gpu = SC.encode(str(tmp_path), clip, w, h, n, "1", enc_mode, tag="_gpu")
    assert ref[0] == 0, ref[3][-600:]
    assert gpu[0] == 0, gpu[3][-600:]
    assert "on the GPU" in gpu[3], gpu[3][-600:]
    assert ref[1] is not None and ref[1] == gpu[1], "bitstreams differ (GPU motion estimation + open-loop intra search)"
    assert ref[2] is not None and ref[2] == gpu[2], "reconstructions differ"
    # motion estimation alone on the GPU, the intra search on the CPU: isolates the two stages when one of them ever differs
    os.environ["SVT_HEVC_B200_OIS"] = "0"
    try:
        me_only = SC.encode(str(tmp_path), clip, w, h, n, "1", enc_mode, tag="_me")
    finally:
        os.environ.pop("SVT_HEVC_B200_OIS", None)
    assert me_only[0] == 0 and me_only[1] == ref[1] and me_only[2] == ref[2], "bitstreams differ (GPU motion estimation only)"
