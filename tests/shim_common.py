"""Helpers of the encoder-in-the-loop tests: the reference's own CLI built WITH the B200 shim (shim/build_shim.sh -> oracle/_ref/SvtHevcEncApp_b200)."""
import hashlib
import os
import subprocess

import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "oracle", "_ref", "SvtHevcEncApp_b200")
LIB = os.path.join(ROOT, "svt-hevc_b200", "libsvthevc_b200.so")


def write_clip(path, w, h, n, seed=21):
    v = common.SynthVideo(w, h, seed=seed)
    with open(path, "wb") as f:
        for i in range(n):
            cb, cr = v.chroma(i)
            f.write(v.luma(i).tobytes()); f.write(cb.tobytes()); f.write(cr.tobytes())


def encode(tmp, clip, w, h, n, b200, enc_mode=10, tag="", extra=()):
    """-> (return code, sha256 of the bitstream, sha256 of the reconstruction, stderr)"""
    out, rec = os.path.join(tmp, f"out{tag}.265"), os.path.join(tmp, f"rec{tag}.yuv")
    env = dict(os.environ)
    env.pop("SVT_HEVC_B200", None)
    if b200 is not None:
        env["SVT_HEVC_B200"] = b200
    env["SVT_HEVC_B200_LIB"] = LIB
    cmd = [APP, "-i", clip, "-w", str(w), "-h", str(h), "-n", str(n), "-encMode", str(enc_mode), "-q", "32", "-fps", "60", "-intra-period", "15",
           "-irefresh-type", "1", "-b", out, "-o", rec, *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    sha = lambda f: hashlib.sha256(open(f, "rb").read()).hexdigest() if os.path.exists(f) and os.path.getsize(f) else None
    return p.returncode, sha(out), sha(rec), p.stderr + p.stdout
