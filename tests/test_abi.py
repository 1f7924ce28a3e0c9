"""CPU tests of the C-ABI boundary: libsvthevc_b200.so loads without a GPU, exports every function include/svt_hevc_b200.h
declares (and nothing is left undeclared), and -- on a box without a CUDA device -- refuses to compute instead of falling
back to a CPU path.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import svtb200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "svt_hevc_b200.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(svtb200_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = S.lib()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/svt_hevc_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", S.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (svtb200_[A-Za-z0-9_]+)", out)))
    assert exported == names, f"exported but undeclared: {set(exported) - set(names)}; declared but missing: {set(names) - set(exported)}"


def test_struct_layouts_match_the_header():
    assert C.sizeof(S.Plane) == 24 and C.sizeof(S.PaPicture) == 72 and C.sizeof(S.YuvPicture) == 72
    assert C.sizeof(S.McJob) == 18 and S.TQ_JOB_DTYPE.itemsize == 16 and S.ME_PU_RESULT_DTYPE.itemsize == 24
    assert S.SAO_STATS_DTYPE.itemsize == 312 and S.SAO_PARAMS_DTYPE.itemsize == 20 and C.sizeof(S.DlfParams) == 10
    # svtb200_me_config: spot-check a few offsets against the C layout through a tiny compiled probe
    assert S.INTRA_JOB_DTYPE.itemsize == 12 and C.sizeof(S.OisConfig) == 12
    src = '#include <stddef.h>\n#include <stdio.h>\n#include "svt_hevc_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu",sizeof(svtb200_me_config),offsetof(svtb200_me_config,lambda),offsetof(svtb200_me_config,mvd_bits),offsetof(svtb200_me_config,ref_poc),sizeof(svtb200_intra_job),sizeof(svtb200_ois_config),offsetof(svtb200_intra_job,left_avail));return 0;}'
    exe = os.path.join(ROOT, "tests", ".abi_probe")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    os.remove(exe)
    assert got == [C.sizeof(S.MeConfig), S.MeConfig.lambda_.offset, S.MeConfig.mvd_bits.offset, S.MeConfig.ref_poc.offset, 12, 12, 8]


def test_no_cpu_fallback_without_a_device():
    if S.device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    a = np.zeros((8, 8), np.uint8)
    with pytest.raises(S.Svtb200Error, match="no CUDA device"):
        S.nxm_sad(a, 8, a, 8, 8, 8)
    assert S.lib().svtb200_init(C.c_int(0)) == -1     # SVTB200_ERR_NO_DEVICE
