"""The preset -> parameter derivation the host side restates (svtb200.me_config / ois_config / picture_signals) equals, field by field, what the
reference's own per-picture signal derivation produces (oracle/refinc_meproc_harness.c: SignalDerivationPreAnalysisOq, SignalDerivationMultiProcessesOq,
SignalDerivationMeKernelOq -> SetMeHmeParamsOq, the lambda tables and the initial mvdBits of the MD rate-estimation tables), for every picture class
the BASELINE.json configs contain."""
import ctypes as C
import itertools
import os

import pytest

import common
import svtb200 as S

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libsvtref_meproc_avx2.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")


class Derived(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("input_resolution", "enable_hme", "enable_hme_l0", "enable_hme_l1", "enable_hme_l2", "use_subpel", "limit_ois_to_dc",
                                          "cu8x8_mode", "cu16x16_mode", "skip_ois_8x8", "n_hme_regions_w", "n_hme_regions_h", "hme_l0_total_w", "hme_l0_total_h")] + \
               [(n, C.c_int32 * 2) for n in ("hme_l0_w", "hme_l0_h", "hme_l1_w", "hme_l1_h", "hme_l2_w", "hme_l2_h")] + \
               [(n, C.c_int32) for n in ("search_area_w", "search_area_h", "one_quadrant_hme", "update_hme_search_center", "fractional_search_method",
                                          "fractional_search_model", "fractional_search_64x64", "ois_kernel_level", "ois_th_set")] + \
               [("lambda_", C.c_uint32), ("mvd_bits", C.c_uint32 * 12)]


SLICE = {"B": 0, "P": 1, "I": 2}
GRID = list(itertools.product(((3840, 2160), (1920, 1080), (7680, 4320)), (7, 10), (0, 1, 2, 3), (30, 50, 60), ("B", "P")))


def derive(w, h, enc_mode, tl, fps, slice_type, levels=3, qp=32):
    lib = C.CDLL(REF)
    d = Derived()
    assert lib.svtref_derive_me(w, h, enc_mode, tl, levels, int(tl < levels), SLICE[slice_type], fps, qp, C.byref(d)) == 0
    return d


@pytest.mark.parametrize("size,enc_mode,tl,fps,slice_type", GRID)
def test_me_config_matches_reference_derivation(size, enc_mode, tl, fps, slice_type):
    w, h = size
    d = derive(w, h, enc_mode, tl, fps, slice_type)
    c = S.me_config(w, h, enc_mode=enc_mode, temporal_layer=tl, num_lists=2, frame_rate=fps, slice_type=slice_type)
    for f in ("enable_hme", "enable_hme_l0", "enable_hme_l1", "enable_hme_l2", "n_hme_regions_w", "n_hme_regions_h", "hme_l0_total_w", "hme_l0_total_h",
              "search_area_w", "search_area_h", "one_quadrant_hme", "update_hme_search_center", "fractional_search_method", "fractional_search_model",
              "cu8x8_mode", "cu16x16_mode"):
        assert getattr(c, f) == getattr(d, f), f
    assert d.fractional_search_64x64 == 0          # (the product rejects fractionalSearch64x64; enc modes >= 3 never set it)
    for f in ("hme_l0_w", "hme_l0_h"):
        assert list(getattr(c, f)) == list(getattr(d, f)), f
    for f, en in (("hme_l1_w", d.enable_hme_l1), ("hme_l1_h", d.enable_hme_l1), ("hme_l2_w", d.enable_hme_l2), ("hme_l2_h", d.enable_hme_l2)):
        if en:                                       # (the tables of a disabled level are never read)
            assert list(getattr(c, f)) == list(getattr(d, f)), f
    assert c.lambda_ == d.lambda_
    assert list(c.mvd_bits) == list(d.mvd_bits)
    sig = S.picture_signals(w, h, enc_mode, tl, slice_type=slice_type)
    assert sig["use_subpel"] == d.use_subpel and (c.fractional_search_model != 2) == bool(d.use_subpel)


@pytest.mark.parametrize("size,enc_mode,tl,slice_type", list(itertools.product(((3840, 2160), (1920, 1080), (7680, 4320)), (7, 10), (0, 1, 3), ("B", "I"))))
def test_ois_config_matches_reference_derivation(size, enc_mode, tl, slice_type):
    w, h = size
    d = derive(w, h, enc_mode, tl, 60, slice_type)
    c = S.ois_config(w, h, enc_mode=enc_mode, temporal_layer=tl, slice_type=slice_type)
    assert (c.ois_kernel_level, c.ois_th_set, c.limit_ois_to_dc_mode, c.skip_ois_8x8, c.cu8x8_mode) == \
           (d.ois_kernel_level, d.ois_th_set, d.limit_ois_to_dc, d.skip_ois_8x8, d.cu8x8_mode)
    assert c.slice_type == SLICE[slice_type]


def test_resolution_class_thresholds():
    assert S.resolution_class(3840, 2160) == "4k" and S.resolution_class(7680, 4320) == "4k" and S.resolution_class(1920, 1080) == "1080p"
    assert derive(3840, 2160, 10, 0, 60, "B").input_resolution == derive(7680, 4320, 10, 0, 60, "B").input_resolution
