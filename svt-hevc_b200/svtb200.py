"""ctypes binding of libsvthevc_b200.so (include/svt_hevc_b200.h) + host-side mirrors of the
reference's per-picture parameter derivation for the hot path.

This module is the product's host side for Python callers (tests, bench.py).  It never imports
anything from oracle/ and raises if the CUDA library is missing: there is no CPU fallback.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVTB200_LIB") or os.path.join(_HERE, "libsvthevc_b200.so")   # the override is for A/B builds of the same ABI

ME_PU_COUNT = 85


class Plane(C.Structure):
    """svtb200_plane == luma view of EbPictureBufferDesc_t (EbPictureBufferDesc.h:26-70)."""
    _fields_ = [("buf", C.c_void_p), ("stride", C.c_uint32), ("origin_x", C.c_uint16),
                ("origin_y", C.c_uint16), ("width", C.c_uint16), ("height", C.c_uint16)]


class PaPicture(C.Structure):
    _fields_ = [("full", Plane), ("quarter", Plane), ("sixteenth", Plane)]


class MeConfig(C.Structure):
    _fields_ = [("pic_width", C.c_uint16), ("pic_height", C.c_uint16),
                ("num_lists", C.c_uint8), ("temporal_layer", C.c_uint8), ("hierarchical_levels", C.c_uint8),
                ("enable_hme", C.c_uint8), ("enable_hme_l0", C.c_uint8), ("enable_hme_l1", C.c_uint8),
                ("enable_hme_l2", C.c_uint8), ("one_quadrant_hme", C.c_uint8),
                ("update_hme_search_center", C.c_uint8),
                ("n_hme_regions_w", C.c_uint8), ("n_hme_regions_h", C.c_uint8),
                ("hme_l0_total_w", C.c_uint16), ("hme_l0_total_h", C.c_uint16),
                ("hme_l0_w", C.c_uint16 * 2), ("hme_l0_h", C.c_uint16 * 2),
                ("hme_l1_w", C.c_uint16 * 2), ("hme_l1_h", C.c_uint16 * 2),
                ("hme_l2_w", C.c_uint16 * 2), ("hme_l2_h", C.c_uint16 * 2),
                ("search_area_w", C.c_uint8), ("search_area_h", C.c_uint8),
                ("fractional_search_model", C.c_uint8),
                ("cu8x8_mode", C.c_uint8), ("cu16x16_mode", C.c_uint8), ("rate_control_mode", C.c_uint8),
                ("fractional_search_method", C.c_uint8),
                ("lambda_", C.c_uint32), ("mvd_bits", C.c_uint32 * 12), ("ref_poc", C.c_uint64 * 2)]


ME_PU_RESULT_DTYPE = np.dtype([("x_mv_l0", "<i2"), ("y_mv_l0", "<i2"), ("x_mv_l1", "<i2"), ("y_mv_l1", "<i2"),
                               ("distortion", "<u4", 3), ("direction", "u1", 3), ("total_me_candidates", "u1")])
assert ME_PU_RESULT_DTYPE.itemsize == 24

_lib = None


def lib():
    """Load the CUDA library; fail loudly if it is absent (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the B200 path has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.svtb200_version.restype = C.c_char_p
        _lib.svtb200_last_error.restype = C.c_char_p
        _lib.svtb200_launch_count.restype = C.c_uint64
    return _lib


class Svtb200Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise Svtb200Error(f"svtb200 error {rc}: {lib().svtb200_last_error().decode()}")


def _u8p(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------------
# picture helpers (host side of PictureAnalysis: pad + point-decimate, EbPictureAnalysisProcess.c:173,4139)
# ------------------------------------------------------------------------------------------------
class HostPaPicture:
    """Padded full / quarter / sixteenth luma planes in host memory with the reference's geometry:
    full-resolution pad 68 (EbEncHandle.c:901-903), 1/4 pad 32 (:912-914), 1/16 pad 16 (:923-925)."""

    def __init__(self, luma, pads=(68, 32, 16)):
        luma = np.ascontiguousarray(luma, dtype=np.uint8)
        self.height, self.width = luma.shape
        self.planes = []
        for step, pad in zip((1, 2, 4), pads):
            dec = luma[::step, ::step]
            self.planes.append((np.ascontiguousarray(np.pad(dec, pad, mode="edge")), pad, dec.shape[1], dec.shape[0]))
        self.struct = PaPicture()
        for name, (arr, pad, w, h) in zip(("full", "quarter", "sixteenth"), self.planes):
            p = getattr(self.struct, name)
            p.buf = arr.ctypes.data
            p.stride = arr.shape[1]
            p.origin_x = pad
            p.origin_y = pad
            p.width = w
            p.height = h


# ------------------------------------------------------------------------------------------------
# preset -> ME / OIS parameters: restatement of the reference's per-picture signal derivation for the presets and resolution classes
# BASELINE.json names (M7 and M10; 1080p and 4K-and-above), PINNED field by field against the reference's own functions by
# tests/test_oracle_preset.py (oracle/refinc_meproc_harness.c):
#   SignalDerivationPreAnalysisOq     EbResourceCoordinationProcess.c:239  (HME enables, tables EbDefinitions.h:1666-1806)
#   SignalDerivationMultiProcessesOq  EbPictureDecisionProcess.c:376       (useSubpelFlag :334, limitOisToDcModeFlag, cu8x8Mode, skipOis8x8)
#   SignalDerivationMeKernelOq        EbMotionEstimationProcess.c:308      (SetMeHmeParamsOq :53, oneQuadrantHME, fractionalSearchModel, ...)
#   lambda / mvdBits                  EbMotionEstimationProcess.c:684-704
# In the encoder the shim reads these values from the encoder's own contexts; this restatement serves bench.py and the tests.
# ------------------------------------------------------------------------------------------------
_ME_TABLE = {
    # (res_class, enc_mode): (l0_total_w, l0_total_h, l0_w, l0_h, en_l1, l1_w, l1_h, en_l2, l2_w, l2_h, sa_w, sa_h)
    ("4k", 10): (64, 32, 32, 16, 0, 0, 0, 0, 0, 0, 8, 7),
    ("4k", 7): (64, 32, 32, 16, 1, 4, 4, 0, 0, 0, 16, 9),
    ("1080p", 10): (48, 40, 24, 20, 1, 4, 4, 0, 0, 0, 8, 7),
    ("1080p", 7): (64, 48, 32, 24, 1, 4, 4, 1, 4, 2, 16, 9),
}
# initial MdRateEstimationContext_t::mvdBits (1 bit = 32768) at picture QP 32, the CQP setting of every BASELINE.json config
# (EncodeCabacContextModelInit + MdRateEstimationContextInit, EbEncodeContext.c:138-143), by slice type; other QPs: pass mvd_bits
MVD_BITS_QP32 = {
    "B": (37716, 28388, 36028, 29856, 39963, 26612, 29856, 36028, 28388, 37716, 31523, 34297),
    "P": (47445, 21645, 42538, 24752, 49833, 20310, 29856, 36028, 28388, 37716, 31523, 34297),
}
DEFAULT_MVD_BITS = MVD_BITS_QP32["B"]
# lambdaModeDecisionRaSad{Base, RefNonBase, NonRef}[32] (EbLambdaRateTables.h:249,280,311); other QPs: pass lambda_
DEFAULT_LAMBDA = {"base": 439162, "ref": 717148, "nonref": 994505}


def resolution_class(width, height):
    """DeriveInputResolution (EbSequenceControlSet.c:288) + the resolutionIndex of the preset tables: '4k' (INPUT_SIZE_4K_RANGE, also 8K) or '1080p'."""
    n = width * height
    if n >= 0x29F630:
        return "4k"
    if n >= 0x1AB3F0:
        return "1080p"
    raise Svtb200Error(f"preset tables are restated for 1080p and 4K-and-above pictures only ({width}x{height})")


def picture_signals(width, height, enc_mode, temporal_layer, hierarchical_levels=3, is_reference=None, slice_type="B"):
    """useSubpelFlag, limitOisToDcModeFlag, cu8x8Mode, skipOis8x8 of one picture (SignalDerivationMultiProcessesOq) for enc modes 7..10."""
    if not 7 <= enc_mode <= 10:
        raise Svtb200Error("picture signals are restated for enc modes 7..10")
    res = resolution_class(width, height)
    if is_reference is None:
        is_reference = temporal_layer < hierarchical_levels
    if enc_mode <= 8:
        use_subpel = 1
    elif enc_mode == 9:
        use_subpel = int(temporal_layer == 0) if res == "4k" else 1
    else:
        use_subpel = int(temporal_layer == 0) if res == "4k" else int(bool(is_reference))
    if enc_mode == 7:
        cu8x8_mode = 1 if res == "4k" else (0 if temporal_layer == 0 else 1)
    else:
        cu8x8_mode = 1
    intra = slice_type == "I"
    return dict(use_subpel=use_subpel, limit_ois_to_dc_mode=int(enc_mode >= 10 and not intra), cu8x8_mode=cu8x8_mode, cu16x16_mode=0,
                skip_ois_8x8=0 if intra else 1, is_reference=int(bool(is_reference)))


def me_config(width, height, enc_mode=10, temporal_layer=0, num_lists=1, hierarchical_levels=3,
              frame_rate=60, lambda_=None, mvd_bits=None, ref_poc=(0, 16), rate_control_mode=0, use_subpel=None, is_reference=None, slice_type="B", qp=32):
    res = resolution_class(width, height)
    if (res, enc_mode) not in _ME_TABLE:
        raise Svtb200Error(f"me_config: enc mode {enc_mode} is not restated (7 and 10 are)")
    t = _ME_TABLE[(res, enc_mode)]
    sig = picture_signals(width, height, enc_mode, temporal_layer, hierarchical_levels, is_reference, slice_type)
    c = MeConfig()
    c.pic_width, c.pic_height = width, height
    c.num_lists, c.temporal_layer, c.hierarchical_levels = num_lists, temporal_layer, hierarchical_levels
    c.enable_hme, c.enable_hme_l0, c.enable_hme_l1, c.enable_hme_l2 = 1, 1, t[4], t[7]
    c.one_quadrant_hme = 1 if (enc_mode >= 8 and res == "4k") else 0
    c.update_hme_search_center = 1       # resolution above 576p and tune 1 (SetMeHmeParamsOq :120)
    c.n_hme_regions_w = c.n_hme_regions_h = 2
    l0tw, l0th, l0w, l0h = t[0], t[1], t[2], t[3]
    if res == "4k" and frame_rate <= 30:   # low frame-rate bump, EbMotionEstimationProcess.c:98-117
        if enc_mode in (6, 7):
            l0tw, l0th, l0w, l0h = max(96, l0tw), max(64, l0th), max(48, l0w), max(32, l0h)
        elif enc_mode >= 8:
            l0tw, l0th, l0w, l0h = max(64, l0tw), max(48, l0th), max(32, l0w), max(24, l0h)
    c.hme_l0_total_w, c.hme_l0_total_h = l0tw, l0th
    for i in range(2):
        c.hme_l0_w[i], c.hme_l0_h[i] = l0w, l0h
        c.hme_l1_w[i], c.hme_l1_h[i] = t[5], t[6]
        c.hme_l2_w[i], c.hme_l2_h[i] = t[8], t[9]
    c.search_area_w, c.search_area_h = t[10], t[11]
    if use_subpel is None:
        use_subpel = sig["use_subpel"]
    # fractionalSearchModel, EbMotionEstimationProcess.c:398-420: 0 all PUs (M<=5, M6 below 4K), 1 selective (SuPelEnable), 2 off
    c.fractional_search_model = 2 if not use_subpel else 1
    c.fractional_search_method = 0       # SUB_SAD_SEARCH for M>=5 (:340-345)
    c.cu8x8_mode, c.cu16x16_mode = sig["cu8x8_mode"], 0
    c.rate_control_mode = rate_control_mode
    if lambda_ is None:
        if qp != 32:
            raise Svtb200Error("me_config: lambda_ is tabulated for QP 32 only; pass lambda_ (lambdaModeDecisionRaSad*[qp])")
        lambda_ = DEFAULT_LAMBDA["base" if temporal_layer == 0 else ("ref" if sig["is_reference"] else "nonref")]
    c.lambda_ = lambda_
    if mvd_bits is None:
        if qp != 32:
            raise Svtb200Error("me_config: mvd_bits are tabulated for QP 32 only; pass mvd_bits (mdRateEstimationArray->mvdBits)")
        mvd_bits = MVD_BITS_QP32["P" if slice_type == "P" else "B"]
    for i, v in enumerate(mvd_bits):
        c.mvd_bits[i] = v
    c.ref_poc[0], c.ref_poc[1] = ref_poc
    return c


def ois_config(width, height, enc_mode=10, temporal_layer=0, hierarchical_levels=3, is_reference=None, slice_type="B"):
    """OisConfig of one picture: oisKernelLevel / oisThSet (SignalDerivationMeKernelOq :352-392), limitOisToDcModeFlag, skipOis8x8, cu8x8Mode."""
    sig = picture_signals(width, height, enc_mode, temporal_layer, hierarchical_levels, is_reference, slice_type)
    c = OisConfig()
    c.pic_width, c.pic_height = width, height
    c.slice_type = {"B": 0, "P": 1, "I": 2}[slice_type]
    c.temporal_layer = temporal_layer
    c.ois_kernel_level, c.ois_th_set, c.set_best_ois_distortion_to_valid = 0, 1, 0      # enc modes >= 7 (:352-392)
    c.limit_ois_to_dc_mode, c.skip_ois_8x8, c.cu8x8_mode = sig["limit_ois_to_dc_mode"], sig["skip_ois_8x8"], sig["cu8x8_mode"]
    return c


# ------------------------------------------------------------------------------------------------
# entry points
# ------------------------------------------------------------------------------------------------
def device_count():
    return lib().svtb200_device_count()


def me_picture(cfg, src, ref0, ref1=None, want_raw=False, want_rc=False):
    """Host-buffer picture ME (copies in, one kernel, copies out). src/ref*: HostPaPicture."""
    n_lcu = ((cfg.pic_width + 63) // 64) * ((cfg.pic_height + 63) // 64)
    res = np.zeros((n_lcu, ME_PU_COUNT), dtype=ME_PU_RESULT_DTYPE)
    bs = np.zeros((n_lcu, 2, 85), dtype=np.uint32) if want_raw else None
    bm = np.zeros((n_lcu, 2, 85), dtype=np.uint32) if want_raw else None
    rc = np.zeros(n_lcu, dtype=np.uint32) if want_rc else None
    check(lib().svtb200_me_picture(C.byref(cfg), C.byref(src.struct), C.byref(ref0.struct),
                                   C.byref(ref1.struct) if ref1 is not None else None,
                                   _u8p(res), _u8p(bs) if want_raw else None, _u8p(bm) if want_raw else None,
                                   _u8p(rc) if want_rc else None))
    return res, bs, bm, rc


def sad_loop_kernel(src, src_stride, ref, ref_off, ref_stride, height, width, raw_stride, saw, sah, x0=0, y0=0):
    best = C.c_uint64(0)
    xs, ys = C.c_int16(x0), C.c_int16(y0)
    check(lib().svtb200_SadLoopKernel(_u8p(src), C.c_uint32(src_stride), C.c_void_p(ref.ctypes.data + ref_off),
                                      C.c_uint32(ref_stride), C.c_uint32(height), C.c_uint32(width),
                                      C.byref(best), C.byref(xs), C.byref(ys), C.c_uint32(raw_stride),
                                      C.c_int16(saw), C.c_int16(sah)))
    return best.value, xs.value, ys.value


def nxm_sad(src, ss, ref, rs, h, w):
    out = C.c_uint32(0)
    check(lib().svtb200_NxMSadKernel(_u8p(src), C.c_uint32(ss), _u8p(ref), C.c_uint32(rs), C.c_uint32(h), C.c_uint32(w), C.byref(out)))
    return out.value


def combined_averaging_sad(src, ss, r1, r1s, r2, r2s, h, w):
    out = C.c_uint32(0)
    check(lib().svtb200_CombinedAveragingSAD(_u8p(src), C.c_uint32(ss), _u8p(r1), C.c_uint32(r1s), _u8p(r2), C.c_uint32(r2s),
                                             C.c_uint32(h), C.c_uint32(w), C.byref(out)))
    return out.value


def fullpel_search_lcu(src, ss, ref, ref_off, rs, xo, yo, saw, sah, best_sad, best_mv):
    check(lib().svtb200_FullPelSearch_LCU(_u8p(src), C.c_uint32(ss), C.c_void_p(ref.ctypes.data + ref_off), C.c_uint32(rs),
                                          C.c_int16(xo), C.c_int16(yo), C.c_uint32(saw), C.c_uint32(sah),
                                          _u8p(best_sad), _u8p(best_mv)))


class DevPaPicture:
    """Device-resident padded full / quarter / sixteenth luma (svtb200_pa_picture_create / _upload)."""

    def __init__(self, width, height):
        self.width, self.height = width, height
        self.struct = PaPicture()
        check(lib().svtb200_pa_picture_create(C.c_uint16(width), C.c_uint16(height), C.byref(self.struct)))

    def upload(self, luma, stream=0):
        """luma: contiguous (pinned) host uint8 array HxW, or an integer host address with stride == width."""
        ptr = luma if isinstance(luma, int) else luma.ctypes.data
        check(lib().svtb200_pa_picture_upload(C.byref(self.struct), C.c_void_p(ptr), C.c_uint32(self.width), C.c_void_p(stream)))

    def build_from_device(self, dev_ptr, dev_stride, stream=0):
        check(lib().svtb200_pa_picture_build_dev(C.byref(self.struct), C.c_void_p(dev_ptr), C.c_uint32(dev_stride), C.c_void_p(stream)))

    def download(self):
        """-> HostPaPicture-shaped numpy planes (for tests)."""
        host = HostPaPicture(np.zeros((self.height, self.width), np.uint8))
        check(lib().svtb200_pa_picture_download(C.byref(self.struct), C.byref(host.struct)))
        return host

    def destroy(self):
        if self.struct.full.buf:
            lib().svtb200_pa_picture_destroy(C.byref(self.struct))

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def me_picture_dev(cfg, src, ref0, ref1, results_ptr, stream=0, best_sad_ptr=None, best_mv_ptr=None, rc_ptr=None):
    """Device-resident picture ME: src/ref*: DevPaPicture; results_ptr: device address of [n_lcu][85] records."""
    check(lib().svtb200_me_picture_dev(C.byref(cfg), C.byref(src.struct), C.byref(ref0.struct),
                                       C.byref(ref1.struct) if ref1 is not None else None,
                                       C.c_void_p(results_ptr), C.c_void_p(best_sad_ptr), C.c_void_p(best_mv_ptr),
                                       C.c_void_p(rc_ptr), C.c_void_p(stream)))


# ================================================================================================
# B. motion-compensated prediction
# ================================================================================================
class YuvPicture(C.Structure):
    _fields_ = [("y", Plane), ("cb", Plane), ("cr", Plane)]


class McJob(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("w", C.c_uint8), ("h", C.c_uint8), ("dir", C.c_uint8),
                ("ref_idx", C.c_uint8 * 2), ("reserved", C.c_uint8), ("mv", (C.c_int16 * 2) * 2)]


MC_JOB_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("w", "u1"), ("h", "u1"), ("dir", "u1"), ("ref_idx", "u1", 2),
                         ("reserved", "u1"), ("mv", "<i2", (2, 2))])
assert MC_JOB_DTYPE.itemsize == C.sizeof(McJob) == 18


class HostYuvPicture:
    """8-bit 4:2:0 picture with the reference's recon-reference geometry: luma pad 80 = 64 + MCPXPaddingOffset 16
    (EbEncHandle.c:867-870), chroma pad 40; edge-replicated like GeneratePadding (EbMcp.c:1017)."""

    def __init__(self, y, cb, cr, pad=80, replicate=True, dtype=np.uint8):
        """dtype np.uint16: 16-bit sample planes (10-bit video); strides and origins stay in samples"""
        self.height, self.width = y.shape
        self.pad = pad
        mode = "edge" if replicate else "constant"
        self.y = np.ascontiguousarray(np.pad(np.asarray(y, dtype), pad, mode=mode))
        self.cb = np.ascontiguousarray(np.pad(np.asarray(cb, dtype), pad // 2, mode=mode))
        self.cr = np.ascontiguousarray(np.pad(np.asarray(cr, dtype), pad // 2, mode=mode))
        self.struct = YuvPicture()
        for name, arr, p, w, h in (("y", self.y, pad, self.width, self.height), ("cb", self.cb, pad // 2, self.width // 2, self.height // 2),
                                   ("cr", self.cr, pad // 2, self.width // 2, self.height // 2)):
            pl = getattr(self.struct, name)
            pl.buf, pl.stride, pl.origin_x, pl.origin_y, pl.width, pl.height = arr.ctypes.data, arr.shape[1], p, p, w, h

    def interior(self):
        p, q = self.pad, self.pad // 2
        return (self.y[p:p + self.height, p:p + self.width], self.cb[q:q + self.height // 2, q:q + self.width // 2],
                self.cr[q:q + self.height // 2, q:q + self.width // 2])


def yuv_array(pics):
    arr = (YuvPicture * len(pics))()
    for i, p in enumerate(pics):
        arr[i] = p.struct
    return arr


def mc_batch(refs, jobs, dst):
    """refs: list of HostYuvPicture; jobs: numpy array of MC_JOB_DTYPE; dst: HostYuvPicture (written in place)."""
    jobs = np.ascontiguousarray(jobs)
    check(lib().svtb200_mc_batch(yuv_array(refs), C.c_int(len(refs)), _u8p(jobs), C.c_int(len(jobs)), C.byref(dst.struct)))


def mc_ifree_batch(refs, jobs, dst):
    """the mode-decision fast loop's interpolation-free prediction (UniPredIFreeRef8Bit / BiPredIFreeRef8Bit); same arguments as mc_batch"""
    jobs = np.ascontiguousarray(jobs)
    check(lib().svtb200_mc_ifree_batch(yuv_array(refs), C.c_int(len(refs)), _u8p(jobs), C.c_int(len(jobs)), C.byref(dst.struct)))


def mc_batch16(refs, jobs, dst):
    """16-bit sample planes (HostYuvPicture(dtype=np.uint16)), 10-bit video"""
    jobs = np.ascontiguousarray(jobs)
    check(lib().svtb200_mc_batch16(yuv_array(refs), C.c_int(len(refs)), _u8p(jobs), C.c_int(len(jobs)), C.byref(dst.struct)))


def luma_interp(frac, ref, ref_off, rs, w, h):
    dst = np.zeros((h, w), np.uint8)
    check(lib().svtb200_LumaInterpolationFilterNew(C.c_uint32(frac), C.c_void_p(ref.ctypes.data + ref_off), C.c_uint32(rs), _u8p(dst),
                                                   C.c_uint32(w), C.c_uint32(w), C.c_uint32(h)))
    return dst


def chroma_interp(fx, fy, ref, ref_off, rs, w, h):
    dst = np.zeros((h, w), np.uint8)
    check(lib().svtb200_ChromaInterpolationFilterNew(C.c_uint32(fx), C.c_uint32(fy), C.c_void_p(ref.ctypes.data + ref_off), C.c_uint32(rs),
                                                     _u8p(dst), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h)))
    return dst


def luma_bipred(f0, r0, r0_off, r0s, f1, r1, r1_off, r1s, w, h):
    dst = np.zeros((h, w), np.uint8)
    check(lib().svtb200_LumaBiPredInterpolation(C.c_uint32(f0), C.c_void_p(r0.ctypes.data + r0_off), C.c_uint32(r0s), C.c_uint32(f1),
                                                C.c_void_p(r1.ctypes.data + r1_off), C.c_uint32(r1s), _u8p(dst), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h)))
    return dst


def chroma_bipred(f0x, f0y, r0, r0_off, r0s, f1x, f1y, r1, r1_off, r1s, w, h):
    dst = np.zeros((h, w), np.uint8)
    check(lib().svtb200_ChromaBiPredInterpolation(C.c_uint32(f0x), C.c_uint32(f0y), C.c_void_p(r0.ctypes.data + r0_off), C.c_uint32(r0s),
                                                  C.c_uint32(f1x), C.c_uint32(f1y), C.c_void_p(r1.ctypes.data + r1_off), C.c_uint32(r1s),
                                                  _u8p(dst), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h)))
    return dst


# ================================================================================================
# C. residual / transform / quant / inverse / recon
# ================================================================================================
SHAPE_DEFAULT, SHAPE_N2, SHAPE_N4, SHAPE_ONLY_DC = 0, 1, 2, 3
TQ_CLEAN_SPARSE, TQ_ENABLE_CBF, TQ_CONTOURING, TQ_SKIP_RECON = 1, 2, 4, 8

TQ_JOB_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("plane", "u1"), ("size", "u1"), ("dst4x4", "u1"), ("shape", "u1"), ("qp", "u1"),
                         ("slice_type", "u1"), ("temporal_layer", "u1"), ("flags", "u1"), ("dz_offset", "<u4")])
assert TQ_JOB_DTYPE.itemsize == 16


class CoeffPicture(C.Structure):
    _fields_ = [("buf", C.c_void_p * 3), ("stride", C.c_uint32 * 3)]


class HostCoeffPicture:
    def __init__(self, width, height):
        self.planes = [np.zeros((height, width), np.int16), np.zeros((height // 2, width // 2), np.int16), np.zeros((height // 2, width // 2), np.int16)]
        self.struct = CoeffPicture()
        for i, p in enumerate(self.planes):
            self.struct.buf[i] = p.ctypes.data
            self.struct.stride[i] = p.shape[1]


def tq_batch(src, pred, recon, coeff, jobs, hbd=False):
    """src/pred/recon: HostYuvPicture (recon may be pred; dtype uint16 with hbd); coeff: HostCoeffPicture; jobs: TQ_JOB_DTYPE array -> nz counts."""
    jobs = np.ascontiguousarray(jobs)
    nz = np.zeros(len(jobs), np.uint32)
    check((lib().svtb200_tq_batch16 if hbd else lib().svtb200_tq_batch)(C.byref(src.struct), C.byref(pred.struct), C.byref(recon.struct), C.byref(coeff.struct), _u8p(jobs),
                                 C.c_int(len(jobs)), _u8p(nz)))
    return nz


# ---- distortion leaves of mode decision --------------------------------------------------------------------------
DIST_JOB_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("plane", "u1"), ("width", "u1"), ("height", "u1"), ("kind", "u1")])
assert DIST_JOB_DTYPE.itemsize == 8
FULLDIST_DEFAULT, FULLDIST_INTRA, FULLDIST_CBF_ZERO = 0, 1, 2
SATD_8X8, SATD_4X4 = 0, 1


def spatial_distortion_batch(inp, recon, jobs):
    """SpatialFullDistortionKernel per job: sum (input - recon)^2 -> uint64[n]"""
    jobs = np.ascontiguousarray(jobs)
    out = np.zeros(len(jobs), np.uint64)
    check(lib().svtb200_spatial_distortion_batch(C.byref(inp.struct), C.byref(recon.struct), _u8p(jobs), C.c_int(len(jobs)), _u8p(out)))
    return out


def full_distortion_batch(coeff, recon_coeff, width, height, jobs):
    """FullDistortionKernel{,Intra,CbfZero}_32bit per job -> uint64[n][2] (distortionResult[0], [1])"""
    jobs = np.ascontiguousarray(jobs)
    out = np.zeros((len(jobs), 2), np.uint64)
    check(lib().svtb200_full_distortion_batch(C.byref(coeff.struct), C.byref(recon_coeff.struct), C.c_uint32(width), C.c_uint32(height), _u8p(jobs),
                                              C.c_int(len(jobs)), _u8p(out)))
    return out


def satd_batch(src, pred, jobs):
    """Compute8x8Satd / Compute4x4Satd of the residual src - pred summed over the block's units -> uint64[n]"""
    jobs = np.ascontiguousarray(jobs)
    out = np.zeros(len(jobs), np.uint64)
    check(lib().svtb200_satd_batch(C.byref(src.struct), C.byref(pred.struct), _u8p(jobs), C.c_int(len(jobs)), _u8p(out)))
    return out


def estimate_transform(res, N, bit_inc=0, dst=0, shape=0):
    res = np.ascontiguousarray(res, np.int16)
    out = np.zeros((N, N), np.int16)
    check(lib().svtb200_EstimateTransform(_u8p(res), C.c_uint32(res.shape[1]), _u8p(out), C.c_uint32(N), C.c_uint32(N), C.c_uint32(bit_inc),
                                          C.c_uint8(dst), C.c_uint8(shape)))
    return out


def quantize_inv_quantize(coeff, qFunc, q_offset, shiftedQBits, shiftedFFunc, iq_offset, shiftNum, area):
    coeff = np.ascontiguousarray(coeff, np.int16)
    q = np.zeros_like(coeff)
    r = np.zeros_like(coeff)
    nz = C.c_uint32(0)
    check(lib().svtb200_QuantizeInvQuantize(_u8p(coeff), C.c_uint32(coeff.shape[1]), _u8p(q), _u8p(r), C.c_uint32(qFunc), C.c_uint32(q_offset),
                                            C.c_int32(shiftedQBits), C.c_int32(shiftedFFunc), C.c_int32(iq_offset), C.c_int32(shiftNum),
                                            C.c_uint32(area), C.byref(nz)))
    return q, r, nz.value


def encode_inv_transform(coeff, N, bit_inc=0, dst=0):
    coeff = np.ascontiguousarray(coeff, np.int16)
    out = np.zeros((N, N), np.int16)
    check(lib().svtb200_EncodeInvTransform(_u8p(coeff), C.c_uint32(coeff.shape[1]), _u8p(out), C.c_uint32(N), C.c_uint32(N), C.c_uint32(bit_inc), C.c_uint8(dst)))
    return out


# ================================================================================================
# E. deblocking
# ================================================================================================
class DlfParams(C.Structure):
    _fields_ = [("pic_width", C.c_uint16), ("pic_height", C.c_uint16), ("qp_array_stride", C.c_uint16), ("tc_offset", C.c_int8),
                ("beta_offset", C.c_int8), ("cb_qp_offset", C.c_int8), ("cr_qp_offset", C.c_int8)]


def dlf_picture(recon, vbs, hbs, qp_array, prm, hbd=False):
    """In-place deblocking of a HostYuvPicture (dtype uint16 with hbd). vbs/hbs: uint8 [n_lcu][256]; qp_array: uint8 [H/8][qp_stride]."""
    check((lib().svtb200_dlf_picture16 if hbd else lib().svtb200_dlf_picture)(C.byref(recon.struct), _u8p(vbs), _u8p(hbs), _u8p(qp_array), C.byref(prm)))


def luma_edge_dlf(buf, off, stride, vertical, tc, beta):
    check(lib().svtb200_Luma4SampleEdgeDLFCore(C.c_void_p(buf.ctypes.data + off), C.c_uint32(stride), C.c_uint8(vertical), C.c_int32(tc), C.c_int32(beta)))


def chroma_edge_dlf(cb, cr, off, stride, vertical, cb_tc, cr_tc):
    check(lib().svtb200_Chroma2SampleEdgeDLFCore(C.c_void_p(cb.ctypes.data + off), C.c_void_p(cr.ctypes.data + off), C.c_uint32(stride),
                                                 C.c_uint8(vertical), C.c_uint8(cb_tc), C.c_uint8(cr_tc)))


# ================================================================================================
# F. SAO
# ================================================================================================
SAO_STATS_DTYPE = np.dtype([("bo_diff", "<i4", 32), ("bo_count", "<u2", 32), ("eo_diff", "<i4", (4, 5)), ("eo_count", "<u2", (4, 5))])
assert SAO_STATS_DTYPE.itemsize == 312
SAO_PARAMS_DTYPE = np.dtype([("type", "u1", 2), ("band_position", "u1", 3), ("offset", "i1", (3, 4)), ("reserved", "u1", 3)])
assert SAO_PARAMS_DTYPE.itemsize == 20


def sao_stats_picture(src, recon, hbd=False):
    n_lcu = ((recon.width + 63) // 64) * ((recon.height + 63) // 64)
    out = np.zeros((n_lcu, 3), dtype=SAO_STATS_DTYPE)
    check((lib().svtb200_sao_stats_picture16 if hbd else lib().svtb200_sao_stats_picture)(C.byref(src.struct), C.byref(recon.struct), _u8p(out)))
    return out


def sao_apply_picture(inp, out, params, sao_flag=(1, 1), hbd=False):
    params = np.ascontiguousarray(params)
    flag = (C.c_uint8 * 2)(*sao_flag)
    check((lib().svtb200_sao_apply_picture16 if hbd else lib().svtb200_sao_apply_picture)(C.byref(inp.struct), C.byref(out.struct), _u8p(params), flag))


# ================================================================================================
# device-resident (_dev) entry points used by bench.py: every picture struct carries DEVICE pointers
# ================================================================================================
def make_yuv_struct(ptrs, strides, pad, width, height):
    """YuvPicture over caller-owned (device or host) planes: ptrs = (y, cb, cr) addresses of the padded allocations."""
    s = YuvPicture()
    for name, ptr, st, p, w, h in (("y", ptrs[0], strides[0], pad, width, height), ("cb", ptrs[1], strides[1], pad // 2, width // 2, height // 2),
                                   ("cr", ptrs[2], strides[2], pad // 2, width // 2, height // 2)):
        pl = getattr(s, name)
        pl.buf, pl.stride, pl.origin_x, pl.origin_y, pl.width, pl.height = ptr, st, p, p, w, h
    return s


def yuv_struct_array(ref_structs):
    """a C array of svtb200_yuv_picture descriptors (build once, reuse per call)"""
    return (YuvPicture * len(ref_structs))(*ref_structs)


def mc_batch_dev(ref_structs, jobs_ptr, n_jobs, dst_struct, stream=0):
    arr = ref_structs if isinstance(ref_structs, C.Array) else yuv_struct_array(ref_structs)
    check(lib().svtb200_mc_batch_dev(arr, C.c_int(len(ref_structs)), C.c_void_p(jobs_ptr), C.c_int(n_jobs), C.byref(dst_struct), C.c_void_p(stream)))


def tq_batch_dev(src, pred, recon, coeff_struct, jobs_ptr, n_jobs, nz_ptr, stream=0):
    check(lib().svtb200_tq_batch_dev(C.byref(src), C.byref(pred), C.byref(recon), C.byref(coeff_struct), C.c_void_p(jobs_ptr), C.c_int(n_jobs),
                                     C.c_void_p(nz_ptr), C.c_void_p(stream)))


CU_INFO_DTYPE = np.dtype([("mv", "<i2", (2, 2)), ("pred_dir", "u1"), ("intra", "u1"), ("log2_cu_size", "u1"), ("reserved", "u1")])
assert CU_INFO_DTYPE.itemsize == 12


def variance_picture(plane):
    """PictureAnalysis' per-LCU mean / variance pyramid of a padded host luma plane (Plane) -> (y_mean[n_lcu][85] u8, variance[n_lcu][85] u16)"""
    n = ((plane.width + 63) // 64) * ((plane.height + 63) // 64)
    ym, vr = np.zeros((n, 85), np.uint8), np.zeros((n, 85), np.uint16)
    check(lib().svtb200_variance_picture(C.byref(plane), _u8p(ym), _u8p(vr)))
    return ym, vr


# ---- peer memory (GOP-boundary hand-off over NVLink) ------------------------------------------------------------------------------------
def peer_alloc(nbytes):
    p = C.c_void_p()
    check(lib().svtb200_peer_alloc(C.c_size_t(nbytes), C.byref(p)))
    return p.value


def peer_free(ptr):
    check(lib().svtb200_peer_free(C.c_void_p(ptr)))


def peer_export(ptr):
    h = (C.c_uint8 * 64)()
    check(lib().svtb200_peer_export(C.c_void_p(ptr), h))
    return bytes(h)


def peer_open(handle):
    h = (C.c_uint8 * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    check(lib().svtb200_peer_open(h, C.byref(p)))
    return p.value


def peer_close(ptr):
    check(lib().svtb200_peer_close(C.c_void_p(ptr)))


def peer_copy_async(dst_ptr, src_ptr, nbytes, stream=0):
    check(lib().svtb200_peer_copy_async(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), C.c_size_t(nbytes), C.c_void_p(stream)))


def dlf_bs_picture(cu_map, cbf_map, width, height, slice_type, ref_poc_l0=0, ref_poc_l1=16):
    """boundary strengths from the picture's final mode / motion / cbf maps -> (vertical_bs, horizontal_bs), each [n_lcu][256] (row N3)"""
    cu_map, cbf_map = np.ascontiguousarray(cu_map), np.ascontiguousarray(cbf_map, np.uint8)
    n_lcu = ((width + 63) // 64) * ((height + 63) // 64)
    v, h = np.zeros((n_lcu, 256), np.uint8), np.zeros((n_lcu, 256), np.uint8)
    check(lib().svtb200_dlf_bs_picture(_u8p(cu_map), C.c_uint32(cu_map.shape[1]), _u8p(cbf_map), C.c_uint32(cbf_map.shape[1]), C.c_uint16(width), C.c_uint16(height),
                                       C.c_uint8(slice_type), C.c_uint64(ref_poc_l0), C.c_uint64(ref_poc_l1), _u8p(v), _u8p(h)))
    return v, h


def dlf_bs_picture_dev(cu_ptr, cu_stride, cbf_ptr, cbf_stride, width, height, slice_type, ref_poc_l0, ref_poc_l1, vbs_ptr, hbs_ptr, stream=0):
    check(lib().svtb200_dlf_bs_picture_dev(C.c_void_p(cu_ptr), C.c_uint32(cu_stride), C.c_void_p(cbf_ptr), C.c_uint32(cbf_stride), C.c_uint16(width), C.c_uint16(height),
                                           C.c_uint8(slice_type), C.c_uint64(ref_poc_l0), C.c_uint64(ref_poc_l1), C.c_void_p(vbs_ptr), C.c_void_p(hbs_ptr), C.c_void_p(stream)))


def dlf_picture_dev(recon, vbs_ptr, hbs_ptr, qp_ptr, prm, stream=0):
    check(lib().svtb200_dlf_picture_dev(C.byref(recon), C.c_void_p(vbs_ptr), C.c_void_p(hbs_ptr), C.c_void_p(qp_ptr), C.byref(prm), C.c_void_p(stream)))


def sao_stats_picture_dev(src, recon, out_ptr, stream=0):
    check(lib().svtb200_sao_stats_picture_dev(C.byref(src), C.byref(recon), C.c_void_p(out_ptr), C.c_void_p(stream)))


def sao_apply_picture_dev(inp, out, prm_ptr, flags=(1, 1), stream=0):
    check(lib().svtb200_sao_apply_picture_dev(C.byref(inp), C.byref(out), C.c_void_p(prm_ptr), (C.c_uint8 * 2)(*flags), C.c_void_p(stream)))


# ---- 16-bit sample planes (10-bit video): device-picture entry points of configs 3 / 4 -----------------------------------------------------
def mc_batch16_dev(ref_structs, jobs_ptr, n_jobs, dst_struct, stream=0):
    arr = ref_structs if isinstance(ref_structs, C.Array) else yuv_struct_array(ref_structs)
    check(lib().svtb200_mc_batch16_dev(arr, C.c_int(len(ref_structs)), C.c_void_p(jobs_ptr), C.c_int(n_jobs), C.byref(dst_struct), C.c_void_p(stream)))


def tq_batch16_dev(src, pred, recon, coeff_struct, jobs_ptr, n_jobs, nz_ptr, stream=0):
    check(lib().svtb200_tq_batch16_dev(C.byref(src), C.byref(pred), C.byref(recon), C.byref(coeff_struct), C.c_void_p(jobs_ptr), C.c_int(n_jobs),
                                       C.c_void_p(nz_ptr), C.c_void_p(stream)))


def dlf_picture16_dev(recon, vbs_ptr, hbs_ptr, qp_ptr, prm, stream=0):
    check(lib().svtb200_dlf_picture16_dev(C.byref(recon), C.c_void_p(vbs_ptr), C.c_void_p(hbs_ptr), C.c_void_p(qp_ptr), C.byref(prm), C.c_void_p(stream)))


def sao_stats_picture16_dev(src, recon, out_ptr, stream=0):
    check(lib().svtb200_sao_stats_picture16_dev(C.byref(src), C.byref(recon), C.c_void_p(out_ptr), C.c_void_p(stream)))


def sao_stats_picture16_h3_dev(src, deblocked, unfiltered, out_ptr, stream=0):
    check(lib().svtb200_sao_stats_picture16_h3_dev(C.byref(src), C.byref(deblocked), C.byref(unfiltered), C.c_void_p(out_ptr), C.c_void_p(stream)))


def sao_stats_picture16_h3(src, deblocked, unfiltered):
    n = ((deblocked.width + 63) // 64) * ((deblocked.height + 63) // 64)
    out = np.zeros((n, 3), dtype=SAO_STATS_DTYPE)
    check(lib().svtb200_sao_stats_picture16_h3(C.byref(src.struct), C.byref(deblocked.struct), C.byref(unfiltered.struct), _u8p(out)))
    return out


def sao_apply_picture16_dev(inp, out, prm_ptr, flags=(1, 1), stream=0):
    check(lib().svtb200_sao_apply_picture16_dev(C.byref(inp), C.byref(out), C.c_void_p(prm_ptr), (C.c_uint8 * 2)(*flags), C.c_void_p(stream)))


# ---- 10-bit pack / unpack -----------------------------------------------------------
PACK_MSB, PACK_COMPRESSED_MSB, PACK_CPACK, PACK_UNPACK, PACK_UNPACK8, PACK_UNPACK_AVG = range(6)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def msb_pack2d(in8, inn, compressed=False):
    """EB_ENC_msbPack2D / CompressedPackmsb on 2-D numpy planes (row strides taken from the arrays)."""
    h, w = in8.shape
    out = np.zeros((h, w), np.uint16)
    f = lib().svtb200_CompressedPackmsb if compressed else lib().svtb200_EB_ENC_msbPack2D
    check(f(_p(in8), C.c_uint32(in8.strides[0]), _p(inn), _p(out), C.c_uint32(inn.strides[0]), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h)))
    return out


def cpack(inn):
    h, w = inn.shape
    out = np.zeros((h, w // 4), np.uint8)
    check(lib().svtb200_CPack(_p(inn), C.c_uint32(inn.strides[0]), _p(out), C.c_uint32(w // 4), None, C.c_uint32(w), C.c_uint32(h)))
    return out


def msb_unpack2d(in16, want_nbit=True):
    h, w = in16.shape
    o8 = np.zeros((h, w), np.uint8)
    if want_nbit:
        on = np.zeros((h, w), np.uint8)
        check(lib().svtb200_EB_ENC_msbUnPack2D(_p(in16), C.c_uint32(in16.strides[0] // 2), _p(o8), _p(on), C.c_uint32(w), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h)))
        return o8, on
    check(lib().svtb200_UnPack8BitData(_p(in16), C.c_uint32(in16.strides[0] // 2), _p(o8), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h)))
    return o8


def unpack_avg(l0, l1):
    h, w = l0.shape
    o = np.zeros((h, w), np.uint8)
    check(lib().svtb200_UnpackAvg(_p(l0), C.c_uint32(l0.strides[0] // 2), _p(l1), C.c_uint32(l1.strides[0] // 2), _p(o), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h)))
    return o


def pack_plane_dev(op, in0, s_in0, in1, s_in1, out0, s_out0, out1, s_out1, width, height, stream=0):
    check(lib().svtb200_pack_plane_dev(C.c_int(op), C.c_void_p(in0), C.c_uint32(s_in0), C.c_void_p(in1), C.c_uint32(s_in1), C.c_void_p(out0), C.c_uint32(s_out0),
                                       C.c_void_p(out1), C.c_uint32(s_out1), C.c_uint32(width), C.c_uint32(height), C.c_void_p(stream)))


# ================================================================================================
# D. intra prediction
# ================================================================================================
INTRA_SKIP, INTRA_TOPLEFT_AVAIL, INTRA_STRONG_SMOOTHING = 0xFF, 1, 2
INTRA_JOB_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("size", "u1"), ("luma_mode", "u1"), ("chroma_mode", "u1"), ("flags", "u1"),
                            ("left_avail", "<u2"), ("top_avail", "<u2")])
assert INTRA_JOB_DTYPE.itemsize == 12
OIS_CU_COUNT = 85


class OisConfig(C.Structure):
    _fields_ = [("pic_width", C.c_uint16), ("pic_height", C.c_uint16), ("slice_type", C.c_uint8), ("temporal_layer", C.c_uint8),
                ("ois_kernel_level", C.c_uint8), ("ois_th_set", C.c_uint8), ("set_best_ois_distortion_to_valid", C.c_uint8),
                ("limit_ois_to_dc_mode", C.c_uint8), ("skip_ois_8x8", C.c_uint8), ("cu8x8_mode", C.c_uint8)]


def intra_prediction(mode, size, is_chroma, line):
    """leaf: one intra mode on an explicit (4*size+1)-sample reference line -> size x size block"""
    line = np.ascontiguousarray(line, np.uint8)
    pred = np.zeros((size, size), np.uint8)
    check(lib().svtb200_IntraPrediction(C.c_uint32(mode), C.c_uint32(size), C.c_int(is_chroma), _u8p(line), _u8p(pred), C.c_uint32(size)))
    return pred


def intra_batch(recon, pred, jobs, hbd=False):
    """recon / pred: HostYuvPicture (may be the same object; dtype uint16 with hbd); jobs: INTRA_JOB_DTYPE array"""
    jobs = np.ascontiguousarray(jobs)
    check((lib().svtb200_intra_batch16 if hbd else lib().svtb200_intra_batch)(C.byref(recon.struct), C.byref(pred.struct), _u8p(jobs), C.c_uint32(len(jobs))))


def intra_batch_dev(recon_struct, pred_struct, jobs_ptr, n_jobs, stream=0):
    check(lib().svtb200_intra_batch_dev(C.byref(recon_struct), C.byref(pred_struct), C.c_void_p(jobs_ptr), C.c_uint32(n_jobs), C.c_void_p(stream)))


def ois_picture(cfg, src_plane, me_results, cap):
    """src_plane: Plane (host luma); me_results: ME_PU_RESULT_DTYPE [n_lcu][85] or None -> (total [n_lcu][85] u8, cand [n_lcu][85][cap] u32)"""
    n_lcu = ((cfg.pic_width + 63) // 64) * ((cfg.pic_height + 63) // 64)
    total = np.zeros((n_lcu, OIS_CU_COUNT), np.uint8)
    cand = np.zeros((n_lcu, OIS_CU_COUNT, cap), np.uint32)
    check(lib().svtb200_ois_picture(C.byref(cfg), C.byref(src_plane), _u8p(me_results) if me_results is not None else None, _u8p(total), _u8p(cand), C.c_uint32(cap)))
    return total, cand


def ois_picture_dev(cfg, src_plane_dev, me_ptr, total_ptr, cand_ptr, cap, stream=0):
    check(lib().svtb200_ois_picture_dev(C.byref(cfg), C.byref(src_plane_dev), C.c_void_p(me_ptr), C.c_void_p(total_ptr), C.c_void_p(cand_ptr), C.c_uint32(cap),
                                        C.c_void_p(stream)))


def tq_pack_coeffs_dev(coeff, jobs_ptr, nz_ptr, n_jobs, stream_ptr, capacity, offsets_ptr, stream=0):
    """coeff: CoeffPicture with device pointers; packs the coefficients of the TUs with nz != 0 into one dense int16 stream"""
    check(lib().svtb200_tq_pack_coeffs_dev(C.byref(coeff), C.c_void_p(jobs_ptr), C.c_void_p(nz_ptr), C.c_int(n_jobs), C.c_void_p(stream_ptr), C.c_uint32(capacity),
                                           C.c_void_p(offsets_ptr), C.c_void_p(stream)))


def yuv_picture_upload(dev_struct, y_ptr, y_stride, cb_ptr, cb_stride, cr_ptr, cr_stride, bps=1, stream=0):
    check(lib().svtb200_yuv_picture_upload(C.byref(dev_struct), C.c_void_p(y_ptr), C.c_uint32(y_stride), C.c_void_p(cb_ptr), C.c_uint32(cb_stride),
                                           C.c_void_p(cr_ptr), C.c_uint32(cr_stride), C.c_uint32(bps), C.c_void_p(stream)))


def yuv_picture_copy_dev(src_struct, dst_struct, stream=0, bytes_per_sample=1):
    check(lib().svtb200_yuv_picture_copy_dev(C.byref(src_struct), C.byref(dst_struct), C.c_uint32(bytes_per_sample), C.c_void_p(stream)))


def copy_async(dst_ptr, src_ptr, nbytes, device_to_host, stream=0):
    check(lib().svtb200_copy_async(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), C.c_size_t(nbytes), C.c_int(1 if device_to_host else 0), C.c_void_p(stream)))


# ------------------------------------------------------------------------------------------------
# SAO rate-distortion decision (row f2 / next-row N3)
# ------------------------------------------------------------------------------------------------
class SaoRateTables(C.Structure):
    _fields_ = [("merge_flag_bits", C.c_uint32 * 2), ("type_index_bits", C.c_uint32 * 6), ("offset_trun_unary_bits", C.c_uint32 * 8)]


SAO_DECISION_DTYPE = np.dtype([("merge_left", np.uint8), ("merge_up", np.uint8), ("type", np.uint8, 2), ("band_position", np.uint8, 3), ("reserved", np.uint8),
                               ("offset", np.int32, (3, 4)), ("luma_cost", np.int64), ("chroma_cost", np.int64)], align=True)


def sao_decision_picture(stats, lcu_cols, lcu_rows, rates, full_lambda, full_chroma_lambda_sao, is10bit=0):
    """Host-buffer SAO decision of a picture: stats[(lcu, component)] -> one SAO_DECISION_DTYPE record per LCU."""
    out = np.zeros(lcu_cols * lcu_rows, dtype=SAO_DECISION_DTYPE)
    r = SaoRateTables()
    C.memmove(C.byref(r), C.byref(rates), C.sizeof(r))
    check(lib().svtb200_sao_decision_picture(_u8p(stats), C.c_int(lcu_cols), C.c_int(lcu_rows), C.byref(r), C.c_uint64(full_lambda),
                                             C.c_uint64(full_chroma_lambda_sao), C.c_int(is10bit), _u8p(out)))
    return out


def sao_decision_picture_dev(stats_ptr, lcu_cols, lcu_rows, rates, full_lambda, full_chroma_lambda_sao, is10bit, out_ptr, params_ptr=0, stream=0):
    """SAO decision of a picture on device buffers: stats -> svtb200_sao_decision records (+ optionally the svtb200_sao_params sao_apply consumes)"""
    r = SaoRateTables()
    C.memmove(C.byref(r), C.byref(rates), C.sizeof(r))
    check(lib().svtb200_sao_decision_picture_dev(C.c_void_p(stats_ptr), C.c_int(lcu_cols), C.c_int(lcu_rows), C.byref(r), C.c_uint64(full_lambda),
                                                 C.c_uint64(full_chroma_lambda_sao), C.c_int(is10bit), C.c_void_p(out_ptr), C.c_void_p(params_ptr or None), C.c_void_p(stream)))


def sao_stats_picture_h3(src, deblocked, unfiltered):
    """H3 statistics (host pictures, HostYuvPicture)"""
    n = ((deblocked.width + 63) // 64) * ((deblocked.height + 63) // 64)
    out = np.zeros((n, 3), dtype=SAO_STATS_DTYPE)
    check(lib().svtb200_sao_stats_picture_h3(C.byref(src.struct), C.byref(deblocked.struct), C.byref(unfiltered.struct), _u8p(out)))
    return out


def sao_stats_picture_h3_dev(src_struct, deblocked_struct, unfiltered_struct, out_ptr, stream=0):
    """H3 statistics (device pictures): `deblocked` except the right / bottom 4-sample strips of every LCU, which come from `unfiltered`."""
    check(lib().svtb200_sao_stats_picture_h3_dev(C.byref(src_struct), C.byref(deblocked_struct), C.byref(unfiltered_struct), C.c_void_p(out_ptr), C.c_void_p(stream)))
