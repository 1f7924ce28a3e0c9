"""Shared test infrastructure: synthetic YUV, loaders for the reference build (oracle/_ref) and the
plain-C port (oracle/liboracle_port.so).  Only tests/, smoke() and bench.py's baseline legs use this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")


def cpu_has(flag):
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return flag in line.split()
    except OSError:
        pass
    return False


_cache = {}


def load_port():
    """The plain-C restatement (oracle/port/*.c).  Built on demand with gcc."""
    if "port" not in _cache:
        so = os.path.join(ORACLE, "liboracle_port.so")
        srcs = [os.path.join(ORACLE, "port", f) for f in os.listdir(os.path.join(ORACLE, "port")) if f.endswith(".c")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", ORACLE, "liboracle_port.so"], stdout=subprocess.DEVNULL)
        _cache["port"] = C.CDLL(so)
    return _cache["port"]


def load_ref(variant=None):
    """The UNMODIFIED reference compiled in place (oracle/_ref/libsvtref_{avx2,avx512}.so), or None."""
    if variant is None:
        variant = "avx512" if cpu_has("avx512bw") else "avx2"
    key = "ref_" + variant
    if key not in _cache:
        so = os.path.join(ORACLE, "_ref", f"libsvtref_{variant}.so")
        if not os.path.exists(so) and os.path.isdir("/root/reference/Source"):
            subprocess.check_call([os.path.join(ORACLE, "build_ref.sh")], stdout=subprocess.DEVNULL)
        _cache[key] = C.CDLL(so) if os.path.exists(so) else None
    return _cache[key]


def ref_asm_simd_ok():
    """The reference's SIMD rows need AVX2 on the host."""
    return cpu_has("avx2")


# ---------------------------------------------------------------------------------------------
# synthetic video (SURVEY §8d recipe): low-pass random texture translating (+2,+1) px/frame, a second
# layer moving (-3,0) in the lower half, plus Gaussian noise sigma 1.5; chroma: smooth fields.
# ---------------------------------------------------------------------------------------------
def _box(a, k):
    c = np.cumsum(np.cumsum(np.pad(a, ((k, 0), (k, 0)), mode="wrap").astype(np.float64), axis=0), axis=1)
    return (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)


class SynthVideo:
    def __init__(self, width, height, seed=0x5F7C0001, texture_box=8):
        self.w, self.h, self.seed = width, height, seed
        rng = np.random.default_rng(seed)
        m = 128
        base = rng.integers(0, 256, (height + 2 * m, width + 2 * m)).astype(np.float64)
        t = _box(base, texture_box)
        t = (t - t.mean()) / (t.std() + 1e-9) * 48 + 128
        self.tex = t
        t2 = _box(rng.integers(0, 256, (height + 2 * m, width + 2 * m)).astype(np.float64), texture_box)
        self.tex2 = (t2 - t2.mean()) / (t2.std() + 1e-9) * 40 + 120
        self.m = m
        cb = _box(rng.integers(0, 256, (height // 2 + 64, width // 2 + 64)).astype(np.float64), 16)
        cr = _box(rng.integers(0, 256, (height // 2 + 64, width // 2 + 64)).astype(np.float64), 16)
        self.cb = np.clip((cb - cb.mean()) / (cb.std() + 1e-9) * 12 + 128, 0, 255)
        self.cr = np.clip((cr - cr.mean()) / (cr.std() + 1e-9) * 12 + 128, 0, 255)

    def luma(self, n):
        m, h, w = self.m, self.h, self.w
        dx, dy = (2 * n) % m, (1 * n) % m
        f = self.tex[m - dy:m - dy + h, m - dx:m - dx + w].copy()
        dx2 = (3 * n) % m
        f[h // 2:, :] = 0.5 * f[h // 2:, :] + 0.5 * self.tex2[m + h // 2:m + h, m + dx2:m + dx2 + w]
        rng = np.random.default_rng(self.seed + 1000 + n)
        f += rng.normal(0.0, 1.5, f.shape)
        return np.clip(np.rint(f), 0, 255).astype(np.uint8)

    def chroma(self, n):
        h, w = self.h // 2, self.w // 2
        d = n % 32
        cb = self.cb[32 - d // 2:32 - d // 2 + h, 32 - d:32 - d + w]
        cr = self.cr[32 - d // 2:32 - d // 2 + h, 32 - d:32 - d + w]
        return np.rint(cb).astype(np.uint8), np.rint(cr).astype(np.uint8)


def adversarial_luma(kind, width, height, seed=1):
    rng = np.random.default_rng(seed)
    if kind == "constant":
        return np.full((height, width), 77, np.uint8)
    if kind == "random":
        return rng.integers(0, 256, (height, width), dtype=np.uint8)
    if kind == "checker":
        y, x = np.mgrid[0:height, 0:width]
        return (((x + y) & 1) * 255).astype(np.uint8)
    if kind == "gradient":
        y, x = np.mgrid[0:height, 0:width]
        return ((x * 3 + y * 5) & 255).astype(np.uint8)
    raise ValueError(kind)


# ---------------------------------------------------------------------------------------------
# ME drivers on the oracle side
# ---------------------------------------------------------------------------------------------
def oracle_me_picture(libhandle, fn_name, cfg, src, ref0, ref1=None, asm_mode=None):
    """Run port_me_picture (asm_mode None) or svtref_me_picture (asm_mode 0/1) with the product's ABI structs."""
    import svtb200 as S
    n_lcu = ((cfg.pic_width + 63) // 64) * ((cfg.pic_height + 63) // 64)
    res = np.zeros((n_lcu, 85), dtype=S.ME_PU_RESULT_DTYPE)
    bs = np.zeros((n_lcu, 2, 85), dtype=np.uint32)
    bm = np.zeros((n_lcu, 2, 85), dtype=np.uint32)
    rc = np.zeros(n_lcu, dtype=np.uint32)
    fn = getattr(libhandle, fn_name)
    fn.restype = C.c_int
    args = [C.byref(cfg), C.byref(src.struct), C.byref(ref0.struct), C.byref(ref1.struct) if ref1 is not None else None,
            res.ctypes.data_as(C.c_void_p), bs.ctypes.data_as(C.c_void_p), bm.ctypes.data_as(C.c_void_p),
            rc.ctypes.data_as(C.c_void_p)]
    if asm_mode is not None:
        args.append(C.c_int(asm_mode))
    r = fn(*args)
    assert r == 0, f"{fn_name} returned {r}"
    return res, bs, bm, rc


def assert_me_equal(a, b, num_lists, what=""):
    """Compare (results, best_sad, best_mv, rc) tuples; list-1 raw arrays only when two lists were searched."""
    ra, rb = a[0], b[0]
    for f in ra.dtype.names:
        if num_lists == 1 and f in ("x_mv_l1", "y_mv_l1"):
            continue
        if not np.array_equal(ra[f], rb[f]):
            bad = np.argwhere(ra[f] != rb[f])[0]
            raise AssertionError(f"{what}: results.{f} differ first at {tuple(bad)}: {ra[f][tuple(bad)]} vs {rb[f][tuple(bad)]}")
    nl = num_lists
    if a[1] is not None and b[1] is not None:
        assert np.array_equal(a[1][:, :nl], b[1][:, :nl]), f"{what}: pLcuBestSad differs"
    if a[2] is not None and b[2] is not None:
        assert np.array_equal(a[2][:, :nl], b[2][:, :nl]), f"{what}: pLcuBestMV differs"
    if a[3] is not None and b[3] is not None:
        assert np.array_equal(a[3], b[3]), f"{what}: rcMEdistortion differs"
