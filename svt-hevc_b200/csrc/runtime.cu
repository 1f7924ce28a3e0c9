// runtime.cu -- library lifetime, per-thread contexts, error reporting.
// Hook point in the reference: EbInitEncoder right after ASM_TYPES is derived
// (Source/Lib/Codec/EbEncHandle.c:702-707); failures map to EB_ErrorInsufficientResources.
#include <stdarg.h>
#include <mutex>

#include "common.cuh"

namespace svtb200 {

std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_device{-1};          // -1: not initialised, -2: no device
static std::mutex g_init_mu;
// the last error is per calling thread (the boundary promises lock-free concurrent callers: 8*u ME threads + EncDec threads);
// svtb200_last_error() returns the calling thread's own message
static thread_local char t_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_plane_tensor_map(CUtensorMap *out, const void *buf, uint32_t pitch, uint32_t rows, uint32_t box_w, uint32_t box_h)
{
    static std::atomic<EncodeTiledFn> g_fn{nullptr};
    EncodeTiledFn fn = g_fn.load(std::memory_order_acquire);
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return SVTB200_ERR_CUDA;
        }
        fn = (EncodeTiledFn)p;
        g_fn.store(fn, std::memory_order_release);
    }
    if (((uintptr_t)buf & 15) || (pitch & 15) || (box_w & 15) || box_w == 0 || box_w > 256 || box_h == 0 || box_h > 256) {
        set_error("tensor map: buffer %p pitch %u box %ux%u violates the 16-byte / 256-element rules", buf, pitch, box_w, box_h);
        return SVTB200_ERR_UNSUPPORTED;
    }
    const cuuint64_t dims[2] = {pitch, rows}, strides[1] = {pitch};
    const cuuint32_t box[2] = {box_w, box_h}, es[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(buf), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for pitch %u rows %u box %ux%u", (int)r, pitch, rows, box_w, box_h); return SVTB200_ERR_CUDA; }
    return SVTB200_OK;
}

int ensure_device()
{
    int d = g_device.load(std::memory_order_acquire);
    if (d == -1) {
        std::lock_guard<std::mutex> lk(g_init_mu);
        d = g_device.load();
        if (d == -1) {
            int n = 0;
            if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
                cudaGetLastError();
                set_error("no CUDA device visible: libsvthevc_b200 has no CPU fallback");
                g_device.store(-2);
                return SVTB200_ERR_NO_DEVICE;
            }
            // one process per GPU: honour LOCAL_RANK (torchrun) unless svtb200_init() chose a device
            int want = 0;
            if (const char *lr = getenv("LOCAL_RANK")) want = atoi(lr) % n;
            g_device.store(want);
            d = want;
        }
    }
    if (d < 0) { set_error("no CUDA device visible: libsvthevc_b200 has no CPU fallback"); return SVTB200_ERR_NO_DEVICE; }
    static thread_local int t_dev = -1;          // the device this thread last selected (svtb200_init may change the ordinal)
    if (t_dev != d) { if (cudaSetDevice(d) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", d); return SVTB200_ERR_CUDA; } t_dev = d; }
    return SVTB200_OK;
}

int ThreadCtx::init()
{
    if (stream) return SVTB200_OK;
    int rc = ensure_device();
    if (rc) return rc;
    SVTB200_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return SVTB200_OK;
}
int ThreadCtx::dev_reserve(size_t bytes)
{
    if (bytes <= dev_cap) return SVTB200_OK;
    if (dev) { SVTB200_CUDA_TRY(cudaStreamSynchronize(stream)); SVTB200_CUDA_TRY(cudaFree(dev)); dev = nullptr; dev_cap = 0; }
    size_t cap = align_up(bytes + (bytes >> 2), 1 << 20);
    SVTB200_CUDA_TRY(cudaMalloc(&dev, cap));
    dev_cap = cap;
    return SVTB200_OK;
}
int ThreadCtx::pin_reserve(size_t bytes)
{
    if (bytes <= pin_cap) return SVTB200_OK;
    if (pin) { SVTB200_CUDA_TRY(cudaStreamSynchronize(stream)); SVTB200_CUDA_TRY(cudaFreeHost(pin)); pin = nullptr; pin_cap = 0; }
    size_t cap = align_up(bytes + (bytes >> 2), 1 << 20);
    SVTB200_CUDA_TRY(cudaMallocHost(&pin, cap));
    pin_cap = cap;
    return SVTB200_OK;
}
int ThreadCtx::dev_alloc(size_t bytes, void **out)
{
    size_t off = align_up(dev_used, 256);
    if (off + bytes > dev_cap) {
        if (dev_used != 0) { set_error("device arena overflow (reserve %zu first)", off + bytes); return SVTB200_ERR_BAD_ARG; }
        int rc = dev_reserve(off + bytes);
        if (rc) return rc;
    }
    *out = dev + off;
    dev_used = off + bytes;
    return SVTB200_OK;
}
int ThreadCtx::pin_alloc(size_t bytes, void **out)
{
    size_t off = align_up(pin_used, 256);
    if (off + bytes > pin_cap) {
        if (pin_used != 0) { set_error("pinned arena overflow (reserve %zu first)", off + bytes); return SVTB200_ERR_BAD_ARG; }
        int rc = pin_reserve(off + bytes);
        if (rc) return rc;
    }
    *out = pin + off;
    pin_used = off + bytes;
    return SVTB200_OK;
}

void ThreadCtx::release()
{
    if (stream) cudaStreamSynchronize(stream);
    if (dev) { cudaFree(dev); dev = nullptr; dev_cap = 0; }
    if (pin) { cudaFreeHost(pin); pin = nullptr; pin_cap = 0; }
    if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
    cudaGetLastError();      // a worker thread may exit after the primary context is gone: nothing to report
}
ThreadCtx::~ThreadCtx() { release(); }

ThreadCtx *tls_ctx()
{
    static thread_local ThreadCtx ctx;      // destroyed (stream, arenas released) when the calling thread exits
    return &ctx;
}

}  // namespace svtb200

using namespace svtb200;

extern "C" {

int svtb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int svtb200_init(int device_ordinal)
{
    int n = svtb200_device_count();
    if (n <= 0) { set_error("no CUDA device visible: libsvthevc_b200 has no CPU fallback"); g_device.store(-2); return SVTB200_ERR_NO_DEVICE; }
    if (device_ordinal < 0 || device_ordinal >= n) { set_error("device ordinal %d out of range (0..%d)", device_ordinal, n - 1); return SVTB200_ERR_BAD_ARG; }
    g_device.store(device_ordinal);
    int rc = ensure_device();
    if (rc) return rc;
    return tls_ctx()->init();
}

void svtb200_shutdown(void)
{
    tls_ctx()->release();      // the calling thread's context; every other thread's is released when that thread exits
}

const char *svtb200_version(void) { return "svt-hevc_b200 0.1 (sm_100a)"; }
const char *svtb200_last_error(void) { return t_err; }
uint64_t svtb200_launch_count(void) { return g_launches.load(); }

}  // extern "C"
