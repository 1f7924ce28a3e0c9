// common.cuh -- shared host/device plumbing for libsvthevc_b200.so (sm_100a only).
//
// Threading model mirrors the reference's: leaf kernels are called concurrently and lock-free
// from many EncDec / ME pthreads (SURVEY §8b "Threading").  Every calling thread gets its own
// CUDA stream and its own grow-only device scratch arena + pinned staging block (thread_local),
// so no entry point takes a lock after first use.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

#include "svt_hevc_b200.h"

namespace svtb200 {

extern std::atomic<uint64_t> g_launches;
void set_error(const char *fmt, ...);
int  ensure_device();   // SVTB200_OK or SVTB200_ERR_NO_DEVICE; selects the device for this thread

#define SVTB200_CUDA_TRY(expr)                                                                  \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::svtb200::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return SVTB200_ERR_CUDA;                                                            \
        }                                                                                       \
    } while (0)

// Per-thread execution context.
struct ThreadCtx {
    cudaStream_t stream = nullptr;
    uint8_t *dev = nullptr;   size_t dev_cap = 0,  dev_used = 0;
    uint8_t *pin = nullptr;   size_t pin_cap = 0,  pin_used = 0;
    int init();
    void release();
    ~ThreadCtx();
    void reset() { dev_used = 0; pin_used = 0; }
    // bump allocators; pointers stay valid until the next reset(); grow = sync + realloc
    int dev_alloc(size_t bytes, void **out);
    int pin_alloc(size_t bytes, void **out);
    // reserve the total up front so later dev_alloc() calls never move the block
    int dev_reserve(size_t bytes);
    int pin_reserve(size_t bytes);
};
ThreadCtx *tls_ctx();

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// 2-D uint8 tensor map over a padded plane buffer (pitch % 16 == 0, base % 16 == 0) with a box of box_w x box_h bytes (box_w % 16 == 0, both <= 256)
int make_plane_tensor_map(CUtensorMap *out, const void *buf, uint32_t pitch, uint32_t rows, uint32_t box_w, uint32_t box_h);

#define SVTB200_COUNT_LAUNCH() (::svtb200::g_launches.fetch_add(1, std::memory_order_relaxed))

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// 4 consecutive bytes starting at an arbitrary byte address of a read-only GLOBAL buffer: two aligned
// 32-bit loads + one funnel shift (SASS: SHF.R.W).
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *p)
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    // every caller reads a picture plane in global memory that the kernel never writes: LDG (read-only path) instead of a generic LD
    const uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    return __funnelshift_r(lo, __ldg(w + 1), sh);
}

// c + the four byte-wise absolute differences of a and b in ONE instruction (VABSDIFF4 with its accumulate operand).  __vsadu4() passes 0 there and
// leaves the addition to a separate IADD3 that ptxas only sometimes folds back (static count in the ME search bodies: 540 IADD3 next to 1152 VABSDIFF4).
__device__ __forceinline__ uint32_t vsad4_acc(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
// division of small non-negative n by a loop-invariant d: a shift when d is a power of two (every size the batch kernel sees), else a
// multiply-high with ceil(2^32 / d) (exact for n * d < 2^32 / d; the leaf entry points accept any multiple of 4)
struct SmallDiv { uint32_t m; int sh; };
__device__ __forceinline__ SmallDiv mk_div(int d)
{
    SmallDiv q;
    q.sh = 31 - __clz(d);
    q.m = (d & (d - 1)) ? 0xFFFFFFFFu / (uint32_t)d + 1u : 0u;
    return q;
}
__device__ __forceinline__ int udiv(int n, const SmallDiv &q) { return q.m ? (int)__umulhi((uint32_t)n, q.m) : n >> q.sh; }

// ---- TMA (cp.async.bulk.tensor) + mbarrier primitives --------------------------------------------------
// Picture planes are described by 2-D uint8 tensor maps over the PADDED buffer (row pitch a multiple of 16 bytes); a box is fetched
// at an ARBITRARY byte coordinate and lands densely packed at a 128-byte-aligned shared-memory address, i.e. the copy engine does the
// byte realignment that the register path (load_window_t) leaves to funnel shifts.  Out-of-buffer parts of a box are zero-filled and
// still counted in the transaction bytes.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// every thread of the CTA polls the phase bit; a transfer that never completes traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (spin > (1 << 22)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

// ---- window staging ------------------------------------------------------------------------
// Copies `rows` rows of `wbytes` bytes starting at global address g (row pitch gstride) into shared memory with row pitch
// wp = win_pitch(wbytes) (s_win 16-byte aligned).  The global byte misalignment `a` is preserved: window byte x lives at
// column a + x, and a is returned (a & 3 is what the funnel-shift consumers specialise on).
//   pitch multiple of 16 : 128-bit loads from the 16-byte-aligned start, a in 0..15   (device-resident pictures)
//   pitch multiple of 4  : 32-bit loads, a in 0..3                                     (host pictures copied as they are)
//   otherwise            : byte loads, a = 0
// A warp covers 32 / P rows per pass, P = the power of two >= the vectors per row, so narrow windows still use most lanes.
template <int NTHREADS>
__device__ __forceinline__ int load_window_t(uint8_t *s_win, int wp, const uint8_t *g, int gstride, int wbytes, int rows)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if ((gstride & 15) == 0) {
        const int a = (int)((uintptr_t)g & 15);
        const uint4 *gv = (const uint4 *)(g - a);
        const int nV = (a + wbytes + 15) >> 4, gs = gstride >> 4, ws = wp >> 4;
        int P = 1;
        while (P < nV) P <<= 1;
        uint4 *sv = (uint4 *)s_win;
        if (P <= 32) {
            const int rpw = 32 / P, sub = lane / P, c = lane & (P - 1);
            for (int r = wid * rpw + sub; r < rows; r += (NTHREADS / 32) * rpw)
                if (c < nV) sv[r * ws + c] = __ldg(gv + (ptrdiff_t)r * gs + c);
        } else {
            for (int r = wid; r < rows; r += NTHREADS / 32)
                for (int c = lane; c < nV; c += 32) sv[r * ws + c] = __ldg(gv + (ptrdiff_t)r * gs + c);
        }
        return a;
    }
    if ((gstride & 3) == 0) {
        const int a = (int)((uintptr_t)g & 3);
        const uint32_t *gw = (const uint32_t *)(g - a);
        const int nW = (a + wbytes + 3) >> 2, gs4 = gstride >> 2, wp4 = wp >> 2;
        uint32_t *sw = (uint32_t *)s_win;
        int P = 1;
        while (P < nW && P < 32) P <<= 1;
        const int rpw = 32 / P, sub = lane / P, c0 = lane & (P - 1);
        for (int r = wid * rpw + sub; r < rows; r += (NTHREADS / 32) * rpw)
            for (int c = c0; c < nW; c += P) sw[r * wp4 + c] = __ldg(gw + (ptrdiff_t)r * gs4 + c);
        return a;
    }
    for (int r = wid; r < rows; r += NTHREADS / 32)
        for (int c = lane; c < wbytes; c += 32) s_win[r * wp + c] = __ldg(g + (ptrdiff_t)r * gstride + c);
    return 0;
}
// row pitch of a staged window: room for the 0..15 byte misalignment, the 16-byte read granularity and the few bytes the
// funnel-shift consumers read past the last needed byte
__device__ __host__ __forceinline__ int win_pitch(int wbytes) { return ((wbytes + 15 + 15) & ~15) + 16; }
#endif

}  // namespace svtb200
