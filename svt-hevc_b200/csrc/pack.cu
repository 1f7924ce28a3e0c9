// pack.cu -- 10-bit sample packing: (8-bit MSB plane, 2-bit LSB plane) <-> 16-bit samples.
//
// Replaces:
//   EB_ENC_msbPack2D / CompressedPackmsb / CPack_C / EB_ENC_msbUnPack2D / UnPack8BitData[SafeSub] / UnpackAvg[SafeSub]
//     C_DEFAULT/EbPackUnPack_C.c:12,44,93,130,153,174,203,226 (+ the AVX2 / AVX512 / SSE2 rows of Codec/EbPackUnPack.h:27-169)
//   callers: Pack2D_SRC / CompressedPackLcu / UnPack2D (Codec/EbPictureOperators.c:512-567), EncodePassPackLcu (EbCodingLoop.c:2867)
//
// Pure streaming byte work, HBM-bound: every thread moves 8 samples (64-bit / 128-bit accesses when all planes are 8-sample aligned,
// element accesses otherwise); one launch covers a whole plane, grid y = rows.  Algorithmic bytes per sample: pack 1 + 1 (or 1/4) in,
// 2 out; unpack the reverse.
#include "common.cuh"

namespace svtb200 {

enum PackOp { OP_PACK = 0, OP_CPACKMSB = 1, OP_CPACK = 2, OP_UNPACK = 3, OP_UNPACK8 = 4, OP_UNPACK_AVG = 5 };

struct PackArgs {
    const void *in0, *in1;   // pack: 8-bit, n-bit; unpack: 16-bit (, second 16-bit for the average)
    void *out0, *out1;
    int s_in0, s_in1, s_out0, s_out1;   // strides in elements
    int width, height;
};

constexpr int PK_NT = 128;

__device__ __forceinline__ uint32_t spread_lo(uint32_t b4, int k) { return ((b4 >> (8 * k)) & 0xffu); }

template <int OP, bool VEC>
__global__ void __launch_bounds__(PK_NT) pack_kernel(const __grid_constant__ PackArgs A)
{
    const int y = blockIdx.y;
    const int x0 = (blockIdx.x * PK_NT + threadIdx.x) * 8;
    if (x0 >= A.width) return;
    const int n = min(8, A.width - x0);
    if (OP == OP_PACK || OP == OP_CPACKMSB) {
        const uint8_t *p8 = (const uint8_t *)A.in0 + (size_t)y * A.s_in0 + x0;
        uint16_t *o = (uint16_t *)A.out0 + (size_t)y * A.s_out0 + x0;
        if (VEC && n == 8) {
            const uint2 m = *(const uint2 *)p8;
            uint32_t lsb[8];
            if (OP == OP_PACK) {
                const uint2 l = *(const uint2 *)((const uint8_t *)A.in1 + (size_t)y * A.s_in1 + x0);
#pragma unroll
                for (int k = 0; k < 4; ++k) { lsb[k] = (l.x >> (8 * k + 6)) & 3u; lsb[4 + k] = (l.y >> (8 * k + 6)) & 3u; }
            } else {
                const uint32_t c = *(const uint16_t *)((const uint8_t *)A.in1 + (size_t)y * A.s_in1 + (x0 >> 2));   // two bytes = 8 pels, first pel in bits 7:6
#pragma unroll
                for (int k = 0; k < 4; ++k) { lsb[k] = (c >> (6 - 2 * k)) & 3u; lsb[4 + k] = (c >> (14 - 2 * k)) & 3u; }
            }
            uint4 r;
            r.x = ((spread_lo(m.x, 0) << 2) | lsb[0]) | (((spread_lo(m.x, 1) << 2) | lsb[1]) << 16);
            r.y = ((spread_lo(m.x, 2) << 2) | lsb[2]) | (((spread_lo(m.x, 3) << 2) | lsb[3]) << 16);
            r.z = ((spread_lo(m.y, 0) << 2) | lsb[4]) | (((spread_lo(m.y, 1) << 2) | lsb[5]) << 16);
            r.w = ((spread_lo(m.y, 2) << 2) | lsb[6]) | (((spread_lo(m.y, 3) << 2) | lsb[7]) << 16);
            *(uint4 *)o = r;
        } else {
            for (int k = 0; k < n; ++k) {
                uint32_t l;
                if (OP == OP_PACK) l = (((const uint8_t *)A.in1)[(size_t)y * A.s_in1 + x0 + k] >> 6) & 3u;
                else l = (((const uint8_t *)A.in1)[(size_t)y * A.s_in1 + ((x0 + k) >> 2)] >> (6 - 2 * ((x0 + k) & 3))) & 3u;
                o[k] = (uint16_t)(((uint32_t)p8[k] << 2) | l);
            }
        }
    } else if (OP == OP_CPACK) {
        // 4 unpacked 2-bit pels (each in bits 7:6 of its byte) -> one byte, first pel in bits 7:6
        const uint8_t *p = (const uint8_t *)A.in0 + (size_t)y * A.s_in0 + x0;
        uint8_t *o = (uint8_t *)A.out0 + (size_t)y * A.s_out0 + (x0 >> 2);
        for (int g = 0; g * 4 < n; ++g) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) v |= ((uint32_t)p[g * 4 + k] >> (2 * k)) & (0xC0u >> (2 * k));
            o[g] = (uint8_t)v;
        }
    } else if (OP == OP_UNPACK || OP == OP_UNPACK8) {
        const uint16_t *p = (const uint16_t *)A.in0 + (size_t)y * A.s_in0 + x0;
        uint8_t *o8 = (uint8_t *)A.out0 + (size_t)y * A.s_out0 + x0;
        uint8_t *on = OP == OP_UNPACK ? (uint8_t *)A.out1 + (size_t)y * A.s_out1 + x0 : nullptr;
        if (VEC && n == 8) {
            const uint4 v = *(const uint4 *)p;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint2 m = {0, 0}, l = {0, 0};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a = w[k] & 0xffffu, b = w[k] >> 16;
                const uint32_t mm = ((a >> 2) & 0xffu) | (((b >> 2) & 0xffu) << 8), ll = ((a << 6) & 0xffu) | (((b << 6) & 0xffu) << 8);
                if (k < 2) { m.x |= mm << (16 * k); l.x |= ll << (16 * k); } else { m.y |= mm << (16 * (k - 2)); l.y |= ll << (16 * (k - 2)); }
            }
            *(uint2 *)o8 = m;
            if (OP == OP_UNPACK) *(uint2 *)on = l;
        } else {
            for (int k = 0; k < n; ++k) {
                const uint32_t a = p[k];
                o8[k] = (uint8_t)(a >> 2);
                if (OP == OP_UNPACK) on[k] = (uint8_t)(a << 6);
            }
        }
    } else {   // OP_UNPACK_AVG
        const uint16_t *p0 = (const uint16_t *)A.in0 + (size_t)y * A.s_in0 + x0, *p1 = (const uint16_t *)A.in1 + (size_t)y * A.s_in1 + x0;
        uint8_t *o = (uint8_t *)A.out0 + (size_t)y * A.s_out0 + x0;
        if (VEC && n == 8) {
            const uint4 a = *(const uint4 *)p0, b = *(const uint4 *)p1;
            const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
            uint32_t ma[2] = {0, 0}, mb[2] = {0, 0};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ma[k >> 1] |= ((((wa[k] & 0xffffu) >> 2) & 0xffu) | ((((wa[k] >> 16) >> 2) & 0xffu) << 8)) << (16 * (k & 1));
                mb[k >> 1] |= ((((wb[k] & 0xffffu) >> 2) & 0xffu) | ((((wb[k] >> 16) >> 2) & 0xffu) << 8)) << (16 * (k & 1));
            }
            uint2 r;
            r.x = __vavgu4(ma[0], mb[0]); r.y = __vavgu4(ma[1], mb[1]);
            *(uint2 *)o = r;
        } else {
            for (int k = 0; k < n; ++k) o[k] = (uint8_t)(((uint32_t)(uint8_t)(p0[k] >> 2) + (uint32_t)(uint8_t)(p1[k] >> 2) + 1) >> 1);
        }
    }
}

template <int OP>
static int launch_pack(const PackArgs &A, cudaStream_t st)
{
    if (A.width <= 0 || A.height <= 0) return SVTB200_OK;
    const uintptr_t al = (uintptr_t)A.in0 | (uintptr_t)A.in1 | (uintptr_t)A.out0 | (uintptr_t)A.out1;
    // 8-sample alignment of every plane: 16-bit planes 16 B, 8-bit planes 8 B, compressed plane 2 B
    const bool vec = (al & 15) == 0 && ((A.s_in0 | A.s_in1 | A.s_out0 | A.s_out1) & 15) == 0;
    dim3 grid((A.width + PK_NT * 8 - 1) / (PK_NT * 8), A.height);
    if (vec) pack_kernel<OP, true><<<grid, PK_NT, 0, st>>>(A); else pack_kernel<OP, false><<<grid, PK_NT, 0, st>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static int run_dev(int op, const PackArgs &A, cudaStream_t st)
{
    switch (op) {
    case OP_PACK: return launch_pack<OP_PACK>(A, st);
    case OP_CPACKMSB: return launch_pack<OP_CPACKMSB>(A, st);
    case OP_CPACK: return launch_pack<OP_CPACK>(A, st);
    case OP_UNPACK: return launch_pack<OP_UNPACK>(A, st);
    case OP_UNPACK8: return launch_pack<OP_UNPACK8>(A, st);
    default: return launch_pack<OP_UNPACK_AVG>(A, st);
    }
}

// element size / width divisor of the four planes of each op: {in0, in1, out0, out1}; 0 = plane unused
struct PlaneDesc { int esz[4]; int wdiv[4]; };
static PlaneDesc plane_desc(int op)
{
    switch (op) {
    case OP_PACK: return {{1, 1, 2, 0}, {1, 1, 1, 1}};
    case OP_CPACKMSB: return {{1, 1, 2, 0}, {1, 4, 1, 1}};
    case OP_CPACK: return {{1, 0, 1, 0}, {1, 1, 4, 1}};
    case OP_UNPACK: return {{2, 0, 1, 1}, {1, 1, 1, 1}};
    case OP_UNPACK8: return {{2, 0, 1, 0}, {1, 1, 1, 1}};
    default: return {{2, 2, 1, 0}, {1, 1, 1, 1}};
    }
}

// host planes: stage in pinned memory with a 16-byte aligned pitch, run, copy back
static int run_host(int op, const char *name, const void *in0, uint32_t s_in0, const void *in1, uint32_t s_in1, void *out0, uint32_t s_out0, void *out1, uint32_t s_out1,
                    uint32_t width, uint32_t height)
{
    const PlaneDesc d = plane_desc(op);
    const void *hp[4] = {in0, in1, out0, out1};
    const uint32_t hs[4] = {s_in0, s_in1, s_out0, s_out1};
    for (int k = 0; k < 4; ++k)
        if (d.esz[k] && height && width && (!hp[k] || hs[k] < (width + d.wdiv[k] - 1) / d.wdiv[k])) { set_error("%s: null plane or stride below the width", name); return SVTB200_ERR_BAD_ARG; }
    if (!width || !height) return SVTB200_OK;
    if ((op == OP_CPACKMSB || op == OP_CPACK) && (width & 3)) { set_error("%s: width must be a multiple of 4", name); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    void *dp[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t pitch[4] = {0, 0, 0, 0};
    size_t total = 0;
    for (int k = 0; k < 4; ++k) if (d.esz[k]) { pitch[k] = align_up((size_t)((width + d.wdiv[k] - 1) / d.wdiv[k]) * d.esz[k], 32); total += pitch[k] * height + 256; }
    if ((rc = c->dev_reserve(total))) return rc;
    for (int k = 0; k < 4; ++k) if (d.esz[k]) { if ((rc = c->dev_alloc(pitch[k] * height, &dp[k]))) return rc; }
    for (int k = 0; k < 2; ++k)
        if (d.esz[k]) SVTB200_CUDA_TRY(cudaMemcpy2DAsync(dp[k], pitch[k], hp[k], (size_t)hs[k] * d.esz[k], (size_t)((width + d.wdiv[k] - 1) / d.wdiv[k]) * d.esz[k], height, cudaMemcpyHostToDevice, c->stream));
    PackArgs A;
    A.in0 = dp[0]; A.in1 = dp[1]; A.out0 = dp[2]; A.out1 = dp[3];
    A.s_in0 = d.esz[0] ? (int)(pitch[0] / d.esz[0]) : 0; A.s_in1 = d.esz[1] ? (int)(pitch[1] / d.esz[1]) : 0;
    A.s_out0 = d.esz[2] ? (int)(pitch[2] / d.esz[2]) : 0; A.s_out1 = d.esz[3] ? (int)(pitch[3] / d.esz[3]) : 0;
    A.width = (int)width; A.height = (int)height;
    if ((rc = run_dev(op, A, c->stream))) return rc;
    for (int k = 2; k < 4; ++k)
        if (d.esz[k]) SVTB200_CUDA_TRY(cudaMemcpy2DAsync((void *)hp[k], (size_t)hs[k] * d.esz[k], dp[k], pitch[k], (size_t)((width + d.wdiv[k] - 1) / d.wdiv[k]) * d.esz[k], height, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

int svtb200_EB_ENC_msbPack2D(const uint8_t *in8BitBuffer, uint32_t in8Stride, const uint8_t *innBitBuffer, uint16_t *out16BitBuffer, uint32_t innStride, uint32_t outStride,
                             uint32_t width, uint32_t height)
{ return run_host(OP_PACK, "EB_ENC_msbPack2D", in8BitBuffer, in8Stride, innBitBuffer, innStride, out16BitBuffer, outStride, nullptr, 0, width, height); }

int svtb200_CompressedPackmsb(const uint8_t *in8BitBuffer, uint32_t in8Stride, const uint8_t *innBitBuffer, uint16_t *out16BitBuffer, uint32_t innStride, uint32_t outStride,
                              uint32_t width, uint32_t height)
{ return run_host(OP_CPACKMSB, "CompressedPackmsb", in8BitBuffer, in8Stride, innBitBuffer, innStride, out16BitBuffer, outStride, nullptr, 0, width, height); }

int svtb200_CPack(const uint8_t *innBitBuffer, uint32_t innStride, uint8_t *inCompnBitBuffer, uint32_t outStride, uint8_t *localCache, uint32_t width, uint32_t height)
{ (void)localCache; return run_host(OP_CPACK, "CPack", innBitBuffer, innStride, nullptr, 0, inCompnBitBuffer, outStride, nullptr, 0, width, height); }

int svtb200_EB_ENC_msbUnPack2D(const uint16_t *in16BitBuffer, uint32_t inStride, uint8_t *out8BitBuffer, uint8_t *outnBitBuffer, uint32_t out8Stride, uint32_t outnStride,
                               uint32_t width, uint32_t height)
{ return run_host(OP_UNPACK, "EB_ENC_msbUnPack2D", in16BitBuffer, inStride, nullptr, 0, out8BitBuffer, out8Stride, outnBitBuffer, outnStride, width, height); }

int svtb200_UnPack8BitData(const uint16_t *in16BitBuffer, uint32_t inStride, uint8_t *out8BitBuffer, uint32_t out8Stride, uint32_t width, uint32_t height)
{ return run_host(OP_UNPACK8, "UnPack8BitData", in16BitBuffer, inStride, nullptr, 0, out8BitBuffer, out8Stride, nullptr, 0, width, height); }

int svtb200_UnpackAvg(const uint16_t *ref16L0, uint32_t refL0Stride, const uint16_t *ref16L1, uint32_t refL1Stride, uint8_t *dstPtr, uint32_t dstStride, uint32_t width, uint32_t height)
{ return run_host(OP_UNPACK_AVG, "UnpackAvg", ref16L0, refL0Stride, ref16L1, refL1Stride, dstPtr, dstStride, nullptr, 0, width, height); }

int svtb200_pack_plane_dev(int op, const void *in0, uint32_t in0_stride, const void *in1, uint32_t in1_stride, void *out0, uint32_t out0_stride, void *out1, uint32_t out1_stride,
                           uint32_t width, uint32_t height, void *cuda_stream)
{
    if (op < 0 || op > OP_UNPACK_AVG) { set_error("pack_plane_dev: unknown op %d", op); return SVTB200_ERR_BAD_ARG; }
    const PlaneDesc d = plane_desc(op);
    const void *p[4] = {in0, in1, out0, out1};
    for (int k = 0; k < 4; ++k) if (d.esz[k] && !p[k] && width && height) { set_error("pack_plane_dev: null plane %d", k); return SVTB200_ERR_BAD_ARG; }
    if ((op == OP_CPACKMSB || op == OP_CPACK) && (width & 3)) { set_error("pack_plane_dev: width must be a multiple of 4"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    PackArgs A;
    A.in0 = in0; A.in1 = d.esz[1] ? in1 : nullptr; A.out0 = out0; A.out1 = d.esz[3] ? out1 : nullptr;
    A.s_in0 = (int)in0_stride; A.s_in1 = d.esz[1] ? (int)in1_stride : 0; A.s_out0 = (int)out0_stride; A.s_out1 = d.esz[3] ? (int)out1_stride : 0;
    A.width = (int)width; A.height = (int)height;
    return run_dev(op, A, cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

}
