// mc.cu -- HEVC motion-compensated prediction (luma 8-tap quarter-pel, chroma 4-tap eighth-pel, uni and bi).
//
// Replaces:
//   LumaInterpolationCopy / LumaInterpolationFilterPos{a..r}New, ...OutRaw, ChromaInterpolation*   C_DEFAULT/EbMcp_C.c:221-1214,3789-5099
//     and their SSSE3 rows (ASM_SSSE3/EbMcp_Intrinsic_SSSE3.c), tables Codec/EbMcpTables.c:14-745
//   BiPredClipping (C_DEFAULT/EbMcp_C.c:21), BiPredClippingOnTheFly (integer/integer shortcut, Codec/EbMcp.c:603)
//   EncodeUniPredInterpolation / EncodeBiPredInterpolation (Codec/EbMcp.c:175,562), EncodePassInterPrediction
//   position clamp (Codec/EbInterPrediction.c:794-808)
//
// Design: one CTA (128 threads) per <= 32x32 luma tile of a PU (+ its two 16x16 chroma tiles).  For each plane and
// list the (tile + taps - 1)^2 reference window is staged once in shared memory with coalesced word loads, then
//   H pass : 4 outputs per thread from funnel-shifted words with DP4A (u8 x s8), s16 rows in shared memory,
//   V pass : a 2x2 output quad per thread with DP2A on PRMT-built vertical s16 pairs.
// Both passes ALWAYS run; integer phases use the identity taps {..,64,..}.  That is bit-exact with the reference's
// 16 (64) specialised functions because their 14-bit intermediates differ from the two-pass value only by an exact
// factor of 64 that the final shift removes (see oracle/port/mc_port.c and tests/test_gpu_mc.py, every phase pair).
// HBM traffic per CTU is the algorithmic minimum + the 7-sample halo (served by L2 between neighbouring CTAs).
#include "common.cuh"

namespace svtb200 {

constexpr int MC_NT = 128;
constexpr int TILE = 32;

struct YuvD {
    const uint8_t *p[3];
    int stride[3];
    int ox, oy, w, h;   // luma origin / size; chroma = half
};
struct McArgs {
    YuvD refs[16];            // by value: a launch carries its own reference table (no staging buffer to race on between back-to-back launches)
    int n_refs;
    const svtb200_mc_job *jobs;
    int n_jobs;
    uint8_t *dst[3];
    int dstride[3];
    int dox, doy;
};

__constant__ int8_t c_luma_taps[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
__constant__ int8_t c_chroma_taps[8][4] = {{0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4}, {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_ss(uint32_t a_s16x2, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_ss(uint32_t a_s16x2, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack4(const int8_t *t) { return (uint8_t)t[0] | ((uint32_t)(uint8_t)t[1] << 8) | ((uint32_t)(uint8_t)t[2] << 16) | ((uint32_t)(uint8_t)t[3] << 24); }

// One plane, one list: window -> H pass -> V pass.  TAPS = 8 (luma) or 4 (chroma).
//   ref   : integer sample of the tile's top-left in the reference plane
//   out8  : != nullptr -> uni-pred, final samples written to global memory (pitch ostride)
//   raw   : != nullptr -> bi-pred, 14-bit intermediates written to shared memory (pitch tw)
// HBD: 16-bit samples (10-bit video).  ref / out8 are byte pointers, rstride / ostride are in SAMPLES.  The same two-pass
// form is exact there as well: first pass (S - 32768) >> 2 kept in 16 bits like LumaInterpolationFilterPos*OutRaw16bit, second pass
// (S2 + 524800) >> 10 (uni) or S2 >> 6 (bi); with identity taps these reduce to the reference's 1-D (S + 32) >> 6, (S - 32768) >> 2 and copy
// forms (Codec/EbDefinitions.h:1041-1059).
template <int TAPS, bool HBD>
__device__ void interp_tile(const uint8_t *ref, int rstride, int fx, int fy, int tw, int th, uint8_t *s_win, int16_t *s_tmp,
                            uint8_t *out8, int ostride, int16_t *raw)
{
    constexpr int HALF = TAPS / 2 - 1, BPS = HBD ? 2 : 1;
    const int wbytes = (tw + TAPS - 1) * BPS, wrows = th + TAPS - 1, wp = win_pitch(wbytes);
    __syncthreads();   // previous users of s_win / s_tmp are done
    const int a = load_window_t<MC_NT>(s_win, wp, ref - (HALF * rstride + HALF) * BPS, rstride * BPS, wbytes, wrows);
    __syncthreads();
    const int8_t *tx = TAPS == 8 ? c_luma_taps[fx] : c_chroma_taps[fx];
    const int8_t *ty = TAPS == 8 ? c_luma_taps[fy] : c_chroma_taps[fy];
    const uint32_t cx0 = pack4(tx), cx1 = TAPS == 8 ? pack4(tx + 4) : 0u;
    const uint32_t cy0 = pack4(ty), cy1 = TAPS == 8 ? pack4(ty + 4) : 0u;
    // ---- H pass: s_tmp[r][x], r in [0, wrows), pitch tw (s16) ----
    if (HBD) {
        // two adjacent outputs per thread from TAPS/2 + 1 words of s16 pairs; DP2A on (sample, sample) x (tap, tap)
        const int ppr = tw >> 1;
        for (int i = threadIdx.x; i < wrows * ppr; i += MC_NT) {
            const int r = i / ppr, x0 = (i - r * ppr) * 2;
            const int c0 = (a >> 1) + x0, par = c0 & 1;
            const uint32_t *row = (const uint32_t *)(s_win + r * wp) + (c0 >> 1);
            uint32_t w[TAPS / 2 + 1];
#pragma unroll
            for (int k = 0; k < TAPS / 2 + 1; ++k) w[k] = row[k];
            int o0 = 0, o1 = 0;
#pragma unroll
            for (int k = 0; k < TAPS / 2; ++k) {
                const uint32_t sh = __funnelshift_r(w[k], w[k + 1], 16);
                const uint32_t p0 = par ? sh : w[k], p1 = par ? w[k + 1] : sh;
                const uint32_t cc = k < 2 ? cx0 : cx1;
                if ((k & 1) == 0) { o0 = dp2a_lo_ss(p0, cc, o0); o1 = dp2a_lo_ss(p1, cc, o1); }
                else { o0 = dp2a_hi_ss(p0, cc, o0); o1 = dp2a_hi_ss(p1, cc, o1); }
            }
            *(uint32_t *)(s_tmp + r * tw + x0) = (uint32_t)(uint16_t)((o0 - 32768) >> 2) | ((uint32_t)(uint16_t)((o1 - 32768) >> 2) << 16);
        }
    }
    const int gpr = HBD ? 0 : tw >> 2;
    for (int i = threadIdx.x; i < wrows * gpr; i += MC_NT) {
        const int r = i / gpr, x0 = (i - r * gpr) * 4;
        const uint32_t *row = (const uint32_t *)(s_win + r * wp);
        const int c0 = a + x0, wi = c0 >> 2, sh = (c0 & 3) * 8;
        uint32_t w[TAPS == 8 ? 4 : 3];
#pragma unroll
        for (int k = 0; k < (TAPS == 8 ? 4 : 3); ++k) w[k] = row[wi + k];
        int o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // bytes c0+k .. c0+k+TAPS-1
            const int s0 = sh + 8 * k;                       // bit offset of byte k inside (w[0],w[1])
            const uint32_t lo = s0 < 32 ? __funnelshift_r(w[0], w[1], s0) : __funnelshift_r(w[1], w[2], s0 - 32);
            int acc = dp4a_us(lo, cx0, 0);
            if (TAPS == 8) {
                const int s1 = s0 + 32;
                const uint32_t hi = s1 < 64 ? __funnelshift_r(w[1], w[2], s1 - 32) : __funnelshift_r(w[2], w[3], s1 - 64);
                acc = dp4a_us(hi, cx1, acc);
            }
            o[k] = acc;
        }
        uint32_t *t = (uint32_t *)(s_tmp + r * tw + x0);
        t[0] = (uint32_t)(uint16_t)o[0] | ((uint32_t)(uint16_t)o[1] << 16);
        t[1] = (uint32_t)(uint16_t)o[2] | ((uint32_t)(uint16_t)o[3] << 16);
    }
    __syncthreads();
    // ---- V pass: 2x2 output quads ----
    const int qpr = tw >> 1;
    for (int i = threadIdx.x; i < (th >> 1) * qpr; i += MC_NT) {
        const int qy = i / qpr, x = (i - qy * qpr) * 2, y = qy * 2;
        const uint32_t *col = (const uint32_t *)(s_tmp + y * tw + x);
        const int p4 = tw >> 1;   // words per row
        uint32_t w[TAPS + 1];
#pragma unroll
        for (int k = 0; k < TAPS + 1; ++k) w[k] = col[k * p4];
        int acc[2][2];   // [row][col]
#pragma unroll
        for (int ry = 0; ry < 2; ++ry) {
            int a0 = 0, a1 = 0;
#pragma unroll
            for (int k = 0; k < TAPS; k += 2) {
                const uint32_t lo = __byte_perm(w[ry + k], w[ry + k + 1], 0x5410);   // (x, row k), (x, row k+1)
                const uint32_t hi = __byte_perm(w[ry + k], w[ry + k + 1], 0x7632);   // (x+1, ..)
                const uint32_t cc = k < 4 ? cy0 : cy1;
                if ((k & 2) == 0) { a0 = dp2a_lo_ss(lo, cc, a0); a1 = dp2a_lo_ss(hi, cc, a1); }
                else { a0 = dp2a_hi_ss(lo, cc, a0); a1 = dp2a_hi_ss(hi, cc, a1); }
            }
            acc[ry][0] = a0; acc[ry][1] = a1;
        }
        if (out8 && HBD) {
#pragma unroll
            for (int ry = 0; ry < 2; ++ry) {
                const int v0 = min(max((acc[ry][0] + 524800) >> 10, 0), 1023), v1 = min(max((acc[ry][1] + 524800) >> 10, 0), 1023);
                uint16_t *o = (uint16_t *)out8 + (size_t)(y + ry) * ostride + x;
                if (((uintptr_t)o & 3) == 0) *(uint32_t *)o = (uint32_t)v0 | ((uint32_t)v1 << 16);
                else { o[0] = (uint16_t)v0; o[1] = (uint16_t)v1; }
            }
        } else if (out8) {
#pragma unroll
            for (int ry = 0; ry < 2; ++ry) {
                const int v0 = min(max((acc[ry][0] + 2048) >> 12, 0), 255), v1 = min(max((acc[ry][1] + 2048) >> 12, 0), 255);
                uint8_t *o = out8 + (size_t)(y + ry) * ostride + x;
                if (((uintptr_t)o & 1) == 0) *(uint16_t *)o = (uint16_t)(v0 | (v1 << 8));
                else { o[0] = (uint8_t)v0; o[1] = (uint8_t)v1; }
            }
        } else {
#pragma unroll
            for (int ry = 0; ry < 2; ++ry)
                *(uint32_t *)(raw + (y + ry) * tw + x) = (uint32_t)(uint16_t)(acc[ry][0] >> 6) | ((uint32_t)(uint16_t)(acc[ry][1] >> 6) << 16);
        }
    }
}

__device__ __forceinline__ int clampi(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

// uni (rp1 == nullptr) or bi prediction of one tile of one plane
template <int TAPS, bool HBD>
__device__ void predict_tile(const uint8_t *rp0, int s0, int f0x, int f0y, const uint8_t *rp1, int s1, int f1x, int f1y, int tw, int th,
                             uint8_t *out, int ostride, uint8_t *s_win, int16_t *s_tmp, int16_t *s_raw0, int16_t *s_raw1)
{
    if (HBD && (f0x | f0y) == 0 && (!rp1 || (f1x | f1y) == 0)) {
        // uniPredCopy16bit (C_DEFAULT/EbMcp_C.c:3827) / BiPredClippingOnTheFly16bit (Codec/EbMcp.c:847): (p0 + p1 + 1) >> 1
        for (int i = threadIdx.x; i < th * (tw >> 1); i += MC_NT) {
            const int y = i / (tw >> 1), x = (i - y * (tw >> 1)) * 2;
            const uint16_t *a = (const uint16_t *)rp0 + (size_t)y * s0 + x;
            uint32_t v = (uint32_t)a[0] | ((uint32_t)a[1] << 16);
            if (rp1) { const uint16_t *b = (const uint16_t *)rp1 + (size_t)y * s1 + x; v = __vavgu2(v, (uint32_t)b[0] | ((uint32_t)b[1] << 16)); }
            uint16_t *o = (uint16_t *)out + (size_t)y * ostride + x;
            o[0] = (uint16_t)v; o[1] = (uint16_t)(v >> 16);
        }
        return;
    }
    if (!HBD && (f0x | f0y) == 0 && (!rp1 || (f1x | f1y) == 0)) {
        // integer motion vectors (every layer-1..3 picture of M10 runs integer-pel ME): LumaInterpolationCopy (C_DEFAULT/EbMcp_C.c:221) /
        // BiPredClippingOnTheFly (Codec/EbMcp.c:603): ((p0 << 6) + (p1 << 6) + 64) >> 7 == (p0 + p1 + 1) >> 1
        for (int i = threadIdx.x; i < th * (tw >> 2); i += MC_NT) {
            const int y = i / (tw >> 2), x = (i - y * (tw >> 2)) * 4;
            uint32_t v = ld_u32_unaligned(rp0 + (size_t)y * s0 + x);
            if (rp1) v = __vavgu4(v, ld_u32_unaligned(rp1 + (size_t)y * s1 + x));
            uint8_t *o = out + (size_t)y * ostride + x;
            if (((uintptr_t)o & 3) == 0) *(uint32_t *)o = v;
            else { o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16); o[3] = (uint8_t)(v >> 24); }
        }
        return;
    }
    if (!rp1) {
        interp_tile<TAPS, HBD>(rp0, s0, f0x, f0y, tw, th, s_win, s_tmp, out, ostride, nullptr);
        return;
    }
    interp_tile<TAPS, HBD>(rp0, s0, f0x, f0y, tw, th, s_win, s_tmp, nullptr, 0, s_raw0);
    interp_tile<TAPS, HBD>(rp1, s1, f1x, f1y, tw, th, s_win, s_tmp, nullptr, 0, s_raw1);
    __syncthreads();
    if (HBD) {   // BiPredClipping16bit, C_DEFAULT/EbMcp_C.c:3265: (l0 + l1 + 16400) >> 5 (16400 = 2 * 8192 bias + rounding)
        for (int i = threadIdx.x; i < th * (tw >> 1); i += MC_NT) {
            const int y = i / (tw >> 1), x = (i - y * (tw >> 1)) * 2;
            const uint32_t a = *(const uint32_t *)(s_raw0 + y * tw + x), b = *(const uint32_t *)(s_raw1 + y * tw + x);
            uint16_t *o = (uint16_t *)out + (size_t)y * ostride + x;
            o[0] = (uint16_t)min(max(((int)(short)(a & 0xffff) + (int)(short)(b & 0xffff) + 16400) >> 5, 0), 1023);
            o[1] = (uint16_t)min(max((((int)a >> 16) + ((int)b >> 16) + 16400) >> 5, 0), 1023);
        }
        return;
    }
    // BiPredClipping, C_DEFAULT/EbMcp_C.c:21 (the -8192 biases of the reference's raw buffers cancel against Offset5)
    for (int i = threadIdx.x; i < th * (tw >> 1); i += MC_NT) {
        const int y = i / (tw >> 1), x = (i - y * (tw >> 1)) * 2;
        const uint32_t a = *(const uint32_t *)(s_raw0 + y * tw + x), b = *(const uint32_t *)(s_raw1 + y * tw + x);
        const int v0 = min(max(((int)(short)(a & 0xffff) + (int)(short)(b & 0xffff) + 64) >> 7, 0), 255);
        const int v1 = min(max((((int)a >> 16) + ((int)b >> 16) + 64) >> 7, 0), 255);
        uint8_t *o = out + (size_t)y * ostride + x;
        o[0] = (uint8_t)v0;
        o[1] = (uint8_t)v1;
    }
}

template <bool HBD>
__global__ void __launch_bounds__(MC_NT) mc_batch_kernel(const __grid_constant__ McArgs A)
{
    constexpr int BPS = HBD ? 2 : 1;
    __shared__ __align__(16) uint8_t s_win[(TILE + 7) * (HBD ? 128 : 96)];   // win_pitch(39) = 80, win_pitch(78) = 112
    __shared__ __align__(16) int16_t s_tmp[(TILE + 7) * TILE];
    __shared__ __align__(16) int16_t s_raw[2][TILE * TILE];
    const int job = blockIdx.x;
    if (job >= A.n_jobs) return;
    const svtb200_mc_job q = A.jobs[job];
    int px[2] = {0, 0}, py[2] = {0, 0};
    const YuvD *r[2] = {nullptr, nullptr};
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        if (!(q.dir == l || q.dir == 2)) continue;
        r[l] = &A.refs[min((int)q.ref_idx[l], A.n_refs - 1)];   // param space
        // absolute quarter-pel position in the padded reference (EB_U16 like the reference); a clamped position has no fraction
        px[l] = (uint16_t)clampi((r[l]->ox - 71) << 2, (r[l]->w + r[l]->ox + 7) << 2, ((q.x + r[l]->ox) << 2) + q.mv[l][0]);
        py[l] = (uint16_t)clampi((r[l]->oy - 71) << 2, (r[l]->h + r[l]->oy + 7) << 2, ((q.y + r[l]->oy) << 2) + q.mv[l][1]);
    }
  // one CTA per PU; PUs larger than 32x32 are walked tile by tile (an early-exit CTA per absent tile cost more than the copy itself)
  for (int ty0 = 0; ty0 < q.h; ty0 += TILE)
   for (int tx0 = 0; tx0 < q.w; tx0 += TILE) {
    const int tw = min(TILE, q.w - tx0), th = min(TILE, q.h - ty0);
    const bool bi = q.dir == 2;
    const int l0 = bi ? 0 : q.dir, l1 = 1;
    // luma
    {
        uint8_t *dy = A.dst[0] + ((size_t)(A.doy + q.y + ty0) * A.dstride[0] + A.dox + q.x + tx0) * BPS;
        const uint8_t *rp0 = r[l0]->p[0] + ((size_t)((py[l0] >> 2) + ty0) * r[l0]->stride[0] + (px[l0] >> 2) + tx0) * BPS;
        const uint8_t *rp1 = bi ? r[l1]->p[0] + ((size_t)((py[l1] >> 2) + ty0) * r[l1]->stride[0] + (px[l1] >> 2) + tx0) * BPS : nullptr;
        predict_tile<8, HBD>(rp0, r[l0]->stride[0], px[l0] & 3, py[l0] & 3, rp1, bi ? r[l1]->stride[0] : 0, px[l1] & 3, py[l1] & 3, tw, th, dy, A.dstride[0],
                        s_win, s_tmp, s_raw[0], s_raw[1]);
    }
    // chroma (Cb then Cr): half-size tile, eighth-pel phases
    const int cw = tw >> 1, ch = th >> 1;
    // chroma destination index exactly as the reference forms it (EbInterPrediction.c:783)
    const size_t cidx = (((size_t)(A.doy + q.y) * A.dstride[1]) >> 1) + ((A.dox + q.x) >> 1) + (size_t)(ty0 >> 1) * A.dstride[1] + (tx0 >> 1);
#pragma unroll 1
    for (int pl = 1; pl < 3; ++pl) {
        const uint8_t *rp0 = r[l0]->p[pl] + ((size_t)((py[l0] >> 3) + (ty0 >> 1)) * r[l0]->stride[pl] + (px[l0] >> 3) + (tx0 >> 1)) * BPS;
        const uint8_t *rp1 = bi ? r[l1]->p[pl] + ((size_t)((py[l1] >> 3) + (ty0 >> 1)) * r[l1]->stride[pl] + (px[l1] >> 3) + (tx0 >> 1)) * BPS : nullptr;
        predict_tile<4, HBD>(rp0, r[l0]->stride[pl], px[l0] & 7, py[l0] & 7, rp1, bi ? r[l1]->stride[pl] : 0, px[l1] & 7, py[l1] & 7, cw, ch, A.dst[pl] + cidx * BPS,
                        A.dstride[pl], s_win, s_tmp, s_raw[0], s_raw[1]);
    }
    __syncthreads();
   }
}

// leaf: one block (w x h, multiples of 4 x 2) through the same tile routine
template <int TAPS>
__global__ void __launch_bounds__(MC_NT) mc_leaf_kernel(const uint8_t *r0, int s0, int f0x, int f0y, const uint8_t *r1, int s1, int f1x, int f1y,
                                                        uint8_t *dst, int ds, int w, int h)
{
    __shared__ __align__(16) uint8_t s_win[(TILE + 7) * 96];   // win_pitch(39) = 80
    __shared__ __align__(16) int16_t s_tmp[(TILE + 7) * TILE];
    __shared__ __align__(16) int16_t s_raw[2][TILE * TILE];
    for (int ty0 = 0; ty0 < h; ty0 += TILE)
        for (int tx0 = 0; tx0 < w; tx0 += TILE) {
            const int tw = min(TILE, w - tx0), th = min(TILE, h - ty0);
            predict_tile<TAPS, false>(r0 + (size_t)ty0 * s0 + tx0, s0, f0x, f0y, r1 ? r1 + (size_t)ty0 * s1 + tx0 : nullptr, s1, f1x, f1y, tw, th,
                               dst + (size_t)ty0 * ds + tx0, ds, s_win, s_tmp, s_raw[0], s_raw[1]);
        }
}

static YuvD to_yuvd(const svtb200_yuv_picture &p, const uint8_t *y, const uint8_t *cb, const uint8_t *cr)
{
    YuvD d;
    d.p[0] = y; d.p[1] = cb; d.p[2] = cr;
    d.stride[0] = (int)p.y.stride; d.stride[1] = (int)p.cb.stride; d.stride[2] = (int)p.cr.stride;
    d.ox = p.y.origin_x; d.oy = p.y.origin_y; d.w = p.y.width; d.h = p.y.height;
    return d;
}

static int check_pic(const svtb200_yuv_picture &p, const char *what)
{
    if (!p.y.buf || !p.cb.buf || !p.cr.buf || p.y.width < 8 || p.y.height < 8 || (p.y.origin_x & 1) || (p.y.origin_y & 1) ||
        p.cb.origin_x * 2 != p.y.origin_x || p.cb.origin_y * 2 != p.y.origin_y) {
        set_error("mc: %s picture malformed (null plane, or chroma padding != luma padding / 2)", what);
        return SVTB200_ERR_BAD_ARG;
    }
    return SVTB200_OK;
}

static int launch_mc(const YuvD *h_refs, int n_refs, const svtb200_mc_job *d_jobs, int n_jobs, const svtb200_yuv_picture &dst,
                     uint8_t *dy, uint8_t *dcb, uint8_t *dcr, cudaStream_t st, bool hbd = false)
{
    McArgs A;
    for (int i = 0; i < 16; ++i) A.refs[i] = h_refs[i < n_refs ? i : 0];
    A.n_refs = n_refs; A.jobs = d_jobs; A.n_jobs = n_jobs;
    A.dst[0] = dy; A.dst[1] = dcb; A.dst[2] = dcr;
    A.dstride[0] = (int)dst.y.stride; A.dstride[1] = (int)dst.cb.stride; A.dstride[2] = (int)dst.cr.stride;
    A.dox = dst.y.origin_x; A.doy = dst.y.origin_y;
    if (n_jobs == 0) return SVTB200_OK;
    if (hbd) mc_batch_kernel<true><<<n_jobs, MC_NT, 0, st>>>(A); else mc_batch_kernel<false><<<n_jobs, MC_NT, 0, st>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static size_t pbytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }

}  // namespace svtb200
using namespace svtb200;

extern "C" {

static int mc_batch_dev_impl(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs_dev, int n_jobs, const svtb200_yuv_picture *dst, void *cuda_stream, bool hbd)
{
    if (!refs || n_refs <= 0 || n_refs > 16 || (!jobs_dev && n_jobs) || n_jobs < 0 || !dst) { set_error("mc_batch_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = check_pic(*dst, "destination");
    for (int i = 0; i < n_refs && !rc; ++i) rc = check_pic(refs[i], "reference");
    if (rc) return rc;
    if ((rc = ensure_device())) return rc;
    YuvD tab[16];
    for (int i = 0; i < n_refs; ++i) tab[i] = to_yuvd(refs[i], refs[i].y.buf, refs[i].cb.buf, refs[i].cr.buf);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    return launch_mc(tab, n_refs, jobs_dev, n_jobs, *dst, dst->y.buf, dst->cb.buf, dst->cr.buf, st, hbd);
}

static int mc_batch_impl(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst, bool hbd)
{
    const size_t bps = hbd ? 2 : 1;
    if (!refs || n_refs <= 0 || n_refs > 16 || (!jobs && n_jobs) || n_jobs < 0 || !dst) { set_error("mc_batch: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = check_pic(*dst, "destination");
    for (int i = 0; i < n_refs && !rc; ++i) rc = check_pic(refs[i], "reference");
    if (rc) return rc;
    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_mc_job &q = jobs[j];
        const bool ok = (q.w == 8 || q.w == 16 || q.w == 32 || q.w == 64) && (q.h == 8 || q.h == 16 || q.h == 32 || q.h == 64) && q.dir <= 2 &&
                        q.x + q.w <= dst->y.width + 63 && q.y + q.h <= dst->y.height + 63 &&
                        (q.dir == 1 || q.ref_idx[0] < n_refs) && (q.dir == 0 || q.ref_idx[1] < n_refs);
        if (!ok) { set_error("mc_batch: job %d malformed", j); return SVTB200_ERR_BAD_ARG; }
    }
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    size_t need = 4096 + sizeof(YuvD) * n_refs + sizeof(svtb200_mc_job) * (size_t)n_jobs;
    for (int i = 0; i < n_refs; ++i) need += align_up(bps * pbytes(refs[i].y), 256) + align_up(bps * pbytes(refs[i].cb), 256) + align_up(bps * pbytes(refs[i].cr), 256);
    need += align_up(bps * pbytes(dst->y), 256) + align_up(bps * pbytes(dst->cb), 256) + align_up(bps * pbytes(dst->cr), 256) + 2048;
    if ((rc = c->dev_reserve(need)) || (rc = c->pin_reserve(sizeof(YuvD) * n_refs + 256))) return rc;
    void *h_tab, *d_tab, *d_jobs;
    if ((rc = c->pin_alloc(sizeof(YuvD) * n_refs, &h_tab)) || (rc = c->dev_alloc(sizeof(YuvD) * n_refs, &d_tab)) ||
        (rc = c->dev_alloc(sizeof(svtb200_mc_job) * (size_t)std::max(n_jobs, 1), &d_jobs))) return rc;
    for (int i = 0; i < n_refs; ++i) {
        void *p[3];
        const svtb200_plane *pl[3] = {&refs[i].y, &refs[i].cb, &refs[i].cr};
        for (int k = 0; k < 3; ++k) {
            if ((rc = c->dev_alloc(bps * pbytes(*pl[k]), &p[k]))) return rc;
            SVTB200_CUDA_TRY(cudaMemcpyAsync(p[k], pl[k]->buf, bps * pbytes(*pl[k]), cudaMemcpyHostToDevice, c->stream));
        }
        ((YuvD *)h_tab)[i] = to_yuvd(refs[i], (uint8_t *)p[0], (uint8_t *)p[1], (uint8_t *)p[2]);
    }
    void *dp[3];
    const svtb200_plane *dpl[3] = {&dst->y, &dst->cb, &dst->cr};
    for (int k = 0; k < 3; ++k) {
        if ((rc = c->dev_alloc(bps * pbytes(*dpl[k]), &dp[k]))) return rc;
        // the destination is updated in place (only PU areas are written): bring the current content over first
        SVTB200_CUDA_TRY(cudaMemcpyAsync(dp[k], dpl[k]->buf, bps * pbytes(*dpl[k]), cudaMemcpyHostToDevice, c->stream));
    }
    if (n_jobs) SVTB200_CUDA_TRY(cudaMemcpyAsync(d_jobs, jobs, sizeof(svtb200_mc_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_mc((const YuvD *)h_tab, n_refs, (const svtb200_mc_job *)d_jobs, n_jobs, *dst, (uint8_t *)dp[0], (uint8_t *)dp[1], (uint8_t *)dp[2], c->stream, hbd))) return rc;
    for (int k = 0; k < 3; ++k) SVTB200_CUDA_TRY(cudaMemcpyAsync(dpl[k]->buf, dp[k], bps * pbytes(*dpl[k]), cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_mc_batch_dev(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs_dev, int n_jobs, const svtb200_yuv_picture *dst, void *cuda_stream)
{ return mc_batch_dev_impl(refs, n_refs, jobs_dev, n_jobs, dst, cuda_stream, false); }
int svtb200_mc_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst)
{ return mc_batch_impl(refs, n_refs, jobs, n_jobs, dst, false); }
int svtb200_mc_batch16_dev(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs_dev, int n_jobs, const svtb200_yuv_picture *dst, void *cuda_stream)
{ return mc_batch_dev_impl(refs, n_refs, jobs_dev, n_jobs, dst, cuda_stream, true); }
int svtb200_mc_batch16(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst)
{ return mc_batch_impl(refs, n_refs, jobs, n_jobs, dst, true); }

// ---- leaf entry points: a single block through the same kernel ----------------------------------------------
static int mc_leaf(int taps, uint32_t f0x, uint32_t f0y, const uint8_t *ref0, uint32_t r0s, bool bi, uint32_t f1x, uint32_t f1y, const uint8_t *ref1,
                   uint32_t r1s, uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h)
{
    const uint32_t maxf = taps == 8 ? 3 : 7;
    if (!ref0 || (bi && !ref1) || !dst || w == 0 || h == 0 || w > 64 || h > 64 || (w & 3) || (h & 1) || f0x > maxf || f0y > maxf || f1x > maxf || f1y > maxf) { set_error("mc leaf: bad argument (w multiple of 4, h even, <= 64)"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    const int half = taps / 2 - 1, ext = taps - 1;
    const size_t b0 = (size_t)(h + ext - 1) * r0s + w + ext, b1 = bi ? (size_t)(h + ext - 1) * r1s + w + ext : 0, bd = (size_t)(h - 1) * ds + w;
    void *d0, *d1 = nullptr, *dd, *hd;
    if ((rc = c->dev_reserve(b0 + b1 + bd + 4096)) || (rc = c->dev_alloc(b0 + 8, &d0)) || (rc = c->dev_alloc(b1 + 8, &d1)) || (rc = c->dev_alloc(bd, &dd)) || (rc = c->pin_alloc(bd, &hd))) return rc;
    const uint8_t *h0 = ref0 - (size_t)half * r0s - half, *h1 = bi ? ref1 - (size_t)half * r1s - half : nullptr;
    const size_t m0 = (uintptr_t)h0 & 3, m1 = bi ? ((uintptr_t)h1 & 3) : 0;   // keep the host misalignment: exercises every window phase
    SVTB200_CUDA_TRY(cudaMemcpyAsync((uint8_t *)d0 + m0, h0, b0, cudaMemcpyHostToDevice, c->stream));
    if (bi) SVTB200_CUDA_TRY(cudaMemcpyAsync((uint8_t *)d1 + m1, h1, b1, cudaMemcpyHostToDevice, c->stream));
    const uint8_t *p0 = (uint8_t *)d0 + m0 + (size_t)half * r0s + half, *p1 = bi ? (uint8_t *)d1 + m1 + (size_t)half * r1s + half : nullptr;
    if (taps == 8) mc_leaf_kernel<8><<<1, MC_NT, 0, c->stream>>>(p0, (int)r0s, (int)f0x, (int)f0y, p1, (int)r1s, (int)f1x, (int)f1y, (uint8_t *)dd, (int)ds, (int)w, (int)h);
    else mc_leaf_kernel<4><<<1, MC_NT, 0, c->stream>>>(p0, (int)r0s, (int)f0x, (int)f0y, p1, (int)r1s, (int)f1x, (int)f1y, (uint8_t *)dd, (int)ds, (int)w, (int)h);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    SVTB200_CUDA_TRY(cudaMemcpyAsync(hd, dd, bd, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (uint32_t y = 0; y < h; ++y) memcpy(dst + (size_t)y * ds, (uint8_t *)hd + (size_t)y * ds, w);
    return SVTB200_OK;
}

int svtb200_LumaInterpolationFilterNew(uint32_t frac, const uint8_t *refPic, uint32_t srcStride, uint8_t *dst, uint32_t dstStride, uint32_t puWidth, uint32_t puHeight)
{
    return mc_leaf(8, frac & 3, frac >> 2, refPic, srcStride, false, 0, 0, nullptr, 0, dst, dstStride, puWidth, puHeight);
}
int svtb200_ChromaInterpolationFilterNew(uint32_t fracPosx, uint32_t fracPosy, const uint8_t *refPic, uint32_t srcStride, uint8_t *dst, uint32_t dstStride,
                                         uint32_t puWidth, uint32_t puHeight)
{
    return mc_leaf(4, fracPosx, fracPosy, refPic, srcStride, false, 0, 0, nullptr, 0, dst, dstStride, puWidth, puHeight);
}
int svtb200_LumaBiPredInterpolation(uint32_t frac0, const uint8_t *ref0, uint32_t ref0Stride, uint32_t frac1, const uint8_t *ref1, uint32_t ref1Stride,
                                    uint8_t *dst, uint32_t dstStride, uint32_t puWidth, uint32_t puHeight)
{
    return mc_leaf(8, frac0 & 3, frac0 >> 2, ref0, ref0Stride, true, frac1 & 3, frac1 >> 2, ref1, ref1Stride, dst, dstStride, puWidth, puHeight);
}
int svtb200_ChromaBiPredInterpolation(uint32_t frac0x, uint32_t frac0y, const uint8_t *ref0, uint32_t ref0Stride, uint32_t frac1x, uint32_t frac1y,
                                      const uint8_t *ref1, uint32_t ref1Stride, uint8_t *dst, uint32_t dstStride, uint32_t puWidth, uint32_t puHeight)
{
    return mc_leaf(4, frac0x, frac0y, ref0, ref0Stride, true, frac1x, frac1y, ref1, ref1Stride, dst, dstStride, puWidth, puHeight);
}

}  // extern "C"
