// mc.cu -- HEVC motion-compensated prediction (luma 8-tap quarter-pel, chroma 4-tap eighth-pel, uni and bi).
//
// Replaces:
//   LumaInterpolationCopy / LumaInterpolationFilterPos{a..r}New, ...OutRaw, ChromaInterpolation*   C_DEFAULT/EbMcp_C.c:221-1214,3789-5099
//     and their SSSE3 rows (ASM_SSSE3/EbMcp_Intrinsic_SSSE3.c), tables Codec/EbMcpTables.c:14-745
//   BiPredClipping (C_DEFAULT/EbMcp_C.c:21), BiPredClippingOnTheFly (integer/integer shortcut, Codec/EbMcp.c:603)
//   EncodeUniPredInterpolation / EncodeBiPredInterpolation (Codec/EbMcp.c:175,562), EncodePassInterPrediction
//   position clamp (Codec/EbInterPrediction.c:794-808)
//
// Design: one CTA (64 threads; most PUs of M10 are integer-vector copies, where the per-PU address derivation every warp repeats dominates) per PU, walked in <= 32x32 luma tiles (+ the two 16x16 chroma tiles).  Per tile ALL reference windows
// (luma / Cb / Cr x list 0 / list 1, (tile + taps - 1)^2 each) are staged in shared memory back to back with coalesced 128-bit loads, then
//   H pass : every window, flattened over the CTA; 4 outputs per thread from funnel-shifted words with DP4A (u8 x s8), s16 rows in smem,
//   V pass : a 2x2 output quad per thread with DP2A on PRMT-built vertical s16 pairs, both lists in registers, final rounding /
//            BiPredClipping fused (three barriers per tile in total).
// Interpolated planes ALWAYS run both passes; integer phases use the identity taps {..,64,..}.  That is bit-exact with the reference's
// 16 (64) specialised functions because their 14-bit intermediates differ from the two-pass value only by an exact
// factor of 64 that the final shift removes (see oracle/port/mc_port.c and tests/test_gpu_mc.py, every phase pair).  Planes whose used
// lists are all at integer positions are copied / averaged straight from global memory.
// HBM traffic per CTU is the algorithmic minimum + the 7-sample halo (served by L2 between neighbouring CTAs).
#include "common.cuh"

namespace svtb200 {

#ifndef MC_NT_V
#define MC_NT_V 64
#endif
constexpr int MC_NT = MC_NT_V;
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
    // TMA path of the 8-bit batch kernel (n_refs <= 4, every plane 16-byte aligned with a 16-byte pitch): per reference a 48 x 32 box of the luma
    // plane and 48 x 19 boxes of Cb / Cr
    int tma;
    CUtensorMap maps[4][3];
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

// ---- staged three-phase tile predictor -------------------------------------------------------------------------------------
// A tile has up to six reference windows ("units"): luma L0/L1 (8 taps) and Cb/Cr x L0/L1 (4 taps).  All windows of a tile are
// fetched back to back (one exposed HBM/L2 latency), then ONE barrier, the H passes of every unit flattened over the CTA, ONE
// barrier, and the V passes: a thread produces the same 2x2 output quad of both lists and finishes uni-pred rounding or
// BiPredClipping in registers (no raw 14-bit buffers, no third pass).  Planes whose used lists all have integer phases skip the
// staging altogether (copy / rounded average straight from global memory).
// HBD: 16-bit samples (10-bit video).  Pointers are byte pointers, strides are in SAMPLES.  The two-pass form is exact there as
// well: first pass (S - 32768) >> 2 kept in 16 bits like LumaInterpolationFilterPos*OutRaw16bit, second pass (S2 + 524800) >> 10 (uni) or
// S2 >> 6 (bi); with identity taps these reduce to the reference's 1-D (S + 32) >> 6, (S - 32768) >> 2 and copy forms (Codec/EbDefinitions.h:1041-1059).
struct McUnit {
    const uint8_t *ref;   // integer sample of the tile's top-left in the reference plane (byte pointer)
    int rstride;          // samples
};
struct McClass {          // the units of one plane class (luma: np = 1; chroma: np = 2, same geometry and phases)
    McUnit u[2][2];       // [plane][list slot]; slot 0 = the only list of a uni-pred PU
    uint8_t *out[2];
    int ostride[2];
    int fx[2], fy[2];     // phases per list slot
    int np, tw, th;
    bool interp;
};

template <int TAPS, bool HBD, int NTH = MC_NT>
__device__ __forceinline__ void mc_stage_load(const McClass &c, int nl, uint8_t *s_win, int win_unit, int *s_a)
{
    constexpr int HALF = TAPS / 2 - 1, BPS = HBD ? 2 : 1;
    const int wbytes = (c.tw + TAPS - 1) * BPS, wrows = c.th + TAPS - 1, wp = win_pitch(wbytes);
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            if (p >= c.np || l >= nl) continue;
            const int a = load_window_t<NTH>(s_win + (p * 2 + l) * win_unit, wp, c.u[p][l].ref - (HALF * c.u[p][l].rstride + HALF) * BPS,
                                               c.u[p][l].rstride * BPS, wbytes, wrows);
            if (threadIdx.x == 0) s_a[p * 2 + l] = a;
        }
}

// H pass of every unit of the class: s_tmp[unit][r][x], r in [0, wrows), pitch tw (s16)
template <int TAPS, bool HBD, int NTH = MC_NT>
__device__ __forceinline__ void mc_stage_h(const McClass &c, int nl, const uint8_t *s_win, int win_unit, const int *s_a, int16_t *s_tmp, int tmp_unit)
{
    const int tid = NTH == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;      // NTH == 32: the warp works alone
    constexpr int BPS = HBD ? 2 : 1;
    const int wbytes = (c.tw + TAPS - 1) * BPS, wrows = c.th + TAPS - 1, wp = win_pitch(wbytes), tw = c.tw;
    const int ipr = HBD ? tw >> 1 : tw >> 2, per_unit = wrows * ipr, total = c.np * nl * per_unit;
    const SmallDiv d_ipr = mk_div(ipr);
    uint32_t kx0[2], kx1[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const int8_t *tx = TAPS == 8 ? c_luma_taps[c.fx[l] & 3] : c_chroma_taps[c.fx[l] & 7];
        kx0[l] = pack4(tx); kx1[l] = TAPS == 8 ? pack4(tx + 4) : 0u;
    }
    for (int i = tid; i < total; i += NTH) {
        const int rr = udiv(i, d_ipr), g = i - rr * ipr;                                                  // row over all units, item in the row
        const int uu = (rr >= wrows) + (rr >= 2 * wrows) + (rr >= 3 * wrows), r = rr - uu * wrows;     // <= 4 units per class
        const int l = nl == 2 ? (uu & 1) : 0, slot = (nl == 2 ? (uu >> 1) : uu) * 2 + l;
        const int a = s_a[slot];
        const uint8_t *rowb = s_win + slot * win_unit + r * wp;
        int16_t *trow = s_tmp + slot * tmp_unit + r * tw;
        const uint32_t cx0 = l ? kx0[1] : kx0[0], cx1 = l ? kx1[1] : kx1[0];
        if (HBD) {
            // two adjacent outputs per thread from TAPS/2 + 1 words of s16 pairs; DP2A on (sample, sample) x (tap, tap)
            const int x0 = g * 2, c0 = (a >> 1) + x0, par = c0 & 1;
            const uint32_t *row = (const uint32_t *)rowb + (c0 >> 1);
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
            *(uint32_t *)(trow + x0) = (uint32_t)(uint16_t)((o0 - 32768) >> 2) | ((uint32_t)(uint16_t)((o1 - 32768) >> 2) << 16);
        } else {
            // four outputs per thread from funnel-shifted words with DP4A (u8 x s8)
            const int x0 = g * 4, c0 = a + x0, wi = c0 >> 2, sh = (c0 & 3) * 8;
            const uint32_t *row = (const uint32_t *)rowb;
            uint32_t w[TAPS == 8 ? 4 : 3];
#pragma unroll
            for (int k = 0; k < (TAPS == 8 ? 4 : 3); ++k) w[k] = row[wi + k];
            int o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int s0 = sh + 8 * k;                       // bit offset of byte c0 + k inside (w[0], w[1])
                const uint32_t lo = s0 < 32 ? __funnelshift_r(w[0], w[1], s0) : __funnelshift_r(w[1], w[2], s0 - 32);
                int acc = dp4a_us(lo, cx0, 0);
                if (TAPS == 8) {
                    const int s1 = s0 + 32;
                    const uint32_t hi = s1 < 64 ? __funnelshift_r(w[1], w[2], s1 - 32) : __funnelshift_r(w[2], w[3], s1 - 64);
                    acc = dp4a_us(hi, cx1, acc);
                }
                o[k] = acc;
            }
            uint32_t *t = (uint32_t *)(trow + x0);
            t[0] = (uint32_t)(uint16_t)o[0] | ((uint32_t)(uint16_t)o[1] << 16);
            t[1] = (uint32_t)(uint16_t)o[2] | ((uint32_t)(uint16_t)o[3] << 16);
        }
    }
}

// V pass: a 2x2 output quad per thread with DP2A on PRMT-built vertical s16 pairs, both lists, then the final rounding:
//   uni 8-bit (S2 + 2048) >> 12, bi 8-bit BiPredClipping (C_DEFAULT/EbMcp_C.c:21; the -8192 biases of the reference's raw buffers cancel
//   against Offset5): ((S2_0 >> 6) + (S2_1 >> 6) + 64) >> 7;  16-bit: (S2 + 524800) >> 10 / BiPredClipping16bit :3265 (l0 + l1 + 16400) >> 5
template <int TAPS, bool HBD, int NTH = MC_NT>
__device__ __forceinline__ void mc_stage_v(const McClass &c, int nl, const int16_t *s_tmp, int tmp_unit)
{
    const int tid = NTH == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    const int tw = c.tw, qpr = tw >> 1, per_plane = (c.th >> 1) * qpr, total = c.np * per_plane, p4 = tw >> 1;
    const SmallDiv d_qpr = mk_div(qpr);
    uint32_t ky0[2], ky1[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const int8_t *ty = TAPS == 8 ? c_luma_taps[c.fy[l] & 3] : c_chroma_taps[c.fy[l] & 7];
        ky0[l] = pack4(ty); ky1[l] = TAPS == 8 ? pack4(ty + 4) : 0u;
    }
    for (int i = tid; i < total; i += NTH) {
        const int p = i >= per_plane ? 1 : 0, j = i - p * per_plane;
        const int qy = udiv(j, d_qpr), x = (j - qy * qpr) * 2, y = qy * 2;
        int acc[2][2][2];   // [list][row][col]
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            if (l >= nl) { acc[l][0][0] = acc[l][0][1] = acc[l][1][0] = acc[l][1][1] = 0; continue; }
            const uint32_t *col = (const uint32_t *)(s_tmp + (p * 2 + l) * tmp_unit + y * tw + x);
            uint32_t w[TAPS + 1];
#pragma unroll
            for (int k = 0; k < TAPS + 1; ++k) w[k] = col[k * p4];
#pragma unroll
            for (int ry = 0; ry < 2; ++ry) {
                int a0 = 0, a1 = 0;
#pragma unroll
                for (int k = 0; k < TAPS; k += 2) {
                    const uint32_t lo = __byte_perm(w[ry + k], w[ry + k + 1], 0x5410);   // (x, row k), (x, row k+1)
                    const uint32_t hi = __byte_perm(w[ry + k], w[ry + k + 1], 0x7632);   // (x+1, ..)
                    const uint32_t cc = k < 4 ? ky0[l] : ky1[l];
                    if ((k & 2) == 0) { a0 = dp2a_lo_ss(lo, cc, a0); a1 = dp2a_lo_ss(hi, cc, a1); }
                    else { a0 = dp2a_hi_ss(lo, cc, a0); a1 = dp2a_hi_ss(hi, cc, a1); }
                }
                acc[l][ry][0] = a0; acc[l][ry][1] = a1;
            }
        }
        uint8_t *outp = p ? c.out[1] : c.out[0];
        const int ostride = p ? c.ostride[1] : c.ostride[0];
#pragma unroll
        for (int ry = 0; ry < 2; ++ry) {
            int v0, v1;
            if (nl == 2) {
                const int s0 = (int)(short)(acc[0][ry][0] >> 6) + (int)(short)(acc[1][ry][0] >> 6), s1 = (int)(short)(acc[0][ry][1] >> 6) + (int)(short)(acc[1][ry][1] >> 6);
                if (HBD) { v0 = min(max((s0 + 16400) >> 5, 0), 1023); v1 = min(max((s1 + 16400) >> 5, 0), 1023); }
                else { v0 = min(max((s0 + 64) >> 7, 0), 255); v1 = min(max((s1 + 64) >> 7, 0), 255); }
            } else if (HBD) {
                v0 = min(max((acc[0][ry][0] + 524800) >> 10, 0), 1023); v1 = min(max((acc[0][ry][1] + 524800) >> 10, 0), 1023);
            } else {
                v0 = min(max((acc[0][ry][0] + 2048) >> 12, 0), 255); v1 = min(max((acc[0][ry][1] + 2048) >> 12, 0), 255);
            }
            if (HBD) {
                uint16_t *o = (uint16_t *)outp + (size_t)(y + ry) * ostride + x;
                if (((uintptr_t)o & 3) == 0) *(uint32_t *)o = (uint32_t)v0 | ((uint32_t)v1 << 16);
                else { o[0] = (uint16_t)v0; o[1] = (uint16_t)v1; }
            } else {
                uint8_t *o = outp + (size_t)(y + ry) * ostride + x;
                if (((uintptr_t)o & 1) == 0) *(uint16_t *)o = (uint16_t)(v0 | (v1 << 8));
                else { o[0] = (uint8_t)v0; o[1] = (uint8_t)v1; }
            }
        }
    }
}

// Integer motion vectors on every used list of the class (every layer-1..3 picture of M10 runs integer-pel ME): LumaInterpolationCopy
// (C_DEFAULT/EbMcp_C.c:221) / BiPredClippingOnTheFly (Codec/EbMcp.c:603): ((p0 << 6) + (p1 << 6) + 64) >> 7 == (p0 + p1 + 1) >> 1;
// 16-bit: uniPredCopy16bit (C_DEFAULT/EbMcp_C.c:3827) / BiPredClippingOnTheFly16bit (Codec/EbMcp.c:847).  No shared memory, no barrier.
template <bool HBD, int NTH = MC_NT>
__device__ __forceinline__ void mc_copy_scalar(const McClass &c, int nl)
{
    const int tid = NTH == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    const int tw = c.tw, ipr = HBD ? tw >> 1 : tw >> 2, per_plane = c.th * ipr, total = c.np * per_plane;
    const SmallDiv d_ipr = mk_div(ipr);
    for (int i = tid; i < total; i += NTH) {
        const int p = i >= per_plane ? 1 : 0, j = i - p * per_plane;
        const int y = udiv(j, d_ipr), g = j - y * ipr;
        const uint8_t *r0 = p ? c.u[1][0].ref : c.u[0][0].ref, *r1 = p ? c.u[1][1].ref : c.u[0][1].ref;
        const int s0 = p ? c.u[1][0].rstride : c.u[0][0].rstride, s1 = p ? c.u[1][1].rstride : c.u[0][1].rstride;
        uint8_t *outp = p ? c.out[1] : c.out[0];
        const int ostride = p ? c.ostride[1] : c.ostride[0];
        if (HBD) {
            const int x = g * 2;
            const uint16_t *a = (const uint16_t *)r0 + (size_t)y * s0 + x;
            uint32_t v = (uint32_t)a[0] | ((uint32_t)a[1] << 16);
            if (nl == 2) { const uint16_t *b = (const uint16_t *)r1 + (size_t)y * s1 + x; v = __vavgu2(v, (uint32_t)b[0] | ((uint32_t)b[1] << 16)); }
            uint16_t *o = (uint16_t *)outp + (size_t)y * ostride + x;
            o[0] = (uint16_t)v; o[1] = (uint16_t)(v >> 16);
        } else {
            const int x = g * 4;
            uint32_t v = ld_u32_unaligned(r0 + (size_t)y * s0 + x);
            if (nl == 2) v = __vavgu4(v, ld_u32_unaligned(r1 + (size_t)y * s1 + x));
            uint8_t *o = outp + (size_t)y * ostride + x;
            if (((uintptr_t)o & 3) == 0) *(uint32_t *)o = v;
            else { o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16); o[3] = (uint8_t)(v >> 24); }
        }
    }
}

// VW consecutive 32-bit words starting at an arbitrary byte address: two aligned 8 / 16-byte loads, a word-granular select chain and VW funnel shifts
template <int VW>
__device__ __forceinline__ void ld_bytes_vec(const uint8_t *p, uint32_t (&o)[VW])
{
    const uintptr_t a = (uintptr_t)p;
    const int off = (int)(a & (4 * VW - 1));
    uint32_t w[2 * VW];
    if (VW == 4) {
        const uint4 *b = (const uint4 *)(a - off);
        const uint4 A = __ldg(b), B = __ldg(b + 1);
        w[0] = A.x; w[1] = A.y; w[2] = A.z; w[3] = A.w; w[4] = B.x; w[5] = B.y; w[6] = B.z; w[7] = B.w;
    } else {
        const uint2 *b = (const uint2 *)(a - off);
        const uint2 A = __ldg(b), B = __ldg(b + 1);
        w[0] = A.x; w[1] = A.y; w[2] = B.x; w[3] = B.y;
    }
    const int q = off >> 2;
    const uint32_t sh = (uint32_t)(off & 3) * 8;
    if (VW == 4 && (q & 2)) {
#pragma unroll
        for (int k = 0; k < 2 * VW - 2; ++k) w[k] = w[k + 2];
    }
    if (q & 1) {
#pragma unroll
        for (int k = 0; k < 2 * VW - 1; ++k) w[k] = w[k + 1];
    }
#pragma unroll
    for (int k = 0; k < VW; ++k) o[k] = __funnelshift_r(w[k], w[k + 1], sh);
}

// the same copy / rounded average with 8- or 16-byte items (row bytes, destination pointers and pitches multiples of 4 * VW)
template <bool HBD, int VW, int NTH = MC_NT>
__device__ __forceinline__ void mc_copy_vec(const McClass &c, int nl)
{
    const int tid = NTH == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    constexpr int BPS = HBD ? 2 : 1, VB = 4 * VW;
    const int ipr = (c.tw * BPS) / VB, lg = 31 - __clz(ipr), per_plane = c.th * ipr, total = c.np * per_plane;   // ipr is a power of two here
#pragma unroll 2
    for (int i = tid; i < total; i += NTH) {
        const int p = i >= per_plane ? 1 : 0, j = i - p * per_plane;
        const int y = j >> lg, g = j & (ipr - 1);
        const uint8_t *r0 = p ? c.u[1][0].ref : c.u[0][0].ref, *r1 = p ? c.u[1][1].ref : c.u[0][1].ref;
        const int s0 = p ? c.u[1][0].rstride : c.u[0][0].rstride, s1 = p ? c.u[1][1].rstride : c.u[0][1].rstride;
        uint8_t *outp = p ? c.out[1] : c.out[0];
        const int ostride = p ? c.ostride[1] : c.ostride[0];
        uint32_t v[VW];
        ld_bytes_vec<VW>(r0 + (size_t)y * s0 * BPS + g * VB, v);
        if (nl == 2) {
            uint32_t u[VW];
            ld_bytes_vec<VW>(r1 + (size_t)y * s1 * BPS + g * VB, u);
#pragma unroll
            for (int k = 0; k < VW; ++k) v[k] = HBD ? __vavgu2(v[k], u[k]) : __vavgu4(v[k], u[k]);
        }
        uint8_t *o = outp + (size_t)y * ostride * BPS + g * VB;
        if (VW == 4) *(uint4 *)o = make_uint4(v[0], v[1], v[2], v[3]);
        else *(uint2 *)o = make_uint2(v[0], v[1]);
    }
}

template <bool HBD, int NTH = MC_NT>
__device__ __forceinline__ void mc_copy(const McClass &c, int nl)
{
    constexpr int BPS = HBD ? 2 : 1;
    const int rb = c.tw * BPS;
    uintptr_t al = (uintptr_t)c.out[0] | (uintptr_t)(c.ostride[0] * BPS) | (uintptr_t)rb;
    if (c.np == 2) al |= (uintptr_t)c.out[1] | (uintptr_t)(c.ostride[1] * BPS);
    const bool pow2 = (rb & (rb - 1)) == 0;
    if (pow2 && (al & 15) == 0) mc_copy_vec<HBD, 4, NTH>(c, nl);
    else if (pow2 && (al & 7) == 0) mc_copy_vec<HBD, 2, NTH>(c, nl);
    else mc_copy_scalar<HBD, NTH>(c, nl);
}

__device__ __forceinline__ int clampi(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

// shared-memory geometry of the staged path: window pitches win_pitch(39) = 80, win_pitch(19) = 64, 16-bit win_pitch(78) = 112, win_pitch(38) = 80
template <bool HBD> struct McSmem {
    static constexpr int WIN_L = (TILE + 7) * (HBD ? 112 : 80), WIN_C = (TILE / 2 + 3) * (HBD ? 80 : 64);
    static constexpr int TMP_L = (TILE + 7) * TILE, TMP_C = (TILE / 2 + 3) * (TILE / 2);   // s16 elements
};

template <bool HBD>
__global__ void __launch_bounds__(MC_NT, 1024 / MC_NT) mc_batch_kernel(const __grid_constant__ McArgs A)
{
    constexpr int BPS = HBD ? 2 : 1;
    using G = McSmem<HBD>;
    static_assert(4 * G::WIN_C <= 2 * G::WIN_L && 4 * G::TMP_C <= 2 * G::TMP_L, "the chroma class reuses the luma class buffers");
    __shared__ __align__(16) uint8_t s_win[2 * G::WIN_L];
    __shared__ __align__(16) int16_t s_tmp[2 * G::TMP_L];
    __shared__ int s_a[4];
    const int job = blockIdx.x;
    if (job >= A.n_jobs) return;
    const svtb200_mc_job q = A.jobs[job];
    const bool bi = q.dir == 2;
    const int nl = bi ? 2 : 1;
    // list slot s -> list: uni-pred PUs use slot 0 for their only list
    int px[2], py[2];
    const YuvD *r[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int l = bi ? s : (q.dir == 1 ? 1 : 0);
        const int ri = l ? q.ref_idx[1] : q.ref_idx[0], mvx = l ? q.mv[1][0] : q.mv[0][0], mvy = l ? q.mv[1][1] : q.mv[0][1];
        r[s] = &A.refs[min(ri, A.n_refs - 1)];   // param space
        // absolute quarter-pel position in the padded reference (EB_U16 like the reference); a clamped position has no fraction
        px[s] = (uint16_t)clampi((r[s]->ox - 71) << 2, (r[s]->w + r[s]->ox + 7) << 2, ((q.x + r[s]->ox) << 2) + mvx);
        py[s] = (uint16_t)clampi((r[s]->oy - 71) << 2, (r[s]->h + r[s]->oy + 7) << 2, ((q.y + r[s]->oy) << 2) + mvy);
    }
    // one CTA per PU; PUs larger than 32x32 are walked tile by tile (an early-exit CTA per absent tile cost more than the copy itself)
    for (int ty0 = 0; ty0 < q.h; ty0 += TILE)
        for (int tx0 = 0; tx0 < q.w; tx0 += TILE) {
            McClass L, K;   // luma, chroma
            L.np = 1; L.tw = min(TILE, q.w - tx0); L.th = min(TILE, q.h - ty0);
            K.np = 2; K.tw = L.tw >> 1; K.th = L.th >> 1;
            L.out[0] = L.out[1] = A.dst[0] + ((size_t)(A.doy + q.y + ty0) * A.dstride[0] + A.dox + q.x + tx0) * BPS;
            L.ostride[0] = L.ostride[1] = A.dstride[0];
            // chroma destination index exactly as the reference forms it (EbInterPrediction.c:783)
            const size_t cidx = (((size_t)(A.doy + q.y) * A.dstride[1]) >> 1) + ((A.dox + q.x) >> 1) + (size_t)(ty0 >> 1) * A.dstride[1] + (tx0 >> 1);
            K.out[0] = A.dst[1] + cidx * BPS; K.out[1] = A.dst[2] + cidx * BPS;
            K.ostride[0] = A.dstride[1]; K.ostride[1] = A.dstride[2];
            int lf = 0, kf = 0;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                L.fx[s] = px[s] & 3; L.fy[s] = py[s] & 3; K.fx[s] = px[s] & 7; K.fy[s] = py[s] & 7;
                if (s < nl) { lf |= L.fx[s] | L.fy[s]; kf |= K.fx[s] | K.fy[s]; }
                L.u[0][s].ref = r[s]->p[0] + ((size_t)((py[s] >> 2) + ty0) * r[s]->stride[0] + (px[s] >> 2) + tx0) * BPS;
                L.u[0][s].rstride = r[s]->stride[0];
                L.u[1][s] = L.u[0][s];
#pragma unroll
                for (int pl = 1; pl < 3; ++pl) {
                    K.u[pl - 1][s].ref = r[s]->p[pl] + ((size_t)((py[s] >> 3) + (ty0 >> 1)) * r[s]->stride[pl] + (px[s] >> 3) + (tx0 >> 1)) * BPS;
                    K.u[pl - 1][s].rstride = r[s]->stride[pl];
                }
            }
            L.interp = lf != 0; K.interp = kf != 0;
            // copies first (no shared memory, no barrier); then the interpolated classes one after the other through the SAME buffers: a PU with
            // fractional vectors pays six barriers instead of three, every PU gets 11 KB instead of 18 KB of shared memory (16 CTAs / SM)
            if (!L.interp) mc_copy<HBD>(L, nl);
            if (!K.interp) mc_copy<HBD>(K, nl);
            if (L.interp) {
                mc_stage_load<8, HBD>(L, nl, s_win, G::WIN_L, s_a);
                __syncthreads();
                mc_stage_h<8, HBD>(L, nl, s_win, G::WIN_L, s_a, s_tmp, G::TMP_L);
                __syncthreads();
                mc_stage_v<8, HBD>(L, nl, s_tmp, G::TMP_L);
                __syncthreads();   // the chroma class / the next tile reuses the buffers
            }
            if (K.interp) {
                mc_stage_load<4, HBD>(K, nl, s_win, G::WIN_C, s_a);
                __syncthreads();
                mc_stage_h<4, HBD>(K, nl, s_win, G::WIN_C, s_a, s_tmp, G::TMP_C);
                __syncthreads();
                mc_stage_v<4, HBD>(K, nl, s_tmp, G::TMP_C);
                __syncthreads();
            }
        }
}

// ---- round-2 batch kernel (8-bit): ONE WARP PER PU, eight PUs per CTA -------------------------------------------------------------------------------
// At M10 every layer-1..3 picture has integer-sample luma vectors, so a PU is a luma copy (or average) plus chroma that is half-sample on the PUs whose
// vector is odd.  That work has no block-wide step: a warp decodes its own job, issues the luma copy loads (16-byte items, no shared memory), filters its
// chroma rows straight from global memory into a per-warp s16 tile (H pass) and finishes them with the V pass of the staged path -- no block barrier, the
// whole picture in flight in about two waves of 256-thread CTAs.  PUs with fractional LUMA vectors (base-layer pictures) are left to phase 2, where the
// whole CTA runs the staged 8-tap path on them one after the other.
constexpr int MC8_NT = 256, MC8_JOBS = MC8_NT / 32;

// H pass of the chroma class without a staging step: work item = one row of one unit (plane x list): tw + 3 bytes -> tw s16 outputs (tw in {4, 8, 16}).
// SMEM = false: rows straight from global memory;  SMEM = true: rows of the TMA-fetched 48-byte-pitch tiles (tile[slot], byte misalignment am[l]).
template <bool SMEM>
__device__ __forceinline__ void mc_chroma_h_rows(const McClass &c, int nl, int16_t *s_tmp, int tmp_unit, const uint8_t *tiles, int tile_bytes, const int (&am)[2])
{
    const int lane = threadIdx.x & 31;
    const int tw = c.tw, wrows = c.th + 3, units = c.np * nl, total = units * wrows;
    const SmallDiv d_rows = mk_div(wrows);
    for (int i = lane; i < total; i += 32) {
        const int uu = udiv(i, d_rows), r = i - uu * wrows;
        const int l = nl == 2 ? (uu & 1) : 0, p = nl == 2 ? (uu >> 1) : uu, slot = p * 2 + l;
        const uint32_t k0 = pack4(c_chroma_taps[(l ? c.fx[1] : c.fx[0]) & 7]);
        uint32_t raw[6], sh;
        if (SMEM) {
            const int a = l ? am[1] : am[0];
            const uint32_t *w = (const uint32_t *)(tiles + slot * tile_bytes + r * 48) + (a >> 2);
            sh = (uint32_t)(a & 3) * 8;
#pragma unroll
            for (int k = 0; k < 6; ++k) raw[k] = (4 * k < tw + 3 + 3) ? w[k] : 0u;
        } else {
            const uint8_t *uref = p ? (l ? c.u[1][1].ref : c.u[1][0].ref) : (l ? c.u[0][1].ref : c.u[0][0].ref);
            const int ustride = p ? (l ? c.u[1][1].rstride : c.u[1][0].rstride) : (l ? c.u[0][1].rstride : c.u[0][0].rstride);
            const uintptr_t ga = (uintptr_t)(uref + (ptrdiff_t)(r - 1) * ustride - 1);
            const uint32_t *w = (const uint32_t *)(ga & ~(uintptr_t)3);
            sh = (uint32_t)(ga & 3) * 8;
#pragma unroll
            for (int k = 0; k < 6; ++k) raw[k] = (4 * k < tw + 3 + 3) ? __ldg(w + k) : 0u;          // tw + 3 bytes from byte offset <= 3
        }
        uint32_t v[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = __funnelshift_r(raw[k], raw[k + 1], sh);                  // bytes 4k .. 4k + 3 of the row
        uint32_t *t = (uint32_t *)(s_tmp + slot * tmp_unit + r * tw);
#pragma unroll
        for (int x = 0; x < 16; x += 2) {
            if (x >= tw) break;
            const uint32_t b0 = (x & 3) ? __funnelshift_r(v[x >> 2], v[(x >> 2) + 1], 8 * (x & 3)) : v[x >> 2];
            const uint32_t b1 = __funnelshift_r(v[(x + 1) >> 2], v[((x + 1) >> 2) + 1], 8 * ((x + 1) & 3));
            t[x >> 1] = (uint32_t)(uint16_t)dp4a_us(b0, k0, 0) | ((uint32_t)(uint16_t)dp4a_us(b1, k0, 0) << 16);
        }
    }
}

__device__ __forceinline__ uint32_t pack_sat_u8x4(int o0, int o1, int o2, int o3)
{
    uint32_t t, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(o3), "r"(o2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(o1), "r"(o0), "r"(t));
    return d;
}
// V pass of the chroma class for the warp path: a 4-column x 4-row output block per lane and plane (the 2 x 16 x 16 samples of a 32x32 PU = one block
// per lane): seven s16 rows of four columns, six vertical pairs per column by PRMT, two DP2A per output, (S2 + 2048) >> 12 or BiPredClipping, then
// cvt.pack.sat to four bytes and ONE 32-bit store per row.  Same arithmetic as mc_stage_v<4, false>.
__device__ __forceinline__ void mc_chroma_v_warp(const McClass &c, int nl, const int16_t *s_tmp, int tmp_unit)
{
    const int lane = threadIdx.x & 31;
    const int tw = c.tw, gpr = tw >> 2, per_plane = gpr * (c.th >> 2), total = 2 * per_plane;
    const int lgg = 31 - __clz(gpr);
    const uint32_t ky[2] = {pack4(c_chroma_taps[c.fy[0] & 7]), pack4(c_chroma_taps[c.fy[1] & 7])};
    const bool al4 = ((((uintptr_t)c.out[0] | (uintptr_t)c.out[1]) | (uintptr_t)(c.ostride[0] | c.ostride[1])) & 3) == 0;
    for (int i = lane; i < total; i += 32) {
        const int p = i >= per_plane ? 1 : 0, j = i - p * per_plane;
        const int rq = j >> lgg, cg = j & (gpr - 1);
        int sum[4][4];      // [row][col]: uni S2 + 2048; bi (S2_0 >> 6) + (S2_1 >> 6)
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            if (l >= nl) continue;
            const uint8_t *base = (const uint8_t *)(s_tmp + (p * 2 + l) * tmp_unit + (rq * 4) * tw + cg * 4);
            uint2 w[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) w[k] = *(const uint2 *)(base + k * tw * 2);
            uint32_t pr[6][4];      // vertical pairs (row k, row k + 1) per column
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                pr[k][0] = __byte_perm(w[k].x, w[k + 1].x, 0x5410); pr[k][1] = __byte_perm(w[k].x, w[k + 1].x, 0x7632);
                pr[k][2] = __byte_perm(w[k].y, w[k + 1].y, 0x5410); pr[k][3] = __byte_perm(w[k].y, w[k + 1].y, 0x7632);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int v = dp2a_hi_ss(pr[r + 2][x], ky[l], dp2a_lo_ss(pr[r][x], ky[l], nl == 2 ? 0 : 2048));
                    if (nl == 2) sum[r][x] = (l ? sum[r][x] : 0) + (int)(short)(v >> 6);
                    else sum[r][x] = v;
                }
        }
        uint8_t *outp = (p ? c.out[1] : c.out[0]) + (size_t)(rq * 4) * (p ? c.ostride[1] : c.ostride[0]) + cg * 4;
        const int ostride = p ? c.ostride[1] : c.ostride[0];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int o[4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
                o[x] = nl == 2 ? (sum[r][x] + 64) >> 7 : sum[r][x] >> 12;
            const uint32_t v = pack_sat_u8x4(o[0], o[1], o[2], o[3]);
            uint8_t *q = outp + (size_t)r * ostride;
            if (al4) *(uint32_t *)q = v;
            else { q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16); q[3] = (uint8_t)(v >> 24); }
        }
    }
}

// copy / rounded average of np planes out of TMA-fetched tiles (pitch 48): plane p, list l at tiles + (p * 2 + l) * tile_bytes, sample (x, y) of the block at
// tile row y + row0, byte am[l] + col0 + x.  8-byte items (block widths are multiples of 8 for luma; chroma widths 4 take the 4-byte form).
__device__ __forceinline__ void mc_copy_tiles(const McClass &c, int nl, const uint8_t *tiles, int tile_bytes, const int (&am)[2], int row0, int col0)
{
    const int lane = threadIdx.x & 31;
    const int tw = c.tw, th = c.th;
    const uintptr_t al = ((uintptr_t)c.out[0] | (uintptr_t)c.ostride[0]) | (c.np == 2 ? ((uintptr_t)c.out[1] | (uintptr_t)c.ostride[1]) : 0);
    if ((tw & 15) == 0 && (al & 15) == 0) {      // 16-byte items: five tile words, four funnel shifts, one 128-bit store
        const int ipr = tw >> 4, lg = 31 - __clz(ipr), per_plane = th * ipr, total = c.np * per_plane;
#pragma unroll 2
        for (int i = lane; i < total; i += 32) {
            const int p = i >= per_plane ? 1 : 0, j = i - p * per_plane;
            const int y = j >> lg, g = j & (ipr - 1);
            uint32_t v[2][4];
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                if (l >= nl) continue;
                const int o = (l ? am[1] : am[0]) + col0 + g * 16;
                const uint32_t *w = (const uint32_t *)(tiles + (p * 2 + l) * tile_bytes + (y + row0) * 48) + (o >> 2);
                const uint32_t sh = (uint32_t)(o & 3) * 8;
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
                v[l][0] = __funnelshift_r(w0, w1, sh); v[l][1] = __funnelshift_r(w1, w2, sh); v[l][2] = __funnelshift_r(w2, w3, sh); v[l][3] = __funnelshift_r(w3, w4, sh);
            }
            if (nl == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[0][k] = __vavgu4(v[0][k], v[1][k]);
            }
            *(uint4 *)((p ? c.out[1] : c.out[0]) + (size_t)y * (p ? c.ostride[1] : c.ostride[0]) + g * 16) = make_uint4(v[0][0], v[0][1], v[0][2], v[0][3]);
        }
        return;
    }
    const bool wide = (tw & 7) == 0 && (al & 7) == 0;
    const int ib = wide ? 8 : 4, ipr = tw / ib, lg = 31 - __clz(ipr), per_plane = th * ipr, total = c.np * per_plane;
#pragma unroll 2
    for (int i = lane; i < total; i += 32) {
        const int p = i >= per_plane ? 1 : 0, j = i - p * per_plane;
        const int y = j >> lg, g = j & (ipr - 1);
        uint32_t v[2][2];
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            if (l >= nl) { v[l][0] = v[l][1] = 0; continue; }
            const int o = (l ? am[1] : am[0]) + col0 + g * ib;
            const uint32_t *w = (const uint32_t *)(tiles + (p * 2 + l) * tile_bytes + (y + row0) * 48) + (o >> 2);
            const uint32_t sh = (uint32_t)(o & 3) * 8;
            const uint32_t w0 = w[0], w1 = w[1], w2 = wide ? w[2] : 0u;
            v[l][0] = __funnelshift_r(w0, w1, sh);
            v[l][1] = __funnelshift_r(w1, w2, sh);
        }
        if (nl == 2) { v[0][0] = __vavgu4(v[0][0], v[1][0]); v[0][1] = __vavgu4(v[0][1], v[1][1]); }
        uint8_t *o = (p ? c.out[1] : c.out[0]) + (size_t)y * (p ? c.ostride[1] : c.ostride[0]) + g * ib;
        if (wide) *(uint2 *)o = make_uint2(v[0][0], v[0][1]);
        else if (((uintptr_t)o & 3) == 0) *(uint32_t *)o = v[0][0];
        else { o[0] = (uint8_t)v[0][0]; o[1] = (uint8_t)(v[0][0] >> 8); o[2] = (uint8_t)(v[0][0] >> 16); o[3] = (uint8_t)(v[0][0] >> 24); }
    }
}

// TMA = true: the windows of a PU arrive by cp.async.bulk.tensor (lane 0 issues up to six boxes on the warp's mbarrier, nothing else of the PU touches
// global memory for reading): one round trip per PU, the address generation is the copy engine's.  Per warp: two 48 x 32 luma tiles (reused by the
// chroma H-pass output) and four 48 x 19 chroma tiles.
constexpr int MC8_LT = 48 * 32, MC8_CT = 1024;        // tile bytes: luma 1536, chroma 48 * 19 = 912 -> 1024 (TMA destinations are 128-byte aligned)
template <bool TMA>
__global__ void __launch_bounds__(MC8_NT, TMA ? 3 : 4) mc_batch8_kernel(const __grid_constant__ McArgs A)
{
    using G = McSmem<false>;
    extern __shared__ __align__(128) uint8_t mc_dyn[];
    // layout: [TMA: per warp 2 luma + 4 chroma tiles | else per warp 4 chroma s16 tiles] [staged path: windows, s16 rows]
    constexpr int PER_WARP = TMA ? 2 * MC8_LT + 4 * MC8_CT : 4 * G::TMP_C * 2;
    uint8_t *s_warp = mc_dyn + (threadIdx.x >> 5) * PER_WARP;
    uint8_t *s_win = mc_dyn + MC8_JOBS * PER_WARP;
    int16_t *s_tmp = (int16_t *)(s_win + 2 * G::WIN_L);
    __shared__ int s_a[4];
    __shared__ int s_luma_frac[MC8_JOBS];
    __shared__ __align__(8) uint64_t s_bar[MC8_JOBS];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t parity = 0;
    if (TMA) {
        if (lane == 0) mbar_init(&s_bar[wid], 1);
        __syncwarp();
    }
    for (int phase = 0; phase < 2; ++phase) {
        // phase 0: every warp its own job (luma copy, chroma);  phase 1: the CTA walks the jobs whose luma needs the staged 8-tap path
        int any = 0;
        if (phase == 1) {
            __syncthreads();
#pragma unroll
            for (int k = 0; k < MC8_JOBS; ++k) any |= s_luma_frac[k];
            if (!any) return;
        }
        for (int jj = 0; jj < (phase ? MC8_JOBS : 1); ++jj) {
            const int slot_j = phase ? jj : wid, job = blockIdx.x * MC8_JOBS + slot_j;
            if (phase == 1 && !s_luma_frac[jj]) continue;
            if (job >= A.n_jobs) { if (phase == 0 && lane == 0) s_luma_frac[wid] = 0; continue; }
            const svtb200_mc_job q = A.jobs[job];
            const bool bi = q.dir == 2;
            const int nl = bi ? 2 : 1;
            int px[2], py[2], rix[2];
            const YuvD *r[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int l = bi ? s : (q.dir == 1 ? 1 : 0);
                const int ri = l ? q.ref_idx[1] : q.ref_idx[0], mvx = l ? q.mv[1][0] : q.mv[0][0], mvy = l ? q.mv[1][1] : q.mv[0][1];
                rix[s] = min(ri, A.n_refs - 1);
                r[s] = &A.refs[rix[s]];   // param space
                px[s] = (uint16_t)clampi((r[s]->ox - 71) << 2, (r[s]->w + r[s]->ox + 7) << 2, ((q.x + r[s]->ox) << 2) + mvx);
                py[s] = (uint16_t)clampi((r[s]->oy - 71) << 2, (r[s]->h + r[s]->oy + 7) << 2, ((q.y + r[s]->oy) << 2) + mvy);
            }
            int lfrac = 0;
            for (int ty0 = 0; ty0 < q.h; ty0 += TILE)
                for (int tx0 = 0; tx0 < q.w; tx0 += TILE) {
                    McClass L, K;   // luma, chroma
                    L.np = 1; L.tw = min(TILE, q.w - tx0); L.th = min(TILE, q.h - ty0);
                    K.np = 2; K.tw = L.tw >> 1; K.th = L.th >> 1;
                    L.out[0] = L.out[1] = A.dst[0] + ((size_t)(A.doy + q.y + ty0) * A.dstride[0] + A.dox + q.x + tx0);
                    L.ostride[0] = L.ostride[1] = A.dstride[0];
                    const size_t cidx = (((size_t)(A.doy + q.y) * A.dstride[1]) >> 1) + ((A.dox + q.x) >> 1) + (size_t)(ty0 >> 1) * A.dstride[1] + (tx0 >> 1);
                    K.out[0] = A.dst[1] + cidx; K.out[1] = A.dst[2] + cidx;
                    K.ostride[0] = A.dstride[1]; K.ostride[1] = A.dstride[2];
                    int lf = 0, kf = 0;
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        L.fx[s] = px[s] & 3; L.fy[s] = py[s] & 3; K.fx[s] = px[s] & 7; K.fy[s] = py[s] & 7;
                        if (s < nl) { lf |= L.fx[s] | L.fy[s]; kf |= K.fx[s] | K.fy[s]; }
                        if (!TMA || phase == 1) {      // (the TMA warp path addresses its windows by tensor coordinates)
                            L.u[0][s].ref = r[s]->p[0] + ((size_t)((py[s] >> 2) + ty0) * r[s]->stride[0] + (px[s] >> 2) + tx0);
                            L.u[0][s].rstride = r[s]->stride[0];
                            L.u[1][s] = L.u[0][s];
#pragma unroll
                            for (int pl = 1; pl < 3; ++pl) {
                                K.u[pl - 1][s].ref = r[s]->p[pl] + ((size_t)((py[s] >> 3) + (ty0 >> 1)) * r[s]->stride[pl] + (px[s] >> 3) + (tx0 >> 1));
                                K.u[pl - 1][s].rstride = r[s]->stride[pl];
                            }
                        }
                    }
                    L.interp = lf != 0; K.interp = kf != 0;
                    lfrac |= lf;
                    if (phase == 0) {
                        if (TMA) {
                            uint8_t *s_lt = s_warp, *s_ct = s_warp + 2 * MC8_LT;
                            int16_t *s_ctmp = (int16_t *)s_warp;                  // the chroma H-pass rows overlay the luma tiles once the luma copy is done
                            int al[2], ac[2];
#pragma unroll
                            for (int s = 0; s < 2; ++s) { al[s] = ((px[s] >> 2) + tx0) & 15; ac[s] = ((px[s] >> 3) + (tx0 >> 1) - 1) & 15; }
                            __syncwarp();                                      // every lane is done with the previous tile's shared memory
                            if (lane == 0) {
                                fence_proxy_async();
                                mbar_expect_tx(&s_bar[wid], (uint32_t)(nl * ((L.interp ? 0 : 48 * 32) + 2 * 48 * 19)));
#pragma unroll
                                for (int s = 0; s < 2; ++s) {
                                    if (s >= nl) continue;
                                    const int lx = (px[s] >> 2) + tx0, ly = (py[s] >> 2) + ty0, cx = (px[s] >> 3) + (tx0 >> 1) - 1, cy = (py[s] >> 3) + (ty0 >> 1) - 1;
                                    if (!L.interp) tma_load_2d(s_lt + s * MC8_LT, &A.maps[rix[s]][0], lx & ~15, ly, &s_bar[wid]);
                                    tma_load_2d(s_ct + (0 * 2 + s) * MC8_CT, &A.maps[rix[s]][1], cx & ~15, cy, &s_bar[wid]);
                                    tma_load_2d(s_ct + (1 * 2 + s) * MC8_CT, &A.maps[rix[s]][2], cx & ~15, cy, &s_bar[wid]);
                                }
                            }
                            mbar_wait(&s_bar[wid], parity);
                            parity ^= 1u;
                            if (!L.interp) mc_copy_tiles(L, nl, s_lt, MC8_LT, al, 0, 0);
                            if (!K.interp) mc_copy_tiles(K, nl, s_ct, MC8_CT, ac, 1, 1);
                            else {
                                __syncwarp();                                  // the luma copy has read its tiles
                                mc_chroma_h_rows<true>(K, nl, s_ctmp, G::TMP_C, s_ct, MC8_CT, ac);
                                __syncwarp();
                                mc_chroma_v_warp(K, nl, s_ctmp, G::TMP_C);
                            }
                        } else {
                            int16_t *s_ctmp = (int16_t *)s_warp;
                            const int dummy[2] = {0, 0};
                            if (!L.interp) mc_copy<false, 32>(L, nl);
                            if (!K.interp) mc_copy<false, 32>(K, nl);
                            else {
                                mc_chroma_h_rows<false>(K, nl, s_ctmp, G::TMP_C, nullptr, 0, dummy);
                                __syncwarp();
                                mc_chroma_v_warp(K, nl, s_ctmp, G::TMP_C);
                                __syncwarp();      // the next tile reuses the per-warp tile
                            }
                        }
                    } else if (L.interp) {
                        mc_stage_load<8, false, MC8_NT>(L, nl, s_win, G::WIN_L, s_a);
                        __syncthreads();
                        mc_stage_h<8, false, MC8_NT>(L, nl, s_win, G::WIN_L, s_a, s_tmp, G::TMP_L);
                        __syncthreads();
                        mc_stage_v<8, false, MC8_NT>(L, nl, s_tmp, G::TMP_L);
                        __syncthreads();
                    }
                }
            if (phase == 0 && lane == 0) s_luma_frac[wid] = lfrac != 0;
        }
    }
}
static size_t mc8_smem(bool tma)
{
    using G = McSmem<false>;
    return (size_t)MC8_JOBS * (tma ? 2 * MC8_LT + 4 * MC8_CT : 4 * G::TMP_C * 2) + 2 * G::WIN_L + 2 * G::TMP_L * 2 + 128;
}

// leaf: one block (w x h, multiples of 4 x 2) through the same three stages, one plane, one or two lists
template <int TAPS>
__global__ void __launch_bounds__(MC_NT) mc_leaf_kernel(const uint8_t *r0, int s0, int f0x, int f0y, const uint8_t *r1, int s1, int f1x, int f1y,
                                                        uint8_t *dst, int ds, int w, int h)
{
    using G = McSmem<false>;
    __shared__ __align__(16) uint8_t s_win[2 * G::WIN_L];
    __shared__ __align__(16) int16_t s_tmp[2 * G::TMP_L];
    __shared__ int s_a[4];
    const int nl = r1 ? 2 : 1;
    for (int ty0 = 0; ty0 < h; ty0 += TILE)
        for (int tx0 = 0; tx0 < w; tx0 += TILE) {
            McClass c;
            c.np = 1; c.tw = min(TILE, w - tx0); c.th = min(TILE, h - ty0);
            c.out[0] = c.out[1] = dst + (size_t)ty0 * ds + tx0; c.ostride[0] = c.ostride[1] = ds;
            c.fx[0] = f0x; c.fy[0] = f0y; c.fx[1] = f1x; c.fy[1] = f1y;
            c.u[0][0].ref = r0 + (size_t)ty0 * s0 + tx0; c.u[0][0].rstride = s0;
            c.u[0][1].ref = r1 ? r1 + (size_t)ty0 * s1 + tx0 : c.u[0][0].ref; c.u[0][1].rstride = r1 ? s1 : s0;
            c.u[1][0] = c.u[0][0]; c.u[1][1] = c.u[0][1];
            c.interp = (f0x | f0y | (r1 ? (f1x | f1y) : 0)) != 0;
            if (!c.interp) { mc_copy<false>(c, nl); continue; }
            mc_stage_load<TAPS, false>(c, nl, s_win, G::WIN_L, s_a);
            __syncthreads();
            mc_stage_h<TAPS, false>(c, nl, s_win, G::WIN_L, s_a, s_tmp, G::TMP_L);
            __syncthreads();
            mc_stage_v<TAPS, false>(c, nl, s_tmp, G::TMP_L);
            __syncthreads();
        }
}

// ---- interpolation-free prediction of the mode-decision fast loop (INTERPOLATION_FREE_PATH, Codec/EbEncDecProcess.c:2728: pictures without sub-pel) ----
// UniPredIFreeRef8Bit / BiPredIFreeRef8Bit (Codec/EbAvcStyleMcp.c:172,321): whatever the vector's fraction, function [0] of the AVC-style table (the
// copy) runs at the integer position the fraction maps to -- luma: position >> 2 plus IntegerPosoffsetTab{X,Y} (:19-27: +1 in x for fraction (3,2), +1
// in y for (1..3,3)); chroma: position >> 3, +1 where the 1/8 fraction exceeds 4 -- and bi-prediction is the rounded average of the two copies
// (BiPredAverageKernel / PictureAverage, :269).  Same position clamp as every PU-level wrapper (Inter2Nx2NPuPredictionInterpolationFree,
// Codec/EbInterPrediction.c:100-115).  One warp per PU, 4 samples per lane and step: a copy kernel, HBM / L2 bound.
__global__ void __launch_bounds__(256) mc_ifree_kernel(const __grid_constant__ McArgs A)
{
    const int job = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (job >= A.n_jobs) return;
    const svtb200_mc_job q = A.jobs[job];
    const bool bi = q.dir == 2;
    int px[2], py[2];
    const YuvD *r[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int l = bi ? s : (q.dir == 1 ? 1 : 0);
        const int ri = l ? q.ref_idx[1] : q.ref_idx[0], mvx = l ? q.mv[1][0] : q.mv[0][0], mvy = l ? q.mv[1][1] : q.mv[0][1];
        r[s] = &A.refs[min(ri, A.n_refs - 1)];
        px[s] = (uint16_t)clampi((r[s]->ox - 71) << 2, (r[s]->w + r[s]->ox + 7) << 2, ((q.x + r[s]->ox) << 2) + mvx);
        py[s] = (uint16_t)clampi((r[s]->oy - 71) << 2, (r[s]->h + r[s]->oy + 7) << 2, ((q.y + r[s]->oy) << 2) + mvy);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int w4 = (c ? q.w >> 1 : q.w) >> 2, h = c ? q.h >> 1 : q.h;
        const uint8_t *s0, *s1;
        {
            const uint8_t *sp[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                int ix, iy;
                if (c == 0) {
                    const int fx = px[s] & 3, fy = py[s] & 3;
                    ix = (px[s] >> 2) + (fx == 3 && fy == 2); iy = (py[s] >> 2) + (fy == 3 && fx != 0);
                } else {
                    ix = (px[s] >> 3) + ((px[s] & 7) > 4); iy = (py[s] >> 3) + ((py[s] & 7) > 4);
                }
                sp[s] = r[s]->p[c] + (ptrdiff_t)iy * r[s]->stride[c] + ix;
            }
            s0 = sp[0]; s1 = sp[1];
        }
        const int st0 = r[0]->stride[c], st1 = r[1]->stride[c];
        // the destination index exactly as the callers form it: luma origin + PU position, chroma = half of both
        uint8_t *d = c == 0 ? A.dst[0] + (size_t)(A.doy + q.y) * A.dstride[0] + A.dox + q.x
                            : A.dst[c] + (size_t)((A.doy + q.y) >> 1) * A.dstride[c] + ((A.dox + q.x) >> 1);
        const int lg = 31 - __clz(w4), n = h << lg;                   // w4 is 1, 2, 4, 8 or 16
        for (int i = lane; i < n; i += 32) {
            const int y = i >> lg, x = i & (w4 - 1);
            uint32_t v = ld_u32_unaligned(s0 + (ptrdiff_t)y * st0 + 4 * x);
            if (bi) v = __vavgu4(v, ld_u32_unaligned(s1 + (ptrdiff_t)y * st1 + 4 * x));
            uint8_t *o = d + (size_t)y * A.dstride[c] + 4 * x;
            if (((uintptr_t)o & 3) == 0) *(uint32_t *)o = v;
            else { o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16); o[3] = (uint8_t)(v >> 24); }
        }
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
                     uint8_t *dy, uint8_t *dcb, uint8_t *dcr, cudaStream_t st, bool hbd = false, bool ifree = false)
{
    McArgs A;
    for (int i = 0; i < 16; ++i) A.refs[i] = h_refs[i < n_refs ? i : 0];
    A.n_refs = n_refs; A.jobs = d_jobs; A.n_jobs = n_jobs;
    A.dst[0] = dy; A.dst[1] = dcb; A.dst[2] = dcr;
    A.dstride[0] = (int)dst.y.stride; A.dstride[1] = (int)dst.cb.stride; A.dstride[2] = (int)dst.cr.stride;
    A.dox = dst.y.origin_x; A.doy = dst.y.origin_y;
    if (n_jobs == 0) return SVTB200_OK;
    if (ifree) { A.tma = 0; mc_ifree_kernel<<<(n_jobs + 7) / 8, 256, 0, st>>>(A); }
    else if (hbd) mc_batch_kernel<true><<<n_jobs, MC_NT, 0, st>>>(A);
    else {
        // TMA path: at most four references, every plane 16-byte aligned with a 16-byte pitch (device pictures of svtb200_yuv_picture_create)
        A.tma = n_refs <= 4;
        for (int i = 0; i < n_refs && A.tma; ++i)
            for (int k = 0; k < 3; ++k) if (((uintptr_t)h_refs[i].p[k] & 15) || (h_refs[i].stride[k] & 15)) A.tma = 0;
        for (int i = 0; i < n_refs && A.tma; ++i) {
            const uint32_t rows_y = (uint32_t)(h_refs[i].h + 2 * h_refs[i].oy), rows_c = (uint32_t)(h_refs[i].h / 2 + h_refs[i].oy);
            int rc = make_plane_tensor_map(&A.maps[i][0], h_refs[i].p[0], (uint32_t)h_refs[i].stride[0], rows_y, 48, 32);
            if (!rc) rc = make_plane_tensor_map(&A.maps[i][1], h_refs[i].p[1], (uint32_t)h_refs[i].stride[1], rows_c, 48, 19);
            if (!rc) rc = make_plane_tensor_map(&A.maps[i][2], h_refs[i].p[2], (uint32_t)h_refs[i].stride[2], rows_c, 48, 19);
            if (rc) return rc;
        }
        const size_t smem = mc8_smem(A.tma != 0);
        static thread_local bool attr_set[2] = {false, false};
        if (!attr_set[A.tma != 0]) {
            if (A.tma) SVTB200_CUDA_TRY(cudaFuncSetAttribute(mc_batch8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else SVTB200_CUDA_TRY(cudaFuncSetAttribute(mc_batch8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[A.tma != 0] = true;
        }
        const int grid = (n_jobs + MC8_JOBS - 1) / MC8_JOBS;
        if (A.tma) mc_batch8_kernel<true><<<grid, MC8_NT, smem, st>>>(A); else mc_batch8_kernel<false><<<grid, MC8_NT, smem, st>>>(A);
    }
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static size_t pbytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }

}  // namespace svtb200
using namespace svtb200;

extern "C" {

static int mc_batch_dev_impl(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs_dev, int n_jobs, const svtb200_yuv_picture *dst, void *cuda_stream, bool hbd, bool ifree = false)
{
    if (!refs || n_refs <= 0 || n_refs > 16 || (!jobs_dev && n_jobs) || n_jobs < 0 || !dst) { set_error("mc_batch_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = check_pic(*dst, "destination");
    for (int i = 0; i < n_refs && !rc; ++i) rc = check_pic(refs[i], "reference");
    if (rc) return rc;
    if ((rc = ensure_device())) return rc;
    YuvD tab[16];
    for (int i = 0; i < n_refs; ++i) tab[i] = to_yuvd(refs[i], refs[i].y.buf, refs[i].cb.buf, refs[i].cr.buf);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    return launch_mc(tab, n_refs, jobs_dev, n_jobs, *dst, dst->y.buf, dst->cb.buf, dst->cr.buf, st, hbd, ifree);
}

static int mc_batch_impl(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst, bool hbd, bool ifree = false)
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
    if ((rc = launch_mc((const YuvD *)h_tab, n_refs, (const svtb200_mc_job *)d_jobs, n_jobs, *dst, (uint8_t *)dp[0], (uint8_t *)dp[1], (uint8_t *)dp[2], c->stream, hbd, ifree))) return rc;
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
int svtb200_mc_ifree_batch_dev(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs_dev, int n_jobs, const svtb200_yuv_picture *dst, void *cuda_stream)
{ return mc_batch_dev_impl(refs, n_refs, jobs_dev, n_jobs, dst, cuda_stream, false, true); }
int svtb200_mc_ifree_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst)
{ return mc_batch_impl(refs, n_refs, jobs, n_jobs, dst, false, true); }
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
    if ((rc = c->dev_reserve(b0 + b1 + bd + 4096)) || (rc = c->dev_alloc(b0 + 48, &d0)) || (rc = c->dev_alloc(b1 + 48, &d1)) || (rc = c->dev_alloc(bd, &dd)) || (rc = c->pin_alloc(bd, &hd))) return rc;
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
