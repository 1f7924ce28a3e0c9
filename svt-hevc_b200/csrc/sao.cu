// sao.cu -- sample adaptive offset: per-LCU statistics and picture-level application, 8-bit 4:2:0.
//
// Replaces:
//   GatherSaoStatisticsLcuLossy_62x62 (C_DEFAULT/EbSampleAdaptiveOffset_C.c:125) and GatherSaoStatisticsLcu_BT_SSE2
//   SAOApplyBO / SAOApplyEO_{0,90,135,45} (C_DEFAULT/EbSampleAdaptiveOffset_C.c:394-861) and their SSSE3 rows,
//   ApplySaoOffsetsLcu / ApplySaoOffsetsPicture (Codec/EbEncDecProcess.c:215,522)
// The RD decision (SaoGenerationDecision, Codec/EbSampleAdaptiveOffsetGenerationDecision.c:647) stays on the host: it
// needs the CABAC rate tables and is a few hundred scalar operations per LCU.
//
// Design.  Statistics: one CTA (128 threads) per (LCU, component); the reconstructed block is staged once in shared memory; edge
// classes are computed SIMD-in-word for 4 samples at a time (edge_idx4: packed compares, no per-sample branches) and accumulated in
// thread-private [row][thread] shared-memory histograms (EO type x edge index, and the 32 bands) packed as (count << 16) + sum -- bank-conflict
// free by construction, no atomics, column-summed at the end.  Application: out = SAO(in), one
// thread per 4 samples, neighbours always read from the unmodified input picture, so the reference's line buffers
// (saoUpBuffer / saoLeftBuffer) have no counterpart.  Both kernels are HBM-bound (18 432 B per CTU).
#include "common.cuh"

namespace svtb200 {

struct PlaneRO { const uint8_t *p; int stride; };
struct StatsArgs {
    PlaneRO src[3], rec[3];      // pointers at the interior origin of each plane
    PlaneRO pre[3];              // H3: the unfiltered reconstruction (p == nullptr: statistics of `rec` as it is)
    svtb200_sao_stats *out;
    int W, H, cols;
};

__device__ __forceinline__ int sgn(int a, int b) { return (a > b) - (a < b); }

// sign(v - n) + sign(v - m) + 2 for four packed unsigned bytes: 0..4 per byte
__device__ __forceinline__ uint32_t edge_idx4(uint32_t v, uint32_t n, uint32_t m)
{
    // __vcmp* give 0xff per true byte (= -1 as s8): sign(v - n) = (v > n) - (v < n) = lt_mask - gt_mask in byte arithmetic
    const uint32_t sa = __vsub4(__vcmpltu4(v, n), __vcmpgtu4(v, n));
    const uint32_t sb = __vsub4(__vcmpltu4(v, m), __vcmpgtu4(v, m));
    return __vadd4(__vadd4(sa, sb), 0x02020202u);
}
__device__ __forceinline__ int dp4a_ss(uint32_t a_s8x4, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s8x4), "r"(b_s8x4), "r"(c));
    return d;
}

#ifndef ST_NT_V
#define ST_NT_V 128
#endif
constexpr int ST_NT = ST_NT_V;
constexpr int ST_ROWS = 20 + 32;      // histogram rows: EO type * 5 + edge index (0..4), then the 32 bands
__global__ void __launch_bounds__(ST_NT) sao_stats_kernel(const __grid_constant__ StatsArgs A)
{
    // reconstructed block with a one-word margin on the left (column 4 = block x 0); pitch 72 bytes = 18 words
    __shared__ __align__(16) uint8_t s_rec[64 * 72];
    // thread-private histograms, [row][thread]: a thread only ever touches its own column, so a warp's 32 accesses always fall on 32 different
    // banks whatever rows they hit -- no atomics, no conflicts.  word = (count << 16) + sum(diff) (both small per thread: <= 31 samples).
    __shared__ int s_h[ST_ROWS * ST_NT];
    const int job = blockIdx.x, lcu = job / 3, c = job - lcu * 3, sh = c ? 1 : 0;
    const int lx = lcu % A.cols, ly = lcu / A.cols;
    const int ox = (lx * 64) >> sh, oy = (ly * 64) >> sh;
    const int lw = min(64, A.W - lx * 64) >> sh, lh = min(64, A.H - ly * 64) >> sh;
    const uint8_t *r = A.rec[c].p + (size_t)oy * A.rec[c].stride + ox, *s = A.src[c].p + (size_t)oy * A.src[c].stride + ox;
    const int wpr = lw >> 2;
    const SmallDiv d_wpr = mk_div(wpr);
    // H3 (rule pinned by tests/test_oracle_sao_h3.py): when the encoder gathers LCU n's statistics its last 4 columns / rows still hold the
    // UNFILTERED reconstruction if a right / bottom neighbour exists (their deblocking waits for that neighbour); everything else is final.
    const uint8_t *u = A.pre[c].p ? A.pre[c].p + (size_t)oy * A.pre[c].stride + ox : nullptr;
    const bool has_right = (lx + 1) * 64 < A.W, has_bottom = (ly + 1) * 64 < A.H;
    for (int i = threadIdx.x; i < lh * wpr; i += ST_NT) {
        const int y = udiv(i, d_wpr), x4 = i - y * wpr;
        const bool strip = u && ((has_right && 4 * x4 >= lw - 4) || (has_bottom && y >= lh - 4));
        ((uint32_t *)(s_rec + y * 72))[1 + x4] = strip ? *(const uint32_t *)(u + (size_t)y * A.pre[c].stride + 4 * x4)
                                                        : *(const uint32_t *)(r + (size_t)y * A.rec[c].stride + 4 * x4);   // block origins are 4-aligned
    }
    int *my = s_h + threadIdx.x;
#pragma unroll 4
    for (int k = 0; k < ST_ROWS; ++k) my[k * ST_NT] = 0;
    __syncthreads();
    const int n_items = (lh - 2) * wpr;
    for (int i = threadIdx.x; i < n_items; i += ST_NT) {
        const int y0 = udiv(i, d_wpr), y = 1 + y0, xw = i - y0 * wpr;
        const uint32_t *row = (const uint32_t *)(s_rec + y * 72) + 1 + xw;
        const uint32_t up0 = row[-19], up1 = row[-18], up2 = row[-17], c0 = row[-1], v = row[0], c2 = row[1], dn0 = row[17], dn1 = row[18], dn2 = row[19];
        const uint32_t sv = *(const uint32_t *)(s + (size_t)y * A.src[c].stride + 4 * xw);
        // valid bytes: the 62x62 interior (GatherSaoStatisticsLcuLossy_62x62 starts at +1,+1 and stops one short)
        uint32_t vm = 0xffffffffu;
        if (xw == 0) vm &= 0xffffff00u;
        if (xw == wpr - 1) vm &= 0x00ffffffu;
        // diff = clipToSigned8Bit(src - rec), packed s8
        const uint32_t dpos = __vminu4(__vsubus4(sv, v), 0x7f7f7f7fu), dneg = __vminu4(__vsubus4(v, sv), 0x80808080u);
        const uint32_t d = __vsub4(dpos, dneg);
        const uint32_t L = __funnelshift_r(c0, v, 24), R = __funnelshift_r(v, c2, 8);
        const uint32_t UL = __funnelshift_r(up0, up1, 24), UR = __funnelshift_r(up1, up2, 8), DL = __funnelshift_r(dn0, dn1, 24), DR = __funnelshift_r(dn1, dn2, 8);
        const uint32_t e[4] = {edge_idx4(v, L, R), edge_idx4(v, up1, dn1), edge_idx4(v, UL, DR), edge_idx4(v, UR, DL)};   // EO-0, 90, 135, 45
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (!((vm >> (8 * b)) & 1)) continue;
            const int add = 65536 + (int)(signed char)(d >> (8 * b));
#pragma unroll
            for (int t = 0; t < 4; ++t) my[(t * 5 + ((e[t] >> (8 * b)) & 7)) * ST_NT] += add;
            my[(20 + ((v >> (8 * b + 3)) & 31)) * ST_NT] += add;
        }
    }
    __syncthreads();
    // column sums: warp w takes rows w, w + 4, ...
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    svtb200_sao_stats *o = A.out + job;
    for (int row = wid; row < ST_ROWS; row += ST_NT / 32) {
        // the packed words add as integers: (sum of counts << 16) + sum of diffs, the latter within s16 for four threads (4 * 32 samples * 127)
        int w = 0;
#pragma unroll
        for (int q = 0; q < ST_NT / 32; ++q) w += s_h[row * ST_NT + lane + 32 * q];
        int dsum = (int)(short)(w & 0xffff), csum = (w - dsum) >> 16;
        dsum = __reduce_add_sync(0xffffffffu, dsum);
        csum = __reduce_add_sync(0xffffffffu, csum);
        if (lane == 0) {
            if (row >= 20) { o->bo_diff[row - 20] = dsum; o->bo_count[row - 20] = (uint16_t)csum; }
            else {
                // categories compacted like the reference: slots 0..3 <- edge index 0, 1, 3, 4; slot 4 is scratch there (we write 0)
                const int t = row / 5, k = row - t * 5;
                if (k == 2) { o->eo_diff[t][4] = 0; o->eo_count[t][4] = 0; }
                else { const int slot = k < 2 ? k : k - 1; o->eo_diff[t][slot] = dsum; o->eo_count[t][slot] = (uint16_t)csum; }
            }
        }
    }
}

// ---- round-2 statistics kernel (8-bit): ONE CTA PER LCU, all three components; a thread owns a 4-sample-wide, 8-row-tall strip ----------------------------
// The per-sample work of GatherSaoStatisticsLcuLossy_62x62 is five histogram updates (four edge types + the band).  Here the edge-type updates never touch
// memory: a thread keeps, per edge type and category, a packed byte counter (<= 8 rows: no carry between bytes) and a 32-bit difference sum, fed by 0/1
// byte masks -- the category masks are LOP3 combinations of the byte compare masks (v < a, v > a, v < b, v > b), the difference sums one DP4A each.  The
// vertical compares of a row serve the next row (sign(v[y] - v[y+1]) = -sign(v[y+1] - v[y])), and the three rows a sample needs live in registers as the
// strip walks down.  Only the band-offset bins keep the thread-private shared-memory histogram.  Warps 0-3: luma (16 word columns x 8 strips), warp 4: Cb,
// warp 5: Cr (8 word columns x 4 strips).  Same 62 x 62 interior rule, same H3 composition, same output layout as sao_stats_kernel.
constexpr int ST2_NT = 192, ST2_P = ST2_NT;      // pitch a multiple of 32: a thread's column stays in ITS bank whatever band a sample falls into
// 0xff where the byte's top bit is set, else 0: PRMT with the sign-replicate bit of every selector nibble (__byte_perm() masks that bit off, hence the PTX)
__device__ __forceinline__ uint32_t msb_mask(uint32_t w)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(d) : "r"(w));
    return d;
}
// byte masks (0xff / 0x00) of a > b and a < b for unsigned bytes: the carry-free per-byte difference trick, both directions from shared terms
__device__ __forceinline__ void cmp4(uint32_t a, uint32_t b, uint32_t &gt, uint32_t &lt)
{
    // a > b  <=>  msb of ((b & ~a) | (~(b ^ a) & (b - a per byte, 7-bit))) ...: compute the 7-bit differences without inter-byte borrow
    const uint32_t al = a | 0x80808080u, bl = b | 0x80808080u, ah = a & 0x7f7f7f7fu, bh = b & 0x7f7f7f7fu;
    const uint32_t dab = al - bh, dba = bl - ah;              // per byte: bit 7 = (a7 >= b7) / (b7 >= a7) on the low 7 bits
    // a < b: msb(a) < msb(b), or equal msbs and low7(a) < low7(b) (bit 7 of dab clear)
    const uint32_t x = a ^ b;
    lt = msb_mask(((~a & b) & x) | (~x & ~dab));
    gt = msb_mask(((a & ~b) & x) | (~x & ~dba));
}
struct EoAcc { uint32_t cnt[4]; int sum[4]; };      // categories 0, 1, 3, 4 of one edge type
__device__ __forceinline__ void eo_accumulate(EoAcc &A, uint32_t la, uint32_t ga, uint32_t lb, uint32_t gb, uint32_t valid01, uint32_t d)
{
    // category masks as 0/1 bytes: 0: v < a & v < b; 4: v > a & v > b; 1: exactly one '<' and no '>'; 3: exactly one '>' and no '<'
    const uint32_t k0 = la & lb & valid01, k4 = ga & gb & valid01;
    const uint32_t k1 = (la ^ lb) & ~(ga | gb) & valid01, k3 = (ga ^ gb) & ~(la | lb) & valid01;
    A.cnt[0] += k0; A.cnt[1] += k1; A.cnt[2] += k3; A.cnt[3] += k4;
    A.sum[0] = dp4a_ss(d, k0, A.sum[0]); A.sum[1] = dp4a_ss(d, k1, A.sum[1]); A.sum[2] = dp4a_ss(d, k3, A.sum[2]); A.sum[3] = dp4a_ss(d, k4, A.sum[3]);
}

__global__ void __launch_bounds__(ST2_NT) sao_stats2_kernel(const __grid_constant__ StatsArgs A)
{
    // reconstructed blocks of the three components with a one-word margin on every side: luma 66 rows x 72 bytes, chroma 34 rows x 40 bytes each
    __shared__ __align__(16) uint8_t s_y[66 * 72];
    __shared__ __align__(16) uint8_t s_c[2][34 * 40];
    __shared__ int s_bo[32 * ST2_P];             // thread-private band histograms [band][thread], word = (count << 16) + sum(diff)
    __shared__ int s_eo[3][32];                  // per component: 4 types x (4 counts, 4 sums)
    const int lcu = blockIdx.x, lx = lcu % A.cols, ly = lcu / A.cols;
    const int wid = threadIdx.x >> 5;
    // ---- stage the three blocks (H3 composition included) ----
    for (int c = 0; c < 3; ++c) {
        const int sh = c ? 1 : 0;
        const int ox = (lx * 64) >> sh, oy = (ly * 64) >> sh;
        const int lw = min(64, A.W - lx * 64) >> sh, lh = min(64, A.H - ly * 64) >> sh;
        const int pitch = c ? 40 : 72;
        uint8_t *dst = c ? s_c[c - 1] : s_y;
        const uint8_t *r = A.rec[c].p + (size_t)oy * A.rec[c].stride + ox;
        const uint8_t *u = A.pre[c].p ? A.pre[c].p + (size_t)oy * A.pre[c].stride + ox : nullptr;
        const bool has_right = (lx + 1) * 64 < A.W, has_bottom = (ly + 1) * 64 < A.H;
        const int wpr = lw >> 2;
        const SmallDiv d_wpr = mk_div(wpr);
        for (int i = threadIdx.x; i < lh * wpr; i += ST2_NT) {
            const int y = udiv(i, d_wpr), x4 = i - y * wpr;
            const bool strip = u && ((has_right && 4 * x4 >= lw - 4) || (has_bottom && y >= lh - 4));
            ((uint32_t *)(dst + (y + 1) * pitch))[1 + x4] = strip ? *(const uint32_t *)(u + (size_t)y * A.pre[c].stride + 4 * x4)
                                                                : *(const uint32_t *)(r + (size_t)y * A.rec[c].stride + 4 * x4);
        }
    }
    int *mybo = s_bo + threadIdx.x;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) mybo[k * ST2_P] = 0;
    if (threadIdx.x < 96) s_eo[threadIdx.x >> 5][threadIdx.x & 31] = 0;
    __syncthreads();
    // ---- strips ----
    const int c = wid < 4 ? 0 : wid - 3, sh = c ? 1 : 0;
    const int ox = (lx * 64) >> sh, oy = (ly * 64) >> sh;
    const int lw = min(64, A.W - lx * 64) >> sh, lh = min(64, A.H - ly * 64) >> sh;
    const int pitch = c ? 40 : 72, wpr = lw >> 2;
    const uint8_t *blk = c ? s_c[c - 1] : s_y;
    const int t = c ? (int)(threadIdx.x & 31) : (int)threadIdx.x;          // thread index within the component
    const int xw = c ? (t & 7) : (t & 15), sy = c ? (t >> 3) : (t >> 4);   // word column, strip index (8 rows each)
    EoAcc E[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) { E[k].cnt[q] = 0; E[k].sum[q] = 0; }
    if (xw < wpr && sy * 8 < lh) {
        const uint8_t *src = A.src[c].p + (size_t)oy * A.src[c].stride + ox;
        const int y0 = sy * 8, y1 = min(lh, y0 + 8);
        // rolling window: (left, centre, right) words of rows y - 1, y, y + 1 (tile row index = y + 1, word index = xw + 1)
        const uint32_t *row = (const uint32_t *)(blk + y0 * pitch) + xw;             // row y0 - 1
        const int p4 = pitch >> 2;
        uint32_t uL = row[0], uC = row[1], uR = row[2];
        uint32_t cL = row[p4], cC = row[p4 + 1], cR = row[p4 + 2];
        // compares of row y against its upper neighbours come from the previous row's compares against ITS lower neighbours
        uint32_t gU, lU, gUL, lUL, gUR, lUR;
        {
            const uint32_t UL = __funnelshift_r(uL, uC, 24), UR = __funnelshift_r(uC, uR, 8);
            cmp4(cC, uC, gU, lU); cmp4(cC, UL, gUL, lUL); cmp4(cC, UR, gUR, lUR);
        }
        uint32_t colmask = 0xffffffffu;                                               // 62 x 62 interior: the first / last sample of a row never counts
        if (xw == 0) colmask &= 0xffffff00u;
        if (xw == wpr - 1) colmask &= 0x00ffffffu;
        const uint8_t *srow = src + (size_t)y0 * A.src[c].stride + 4 * xw;
        uint32_t sv_next = *(const uint32_t *)srow;                                   // the source word of a row is fetched one row ahead
        for (int y = y0; y < y1; ++y) {
            const uint32_t sv = sv_next;
            if (y + 1 < y1) sv_next = *(const uint32_t *)(srow + (size_t)(y + 1 - y0) * A.src[c].stride);
            const uint32_t *nr = (const uint32_t *)(blk + (y + 2) * pitch) + xw;      // row y + 1
            const uint32_t dL = nr[0], dC = nr[1], dR = nr[2];
            const uint32_t v = cC;
            const uint32_t Lw = __funnelshift_r(cL, cC, 24), Rw = __funnelshift_r(cC, cR, 8);
            const uint32_t DL = __funnelshift_r(dL, dC, 24), DR = __funnelshift_r(dC, dR, 8);
            uint32_t gL, lL, gR, lR, gD, lD, gDL, lDL, gDR, lDR;
            cmp4(v, Lw, gL, lL); cmp4(v, Rw, gR, lR); cmp4(v, dC, gD, lD); cmp4(v, DL, gDL, lDL); cmp4(v, DR, gDR, lDR);
            if (y >= 1 && y < lh - 1) {
                const uint32_t valid01 = colmask & 0x01010101u;
                // diff = clipToSigned8Bit(src - rec), packed s8
                const uint32_t dpos = __vminu4(__vsubus4(sv, v), 0x7f7f7f7fu), dneg = __vminu4(__vsubus4(v, sv), 0x80808080u);
                const uint32_t d = __vsub4(dpos, dneg);
                eo_accumulate(E[0], lL, gL, lR, gR, valid01, d);         // EO-0: left / right
                eo_accumulate(E[1], lU, gU, lD, gD, valid01, d);         // EO-90: up / down
                eo_accumulate(E[2], lUL, gUL, lDR, gDR, valid01, d);     // EO-135: up-left / down-right
                eo_accumulate(E[3], lUR, gUR, lDL, gDL, valid01, d);     // EO-45: up-right / down-left
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (!((colmask >> (8 * b)) & 1)) continue;
                    mybo[((v >> (8 * b + 3)) & 31) * ST2_P] += 65536 + (int)(signed char)(d >> (8 * b));
                }
            }
            // next row: its upper compares are this row's lower compares mirrored (v[y+1] > v[y] <=> v[y] < v[y+1]); the diagonals swap sides with the
            // one-sample shift: (y+1, x) vs (y, x-1) [its UL] is the DR compare of (y, x-1), which lives in the left neighbour's word -- recomputed below
            gU = lD; lU = gD;
            uL = cL; uC = cC; uR = cR; cL = dL; cC = dC; cR = dR;
            {
                const uint32_t UL = __funnelshift_r(uL, uC, 24), UR = __funnelshift_r(uC, uR, 8);
                cmp4(cC, UL, gUL, lUL); cmp4(cC, UR, gUR, lUR);
            }
        }
    }
    // ---- reduce: byte counters -> counts, warp sums, one shared-memory atomic per (warp, bin) ----
    {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int cn = __reduce_add_sync(0xffffffffu, dp4a_ss(E[k].cnt[q], 0x01010101u, 0));
                const int sm = __reduce_add_sync(0xffffffffu, E[k].sum[q]);
                if (lane == 0) { atomicAdd(&s_eo[c][k * 8 + q], cn); atomicAdd(&s_eo[c][k * 8 + 4 + q], sm); }
            }
    }
    __syncthreads();
    // ---- outputs: per component, EO slots 0..3 <- categories 0, 1, 3, 4 (slot 4 = 0), then the 32 bands ----
    if (threadIdx.x < 96) {
        const int cc = threadIdx.x >> 5, k = threadIdx.x & 31;
        svtb200_sao_stats *o = A.out + (size_t)lcu * 3 + cc;
        const int tt = k >> 3, q = k & 7;
        if (q < 4) o->eo_count[tt][q] = (uint16_t)s_eo[cc][k]; else o->eo_diff[tt][q - 4] = s_eo[cc][k];
        if (k < 4) { o->eo_diff[k][4] = 0; o->eo_count[k][4] = 0; }
    }
    // band histograms: component cc owns the thread columns [first, first + n): luma 0..127, Cb 128..159, Cr 160..191.  One thread per (component,
    // band) walks its row of the [band][thread] table, starting at a column skewed by its band so that the 32 rows a warp reads fall into 32 different
    // banks, and adds four packed words at a time before unpacking (4 x 32 samples: count <= 128, |sum of differences| <= 16 256, inside their half word).
    if (threadIdx.x < 96) {
        const int cc = threadIdx.x >> 5, band = threadIdx.x & 31;
        const int first = cc == 0 ? 0 : 96 + 32 * cc, n = cc == 0 ? 128 : 32;
        const int *row = s_bo + band * ST2_P + first;
        int dsum = 0, csum = 0;
        for (int q0 = 0; q0 < n; q0 += 4) {
            const int q = (q0 + band) & (n - 1);      // n is 32 or 128
            const int w = row[q] + row[(q + 1) & (n - 1)] + row[(q + 2) & (n - 1)] + row[(q + 3) & (n - 1)], dd = (int)(short)(w & 0xffff);
            dsum += dd; csum += (w - dd) >> 16;
        }
        svtb200_sao_stats *o = A.out + (size_t)lcu * 3 + cc;
        o->bo_diff[band] = dsum; o->bo_count[band] = (uint16_t)csum;
    }
}

struct ApplyArgs {
    PlaneRO in[3];
    uint8_t *out[3];
    int ostride[3];
    const svtb200_sao_params *prm;
    int W, H, cols;
    int flag[2];
};

// one thread per 4 horizontally adjacent samples of one plane (blockIdx.z = plane)
__global__ void __launch_bounds__(256) sao_apply_kernel(const __grid_constant__ ApplyArgs A)
{
    const int c = blockIdx.z, sh = c ? 1 : 0, pw = A.W >> sh, ph = A.H >> sh;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= pw || y >= ph) return;
    const svtb200_sao_params *pp = A.prm + (size_t)((y << sh) >> 6) * A.cols + ((x0 << sh) >> 6);   // 4 samples never straddle an LCU
    const int t = A.flag[c ? 1 : 0] ? pp->type[c ? 1 : 0] : 0;
    // the four offsets of this plane in registers (a struct copy indexed by the class would live in local memory)
    const int8_t *po = pp->offset[c];
    const int o0 = po[0], o1 = po[1], o2 = po[2], o3 = po[3];
    const int is = A.in[c].stride;
    const uint8_t *row = A.in[c].p + (size_t)y * is + x0;
    const uint32_t cur = *(const uint32_t *)row;
    uint32_t res = cur;
    if (t == 5) {
        const int pos = pp->band_position[c];
        res = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int v = (cur >> (8 * k)) & 255;
            const int b = (v >> 3) - pos;
            if (b >= 0 && b <= 3) v = min(max(v + (b == 0 ? o0 : b == 1 ? o1 : b == 2 ? o2 : o3), 0), 255);
            res |= (uint32_t)v << (8 * k);
        }
    } else if (t >= 1 && t <= 4) {
        const int dx = (t == 2) ? 0 : (t == 4 ? 1 : -1);      // first neighbour: EO-0 (-1,0) EO-90 (0,-1) EO-135 (-1,-1) EO-45 (+1,-1)
        const int dy = (t == 1) ? 0 : -1;
        const bool skip_row = (t != 1) && (y == 0 || y == ph - 1);
        if (!skip_row) {
            const uint8_t *ra = row + (ptrdiff_t)dy * is, *rb = row - (ptrdiff_t)dy * is;
            // bytes x0-1 .. x0+4 of the two neighbour rows
            const uint32_t a0 = *(const uint32_t *)(ra - 4), a1 = *(const uint32_t *)ra, a2 = *(const uint32_t *)(ra + 4);
            const uint32_t b0 = *(const uint32_t *)(rb - 4), b1 = *(const uint32_t *)rb, b2 = *(const uint32_t *)(rb + 4);
            const uint32_t na = dx == 0 ? a1 : (dx < 0 ? __funnelshift_r(a0, a1, 24) : __funnelshift_r(a1, a2, 8));
            const uint32_t nb = dx == 0 ? b1 : (dx < 0 ? __funnelshift_r(b1, b2, 8) : __funnelshift_r(b0, b1, 24));
            res = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int v = (cur >> (8 * k)) & 255;
                const int x = x0 + k;
                if (!((t != 2) && (x == 0 || x == pw - 1))) {
                    const int idx = sgn(v, (int)((na >> (8 * k)) & 255)) + sgn(v, (int)((nb >> (8 * k)) & 255)) + 2;
                    const int off = idx == 0 ? o0 : idx == 1 ? o1 : idx == 3 ? o2 : idx == 4 ? o3 : 0;
                    v = min(max(v + off, 0), 255);
                }
                res |= (uint32_t)v << (8 * k);
            }
        }
    }
    *(uint32_t *)(A.out[c] + (size_t)y * A.ostride[c] + x0) = res;
}


// ---- 16-bit sample planes (10-bit video): GatherSaoStatisticsLcu_62x62_16bit (C_DEFAULT/EbSampleAdaptiveOffset_C.c:23: no diff clipping, band =
// sample >> 5), SAOApplyBO16bit / SAOApplyEO_*16bit (:431 ff, clip to 1023), ApplySaoOffsetsPicture16bit (Codec/EbEncDecProcess.c:1063).
// One sample per thread-iteration, neighbours straight from L1/L2; same thread-private [row][thread] histograms as the 8-bit kernel.
__global__ void __launch_bounds__(ST_NT) sao_stats16_kernel(const __grid_constant__ StatsArgs A)
{
    __shared__ int s_h[ST_ROWS * ST_NT];
    const int job = blockIdx.x, lcu = job / 3, c = job - lcu * 3, sh = c ? 1 : 0;
    const int lx = lcu % A.cols, ly = lcu / A.cols;
    const int ox = (lx * 64) >> sh, oy = (ly * 64) >> sh;
    const int lw = min(64, A.W - lx * 64) >> sh, lh = min(64, A.H - ly * 64) >> sh;
    const int rs = A.rec[c].stride, ss = A.src[c].stride;
    const uint16_t *r = (const uint16_t *)A.rec[c].p + (size_t)oy * rs + ox, *s = (const uint16_t *)A.src[c].p + (size_t)oy * ss + ox;
    // H3 (same rule as the 8-bit kernel, pinned for 16-bit planes by port_sao_h3_compose16): the last 4 columns / rows of an LCU with a right / bottom
    // neighbour are read from the unfiltered reconstruction
    const int us = A.pre[c].stride;
    const uint16_t *u = A.pre[c].p ? (const uint16_t *)A.pre[c].p + (size_t)oy * us + ox : nullptr;
    const bool has_right = (lx + 1) * 64 < A.W, has_bottom = (ly + 1) * 64 < A.H;
    auto at = [&](int x, int y) -> int {
        const bool strip = u && ((has_right && x >= lw - 4) || (has_bottom && y >= lh - 4));
        return strip ? u[(size_t)y * us + x] : r[(size_t)y * rs + x];
    };
    int *my = s_h + threadIdx.x;
#pragma unroll 4
    for (int k = 0; k < ST_ROWS; ++k) my[k * ST_NT] = 0;
    const int iw = lw - 2, n_items = (lh - 2) * iw;
    for (int i = threadIdx.x; i < n_items; i += ST_NT) {
        const int y = 1 + i / iw, x = 1 + i % iw;
        const int v = at(x, y), d = (int)s[(size_t)y * ss + x] - v;
        const int add = 65536 + d;       // |sum of <= 31 differences| < 2^15
        const int e0 = sgn(v, at(x - 1, y)) + sgn(v, at(x + 1, y)) + 2, e1 = sgn(v, at(x, y - 1)) + sgn(v, at(x, y + 1)) + 2;
        const int e2 = sgn(v, at(x - 1, y - 1)) + sgn(v, at(x + 1, y + 1)) + 2, e3 = sgn(v, at(x + 1, y - 1)) + sgn(v, at(x - 1, y + 1)) + 2;
        my[(0 + e0) * ST_NT] += add; my[(5 + e1) * ST_NT] += add; my[(10 + e2) * ST_NT] += add; my[(15 + e3) * ST_NT] += add;
        my[(20 + min(v >> 5, 31)) * ST_NT] += add;      // (samples above 10 bits -- a caller bug -- must not leave the histogram)
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    svtb200_sao_stats *o = A.out + job;
    for (int row = wid; row < ST_ROWS; row += ST_NT / 32) {
        int dsum = 0, csum = 0;
#pragma unroll
        for (int qd = 0; qd < ST_NT / 32; ++qd) {
            const int w = s_h[row * ST_NT + lane + 32 * qd], dd = (int)(short)(w & 0xffff);
            dsum += dd;
            csum += (w - dd) >> 16;
        }
        dsum = __reduce_add_sync(0xffffffffu, dsum);
        csum = __reduce_add_sync(0xffffffffu, csum);
        if (lane == 0) {
            if (row >= 20) { o->bo_diff[row - 20] = dsum; o->bo_count[row - 20] = (uint16_t)csum; }
            else {
                const int t = row / 5, k = row - t * 5;
                if (k == 2) { o->eo_diff[t][4] = 0; o->eo_count[t][4] = 0; }
                else { const int slot = k < 2 ? k : k - 1; o->eo_diff[t][slot] = dsum; o->eo_count[t][slot] = (uint16_t)csum; }
            }
        }
    }
}

__global__ void __launch_bounds__(256) sao_apply16_kernel(const __grid_constant__ ApplyArgs A)
{
    const int c = blockIdx.z, sh = c ? 1 : 0, pw = A.W >> sh, ph = A.H >> sh;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= pw || y >= ph) return;
    const svtb200_sao_params *pp = A.prm + (size_t)((y << sh) >> 6) * A.cols + ((x << sh) >> 6);
    const int t = A.flag[c ? 1 : 0] ? pp->type[c ? 1 : 0] : 0;
    // the four offsets of this plane in registers (a struct copy indexed by the class would live in local memory)
    const int8_t *po = pp->offset[c];
    const int o0 = po[0], o1 = po[1], o2 = po[2], o3 = po[3];
    const int is = A.in[c].stride;
    const uint16_t *q = (const uint16_t *)A.in[c].p + (size_t)y * is + x;
    int v = q[0];
    if (t == 5) {
        const int b = min(v >> 5, 31) - pp->band_position[c];
        if (b >= 0 && b <= 3) v = min(max(v + (b == 0 ? o0 : b == 1 ? o1 : b == 2 ? o2 : o3), 0), 1023);
    } else if (t >= 1 && t <= 4) {
        const int dx = (t == 2) ? 0 : (t == 4 ? 1 : -1), dy = (t == 1) ? 0 : -1;
        const bool skip = ((t != 1) && (y == 0 || y == ph - 1)) || ((t != 2) && (x == 0 || x == pw - 1));
        if (!skip) {
            const int idx = sgn(v, (int)q[(ptrdiff_t)dy * is + dx]) + sgn(v, (int)q[-(ptrdiff_t)dy * is - dx]) + 2;
            const int off = idx == 0 ? o0 : idx == 1 ? o1 : idx == 3 ? o2 : idx == 4 ? o3 : 0;
            v = min(max(v + off, 0), 1023);
        }
    }
    ((uint16_t *)A.out[c])[(size_t)y * A.ostride[c] + x] = (uint16_t)v;
}

static const svtb200_plane *plane_of(const svtb200_yuv_picture *p, int k) { return k == 0 ? &p->y : k == 1 ? &p->cb : &p->cr; }
static size_t pbytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }

static int sao_check(const svtb200_yuv_picture *a, const svtb200_yuv_picture *b, const char *what)
{
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane *pa = plane_of(a, k), *pb = plane_of(b, k);
        if (!pa->buf || !pb->buf || pa->width != pb->width || pa->height != pb->height || (pa->stride & 3) || (pb->stride & 3) || (pa->origin_x & 3) || (pb->origin_x & 3) ||
            pa->origin_x < 4 || pb->origin_x < 4 || pa->origin_y < 1 || pb->origin_y < 1) {
            set_error("%s: planes must share their size, have pitch and x-origin multiples of 4 and >= 4 samples of padding", what);
            return SVTB200_ERR_BAD_ARG;
        }
    }
    if ((a->y.width & 7) || (a->y.height & 7) || a->y.width < 16 || a->y.height < 16) { set_error("%s: picture size must be a multiple of 8", what); return SVTB200_ERR_BAD_ARG; }
    return SVTB200_OK;
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

static int sao_stats_dev_impl(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out_dev, void *cuda_stream, bool hbd,
                              const svtb200_yuv_picture *unfiltered = nullptr)
{
    const size_t bps = hbd ? 2 : 1;
    if (!src || !recon || !out_dev) { set_error("sao_stats_picture_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = sao_check(src, recon, "sao_stats_picture");
    if (!rc && unfiltered) rc = sao_check(unfiltered, recon, "sao_stats_picture_h3");
    if (rc || (rc = ensure_device())) return rc;
    StatsArgs A;
    for (int k = 0; k < 3; ++k) {
        A.pre[k] = {nullptr, 0};
        if (unfiltered) {
            const svtb200_plane *q = plane_of(unfiltered, k);
            A.pre[k] = {q->buf + ((size_t)q->origin_y * q->stride + q->origin_x) * bps, (int)q->stride};
            if (((uintptr_t)A.pre[k].p) & 3) { set_error("sao_stats_picture_h3: plane interiors must be 4-byte aligned"); return SVTB200_ERR_BAD_ARG; }
        }
        const svtb200_plane *s = plane_of(src, k), *r = plane_of(recon, k);
        A.src[k] = {s->buf + ((size_t)s->origin_y * s->stride + s->origin_x) * bps, (int)s->stride};
        A.rec[k] = {r->buf + ((size_t)r->origin_y * r->stride + r->origin_x) * bps, (int)r->stride};
        if ((((uintptr_t)A.rec[k].p) & 3) || (((uintptr_t)A.src[k].p) & 3)) { set_error("sao_stats_picture: plane interiors must be 4-byte aligned"); return SVTB200_ERR_BAD_ARG; }
    }
    A.out = out_dev; A.W = recon->y.width; A.H = recon->y.height; A.cols = (A.W + 63) / 64;
    const int n = A.cols * ((A.H + 63) / 64) * 3;
    if (hbd) sao_stats16_kernel<<<n, ST_NT, 0, (cudaStream_t)cuda_stream>>>(A); else sao_stats2_kernel<<<n / 3, ST2_NT, 0, (cudaStream_t)cuda_stream>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static int sao_apply_dev_impl(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params_dev,
                              const uint8_t sao_flag[2], void *cuda_stream, bool hbd)
{
    const size_t bps = hbd ? 2 : 1;
    if (!in || !out || !lcu_params_dev || !sao_flag) { set_error("sao_apply_picture_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = sao_check(in, out, "sao_apply_picture");
    if (rc || (rc = ensure_device())) return rc;
    if (in->y.buf == out->y.buf || in->cb.buf == out->cb.buf || in->cr.buf == out->cr.buf) { set_error("sao_apply_picture: in and out must be different pictures"); return SVTB200_ERR_BAD_ARG; }
    ApplyArgs A;
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane *i = plane_of(in, k), *o = plane_of(out, k);
        A.in[k] = {i->buf + ((size_t)i->origin_y * i->stride + i->origin_x) * bps, (int)i->stride};
        A.out[k] = o->buf + ((size_t)o->origin_y * o->stride + o->origin_x) * bps;
        A.ostride[k] = (int)o->stride;
        if ((((uintptr_t)A.in[k].p) & 3) || (((uintptr_t)A.out[k]) & 3)) { set_error("sao_apply_picture: plane interiors must be 4-byte aligned"); return SVTB200_ERR_BAD_ARG; }
    }
    A.prm = lcu_params_dev; A.W = in->y.width; A.H = in->y.height; A.cols = (A.W + 63) / 64; A.flag[0] = sao_flag[0]; A.flag[1] = sao_flag[1];
    if (hbd) { dim3 grid((A.W + 255) / 256, A.H, 3); sao_apply16_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(A); }
    else { dim3 grid((A.W / 4 + 255) / 256, A.H, 3); sao_apply_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(A); }
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static int upload_pic(ThreadCtx *c, const svtb200_yuv_picture *h, svtb200_yuv_picture *d, bool copy, size_t bps = 1)
{
    *d = *h;
    svtb200_plane *dp[3] = {&d->y, &d->cb, &d->cr};
    for (int k = 0; k < 3; ++k) {
        void *p;
        int rc = c->dev_alloc(bps * pbytes(*dp[k]), &p);
        if (rc) return rc;
        if (copy) SVTB200_CUDA_TRY(cudaMemcpyAsync(p, plane_of(h, k)->buf, bps * pbytes(*dp[k]), cudaMemcpyHostToDevice, c->stream));
        dp[k]->buf = (uint8_t *)p;
    }
    return SVTB200_OK;
}
static size_t pic_bytes(const svtb200_yuv_picture *p) { return align_up(pbytes(p->y), 256) + align_up(pbytes(p->cb), 256) + align_up(pbytes(p->cr), 256); }

static int sao_stats_impl(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out, bool hbd, const svtb200_yuv_picture *unfiltered = nullptr)
{
    const size_t bps = hbd ? 2 : 1;
    if (!src || !recon || !out) { set_error("sao_stats_picture: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = sao_check(src, recon, "sao_stats_picture");
    if (!rc && unfiltered) rc = sao_check(recon, unfiltered, "sao_stats_picture_h3");
    if (rc) return rc;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t n = (size_t)((recon->y.width + 63) / 64) * ((recon->y.height + 63) / 64) * 3, ob = n * sizeof(svtb200_sao_stats);
    if ((rc = c->dev_reserve(bps * (pic_bytes(src) + pic_bytes(recon) + (unfiltered ? pic_bytes(unfiltered) : 0)) + ob + 12288))) return rc;
    svtb200_yuv_picture ds, dr, du;
    void *d_out;
    if ((rc = upload_pic(c, src, &ds, true, bps)) || (rc = upload_pic(c, recon, &dr, true, bps)) || (unfiltered && (rc = upload_pic(c, unfiltered, &du, true, bps))) ||
        (rc = c->dev_alloc(ob, &d_out))) return rc;
    if ((rc = sao_stats_dev_impl(&ds, &dr, (svtb200_sao_stats *)d_out, c->stream, hbd, unfiltered ? &du : nullptr))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

static int sao_apply_impl(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params, const uint8_t sao_flag[2], bool hbd)
{
    const size_t bps = hbd ? 2 : 1;
    if (!in || !out || !lcu_params || !sao_flag) { set_error("sao_apply_picture: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = sao_check(in, out, "sao_apply_picture");
    if (rc) return rc;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t n = (size_t)((in->y.width + 63) / 64) * ((in->y.height + 63) / 64), pb = n * sizeof(svtb200_sao_params);
    if ((rc = c->dev_reserve(bps * (pic_bytes(in) + pic_bytes(out)) + pb + 8192))) return rc;
    svtb200_yuv_picture di, dout;
    void *d_prm;
    // the output keeps whatever the caller had in its padding: bring it over, the kernel rewrites the interior only
    if ((rc = upload_pic(c, in, &di, true, bps)) || (rc = upload_pic(c, out, &dout, true, bps)) || (rc = c->dev_alloc(pb, &d_prm))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_prm, lcu_params, pb, cudaMemcpyHostToDevice, c->stream));
    if ((rc = sao_apply_dev_impl(&di, &dout, (const svtb200_sao_params *)d_prm, sao_flag, c->stream, hbd))) return rc;
    const svtb200_plane *hp[3] = {&out->y, &out->cb, &out->cr}, *dp[3] = {&dout.y, &dout.cb, &dout.cr};
    for (int k = 0; k < 3; ++k) SVTB200_CUDA_TRY(cudaMemcpyAsync(hp[k]->buf, dp[k]->buf, bps * pbytes(*hp[k]), cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_sao_stats_picture_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out_dev, void *cuda_stream)
{ return sao_stats_dev_impl(src, recon, out_dev, cuda_stream, false); }
int svtb200_sao_stats_picture(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out) { return sao_stats_impl(src, recon, out, false); }
int svtb200_sao_stats_picture_h3_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered, svtb200_sao_stats *out_dev,
                                     void *cuda_stream)
{
    if (!unfiltered) { set_error("sao_stats_picture_h3_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    return sao_stats_dev_impl(src, deblocked, out_dev, cuda_stream, false, unfiltered);
}
int svtb200_sao_stats_picture_h3(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered, svtb200_sao_stats *out)
{
    if (!unfiltered) { set_error("sao_stats_picture_h3: bad argument"); return SVTB200_ERR_BAD_ARG; }
    return sao_stats_impl(src, deblocked, out, false, unfiltered);
}
int svtb200_sao_apply_picture_dev(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params_dev, const uint8_t sao_flag[2], void *cuda_stream)
{ return sao_apply_dev_impl(in, out, lcu_params_dev, sao_flag, cuda_stream, false); }
int svtb200_sao_apply_picture(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params, const uint8_t sao_flag[2])
{ return sao_apply_impl(in, out, lcu_params, sao_flag, false); }
int svtb200_sao_stats_picture16_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out_dev, void *cuda_stream)
{ return sao_stats_dev_impl(src, recon, out_dev, cuda_stream, true); }
int svtb200_sao_stats_picture16_h3_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered, svtb200_sao_stats *out_dev,
                                       void *cuda_stream)
{
    if (!unfiltered) { set_error("sao_stats_picture16_h3_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    return sao_stats_dev_impl(src, deblocked, out_dev, cuda_stream, true, unfiltered);
}
int svtb200_sao_stats_picture16_h3(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered, svtb200_sao_stats *out)
{
    if (!unfiltered) { set_error("sao_stats_picture16_h3: bad argument"); return SVTB200_ERR_BAD_ARG; }
    return sao_stats_impl(src, deblocked, out, true, unfiltered);
}
int svtb200_sao_stats_picture16(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out) { return sao_stats_impl(src, recon, out, true); }
int svtb200_sao_apply_picture16_dev(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params_dev, const uint8_t sao_flag[2], void *cuda_stream)
{ return sao_apply_dev_impl(in, out, lcu_params_dev, sao_flag, cuda_stream, true); }
int svtb200_sao_apply_picture16(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params, const uint8_t sao_flag[2])
{ return sao_apply_impl(in, out, lcu_params, sao_flag, true); }

}  // extern "C"
