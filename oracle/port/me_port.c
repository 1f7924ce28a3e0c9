/*
 * oracle/port/me_port.c -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by
 * the product path (svt-hevc_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C restatement of the reference's motion-estimation arithmetic, written from the
 * reference's behaviour (file:line cited per function), single-threaded, no SIMD.
 * Pinned against the reference compiled in place (oracle/_ref/libsvtref_*.so) by
 * tests/test_oracle_me.py: leaf functions against both the C_DEFAULT and the AVX2/AVX-512 rows,
 * and the picture driver against the reference's own MotionEstimateLcu (oracle/ref_me_harness.c).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "svt_hevc_b200.h"

#define LCU 64
#define MAX_SAD_INIT (64 * 64 * 255) /* MAX_SAD_VALUE, EbMotionEstimation.c:99 */

static inline uint32_t absdiff(uint8_t a, uint8_t b) { return a > b ? (uint32_t)(a - b) : (uint32_t)(b - a); }

/* FastLoop_NxMSadKernel, C_DEFAULT/EbComputeSAD_C.c:141 */
uint32_t port_sad_nxm(const uint8_t *src, uint32_t sstride, const uint8_t *ref, uint32_t rstride,
                      uint32_t h, uint32_t w)
{
    uint32_t acc = 0;
    for (uint32_t r = 0; r < h; ++r)
        for (uint32_t c = 0; c < w; ++c)
            acc += absdiff(src[r * sstride + c], ref[r * rstride + c]);
    return acc;
}

/* CombinedAveragingSAD, C_DEFAULT/EbComputeSAD_C.c:14 : SAD(src, (r1 + r2 + 1) >> 1) */
uint32_t port_sad_avg(const uint8_t *src, uint32_t sstride, const uint8_t *r1, uint32_t r1s,
                      const uint8_t *r2, uint32_t r2s, uint32_t h, uint32_t w)
{
    uint32_t acc = 0;
    for (uint32_t r = 0; r < h; ++r)
        for (uint32_t c = 0; c < w; ++c)
            acc += absdiff(src[r * sstride + c], (uint8_t)((r1[r * r1s + c] + r2[r * r2s + c] + 1) >> 1));
    return acc;
}

/* SadLoopKernel, C_DEFAULT/EbComputeSAD_C.c:165.
 * Raster scan of the search window, strict '<' so the FIRST minimum wins, start value 0xffffff.
 * Block rows use refStride (row-subsampled by the caller); search rows advance by rawStride. */
void port_sad_loop(const uint8_t *src, uint32_t sstride, const uint8_t *ref, uint32_t rstride,
                   uint32_t h, uint32_t w, uint64_t *best, int16_t *bx, int16_t *by,
                   uint32_t rawStride, int16_t saw, int16_t sah)
{
    uint32_t lo = 0xffffff;
    for (int y = 0; y < sah; ++y)
        for (int x = 0; x < saw; ++x) {
            uint32_t s = port_sad_nxm(src, sstride, ref + (size_t)y * rawStride + x, rstride, h, w);
            if (s < lo) { lo = s; *bx = (int16_t)x; *by = (int16_t)y; }
        }
    *best = lo;
}

/* ---- 85-PU bookkeeping -------------------------------------------------------------------
 * n-order of pLcuBestSad / pLcuBestMV (ME_TIER_ZERO_PU_*, EbMotionEstimationContext.h):
 *   [0] 64x64, [1..4] 32x32 raster, [5..20] 16x16 in z-order of 32x32 quadrants,
 *   [21..84] 8x8: four per 16x16, 16x16 blocks in the same z-order.
 * z16(bx,by): index of the 16x16 block at column bx,row by (0..3) in that order
 * (the 16x16_n pictures in C_DEFAULT/EbComputeSAD_C.c:282-292). */
static inline int z16(int bx, int by) { return ((by >> 1) * 2 + (bx >> 1)) * 4 + (by & 1) * 2 + (bx & 1); }

/* 8 rows-subsampled SAD of an 8x8: rows 0,2,4,6 (Subsad8x8, C_DEFAULT/EbComputeSAD_C.c:214;
 * Compute8x4SAD_Kernel :47 with doubled strides in SadCalculation_8x8_16x16) */
static uint32_t sad8x4_sub(const uint8_t *s, uint32_t ss, const uint8_t *r, uint32_t rs)
{
    return port_sad_nxm(s, 2 * ss, r, 2 * rs, 4, 8);
}

static inline uint32_t pack_mv(int16_t x, int16_t y) { return ((uint32_t)(uint16_t)y << 16) | (uint16_t)x; }

/* One search position of the 85-PU update.
 * eight_pos_path selects the two reference code paths, which differ in details that matter:
 *   1: GetEightHorizontalSearchPointResults_8x8_16x16_PU / _32x32_64x64
 *      (C_DEFAULT/EbComputeSAD_C.c:244,355): 16x16 partial sum truncated to u16 before use,
 *      64x64 compare is '<=' (:439);
 *   0: SadCalculation_8x8_16x16 / _32x32_64x64 (C_DEFAULT/EbMeSadCalculation_C.c:16,73):
 *      everything pre-doubled in wide integers, all compares strict '<'. */
static void pu85_update(const uint8_t *src, uint32_t ss, const uint8_t *ref, uint32_t rs,
                        uint32_t mv, int eight_pos_path, uint32_t *bsad, uint32_t *bmv)
{
    uint32_t s16[16];
    for (int by = 0; by < 4; ++by)
        for (int bx = 0; bx < 4; ++bx) {
            const int n = z16(bx, by);
            uint32_t tot = 0;
            for (int q = 0; q < 4; ++q) {
                const int ox = bx * 16 + (q & 1) * 8, oy = by * 16 + (q >> 1) * 8;
                const uint32_t s = sad8x4_sub(src + oy * ss + ox, ss, ref + oy * rs + ox, rs);
                tot += s;
                if (2 * s < bsad[21 + n * 4 + q]) { bsad[21 + n * 4 + q] = 2 * s; bmv[21 + n * 4 + q] = mv; }
            }
            if (eight_pos_path) tot = (uint16_t)tot;
            s16[n] = tot;
            if (2 * tot < bsad[5 + n]) { bsad[5 + n] = 2 * tot; bmv[5 + n] = mv; }
        }
    uint32_t s64 = 0;
    for (int k = 0; k < 4; ++k) {
        const uint32_t s32 = s16[4 * k] + s16[4 * k + 1] + s16[4 * k + 2] + s16[4 * k + 3];
        s64 += s32;
        if (2 * s32 < bsad[1 + k]) { bsad[1 + k] = 2 * s32; bmv[1 + k] = mv; }
    }
    if (eight_pos_path ? (2 * s64 <= bsad[0]) : (2 * s64 < bsad[0])) { bsad[0] = 2 * s64; bmv[0] = mv; }
}

/* FullPelSearch_LCU, EbMotionEstimation.c:586; MV packing currMV = (y << 18) | (u16)(x << 2)
 * from GetEightHorizontalSearchPointResultsAll85PUs_C (:158) / GetSearchPointResults (:453).
 * ref points at the top-left search position (integerBufferPtr + filter-tap offset). */
void port_fullpel_search_lcu(const uint8_t *src, uint32_t ss, const uint8_t *ref, uint32_t rs,
                             int16_t xo, int16_t yo, uint32_t saw, uint32_t sah,
                             uint32_t *bsad, uint32_t *bmv)
{
    const uint32_t w8 = saw & ~7u;
    for (uint32_t y = 0; y < sah; ++y)
        for (uint32_t x = 0; x < saw; ++x) {
            const int16_t mx = (int16_t)((int16_t)(x + xo) * 4), my = (int16_t)((int16_t)(y + yo) * 4);
            pu85_update(src, ss, ref + (size_t)y * rs + x, rs, pack_mv(mx, my), x < w8, bsad, bmv);
        }
}


typedef struct { const uint8_t *buf; int stride, ox, oy, w, h; } plane_t;
static plane_t mkplane(const svtb200_plane *p)
{
    plane_t q = { p->buf, (int)p->stride, p->origin_x, p->origin_y, p->width, p->height };
    return q;
}
static inline const uint8_t *at(const plane_t *p, int x, int y) { return p->buf + (size_t)(p->oy + y) * p->stride + p->ox + x; }

/* ---- sub-pel refinement --------------------------------------------------------------------
 * EbHevcInterpolateSearchRegionAVC :645, PU_HalfPelRefinement :733, EbHevcHalfPelSearch_LCU :1036,
 * SetQuarterPelRefinementInputsOnTheFly :1540, PU_QuarterPelRefinementOnTheFly :1226, QuarterPelSearch_LCU :1623,
 * SuPelEnable :3035; half-sample filter {-2, 18, 18, -2} (+16) >> 5 of C_DEFAULT/EbAvcStyleMcp_C.c:12-82.
 * Plane coordinates, relative to the search-area origin O of the list:
 *   F(x, y) integer sample O + (x, y);  B(c, y) = sample at (c - 1/2, y);  H(x, r) = (x, r - 1/2);  J(c, r) = (c - 1/2, r - 1/2),
 * J being the vertical filter of the ROUNDED B samples, as in the reference.  Stored with a 2-sample apron. */
typedef struct { int saw, sah, sox, soy, pitch; uint8_t *F, *B, *H, *J; } subpel_t;
#define SP_MAX (127 + 63 + 8)
static inline uint8_t *sp_at(uint8_t *pl, int pitch, int x, int y) { return pl + (size_t)(y + 2) * pitch + (x + 2); }
static inline int clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static inline int half4(int a, int b, int c, int d) { return clip255((-2 * a + 18 * b + 18 * c - 2 * d + 16) >> 5); }

static void subpel_build(subpel_t *s, const plane_t *ref, int ox, int oy)
{
    const int pw = s->pitch;
    for (int y = -2; y <= s->sah + 65; ++y)
        for (int x = -2; x <= s->saw + 65; ++x) *sp_at(s->F, pw, x, y) = *at(ref, ox + s->sox + x, oy + s->soy + y);
    for (int y = -2; y <= s->sah + 64; ++y)
        for (int c = 0; c <= s->saw + 64; ++c)
            *sp_at(s->B, pw, c, y) = (uint8_t)half4(*sp_at(s->F, pw, c - 2, y), *sp_at(s->F, pw, c - 1, y), *sp_at(s->F, pw, c, y), *sp_at(s->F, pw, c + 1, y));
    for (int r = 0; r <= s->sah + 63; ++r) {
        for (int x = -1; x <= s->saw + 64; ++x)
            *sp_at(s->H, pw, x, r) = (uint8_t)half4(*sp_at(s->F, pw, x, r - 2), *sp_at(s->F, pw, x, r - 1), *sp_at(s->F, pw, x, r), *sp_at(s->F, pw, x, r + 1));
        for (int c = 0; c <= s->saw + 64; ++c)
            *sp_at(s->J, pw, c, r) = (uint8_t)half4(*sp_at(s->B, pw, c, r - 2), *sp_at(s->B, pw, c, r - 1), *sp_at(s->B, pw, c, r), *sp_at(s->B, pw, c, r + 1));
    }
}

/* SAD of the PU against one plane position / the average of two; method 0 = SUB_SAD_SEARCH (even rows, doubled), 1 = FULL_SAD_SEARCH */
static uint32_t sp_sad(const uint8_t *src, int ss, const uint8_t *p, int ps, int size, int method)
{
    return method == 0 ? port_sad_nxm(src, 2 * ss, p, 2 * ps, size >> 1, size) << 1 : port_sad_nxm(src, ss, p, ps, size, size);
}
static uint32_t sp_sad_avg(const uint8_t *src, int ss, const uint8_t *p1, const uint8_t *p2, int ps, int size, int method)
{
    return method == 0 ? port_sad_avg(src, 2 * ss, p1, 2 * ps, p2, 2 * ps, size >> 1, size) << 1 : port_sad_avg(src, ss, p1, ps, p2, ps, size, size);
}

enum { DIR_TL = 0, DIR_T = 1, DIR_TR = 2, DIR_R = 3, DIR_BR = 4, DIR_B = 5, DIR_BL = 6, DIR_L = 7 };   /* *_POSITION, EbMotionEstimation.c:68-75 */

static void halfpel_pu(subpel_t *s, const uint8_t *src, int ss, int px, int py, int size, int method, uint32_t *bsad, uint32_t *bmv, uint8_t *dir)
{
    const int16_t xmv = (int16_t)(*bmv & 0xffff), ymv = (int16_t)(*bmv >> 16);
    const int xs = (xmv >> 2) - s->sox + px, ys = (ymv >> 2) - s->soy + py, pw = s->pitch;
    /* evaluation order of the reference: L R T B TL TR BR BL */
    const uint8_t *pos[8] = { sp_at(s->B, pw, xs, ys), sp_at(s->B, pw, xs + 1, ys), sp_at(s->H, pw, xs, ys), sp_at(s->H, pw, xs, ys + 1),
                              sp_at(s->J, pw, xs, ys), sp_at(s->J, pw, xs + 1, ys), sp_at(s->J, pw, xs + 1, ys + 1), sp_at(s->J, pw, xs, ys + 1) };
    static const int dx[8] = { -2, 2, 0, 0, -2, 2, 2, -2 }, dy[8] = { 0, 0, -2, 2, -2, -2, 2, 2 };
    uint32_t d[8];
    for (int k = 0; k < 8; ++k) {
        d[k] = sp_sad(src, ss, pos[k], pw, size, method);
        if (d[k] < *bsad) { *bsad = d[k]; *bmv = pack_mv((int16_t)(xmv + dx[k]), (int16_t)(ymv + dy[k])); }
    }
    uint32_t best = d[0];
    for (int k = 1; k < 8; ++k) if (d[k] < best) best = d[k];
    /* first match in the order L R T B TL TR BL BR (:1000-1023) */
    static const int order[8] = { 0, 1, 2, 3, 4, 5, 7, 6 };
    static const uint8_t code[8] = { DIR_L, DIR_R, DIR_T, DIR_B, DIR_TL, DIR_TR, DIR_BR, DIR_BL };
    for (int k = 0; k < 8; ++k) if (d[order[k]] == best) { *dir = code[order[k]]; break; }
}

/* the two planes averaged for each of the 8 quarter positions (order L R T B TL TR BR BL), by the kind of the current best MV:
 * 0 integer, 1 horizontal half, 2 vertical half, 3 diagonal half.  {plane, dx, dy} x 2; planes 0 F, 1 B, 2 H, 3 J. */
static const int8_t QPEL[4][8][6] = {
    { {1,0,0, 0,0,0}, {0,0,0, 1,1,0}, {2,0,0, 0,0,0}, {0,0,0, 2,0,1}, {1,0,0, 2,0,0}, {2,0,0, 1,1,0}, {2,0,1, 1,1,0}, {1,0,0, 2,0,1} },
    { {0,-1,0, 1,0,0}, {1,0,0, 0,0,0}, {3,0,0, 1,0,0}, {1,0,0, 3,0,1}, {2,-1,0, 1,0,0}, {1,0,0, 2,0,0}, {1,0,0, 2,0,1}, {2,-1,1, 1,0,0} },
    { {3,0,0, 2,0,0}, {2,0,0, 3,1,0}, {0,0,-1, 2,0,0}, {2,0,0, 0,0,0}, {1,0,-1, 2,0,0}, {2,0,0, 1,1,-1}, {2,0,0, 1,1,0}, {1,0,0, 2,0,0} },
    { {2,-1,0, 3,0,0}, {3,0,0, 2,0,0}, {1,0,-1, 3,0,0}, {3,0,0, 1,0,0}, {2,-1,0, 1,0,-1}, {1,0,-1, 2,0,0}, {1,0,0, 2,0,0}, {2,-1,0, 1,0,0} } };

static void quarterpel_pu(subpel_t *s, const uint8_t *src, int ss, int px, int py, int size, int method, uint32_t *bsad, uint32_t *bmv, uint8_t dir)
{
    const int16_t xmv = (int16_t)(*bmv & 0xffff), ymv = (int16_t)(*bmv >> 16);
    const int xs = ((xmv + 2) >> 2) - s->sox + px, ys = ((ymv + 2) >> 2) - s->soy + py, pw = s->pitch;
    const int kind = (ymv & 2) + ((xmv & 2) >> 1);
    uint8_t *pl[4] = { s->F, s->B, s->H, s->J };
    /* which of the 8 quarter positions lie towards the best half position (:1260-1285): three neighbours of the direction,
     * mirrored when the current best is itself a half position */
    static const uint8_t ring[8] = { DIR_TL, DIR_T, DIR_TR, DIR_R, DIR_BR, DIR_B, DIR_BL, DIR_L };
    static const uint8_t pos_dir[8] = { DIR_L, DIR_R, DIR_T, DIR_B, DIR_TL, DIR_TR, DIR_BR, DIR_BL };
    static const int dx[8] = { -1, 1, 0, 0, -1, 1, 1, -1 }, dy[8] = { 0, 0, -1, 1, -1, -1, 1, 1 };
    (void)ring;
    for (int k = 0; k < 8; ++k) {
        /* position p is valid when the half-pel direction is p itself or one of its two ring neighbours (integer best),
         * or the opposite of that (half best) */
        int want = pos_dir[k];
        if (kind) want = (want + 4) & 7;
        const int diff = (dir - want) & 7;
        if (!(diff == 0 || diff == 1 || diff == 7)) continue;
        const int8_t *q = QPEL[kind][k];
        const uint32_t d = sp_sad_avg(src, ss, sp_at(pl[q[0]], pw, xs + q[1], ys + q[2]), sp_at(pl[q[3]], pw, xs + q[4], ys + q[5]), pw, size, method);
        if (d < *bsad) { *bsad = d; *bmv = pack_mv((int16_t)(xmv + dx[k]), (int16_t)(ymv + dy[k])); }
    }
}

/* SuPelEnable :3035: which PU sizes get sub-pel refinement on this LCU / list (fractionalSearchModel 1) */
static void supel_enable(const uint32_t *bsad, const uint32_t *bmv, int tl, int *e32, int *e16, int *e8)
{
    int32_t sx[3] = { 0, 0, 0 }, sy[3] = { 0, 0, 0 };
    uint32_t sd[3] = { 0, 0, 0 };
    static const int first[3] = { 1, 5, 21 }, cnt[3] = { 4, 16, 64 }, sh[3] = { 2, 4, 6 };
    uint32_t mag[3], avg[3];
    for (int l = 0; l < 3; ++l) {
        for (int k = 0; k < cnt[l]; ++k) { sx[l] += (int16_t)(bmv[first[l] + k] & 0xffff); sy[l] += (int16_t)(bmv[first[l] + k] >> 16); sd[l] += bsad[first[l] + k]; }
        const uint32_t ax = (uint32_t)(sx[l] >> sh[l]), ay = (uint32_t)(sy[l] >> sh[l]);
        mag[l] = ax * ax + ay * ay;
        avg[l] = sd[l] >> sh[l];
    }
    const uint32_t thr = tl == 1 ? 32 * 32 : tl == 2 ? 80 * 80 : 48 * 48;
    const int m32 = mag[0] < thr, s32 = avg[0] < 32 * 32 * 6, m16 = mag[1] < thr, s16 = avg[1] < 16 * 16 * 2, m8 = mag[2] < thr, s8 = avg[2] < 8 * 8 * 2;
    /* class = 0 (small MV, small SAD), 1 (small MV, large SAD), 2 (large MV, small SAD), 3 */
    const int c32 = (m32 ? 0 : 2) + (s32 ? 0 : 1), c16 = (m16 ? 0 : 2) + (s16 ? 0 : 1), c8 = (m8 ? 0 : 2) + (s8 ? 0 : 1);
    static const uint8_t T32[4][4] = { { 1, 0, 1, 0 }, { 1, 0, 1, 1 }, { 1, 0, 1, 0 }, { 1, 1, 1, 0 } };   /* [layer 0,1,2,other][class] */
    static const uint8_t T16[4] = { 0, 1, 0, 1 };
    const int row = tl > 2 ? 3 : tl;
    *e32 = T32[row][c32];
    *e16 = T16[c16];
    *e8 = (row == 3 && c8 == 3) ? 0 : T16[c8];
}

/* ---- picture driver ------------------------------------------------------------------- */


/* MeEbHevcGetMvdFractionBits, EbMdRateEstimation.c:172 (+ ExponentialGolombBits for k=1) */
static uint32_t egk1_bits(int32_t sym)
{
    /* number of bins of a 1st-order Exp-Golomb code: unary prefix grows the order */
    uint32_t k = 1, n = 0;
    while (sym >= (int32_t)(1u << k)) { sym -= (1 << k); ++k; ++n; }
    return n + 1 + k;
}
static uint32_t mvd_bits(int32_t mx, int32_t my, const uint32_t *t)
{
    const uint32_t ax = (uint32_t)abs(mx), ay = (uint32_t)abs(my);
    const uint32_t xn = ax != 0, yn = ay != 0, xg = ax > 1, yg = ay > 1;
    uint32_t b = t[xn] + t[yn + (2u << xn)];
    if (xn) b += t[xg + 6];
    if (yn) b += t[yg + 6 + (2u << xg)];
    if (xn) b += (xg ? egk1_bits((int32_t)ax - 2) * 32768u : 0) + 32768u;
    if (yn) b += (yg ? egk1_bits((int32_t)ay - 2) * 32768u : 0) + 32768u;
    return b;
}

/* the four-sided window clamp repeated all over EbMotionEstimation.c (e.g. :1892-1930) */
static void clamp_window(int org, int pad, int limit, int *o, int *len)
{
    /* left/top: the reference moves the origin first and then re-tests the SAME condition with the moved
     * origin (e.g. :1892-1900), so the window is shifted but never shortened on this side */
    if (org + *o < -pad) *o = -pad - org;
    if (org + *o > limit - 1) *o -= (org + *o) - (limit - 1);
    if (org + *o + *len > limit) { int l = *len - ((org + *o + *len) - limit); *len = l < 1 ? 1 : l; }
}
/* centre clamp used by TestSearchAreaBounds / EbHevcCheckZeroZeroCenter (:2979-3000) */
static int clamp_center(int org, int c, int pad, int limit)
{
    if (org + c < -pad) c = -pad - org;
    if (org + c > limit - 1) c -= (org + c) - (limit - 1);
    return c;
}

static const int HME_MULT[6][6] = { /* HME_LEVEL_0_SEARCH_AREA_MULTIPLIER_X == _Y, EbDefinitions.h:1322-1340 */
    { 100 }, { 100, 100 }, { 100, 100, 100 }, { 200, 140, 100, 70 },
    { 350, 200, 100, 100, 100 }, { 525, 350, 200, 100, 100, 100 } };

typedef struct { uint64_t sad; int16_t x, y; } hme_res;

/* LCU-sized SAD on every other row, doubled: the "subsampleSad = 1" calls in
 * TestSearchAreaBounds (:3397) and EbHevcCheckZeroZeroCenter (:2964) */
static uint64_t lcu_sad_sub(const plane_t *s, const plane_t *r, int ox, int oy, int lw, int lh, int dx, int dy)
{
    return (uint64_t)port_sad_nxm(at(s, ox, oy), 2 * s->stride, at(r, ox + dx, oy + dy), 2 * r->stride, lh >> 1, lw) << 1;
}

int port_me_picture(const svtb200_me_config *cfg, const svtb200_pa_picture *srcp,
                    const svtb200_pa_picture *ref0p, const svtb200_pa_picture *ref1p,
                    svtb200_me_pu_result *results, uint32_t *lcu_best_sad, uint32_t *lcu_best_mv,
                    uint32_t *rc_me_distortion)
{
    static const int tab32[16] = { 0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15 }; /* tab32x32, :101 */
    static const int tab8[64] = { 0, 1, 4, 5, 16, 17, 20, 21, 2, 3, 6, 7, 18, 19, 22, 23, 8, 9, 12, 13, 24, 25, 28, 29,
        10, 11, 14, 15, 26, 27, 30, 31, 32, 33, 36, 37, 48, 49, 52, 53, 34, 35, 38, 39, 50, 51, 54, 55,
        40, 41, 44, 45, 56, 57, 60, 61, 42, 43, 46, 47, 58, 59, 62, 63 };                   /* tab8x8, :102 */
    if (cfg->fractional_search_model > 2 || cfg->fractional_search_method > 1) return SVTB200_ERR_UNSUPPORTED;   /* SSD_SEARCH (enc modes <= 4) is not restated */
    const int subpel = cfg->fractional_search_model != 2, spm = cfg->fractional_search_method;
    const int W = cfg->pic_width, H = cfg->pic_height;
    const int cols = (W + 63) / 64, rows = (H + 63) / 64;
    const plane_t sF = mkplane(&srcp->full), sQ = mkplane(&srcp->quarter), sS = mkplane(&srcp->sixteenth);
    const svtb200_pa_picture *refs[2] = { ref0p, ref1p ? ref1p : ref0p };
    const int nl = cfg->num_lists;
    const int mult = HME_MULT[cfg->hierarchical_levels][cfg->temporal_layer];
    /* per-LCU state persists across LCUs exactly like MeContext_t does (stale list-1 values on P pictures) */
    uint32_t bsad[2][85], bmv[2][85], bisad[85];
    memset(bsad, 0, sizeof bsad); memset(bmv, 0, sizeof bmv); memset(bisad, 0, sizeof bisad);
    subpel_t sp[2];
    uint8_t spdir[85];      /* psubPelDirection*: one set shared by both lists, like MeContext_t */
    memset(spdir, 0, sizeof spdir);
    memset(sp, 0, sizeof sp);
    for (int li = 0; li < 2 && subpel; ++li) {
        sp[li].pitch = SP_MAX;
        sp[li].F = (uint8_t *)calloc(4, (size_t)SP_MAX * SP_MAX);
        sp[li].B = sp[li].F + (size_t)SP_MAX * SP_MAX; sp[li].H = sp[li].B + (size_t)SP_MAX * SP_MAX; sp[li].J = sp[li].H + (size_t)SP_MAX * SP_MAX;
    }

    for (int ly = 0; ly < rows; ++ly)
    for (int lx = 0; lx < cols; ++lx) {
        const int lcu = ly * cols + lx, ox = lx * 64, oy = ly * 64;
        const int lw = W - ox < 64 ? W - ox : 64, lh = H - oy < 64 ? H - oy : 64;
        int sa_ox[2] = { 0, 0 }, sa_oy[2] = { 0, 0 };
        plane_t rFs[2];

        for (int li = 0; li < nl; ++li) {
            const plane_t rF = mkplane(&refs[li]->full), rQ = mkplane(&refs[li]->quarter), rS = mkplane(&refs[li]->sixteenth);
            rFs[li] = rF;
            int cx = 0, cy = 0;
            if (cfg->temporal_layer > 0 || li == 0) {
                if (cfg->update_hme_search_center) {
                    /* TestSearchAreaBounds, EbMotionEstimation.c:3363-3665.  Candidates: zero, A(-w,0), B(+w,0),
                     * C(0,-h), D(0,+h), and for list 1 the mirrored list-0 64x64 MV.  The SAD is taken at the
                     * CLAMPED position but the UNCLAMPED vector is returned; cost = sad<<8 (+ MD_OFFSET>>MD_SHIFT
                     * = 0 for the non-zero candidates); ties resolve in the order zero,A,B,C,direct,D. */
                    const int tw = cfg->hme_l0_total_w, th = cfg->hme_l0_total_h;
                    int cand[6][2] = { { 0, 0 }, { -tw, 0 }, { tw, 0 }, { 0, -th }, { 0, 0 }, { 0, th } };
                    uint64_t cost[6];
                    if (li == 1) {
                        cand[4][0] = 0 - ((int16_t)(bmv[0][0] & 0xffff) >> 2);
                        cand[4][1] = 0 - ((int16_t)(bmv[0][0] >> 16) >> 2);
                    }
                    for (int k = 0; k < 6; ++k) {
                        if (k == 4 && li == 0) { cost[k] = 0xFFFFFFFFFFFFFull; continue; }
                        const int dx = k ? clamp_center(ox, cand[k][0], 63, rF.w) : 0;
                        const int dy = k ? clamp_center(oy, cand[k][1], 63, rF.h) : 0;
                        cost[k] = lcu_sad_sub(&sF, &rF, ox, oy, lw, lh, dx, dy) << 8;
                    }
                    int best = 0;
                    for (int k = 1; k < 6; ++k) if (cost[k] < cost[best]) best = k;
                    cx = cand[best][0]; cy = cand[best][1];
                }
                if (cfg->enable_hme && lh == 64) {
                    /* HME, EbMotionEstimation.c:3807-4062 */
                    hme_res l0[2][2], l1[2][2], l2[2][2];
                    const int nw = cfg->n_hme_regions_w < 2 ? cfg->n_hme_regions_w : 2;
                    const int nh = cfg->n_hme_regions_h < 2 ? cfg->n_hme_regions_h : 2;
                    const int sh0 = cfg->update_hme_search_center ? 2 : 0, sh1 = cfg->update_hme_search_center ? 1 : 0;
                    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
                        l0[a][b].sad = l1[a][b].sad = l2[a][b].sad = 0;
                        l0[a][b].x = (int16_t)(cx >> sh0); l0[a][b].y = (int16_t)(cy >> sh0);
                        l1[a][b].x = (int16_t)(cx >> sh1); l1[a][b].y = (int16_t)(cy >> sh1);
                        l2[a][b].x = (int16_t)cx; l2[a][b].y = (int16_t)cy;
                    }
                    /* sixteenth-resolution source block: every other row, pitch 16 (EbMotionEstimationProcess.c:757-771) */
                    uint8_t s16buf[16 * 8];
                    memset(s16buf, 0, sizeof s16buf);
                    for (int r = 0; r < (lh >> 2); r += 2) memcpy(s16buf + (r >> 1) * 16, at(&sS, ox >> 2, (oy >> 2) + r), (size_t)(lw >> 2));
                    uint8_t qbuf[32 * 32];
                    memset(qbuf, 0, sizeof qbuf);
                    for (int r = 0; r < (lh >> 1); ++r) memcpy(qbuf + r * 32, at(&sQ, ox >> 1, (oy >> 1) + r), (size_t)(lw >> 1));

                    if (cfg->enable_hme_l0) {
                        const int one = cfg->one_quadrant_hme && !cfg->enable_hme_l1 && !cfg->enable_hme_l2;
                        for (int qy = 0; qy < (one ? 1 : nh); ++qy) for (int qx = 0; qx < (one ? 1 : nw); ++qx) {
                            /* EbHevcHmeOneQuadrantLevel0 :1847 / EbHevcHmeLevel0 :2012 */
                            const int o4x = ox >> 2, o4y = oy >> 2;
                            int saw, sah, sox, soy;
                            if (one) {
                                saw = (cfg->hme_l0_total_w * mult) / 100; sah = (cfg->hme_l0_total_h * mult) / 100;
                                sox = -(saw >> 1) + (cx >> 2); soy = -(sah >> 1) + (cy >> 2);
                            } else {
                                saw = (cfg->hme_l0_w[qx] * mult) / 100; sah = (cfg->hme_l0_h[qy] * mult) / 100;
                                int dx = cx >> 2, dy = cy >> 2;
                                for (int k = 0; k < qx; ++k) dx += (cfg->hme_l0_w[k] * mult) / 100;
                                for (int k = 0; k < qy; ++k) dy += (cfg->hme_l0_h[k] * mult) / 100;
                                sox = -(((cfg->hme_l0_total_w * mult) / 100) >> 1) + dx;
                                soy = -(((cfg->hme_l0_total_h * mult) / 100) >> 1) + dy;
                            }
                            sox = (int16_t)sox; soy = (int16_t)soy;
                            clamp_window(o4x, rS.ox - 1, rS.w, &sox, &saw);
                            clamp_window(o4y, rS.oy - 1, rS.h, &soy, &sah);
                            if (one && (saw & 15)) saw = (saw >> 4) << 4; /* :1933-1936 */
                            hme_res *o = &l0[qx][qy];
                            port_sad_loop(s16buf, 16, at(&rS, o4x + sox, o4y + soy), 2 * rS.stride, (uint32_t)(lh >> 2) >> 1, (uint32_t)(lw >> 2),
                                          &o->sad, &o->x, &o->y, (uint32_t)rS.stride, (int16_t)saw, (int16_t)sah);
                            o->sad *= 2;
                            o->x = (int16_t)((int16_t)(o->x + sox) * 4); o->y = (int16_t)((int16_t)(o->y + soy) * 4);
                        }
                    }
                    if (cfg->enable_hme_l1)
                        for (int qy = 0; qy < nh; ++qy) for (int qx = 0; qx < nw; ++qx) {
                            /* EbHevcHmeLevel1 :2194 */
                            const int o2x = ox >> 1, o2y = oy >> 1;
                            int w1 = cfg->hme_l1_w[qx];
                            int saw = w1 < 8 ? 8 : (w1 & 7) ? w1 + (w1 - ((w1 >> 3) << 3)) : w1, sah = cfg->hme_l1_h[qy];
                            int sox = -(saw >> 1) + (l0[qx][qy].x >> 1), soy = -(sah >> 1) + (l0[qx][qy].y >> 1);
                            sox = (int16_t)sox; soy = (int16_t)soy;
                            clamp_window(o2x, rQ.ox - 1, rQ.w, &sox, &saw);
                            clamp_window(o2y, rQ.oy - 1, rQ.h, &soy, &sah);
                            hme_res *o = &l1[qx][qy];
                            port_sad_loop(qbuf, 64, at(&rQ, o2x + sox, o2y + soy), 2 * rQ.stride, (uint32_t)(lh >> 1) >> 1, (uint32_t)(lw >> 1),
                                          &o->sad, &o->x, &o->y, (uint32_t)rQ.stride, (int16_t)saw, (int16_t)sah);
                            o->sad *= 2;
                            o->x = (int16_t)((int16_t)(o->x + sox) * 2); o->y = (int16_t)((int16_t)(o->y + soy) * 2);
                        }
                    if (cfg->enable_hme_l2)
                        for (int qy = 0; qy < nh; ++qy) for (int qx = 0; qx < nw; ++qx) {
                            /* EbHevcHmeLevel2 :2315 */
                            int w2 = cfg->hme_l2_w[qx];
                            int saw = w2 < 8 ? 8 : (w2 & 7) ? w2 + (w2 - ((w2 >> 3) << 3)) : w2, sah = cfg->hme_l2_h[qy];
                            int sox = -(saw >> 1) + l1[qx][qy].x, soy = -(sah >> 1) + l1[qx][qy].y;
                            sox = (int16_t)sox; soy = (int16_t)soy;
                            clamp_window(ox, 63, rF.w, &sox, &saw);
                            clamp_window(oy, 63, rF.h, &soy, &sah);
                            hme_res *o = &l2[qx][qy];
                            port_sad_loop(at(&sF, ox, oy), 2 * sF.stride, at(&rF, ox + sox, oy + soy), 2 * rF.stride, (uint32_t)lh >> 1, (uint32_t)lw,
                                          &o->sad, &o->x, &o->y, (uint32_t)rF.stride, (int16_t)saw, (int16_t)sah);
                            o->sad *= 2;
                            o->x = (int16_t)(o->x + sox); o->y = (int16_t)(o->y + soy);
                        }
                    /* pick the best region of the last enabled level, first-minimum in width-fastest order (:3959-4030) */
                    int hx = 0, hy = 0;
                    hme_res (*lvl)[2] = cfg->enable_hme_l2 ? l2 : cfg->enable_hme_l1 ? l1 : (cfg->enable_hme_l0 ? l0 : NULL);
                    if (lvl) {
                        uint64_t bs = lvl[0][0].sad; hx = lvl[0][0].x; hy = lvl[0][0].y;
                        const int onlyone = (lvl == l0) && cfg->one_quadrant_hme;
                        if (!onlyone)
                            for (int qy = 0; qy < cfg->n_hme_regions_h; ++qy)
                                for (int qx = (qy == 0 ? 1 : 0); qx < cfg->n_hme_regions_w; ++qx)
                                    if (lvl[qx][qy].sad < bs) { bs = lvl[qx][qy].sad; hx = lvl[qx][qy].x; hy = lvl[qx][qy].y; }
                        if (lvl == l2 && cfg->ref_poc[0] == cfg->ref_poc[1] && li == 1 &&
                            cfg->n_hme_regions_w * cfg->n_hme_regions_h > 1) {
                            /* same-POC second list: take the 2nd best region after a selection sort whose
                             * [q / nw][q % nw] indexing is kept as in the reference (:4032-4060) */
                            const int nq = cfg->n_hme_regions_w * cfg->n_hme_regions_h, nqw = cfg->n_hme_regions_w;
                            for (int a = 0; a < nq - 1; ++a)
                                for (int b = a + 1; b < nq; ++b)
                                    if (l2[a / nqw][a % nqw].sad > l2[b / nqw][b % nqw].sad) {
                                        hme_res t = l2[a / nqw][a % nqw]; l2[a / nqw][a % nqw] = l2[b / nqw][b % nqw]; l2[b / nqw][b % nqw] = t;
                                    }
                            hx = l2[0][1].x; hy = l2[0][1].y;
                        }
                    }
                    cx = hx; cy = hy;
                }
            }
            int saw = cfg->search_area_w < 127 ? cfg->search_area_w : 127, sah = cfg->search_area_h < 127 ? cfg->search_area_h : 127;
            if (cx != 0 || cy != 0) {
                /* EbHevcCheckZeroZeroCenter :2946 */
                const uint64_t zc = lcu_sad_sub(&sF, &rF, ox, oy, lw, lh, 0, 0) << 8;
                if (cfg->update_hme_search_center) { cx = clamp_center(ox, cx, 63, rF.w); cy = clamp_center(oy, cy, 63, rF.h); }
                const uint64_t hs = lcu_sad_sub(&sF, &rF, ox, oy, lw, lh, cx, cy);
                /* the reference shifts first and takes ABS afterwards; mvd_bits() only uses |.| */
                const uint32_t rate = mvd_bits((int32_t)((int16_t)cx << 2), (int32_t)((int16_t)cy << 2), cfg->mvd_bits);
                const uint64_t hc = (hs << 8) + (((uint64_t)cfg->lambda * rate + (1u << 22)) >> 23); /* MD_OFFSET, MD_SHIFT = 15 + 16 - 8 */
                if ((zc < hc ? zc : hc) == zc) { cx = 0; cy = 0; }
            }
            int sox = (int16_t)(cx - (saw >> 1)), soy = (int16_t)(cy - (sah >> 1));
            clamp_window(ox, 63, W, &sox, &saw);
            clamp_window(oy, 63, H, &soy, &sah);
            sa_ox[li] = sox; sa_oy[li] = soy;
            for (int k = 0; k < 85; ++k) bsad[li][k] = MAX_SAD_INIT;
            port_fullpel_search_lcu(at(&sF, ox, oy), (uint32_t)sF.stride, at(&rF, ox + sox, oy + soy), (uint32_t)rF.stride,
                                    (int16_t)sox, (int16_t)soy, (uint32_t)saw, (uint32_t)sah, bsad[li], bmv[li]);
            if (subpel) {
                /* MotionEstimateLcu :4242-4316 */
                int e32 = 1, e16 = 1, e8 = 1;
                if (cfg->fractional_search_model == 1) supel_enable(bsad[li], bmv[li], cfg->temporal_layer, &e32, &e16, &e8);
                e16 = e16 && cfg->cu16x16_mode == 0;
                const int no8 = cfg->cu8x8_mode == 1;
                subpel_t *S = &sp[li];
                S->saw = saw; S->sah = sah; S->sox = sox; S->soy = soy;
                subpel_build(S, &rF, ox, oy);
                const uint8_t *src = at(&sF, ox, oy);
                for (int pass = 0; pass < 2; ++pass) {       /* all half-pel refinements, then all quarter-pel ones */
                    for (int k = 0; k < 84; ++k) {
                        int size, px, py, n;
                        if (k < 4) { if (!e32) continue; size = 32; px = (k & 1) * 32; py = (k >> 1) * 32; n = 1 + k; }
                        else if (k < 20) { if (!e16) continue; const int q = k - 4; size = 16; px = (q & 3) * 16; py = (q >> 2) * 16; n = 5 + tab32[q]; }
                        else { if (!e8 || no8) continue; const int q = k - 20; size = 8; px = (q & 7) * 8; py = (q >> 3) * 8; n = 21 + tab8[q]; }
                        const uint8_t *ps = src + (size_t)py * sF.stride + px;
                        if (pass == 0) halfpel_pu(S, ps, sF.stride, px, py, size, spm, &bsad[li][n], &bmv[li][n], &spdir[n]);
                        else quarterpel_pu(S, ps, sF.stride, px, py, size, spm, &bsad[li][n], &bmv[li][n], spdir[n]);
                    }
                }
            }
        }
        (void)sa_ox; (void)sa_oy;

        /* candidates per PU, EbMotionEstimation.c:4321-4438 */
        static const uint8_t pux21[21] = { 0, 0,32,0,32, 0,16,32,48,0,16,32,48,0,16,32,48,0,16,32,48 }, puy21[21] = { 0, 0,0,32,32, 0,0,0,0,16,16,16,16,32,32,32,32,48,48,48,48 };
        for (int pu = 0; pu < 85; ++pu) {
            const int n = pu > 20 ? tab8[pu - 21] + 21 : pu > 4 ? tab32[pu - 5] + 5 : pu;
            svtb200_me_pu_result *o = &results[(size_t)lcu * 85 + pu];
            int total = nl;
            if (nl == 2 && (cfg->cu8x8_mode == 0 || pu < 21) && (cfg->cu16x16_mode == 0 || pu < 5)) {
                /* integer-pel EbHevcBiPredictionCompensation :2742 -> EbHevcBiPredAverging :2600 (SUB_SAD_SEARCH) */
                const int sz = pu == 0 ? 64 : pu < 5 ? 32 : pu < 21 ? 16 : 8;
                const int px = pu < 21 ? pux21[pu] : ((pu - 21) & 7) * 8, py = pu < 21 ? puy21[pu] : ((pu - 21) >> 3) * 8;
                const int x0 = (int16_t)(bmv[0][n] & 0xffff) >> 2, y0 = (int16_t)(bmv[0][n] >> 16) >> 2;
                const int x1 = (int16_t)(bmv[1][n] & 0xffff) >> 2, y1 = (int16_t)(bmv[1][n] >> 16) >> 2;
                if (!subpel) {
                    const uint8_t *ps = at(&sF, ox + px, oy + py), *p0 = at(&rFs[0], ox + px + x0, oy + py + y0), *p1 = at(&rFs[1], ox + px + x1, oy + py + y1);
                    bisad[n] = spm == 0 ? port_sad_avg(ps, 2 * sF.stride, p0, 2 * rFs[0].stride, p1, 2 * rFs[1].stride, (uint32_t)sz >> 1, (uint32_t)sz) << 1
                                        : port_sad_avg(ps, sF.stride, p0, rFs[0].stride, p1, rFs[1].stride, (uint32_t)sz, (uint32_t)sz);
                } else {
                    /* EbHevcBiPredictionCompensation :2742, SelectBuffer :2440, QuarterPelCompensation :2490: each list's prediction is a
                     * plane or the rounded average of two planes, chosen by the fractional position */
                    static uint8_t tmp[2][64 * 64];
                    const uint8_t *pp[2]; int pst[2];
                    for (int l = 0; l < 2; ++l) {
                        const int16_t mx = (int16_t)(bmv[l][n] & 0xffff), my = (int16_t)(bmv[l][n] >> 16);
                        const int ix = (mx >> 2) - sp[l].sox + px, iy = (my >> 2) - sp[l].soy + py, fp = (mx & 3) + ((my & 3) << 2), pw = sp[l].pitch;
                        uint8_t *F = sp[l].F, *B = sp[l].B, *Hh = sp[l].H, *J = sp[l].J;
                        const uint8_t *a = NULL, *b = NULL;
                        switch (fp) {
                        case 0: a = sp_at(F, pw, ix, iy); break;
                        case 2: a = sp_at(B, pw, ix + 1, iy); break;
                        case 8: a = sp_at(Hh, pw, ix, iy + 1); break;
                        case 10: a = sp_at(J, pw, ix + 1, iy + 1); break;
                        case 1: a = sp_at(F, pw, ix, iy); b = sp_at(B, pw, ix + 1, iy); break;
                        case 3: a = sp_at(B, pw, ix + 1, iy); b = sp_at(F, pw, ix + 1, iy); break;
                        case 4: a = sp_at(F, pw, ix, iy); b = sp_at(Hh, pw, ix, iy + 1); break;
                        case 5: a = sp_at(B, pw, ix + 1, iy); b = sp_at(Hh, pw, ix, iy + 1); break;
                        case 6: a = sp_at(B, pw, ix + 1, iy); b = sp_at(J, pw, ix + 1, iy + 1); break;
                        case 7: a = sp_at(B, pw, ix + 1, iy); b = sp_at(Hh, pw, ix + 1, iy + 1); break;
                        case 9: a = sp_at(Hh, pw, ix, iy + 1); b = sp_at(J, pw, ix + 1, iy + 1); break;
                        case 11: a = sp_at(J, pw, ix + 1, iy + 1); b = sp_at(Hh, pw, ix + 1, iy + 1); break;
                        case 12: a = sp_at(Hh, pw, ix, iy + 1); b = sp_at(F, pw, ix, iy + 1); break;
                        case 13: a = sp_at(Hh, pw, ix, iy + 1); b = sp_at(B, pw, ix + 1, iy + 1); break;
                        case 14: a = sp_at(J, pw, ix + 1, iy + 1); b = sp_at(B, pw, ix + 1, iy + 1); break;
                        default: a = sp_at(Hh, pw, ix + 1, iy + 1); b = sp_at(B, pw, ix + 1, iy + 1); break;
                        }
                        if (!b) { pp[l] = a; pst[l] = pw; }
                        else {
                            for (int r = 0; r < sz; ++r)
                                for (int c = 0; c < sz; ++c) tmp[l][r * 64 + c] = (uint8_t)((a[(size_t)r * pw + c] + b[(size_t)r * pw + c] + 1) >> 1);
                            pp[l] = tmp[l]; pst[l] = 64;
                        }
                    }
                    const uint8_t *ps = at(&sF, ox + px, oy + py);
                    bisad[n] = spm == 0 ? port_sad_avg(ps, 2 * sF.stride, pp[0], 2 * pst[0], pp[1], 2 * pst[1], (uint32_t)sz >> 1, (uint32_t)sz) << 1
                                        : port_sad_avg(ps, sF.stride, pp[0], pst[0], pp[1], pst[1], (uint32_t)sz, (uint32_t)sz);
                }
                total = 3;
            }
            o->total_me_candidates = (uint8_t)total;
            o->x_mv_l0 = (int16_t)(bmv[0][n] & 0xffff); o->y_mv_l0 = (int16_t)(bmv[0][n] >> 16);
            o->x_mv_l1 = (int16_t)(bmv[1][n] & 0xffff); o->y_mv_l1 = (int16_t)(bmv[1][n] >> 16);
            const uint32_t a = bsad[0][n], b = bsad[1][n], c = bisad[n];
            /* untouched slots keep whatever the previous picture left there in the reference; the
             * harness starts from zeroed storage, so zero is the comparable value */
            uint32_t d[3] = { 0, 0, 0 }; uint8_t dir[3] = { 0, 0, 0 };
            if (total == 3) {
                /* Sort3Elements :2919 */
                if (a <= b && a <= c) { d[0] = a; dir[0] = 0; if (b <= c) { d[1] = b; dir[1] = 1; d[2] = c; dir[2] = 2; } else { d[1] = c; dir[1] = 2; d[2] = b; dir[2] = 1; } }
                else if (b <= a && b <= c) { d[0] = b; dir[0] = 1; if (a <= c) { d[1] = a; dir[1] = 0; d[2] = c; dir[2] = 2; } else { d[1] = c; dir[1] = 2; d[2] = a; dir[2] = 0; } }
                else if (a <= b) { d[0] = c; dir[0] = 2; d[1] = a; dir[1] = 0; d[2] = b; dir[2] = 1; }
                else { d[0] = c; dir[0] = 2; d[1] = b; dir[1] = 1; d[2] = a; dir[2] = 0; }
            } else if (total == 2) {
                if (a <= b) { d[0] = a; dir[0] = 0; d[1] = b; dir[1] = 1; } else { d[0] = b; dir[0] = 1; d[1] = a; dir[1] = 0; }
            } else { d[0] = a; dir[0] = 0; }
            for (int k = 0; k < 3; ++k) { o->distortion[k] = d[k]; o->direction[k] = dir[k]; }
        }
        if (rc_me_distortion) {
            rc_me_distortion[lcu] = 0;
            if (cfg->rate_control_mode) for (int i = 0; i < 16; ++i) rc_me_distortion[lcu] += results[(size_t)lcu * 85 + 5 + i].distortion[0];
        }
        for (int li = 0; li < 2; ++li) {
            if (lcu_best_sad) memcpy(&lcu_best_sad[((size_t)lcu * 2 + li) * 85], bsad[li], sizeof bsad[li]);
            if (lcu_best_mv) memcpy(&lcu_best_mv[((size_t)lcu * 2 + li) * 85], bmv[li], sizeof bmv[li]);
        }
    }
    for (int li = 0; li < 2 && subpel; ++li) free(sp[li].F);
    return 0;
}
