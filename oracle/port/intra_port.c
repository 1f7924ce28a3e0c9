/*
 * oracle/port/intra_port.c -- TEST INFRASTRUCTURE ONLY (see me_port.c header for the rules).
 *
 * Plain-C restatement of the reference's intra prediction, 8-bit 4:2:0, written sample by sample from the HEVC equations:
 *   mode kernels      IntraModePlanar / DCLuma / DCChroma / Vertical* / Horizontal* / Angular_{2,18,34} /
 *                     Angular_{Vertical,Horizontal}_Kernel        C_DEFAULT/EbIntraPrediction_C.c:15-1015
 *                     IntraModeAngular_all and the 3To9 / 11To17 / 19To25 / 27To33 projections   Codec/EbIntraPrediction.c:3080-3600
 *   encode pass       GenerateIntraReferenceSamplesEncodePass (:212; substitution, [1 2 1] / strong smoothing),
 *                     EncodePassIntraPrediction (:4395; filtered-line choice intraLumaFilterTable :60, chroma mode mapping)
 *   open loop         UpdateNeighborSamplesArrayOpenLoop (:5222), IntraPredictionOpenLoop (:5339),
 *                     OpenLoopIntraSearchLcu and helpers           Codec/EbMotionEstimation.c:4458-5330
 * Reference line layout used throughout ("line"): 4N+1 samples, [0] = bottom-most bottom-left sample ... [2N-1] = left of row 0,
 * [2N] = top-left, [2N+1 .. 4N] = top row, then top-right (IntraReferenceSamples_t::yIntraReferenceArray).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "svt_hevc_b200.h"

static const int ANG[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };              /* intraModeAngularTable, EbIntraPrediction.c:25 */
static const int INV_ANG[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };   /* invIntraModeAngularTable :48 */
static const int FILT_THR[4] = { 35, 7, 1, 0 };                             /* intraLumaFilterTable :60 (4x4 .. 32x32) */

static inline int ilog2(int v) { int l = 0; while (v > 1) { v >>= 1; ++l; } return l; }
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clipv(int v, int maxv) { return v < 0 ? 0 : v > maxv ? maxv : v; }

/* sample m of the "main" reference of an angular mode: m = 0 top-left, m > 0 the main side, m < 0 projected from the other side */
static inline int main_ref(const uint16_t *line, int N, int vertical, int m, int inv)
{
    const uint16_t *tl = line + 2 * N;
    if (m >= 0) return vertical ? tl[m] : tl[-m];
    const int j = (128 + (-m) * inv) >> 8;          /* side sample index, 0 = top-left */
    return vertical ? tl[-j] : tl[j];
}

/* one predicted sample; edge = luma block smaller than 32 (DC / horizontal / vertical boundary smoothing) */
static int intra_pel(const uint16_t *line, int N, int mode, int x, int y, int edge, int dc, int maxv)
{
    const uint16_t *tl = line + 2 * N;
#define TOP(i) ((int)tl[1 + (i)])
#define LEFT(i) ((int)tl[-1 - (i)])
    if (mode == 0) {
        const int sh = ilog2(N) + 1;
        return ((N - 1 - x) * LEFT(y) + (x + 1) * TOP(N) + (N - 1 - y) * TOP(x) + (y + 1) * LEFT(N) + N) >> sh;
    }
    if (mode == 1) {
        if (edge) {
            if (x == 0 && y == 0) return (LEFT(0) + TOP(0) + 2 * dc + 2) >> 2;
            if (y == 0) return (TOP(x) + 3 * dc + 2) >> 2;
            if (x == 0) return (LEFT(y) + 3 * dc + 2) >> 2;
        }
        return dc;
    }
    if (mode == 26) return (edge && x == 0) ? clipv(TOP(x) + ((LEFT(y) - (int)tl[0]) >> 1), maxv) : TOP(x);
    if (mode == 10) return (edge && y == 0) ? clipv(LEFT(y) + ((TOP(x) - (int)tl[0]) >> 1), maxv) : LEFT(y);
    const int vertical = mode >= 18;
    const int d = vertical ? mode - 26 : 10 - mode;                 /* 2: +8 ... 18: -8 (either family) */
    const int ang = d < 0 ? -ANG[-d] : ANG[d], inv = INV_ANG[iabs(d)];
    const int along = vertical ? x : y, across = vertical ? y : x;
    const int delta = (across + 1) * ang, di = delta >> 5, df = delta & 31;
    const int a = main_ref(line, N, vertical, along + di + 1, inv);
    if (df == 0) return a;
    const int b = main_ref(line, N, vertical, along + di + 2, inv);
    return ((32 - df) * a + df * b + 16) >> 5;
#undef TOP
#undef LEFT
}

static int dc_value(const uint16_t *line, int N)
{
    int sum = 0;
    for (int i = 0; i < N; ++i) sum += line[2 * N + 1 + i] + line[2 * N - 1 - i];
    return (sum + N) >> (ilog2(N) + 1);
}

/* block into a scratch int array (pitch N) */
static void predict_block(int mode, int N, int is_chroma, const uint16_t *line, int maxv, uint16_t *blk)
{
    const int edge = !is_chroma && N < 32, dc = mode == 1 ? dc_value(line, N) : 0;
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) blk[y * N + x] = (uint16_t)intra_pel(line, N, mode, x, y, edge, dc, maxv);
}

void port_intra_predict(uint32_t mode, uint32_t size, int is_chroma, const uint8_t *line, uint8_t *pred, uint32_t stride)
{
    uint16_t l16[4 * 32 + 1], blk[32 * 32];
    const int N = (int)size;
    for (int i = 0; i < 4 * N + 1; ++i) l16[i] = line[i];
    predict_block((int)mode, N, is_chroma, l16, 255, blk);
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) pred[(size_t)y * stride + x] = (uint8_t)blk[y * N + x];
}

/* ---- encode pass: reference line from a reconstructed plane with per-unit availability ---------------------------------- */
/* unit = samples per availability bit in this plane (4 luma, 2 chroma 4:2:0); bps = bytes per sample (2: 16-bit planes of 10-bit video) */
static void build_line(const uint8_t *plane, uint32_t stride, int bps, int x, int y, int N, int unit, uint32_t left_avail, uint32_t top_avail, int tl_avail, uint16_t *line)
{
    const int nu = 2 * N / unit;            /* units per side */
#define PX(dx, dy) (bps == 2 ? (int)((const uint16_t *)plane)[(ptrdiff_t)(y + (dy)) * (ptrdiff_t)stride + x + (dx)] : (int)plane[(ptrdiff_t)(y + (dy)) * (ptrdiff_t)stride + x + (dx)])
    /* scan order: left units bottom-up, top-left, top units left to right */
    int have = 0, last = bps == 2 ? 512 : 128;          /* the first available sample in scan order pads everything before it */
    for (int u = 0; u < nu && !have; ++u)
        if ((left_avail >> u) & 1) { have = 1; last = PX(-1, 2 * N - 1 - u * unit); }
    if (!have && tl_avail) { have = 1; last = PX(-1, -1); }
    for (int u = 0; u < nu && !have; ++u)
        if ((top_avail >> u) & 1) { have = 1; last = PX(u * unit, -1); }
    if (!have) { for (int i = 0; i < 4 * N + 1; ++i) line[i] = (uint16_t)last; return; }
    for (int i = 0; i < 2 * N; ++i) {                 /* line[i]: left sample of row 2N-1-i */
        if ((left_avail >> (i / unit)) & 1) last = PX(-1, 2 * N - 1 - i);
        line[i] = (uint16_t)last;
    }
    if (tl_avail) last = PX(-1, -1);
    line[2 * N] = (uint16_t)last;
    for (int i = 0; i < 2 * N; ++i) {
        if ((top_avail >> (i / unit)) & 1) last = PX(i, -1);
        line[2 * N + 1 + i] = (uint16_t)last;
    }
#undef PX
}

/* thr: SMOOTHING_THRESHOLD 8 / SMOOTHING_THRESHOLD_10BIT 32 (Codec/EbIntraPrediction.h:119-120) */
static void filter_line(const uint16_t *in, int N, int strong_enabled, int thr, uint16_t *out)
{
    const int bl = in[0], tl = in[2 * N], tr = in[4 * N];
    if (strong_enabled && N >= 32 && iabs(bl + tl - 2 * in[N]) < thr && iabs(tl + tr - 2 * in[3 * N]) < thr) {
        const int sh = ilog2(N) + 1;
        out[0] = in[0]; out[2 * N] = in[2 * N]; out[4 * N] = in[4 * N];
        for (int i = 1; i < 2 * N; ++i) {
            out[i] = (uint16_t)(((2 * N - i) * bl + i * tl + N) >> sh);
            out[2 * N + i] = (uint16_t)(((2 * N - i) * tl + i * tr + N) >> sh);
        }
        return;
    }
    out[0] = in[0]; out[4 * N] = in[4 * N];
    for (int i = 1; i < 4 * N; ++i) out[i] = (uint16_t)((in[i - 1] + 2 * in[i] + in[i + 1] + 2) >> 2);
}

static int chroma_mode_to_luma(int cm, int luma_mode)
{
    switch (cm) { case 0: return 0; case 1: return 26; case 2: return 10; case 3: return 1; default: return luma_mode; }
}

/* hbd: GenerateIntraReference16bitSamplesEncodePass / EncodePassIntraPrediction16bit (Codec/EbIntraPrediction.c:1688,4676): the same process on
 * 16-bit sample planes, midrange 512, smoothing threshold 32, edge filters clipped to 1023 */
static int intra_batch_impl(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs, int hbd)
{
    uint16_t line[4 * 32 + 1], filt[4 * 32 + 1], blk[32 * 32];
    const int bps = hbd ? 2 : 1, maxv = hbd ? 1023 : 255, thr = hbd ? 32 : 8;
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const svtb200_intra_job *t = &jobs[j];
        const int tl = t->flags & SVTB200_INTRA_TOPLEFT_AVAIL, strong = (t->flags & SVTB200_INTRA_STRONG_SMOOTHING) != 0;
        for (int c = 0; c < 3; ++c) {
            const int luma = c == 0;
            if (luma ? t->luma_mode == SVTB200_INTRA_SKIP : t->chroma_mode == SVTB200_INTRA_SKIP) continue;
            const svtb200_plane *rp = c == 0 ? &recon->y : c == 1 ? &recon->cb : &recon->cr, *pp = c == 0 ? &pred->y : c == 1 ? &pred->cb : &pred->cr;
            const int N = luma ? t->size : t->size >> 1, x = luma ? t->x : t->x >> 1, y = luma ? t->y : t->y >> 1;
            const int mode = luma ? t->luma_mode : chroma_mode_to_luma(t->chroma_mode, t->luma_mode);
            build_line(rp->buf + ((size_t)rp->origin_y * rp->stride + rp->origin_x) * bps, rp->stride, bps, x, y, N, luma ? 4 : 2, t->left_avail, t->top_avail, tl, line);
            const uint16_t *use = line;
            if (luma) {
                const int dmin = iabs(mode - 10) < iabs(mode - 26) ? iabs(mode - 10) : iabs(mode - 26);
                if (dmin > FILT_THR[ilog2(N) - 2] && mode != 1) { filter_line(line, N, strong, thr, filt); use = filt; }
            }
            predict_block(mode, N, !luma, use, maxv, blk);      /* into a scratch block: pred may alias recon */
            uint8_t *d = pp->buf + ((size_t)(pp->origin_y + y) * pp->stride + pp->origin_x + x) * bps;
            for (int r = 0; r < N; ++r)
                for (int q = 0; q < N; ++q) {
                    if (hbd) ((uint16_t *)d)[(size_t)r * pp->stride + q] = blk[r * N + q];
                    else d[(size_t)r * pp->stride + q] = (uint8_t)blk[r * N + q];
                }
        }
    }
    return 0;
}
int port_intra_batch(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs)
{ return intra_batch_impl(recon, pred, jobs, n_jobs, 0); }
int port_intra_batch16(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs)
{ return intra_batch_impl(recon, pred, jobs, n_jobs, 1); }

/* ---- open loop ---------------------------------------------------------------------------------------------------------- */
static void open_loop_line(const svtb200_plane *src, int W, int H, int x, int y, int N, uint16_t *line)
{
    const uint8_t *o = src->buf + (size_t)(src->origin_y + y) * src->stride + src->origin_x + x;
    for (int i = 0; i < 4 * N + 1; ++i) line[i] = 128;
    if (x != 0) {
        const int cnt = y + 2 * N > H ? H - y : 2 * N;
        for (int i = 0; i < cnt; ++i) line[2 * N - 1 - i] = o[(ptrdiff_t)i * (ptrdiff_t)src->stride - 1];
    }
    if (x != 0 && y != 0) line[2 * N] = o[-(ptrdiff_t)src->stride - 1];
    if (y != 0) {
        const int cnt = x + 2 * N > W ? W - x : 2 * N;
        for (int i = 0; i < cnt; ++i) line[2 * N + 1 + i] = o[(ptrdiff_t)i - (ptrdiff_t)src->stride];
    }
}

static uint32_t block_sad(const svtb200_plane *src, int x, int y, int N, const uint16_t *blk)
{
    const uint8_t *o = src->buf + (size_t)(src->origin_y + y) * src->stride + src->origin_x + x;
    uint32_t s = 0;
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) s += (uint32_t)iabs((int)o[(size_t)r * src->stride + c] - (int)blk[r * N + c]);
    return s;
}

static uint32_t mode_sad(const svtb200_plane *src, const uint16_t *line, int x, int y, int N, int mode)
{
    uint16_t blk[32 * 32];
    predict_block(mode, N, 0, line, 255, blk);
    return block_sad(src, x, y, N, blk);
}

#define CAND(dist, valid, mode) (((uint32_t)(dist) & 0xFFFFFu) | ((uint32_t)((valid) != 0) << 20) | ((uint32_t)(mode) << 24))
#define SET_MODE(c, m) ((c) = ((c) & 0x00FFFFFFu) | ((uint32_t)(m) << 24))
#define SET_DIST(c, d) ((c) = ((c) & ~0xFFFFFu) | ((uint32_t)(d) & 0xFFFFFu))
#define SET_VALID(c, v) ((c) = ((c) & ~(1u << 20)) | ((uint32_t)((v) != 0) << 20))
#define GET_DIST(c) ((c) & 0xFFFFFu)
#define GET_MODE(c) ((c) >> 24)

static const int ISLICE_MODES[7] = { 0, 1, 10, 26, 2, 18, 34 };            /* iSliceModesArray, EbMotionEstimation.c:148 (first MAX_OIS_0) */
static const int STAGE1_MODES[9] = { 10, 26, 2, 18, 34, 6, 14, 22, 30 };   /* stage1ModesArray :149 */
static const int OIS_TH[3][6][4] = {                                       /* EbHevcOisPointTh :26 */
    { { -20, 50, 150, 200 }, { -20, 50, 150, 200 }, { -20, 50, 100, 150 }, { -20, 50, 200, 300 }, { -20, 50, 200, 300 }, { -20, 50, 200, 300 } },
    { { -150, 0, 150, 200 }, { -150, 0, 150, 200 }, { -125, 0, 100, 150 }, { -50, 50, 200, 300 }, { -50, 50, 200, 300 }, { -50, 50, 200, 300 } },
    { { -400, -300, -200, 0 }, { -400, -300, -200, 0 }, { -400, -300, -200, 0 }, { -400, -300, -200, 0 }, { -400, -300, -200, 0 }, { -400, -300, -200, 0 } } };
/* InjectIntraCandidatesBasedOnBestMode (:4562): [0] = index into stage1SadArray, [1..9] = injected modes */
static const int INJECT[9][10] = {
    { 2, 2, 1, 0, 3, 4, 5, 7, 8, 9 },          /* best = 2  */
    { 0, 10, 1, 0, 9, 11, 8, 12, 7, 13 },      /* best = 10 */
    { 3, 18, 1, 0, 17, 19, 16, 20, 15, 21 },   /* best = 18 */
    { 1, 26, 1, 0, 25, 27, 24, 28, 23, 29 },   /* best = 26 */
    { 4, 34, 1, 0, 33, 32, 29, 31, 27, 28 },   /* best = 34 */
    { 5, 6, 1, 0, 7, 5, 4, 8, 3, 9 },          /* best = 6  */
    { 6, 14, 1, 0, 13, 15, 12, 16, 11, 17 },   /* best = 14 */
    { 7, 22, 1, 0, 21, 23, 20, 24, 19, 25 },   /* best = 22 */
    { 8, 30, 1, 0, 29, 31, 28, 32, 27, 33 } }; /* best = 30 and every other value (default branch) */

static int inject_row(int best)
{
    switch (best) { case 2: return 0; case 10: return 1; case 18: return 2; case 26: return 3; case 34: return 4; case 6: return 5; case 14: return 6; case 22: return 7; default: return 8; }
}

int port_ois_picture(const svtb200_ois_config *cfg, const svtb200_plane *src, const svtb200_me_pu_result *me, uint8_t *total, uint32_t *cand, uint32_t cap)
{
    const int W = cfg->pic_width, H = cfg->pic_height, cols = (W + 63) / 64, rows = (H + 63) / 64;
    const int is_i = cfg->slice_type == 2;
    const uint32_t need = (!is_i && cfg->ois_kernel_level) ? 18 : (!is_i && cfg->limit_ois_to_dc_mode) ? 1 : 9;
    if (cap < need) return -1;
    memset(total, 0, (size_t)cols * rows * 85);
    memset(cand, 0, (size_t)cols * rows * 85 * cap * sizeof(uint32_t));
    uint16_t line[4 * 32 + 1];
    for (int lcu = 0; lcu < cols * rows; ++lcu) {
        const int ox = (lcu % cols) * 64, oy = (lcu / cols) * 64;
        uint32_t stage1[11] = { 0 };     /* stage1SadArray and bestMode live across the CUs of an LCU (:5069,5073) */
        int best_mode = 0;
        const int max_cu = is_i ? 84 : ((cfg->skip_ois_8x8 || cfg->cu8x8_mode == 1) ? 20 : 84);
        for (int cu = 1; cu <= max_cu; ++cu) {
            const int N = cu < 5 ? 32 : cu < 21 ? 16 : 8, k = cu < 5 ? cu - 1 : cu < 21 ? cu - 5 : cu - 21, per = 64 / N;
            const int x = ox + (k % per) * N, y = oy + (k / per) * N, depth = cu < 5 ? 1 : cu < 21 ? 2 : 3;
            uint32_t *c = cand + ((size_t)lcu * 85 + cu) * cap;
            uint8_t *tot = total + (size_t)lcu * 85 + cu;
            if (x + N > W || y + N > H) continue;     /* rasterScanCuValidity */
            open_loop_line(src, W, H, x, y, N, line);
            if (is_i) {
                if (N == 32) { c[0] = CAND(mode_sad(src, line, x, y, N, 0), 1, 0); continue; }
                uint32_t best_sad = 32 * 32 * 255;
                for (int m = 0; m < 7; ++m) {
                    stage1[m] = mode_sad(src, line, x, y, N, ISLICE_MODES[m]);
                    if (stage1[m] < best_sad) { best_mode = ISLICE_MODES[m]; best_sad = stage1[m]; }
                }
                int cnt = 0;
                SET_VALID(c[0], 1); SET_DIST(c[0], stage1[0]);
                SET_MODE(c[cnt], 0); ++cnt; SET_MODE(c[cnt], 1); ++cnt;
                static const int ISL[5][3] = { { 2, 4, 6 }, { 10, 6, 14 }, { 18, 14, 22 }, { 26, 22, 30 }, { 34, 32, 30 } };
                if (best_mode != 0 && best_mode != 1) {
                    const int r = best_mode == 2 ? 0 : best_mode == 10 ? 1 : best_mode == 18 ? 2 : best_mode == 26 ? 3 : 4;
                    for (int q = 0; q < 3; ++q) { SET_MODE(c[cnt], ISL[r][q]); ++cnt; }
                }
                *tot = (uint8_t)cnt;
                continue;
            }
            if (cfg->ois_kernel_level) {
                for (int m = 0; m < 35; ++m) {      /* SortIntraModesOpenLoop :4871: keep the best 18 */
                    const uint32_t sad = mode_sad(src, line, x, y, N, m);
                    if (m < 18) { c[m] = CAND(sad, 0, m); continue; }
                    uint32_t worst = GET_DIST(c[0]); int wi = 0;
                    for (int q = 1; q < 18; ++q) if (GET_DIST(c[q]) > worst) { worst = GET_DIST(c[q]); wi = q; }
                    if (sad < worst) { SET_DIST(c[wi], sad); SET_MODE(c[wi], m); }
                }
                for (int a = 0; a < 18; ++a)        /* SortOisCandidateOpenLoop :4843 */
                    for (int b = a; b < 18; ++b)
                        if (GET_DIST(c[a]) > GET_DIST(c[b])) {
                            const uint32_t da = GET_DIST(c[a]), ma = GET_MODE(c[a]);
                            SET_MODE(c[a], GET_MODE(c[b])); SET_DIST(c[a], GET_DIST(c[b]));
                            SET_MODE(c[b], ma); SET_DIST(c[b], da);
                        }
                *tot = 18;
                continue;
            }
            if (cfg->limit_ois_to_dc_mode) { c[0] = CAND(mode_sad(src, line, x, y, N, 1), 1, 1); *tot = 1; continue; }
            const uint32_t me_sad = me[(size_t)lcu * 85 + cu].distortion[0];
            stage1[0] = mode_sad(src, line, x, y, N, 1);
            const int32_t diff = (int32_t)(me_sad - stage1[0]) * 100;
            const int32_t dist = stage1[0] ? diff / (int32_t)stage1[0] : 0;
            const int *th = OIS_TH[cfg->ois_th_set][cfg->temporal_layer];
            int point = 4;
            if (stage1[0] == 0 || me_sad == 0 || dist <= th[0]) point = 0;
            else if (dist <= th[1]) point = 1;
            else if (dist <= th[2]) point = 2;
            else if (dist <= th[3]) point = 3;
            if (point == 0) { SET_MODE(c[0], 1); SET_DIST(c[0], stage1[0]); *tot = 1; continue; }
            const int n = 1 + 2 * point;
            uint32_t best_sad = 32 * 32 * 255;
            for (int m = 0; m < n; ++m) {
                stage1[m] = mode_sad(src, line, x, y, N, STAGE1_MODES[m]);
                if (stage1[m] < best_sad) { best_mode = STAGE1_MODES[m]; best_sad = stage1[m]; }
            }
            const int *inj = INJECT[inject_row(best_mode)];
            SET_DIST(c[0], stage1[inj[0]]);
            SET_VALID(c[0], cfg->set_best_ois_distortion_to_valid);
            for (int q = 0; q < 9; ++q) SET_MODE(c[q], inj[1 + q]);
            *tot = (uint8_t)n;       /* totalIntraLumaMode[oisPoint][depth >= 1] */
            (void)depth;
        }
    }
    return 0;
}
