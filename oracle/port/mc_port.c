/*
 * oracle/port/mc_port.c -- TEST INFRASTRUCTURE ONLY (see me_port.c header for the rules).
 *
 * Plain-C restatement of the reference's HEVC inter-prediction sample interpolation, 8-bit 4:2:0:
 *   luma  8-tap quarter-pel  (C_DEFAULT/EbMcp_C.c:221-1214, tables Codec/EbMcpTables.c:28-151)
 *   chroma 4-tap eighth-pel  (C_DEFAULT/EbMcp_C.c:11-19, :3789, :4888-5099)
 *   bi-pred combine          (BiPredClipping, C_DEFAULT/EbMcp_C.c:21; integer/integer shortcut Codec/EbMcp.c:603)
 *   PU driver                (EncodePassInterPrediction Codec/EbInterPrediction.c:761,
 *                             EncodeUniPredInterpolation / EncodeBiPredInterpolation Codec/EbMcp.c:175,562)
 * The reference keeps its 14-bit intermediates biased by -8192 (MinusOffset1 / Offset4 / Offset5,
 * Codec/EbDefinitions.h:989-1039); the bias cancels exactly, so this file uses the unbiased textbook form
 * and tests/test_oracle_mc.py checks it sample by sample against the reference's kernels.
 */
#include <stdint.h>
#include <stdlib.h>

#include "svt_hevc_b200.h"

static const int LUMA_TAPS[4][8] = {
    { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 }, { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
static const int CHROMA_TAPS[8][4] = {
    { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 }, { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

static inline uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* Interpolate one w x h block.  p = integer sample of the block origin, taps = 8 (luma) or 4 (chroma),
 * fx/fy = fractional phase index into the tap table.  uni != 0: final 8-bit samples into dst8;
 * uni == 0: 14-bit intermediate (unbiased) into raw (pitch w). */
static void interp_block(const uint8_t *p, int stride, int taps, const int *cx, const int *cy, int fx, int fy,
                         int w, int h, int uni, uint8_t *dst8, int dstride, int16_t *raw)
{
    const int half = taps / 2 - 1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int v;
            if (!fx && !fy) v = p[y * stride + x] << 6;
            else if (!fy) { v = 0; for (int t = 0; t < taps; ++t) v += cx[t] * p[y * stride + x + t - half]; }
            else if (!fx) { v = 0; for (int t = 0; t < taps; ++t) v += cy[t] * p[(y + t - half) * stride + x]; }
            else {
                int acc = 0;
                for (int s = 0; s < taps; ++s) {
                    int hrow = 0;
                    for (int t = 0; t < taps; ++t) hrow += cx[t] * p[(y + s - half) * stride + x + t - half];
                    acc += cy[s] * hrow;   /* first pass keeps full precision at 8-bit (Shift1 == 0) */
                }
                if (uni) { dst8[y * dstride + x] = clip8((acc + (1 << 11)) >> 12); continue; }   /* Shift4 / Offset4 */
                v = acc >> 6;                                                                       /* Shift2 */
            }
            if (uni) dst8[y * dstride + x] = clip8((v + 32) >> 6);   /* Shift3 / Offset3 */
            else raw[y * w + x] = (int16_t)v;
        }
}

void port_luma_uni(uint32_t frac, const uint8_t *ref, uint32_t rs, uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h)
{
    interp_block(ref, (int)rs, 8, LUMA_TAPS[frac & 3], LUMA_TAPS[frac >> 2], frac & 3, frac >> 2, (int)w, (int)h, 1, dst, (int)ds, NULL);
}
void port_chroma_uni(uint32_t fx, uint32_t fy, const uint8_t *ref, uint32_t rs, uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h)
{
    interp_block(ref, (int)rs, 4, CHROMA_TAPS[fx], CHROMA_TAPS[fy], (int)fx, (int)fy, (int)w, (int)h, 1, dst, (int)ds, NULL);
}
static void bi_combine(const int16_t *a, const int16_t *b, uint8_t *dst, int ds, int w, int h)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) dst[y * ds + x] = clip8((a[y * w + x] + b[y * w + x] + 64) >> 7);   /* Shift5 / Offset5 */
}
void port_luma_bi(uint32_t f0, const uint8_t *r0, uint32_t r0s, uint32_t f1, const uint8_t *r1, uint32_t r1s,
                  uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h)
{
    int16_t *a = (int16_t *)malloc(2 * w * h * sizeof(int16_t)), *b = a + w * h;
    interp_block(r0, (int)r0s, 8, LUMA_TAPS[f0 & 3], LUMA_TAPS[f0 >> 2], f0 & 3, f0 >> 2, (int)w, (int)h, 0, NULL, 0, a);
    interp_block(r1, (int)r1s, 8, LUMA_TAPS[f1 & 3], LUMA_TAPS[f1 >> 2], f1 & 3, f1 >> 2, (int)w, (int)h, 0, NULL, 0, b);
    bi_combine(a, b, dst, (int)ds, (int)w, (int)h);
    free(a);
}
void port_chroma_bi(uint32_t f0x, uint32_t f0y, const uint8_t *r0, uint32_t r0s, uint32_t f1x, uint32_t f1y, const uint8_t *r1,
                    uint32_t r1s, uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h)
{
    int16_t *a = (int16_t *)malloc(2 * w * h * sizeof(int16_t)), *b = a + w * h;
    interp_block(r0, (int)r0s, 4, CHROMA_TAPS[f0x], CHROMA_TAPS[f0y], (int)f0x, (int)f0y, (int)w, (int)h, 0, NULL, 0, a);
    interp_block(r1, (int)r1s, 4, CHROMA_TAPS[f1x], CHROMA_TAPS[f1y], (int)f1x, (int)f1y, (int)w, (int)h, 0, NULL, 0, b);
    bi_combine(a, b, dst, (int)ds, (int)w, (int)h);
    free(a);
}

static int clamp_i(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }

int port_mc_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst)
{
    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_mc_job *q = &jobs[j];
        int px[2] = { 0, 0 }, py[2] = { 0, 0 };
        const svtb200_yuv_picture *r[2] = { NULL, NULL };
        for (int l = 0; l < 2; ++l) {
            if (!(q->dir == l || q->dir == 2)) continue;
            if (q->ref_idx[l] >= n_refs) return -1;
            r[l] = &refs[q->ref_idx[l]];
            /* absolute quarter-pel position in the PADDED reference, clamped so that the 8-tap window stays inside the
             * padding; a clamped position loses its fraction (EbInterPrediction.c:794-808); EB_U16 storage */
            px[l] = (uint16_t)clamp_i((r[l]->y.origin_x - 71) << 2, (r[l]->y.width + r[l]->y.origin_x + 7) << 2, ((q->x + r[l]->y.origin_x) << 2) + q->mv[l][0]);
            py[l] = (uint16_t)clamp_i((r[l]->y.origin_y - 71) << 2, (r[l]->y.height + r[l]->y.origin_y + 7) << 2, ((q->y + r[l]->y.origin_y) << 2) + q->mv[l][1]);
        }
        uint8_t *dy = dst->y.buf + (size_t)(dst->y.origin_y + q->y) * dst->y.stride + dst->y.origin_x + q->x;
        /* chroma index exactly as the reference forms it (EbInterPrediction.c:783) */
        const size_t cidx = (((size_t)(dst->y.origin_y + q->y) * dst->cb.stride) >> 1) + ((dst->y.origin_x + q->x) >> 1);
        const int cw = q->w >> 1, ch = q->h >> 1;
        if (q->dir != 2) {
            const int l = q->dir;
            port_luma_uni((uint32_t)((px[l] & 3) + ((py[l] & 3) << 2)), r[l]->y.buf + (px[l] >> 2) + (size_t)(py[l] >> 2) * r[l]->y.stride, r[l]->y.stride,
                          dy, dst->y.stride, q->w, q->h);
            const size_t ro = (size_t)(px[l] >> 3) + (size_t)(py[l] >> 3) * r[l]->cb.stride;
            port_chroma_uni((uint32_t)(px[l] & 7), (uint32_t)(py[l] & 7), r[l]->cb.buf + ro, r[l]->cb.stride, dst->cb.buf + cidx, dst->cb.stride, (uint32_t)cw, (uint32_t)ch);
            port_chroma_uni((uint32_t)(px[l] & 7), (uint32_t)(py[l] & 7), r[l]->cr.buf + ro, r[l]->cr.stride, dst->cr.buf + cidx, dst->cr.stride, (uint32_t)cw, (uint32_t)ch);
        } else {
            port_luma_bi((uint32_t)((px[0] & 3) + ((py[0] & 3) << 2)), r[0]->y.buf + (px[0] >> 2) + (size_t)(py[0] >> 2) * r[0]->y.stride, r[0]->y.stride,
                         (uint32_t)((px[1] & 3) + ((py[1] & 3) << 2)), r[1]->y.buf + (px[1] >> 2) + (size_t)(py[1] >> 2) * r[1]->y.stride, r[1]->y.stride,
                         dy, dst->y.stride, q->w, q->h);
            const size_t o0 = (size_t)(px[0] >> 3) + (size_t)(py[0] >> 3) * r[0]->cb.stride, o1 = (size_t)(px[1] >> 3) + (size_t)(py[1] >> 3) * r[1]->cb.stride;
            port_chroma_bi((uint32_t)(px[0] & 7), (uint32_t)(py[0] & 7), r[0]->cb.buf + o0, r[0]->cb.stride, (uint32_t)(px[1] & 7), (uint32_t)(py[1] & 7),
                           r[1]->cb.buf + o1, r[1]->cb.stride, dst->cb.buf + cidx, dst->cb.stride, (uint32_t)cw, (uint32_t)ch);
            port_chroma_bi((uint32_t)(px[0] & 7), (uint32_t)(py[0] & 7), r[0]->cr.buf + o0, r[0]->cr.stride, (uint32_t)(px[1] & 7), (uint32_t)(py[1] & 7),
                           r[1]->cr.buf + o1, r[1]->cr.stride, dst->cr.buf + cidx, dst->cr.stride, (uint32_t)cw, (uint32_t)ch);
        }
    }
    return 0;
}

/* ---- interpolation-free prediction of the mode-decision fast loop ---------------------------------------------------
 * UniPredIFreeRef8Bit / BiPredIFreeRef8Bit (Codec/EbAvcStyleMcp.c:172,321): function [0] of AvcStyleUniPredLumaIFFunctionPtrArray (the copy) at the
 * integer position the fraction maps to (IntegerPosoffsetTabX/Y :19-27; chroma: +1 where the 1/8 fraction exceeds 4), bi-prediction =
 * (a + b + 1) >> 1 of the two copies (BiPredAverageKernel_C :269).  Position clamp of the caller, Codec/EbInterPrediction.c:100-115. */
int port_mc_ifree_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst)
{
    static const uint8_t offx[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0 }, offy[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1 };
    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_mc_job *q = &jobs[j];
        const uint8_t *src[2][3] = { { 0 } };
        int stride[2][3] = { { 0 } };
        for (int l = 0; l < 2; ++l) {
            if (!(q->dir == l || q->dir == 2)) continue;
            if (q->ref_idx[l] >= n_refs) return -1;
            const svtb200_yuv_picture *r = &refs[q->ref_idx[l]];
            const int px = (uint16_t)clamp_i((r->y.origin_x - 71) << 2, (r->y.width + r->y.origin_x + 7) << 2, ((q->x + r->y.origin_x) << 2) + q->mv[l][0]);
            const int py = (uint16_t)clamp_i((r->y.origin_y - 71) << 2, (r->y.height + r->y.origin_y + 7) << 2, ((q->y + r->y.origin_y) << 2) + q->mv[l][1]);
            const int f = (px & 3) + ((py & 3) << 2);
            src[l][0] = r->y.buf + (size_t)((py >> 2) + offy[f]) * r->y.stride + (px >> 2) + offx[f];
            const int cx = (px >> 3) + ((px & 7) > 4), cy = (py >> 3) + ((py & 7) > 4);
            src[l][1] = r->cb.buf + (size_t)cy * r->cb.stride + cx;
            src[l][2] = r->cr.buf + (size_t)cy * r->cr.stride + cx;
            stride[l][0] = (int)r->y.stride; stride[l][1] = (int)r->cb.stride; stride[l][2] = (int)r->cr.stride;
        }
        uint8_t *d[3] = { dst->y.buf + (size_t)(dst->y.origin_y + q->y) * dst->y.stride + dst->y.origin_x + q->x, NULL, NULL };
        const size_t cidx = (((size_t)(dst->y.origin_y + q->y) * dst->cb.stride) >> 1) + ((dst->y.origin_x + q->x) >> 1);
        d[1] = dst->cb.buf + cidx; d[2] = dst->cr.buf + cidx;
        const int ds[3] = { (int)dst->y.stride, (int)dst->cb.stride, (int)dst->cr.stride };
        for (int c = 0; c < 3; ++c) {
            const int w = c ? q->w >> 1 : q->w, h = c ? q->h >> 1 : q->h;
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x) {
                    if (q->dir == 2) d[c][(size_t)y * ds[c] + x] = (uint8_t)((src[0][c][(size_t)y * stride[0][c] + x] + src[1][c][(size_t)y * stride[1][c] + x] + 1) >> 1);
                    else d[c][(size_t)y * ds[c] + x] = src[q->dir][c][(size_t)y * stride[q->dir][c] + x];
                }
        }
    }
    return 0;
}

/* ---- 16-bit samples (10-bit video): SURVEY §8 row b5 -------------------------------------------------------------------
 * uniPred / biPred ...New16bit / ...OutRaw16bit (C_DEFAULT/EbMcp_C.c:3121-4861), BiPredClipping16bit :3265, integer/integer
 * shortcut BiPredClippingOnTheFly16bit (Codec/EbMcp.c:847), drivers UniPredInterpolation16bit / BiPredInterpolation16bit
 * (Codec/EbMcp.c:249,804), EncodePassInterPrediction16bit (Codec/EbInterPrediction.c:928).  Constants Codec/EbDefinitions.h:1041-1059:
 *   1-D      (S + 32) >> 6                                  raw: (S - 32768) >> 2
 *   2-D      first (S - 32768) >> 2 (s16), then (S2 + 524800) >> 10      raw: S2 >> 6
 *   copy     p                                              raw: (p << 4) - 8192
 *   bi       (a + b + 16400) >> 5, clip to 1023
 * Written case by case like the reference (no unification), so that it checks the two-pass form the CUDA kernel uses. */
static inline uint16_t clip10(int v) { return (uint16_t)(v < 0 ? 0 : v > 1023 ? 1023 : v); }

static void interp_block16(const uint16_t *p, int stride, int taps, const int *cx, const int *cy, int fx, int fy, int w, int h, int uni, uint16_t *dst,
                           int dstride, int16_t *raw)
{
    const int half = taps / 2 - 1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int o_uni, o_raw;
            if (!fx && !fy) { o_uni = p[y * stride + x]; o_raw = (int16_t)(p[y * stride + x] << 4) - 8192; }
            else if (!fy || !fx) {
                int s = 0;
                for (int t = 0; t < taps; ++t) s += !fy ? cx[t] * p[y * stride + x + t - half] : cy[t] * p[(y + t - half) * stride + x];
                o_uni = clip10((s + 32) >> 6); o_raw = (int16_t)((s - 32768) >> 2);
            } else {
                int acc = 0;
                for (int r = 0; r < taps; ++r) {
                    int hrow = 0;
                    for (int t = 0; t < taps; ++t) hrow += cx[t] * p[(y + r - half) * stride + x + t - half];
                    acc += cy[r] * (int16_t)((hrow - 32768) >> 2);
                }
                o_uni = clip10((acc + 524800) >> 10); o_raw = (int16_t)(acc >> 6);
            }
            if (uni) dst[y * dstride + x] = (uint16_t)o_uni; else raw[y * w + x] = (int16_t)o_raw;
        }
}

static void pred16(const uint16_t *r0, int s0, int f0x, int f0y, const uint16_t *r1, int s1, int f1x, int f1y, int taps, const int (*tab)[8], const int (*tab4)[4],
                   uint16_t *dst, int ds, int w, int h)
{
    const int *c0x = taps == 8 ? tab[f0x] : tab4[f0x], *c0y = taps == 8 ? tab[f0y] : tab4[f0y];
    if (!r1) { interp_block16(r0, s0, taps, c0x, c0y, f0x, f0y, w, h, 1, dst, ds, NULL); return; }
    const int *c1x = taps == 8 ? tab[f1x] : tab4[f1x], *c1y = taps == 8 ? tab[f1y] : tab4[f1y];
    int16_t *a = (int16_t *)malloc((size_t)2 * w * h * sizeof(int16_t)), *b = a + w * h;
    interp_block16(r0, s0, taps, c0x, c0y, f0x, f0y, w, h, 0, NULL, 0, a);
    interp_block16(r1, s1, taps, c1x, c1y, f1x, f1y, w, h, 0, NULL, 0, b);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) dst[y * ds + x] = clip10((a[y * w + x] + b[y * w + x] + 16400) >> 5);
    free(a);
}

/* planes hold 16-bit samples: svtb200_plane::buf points at them, stride / origins in samples */
int port_mc_batch16(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst)
{
    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_mc_job *q = &jobs[j];
        int px[2] = { 0, 0 }, py[2] = { 0, 0 };
        const svtb200_yuv_picture *r[2] = { NULL, NULL };
        for (int l = 0; l < 2; ++l) {
            if (!(q->dir == l || q->dir == 2)) continue;
            if (q->ref_idx[l] >= n_refs) return -1;
            r[l] = &refs[q->ref_idx[l]];
            px[l] = (uint16_t)clamp_i((r[l]->y.origin_x - 71) << 2, (r[l]->y.width + r[l]->y.origin_x + 7) << 2, ((q->x + r[l]->y.origin_x) << 2) + q->mv[l][0]);
            py[l] = (uint16_t)clamp_i((r[l]->y.origin_y - 71) << 2, (r[l]->y.height + r[l]->y.origin_y + 7) << 2, ((q->y + r[l]->y.origin_y) << 2) + q->mv[l][1]);
        }
        const int a = q->dir == 1 ? 1 : 0, b = q->dir == 2 ? 1 : -1;
        uint16_t *dy = (uint16_t *)dst->y.buf + (size_t)(dst->y.origin_y + q->y) * dst->y.stride + dst->y.origin_x + q->x;
        const size_t cidx = (((size_t)(dst->y.origin_y + q->y) * dst->cb.stride) >> 1) + ((dst->y.origin_x + q->x) >> 1);
        const uint16_t *ya = (const uint16_t *)r[a]->y.buf + (px[a] >> 2) + (size_t)(py[a] >> 2) * r[a]->y.stride;
        const uint16_t *yb = b < 0 ? NULL : (const uint16_t *)r[b]->y.buf + (px[b] >> 2) + (size_t)(py[b] >> 2) * r[b]->y.stride;
        pred16(ya, (int)r[a]->y.stride, px[a] & 3, py[a] & 3, yb, b < 0 ? 0 : (int)r[b]->y.stride, b < 0 ? 0 : px[b] & 3, b < 0 ? 0 : py[b] & 3, 8, LUMA_TAPS, CHROMA_TAPS,
               dy, (int)dst->y.stride, q->w, q->h);
        for (int c = 0; c < 2; ++c) {
            const svtb200_plane *pa = c ? &r[a]->cr : &r[a]->cb, *pb = b < 0 ? NULL : (c ? &r[b]->cr : &r[b]->cb), *pd = c ? &dst->cr : &dst->cb;
            const uint16_t *ca = (const uint16_t *)pa->buf + (px[a] >> 3) + (size_t)(py[a] >> 3) * pa->stride;
            const uint16_t *cb = pb ? (const uint16_t *)pb->buf + (px[b] >> 3) + (size_t)(py[b] >> 3) * pb->stride : NULL;
            pred16(ca, (int)pa->stride, px[a] & 7, py[a] & 7, cb, pb ? (int)pb->stride : 0, pb ? px[b] & 7 : 0, pb ? py[b] & 7 : 0, 4, LUMA_TAPS, CHROMA_TAPS,
                   (uint16_t *)pd->buf + cidx, (int)pd->stride, q->w >> 1, q->h >> 1);
        }
    }
    return 0;
}
