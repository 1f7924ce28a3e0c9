/*
 * oracle/port/sao_port.c -- TEST INFRASTRUCTURE ONLY (see me_port.c header for the rules).
 *
 * Plain-C restatement of the reference's SAO statistics and SAO application, 8-bit 4:2:0:
 *   statistics   GatherSaoStatisticsLcuLossy_62x62            C_DEFAULT/EbSampleAdaptiveOffset_C.c:125
 *                (call site SaoGenerationDecision, Codec/EbSampleAdaptiveOffsetGenerationDecision.c:647-760)
 *   application  SAOApplyBO / SAOApplyEO_{0,90,135,45}         C_DEFAULT/EbSampleAdaptiveOffset_C.c:394-861
 *                ApplySaoOffsetsLcu / ApplySaoOffsetsPicture   Codec/EbEncDecProcess.c:215,522
 * The reference filters in place and keeps un-filtered neighbours in line buffers; here the same result is written
 * out of place from an unmodified input picture.
 */
#include <stdint.h>
#include <string.h>

#include "svt_hevc_b200.h"

static const svtb200_plane *pl(const svtb200_yuv_picture *p, int k) { return k == 0 ? &p->y : k == 1 ? &p->cb : &p->cr; }
static inline int sgn(int a, int b) { return a > b ? 1 : a < b ? -1 : 0; }
static inline int clip_s8(int v) { return v < -128 ? -128 : v > 127 ? 127 : v; }

/* hbd: GatherSaoStatisticsLcu_62x62_16bit (C_DEFAULT/EbSampleAdaptiveOffset_C.c:23): 16-bit planes, differences NOT clipped, band = sample >> 5 */
#define LDS(ptr, i) (hbd ? (int)((const uint16_t *)(ptr))[i] : (int)((const uint8_t *)(ptr))[i])
static int sao_stats_impl(const svtb200_yuv_picture *src, const svtb200_yuv_picture *rec, svtb200_sao_stats *out, int hbd)
{
    const int bps = hbd ? 2 : 1, bsh = hbd ? 5 : 3;
    const int W = rec->y.width, H = rec->y.height, cols = (W + 63) / 64, rows = (H + 63) / 64;
    for (int ly = 0; ly < rows; ++ly)
        for (int lx = 0; lx < cols; ++lx)
            for (int c = 0; c < 3; ++c) {
                const int sh = c ? 1 : 0;
                const svtb200_plane *sp = pl(src, c), *rp = pl(rec, c);
                const int ox = (lx * 64) >> sh, oy = (ly * 64) >> sh;
                const int lw = (W - lx * 64 < 64 ? W - lx * 64 : 64) >> sh, lh = (H - ly * 64 < 64 ? H - ly * 64 : 64) >> sh;
                svtb200_sao_stats *o = &out[(size_t)(ly * cols + lx) * 3 + c];
                memset(o, 0, sizeof *o);
                const uint8_t *s = sp->buf + ((size_t)(sp->origin_y + oy) * sp->stride + sp->origin_x + ox) * bps;
                const uint8_t *r = rp->buf + ((size_t)(rp->origin_y + oy) * rp->stride + rp->origin_x + ox) * bps;
                const int ss = (int)sp->stride, rs = (int)rp->stride;
                for (int y = 1; y < lh - 1; ++y)
                    for (int x = 1; x < lw - 1; ++x) {
                        const int v = LDS(r, y * rs + x), d = hbd ? LDS(s, y * ss + x) - v : clip_s8(LDS(s, y * ss + x) - v);
                        o->bo_diff[v >> bsh] += d; o->bo_count[v >> bsh]++;
                        const int e[4] = { sgn(v, LDS(r, y * rs + x - 1)) + sgn(v, LDS(r, y * rs + x + 1)) + 2,             /* EO-0   */
                                           sgn(v, LDS(r, (y - 1) * rs + x)) + sgn(v, LDS(r, (y + 1) * rs + x)) + 2,         /* EO-90  */
                                           sgn(v, LDS(r, (y - 1) * rs + x - 1)) + sgn(v, LDS(r, (y + 1) * rs + x + 1)) + 2, /* EO-135 */
                                           sgn(v, LDS(r, (y - 1) * rs + x + 1)) + sgn(v, LDS(r, (y + 1) * rs + x - 1)) + 2 };/* EO-45  */
                        for (int t = 0; t < 4; ++t) { o->eo_diff[t][e[t]] += d; o->eo_count[t][e[t]]++; }
                    }
                for (int t = 0; t < 4; ++t) {   /* drop the "no edge" category: [2] <- [3], [3] <- [4] (:216-221) */
                    o->eo_diff[t][2] = o->eo_diff[t][3]; o->eo_diff[t][3] = o->eo_diff[t][4];
                    o->eo_count[t][2] = o->eo_count[t][3]; o->eo_count[t][3] = o->eo_count[t][4];
                    o->eo_diff[t][4] = 0; o->eo_count[t][4] = 0;   /* scratch slot: the C row leaves a copy, the SSE2 row zero */
                }
            }
    return 0;
}

int port_sao_stats_picture(const svtb200_yuv_picture *src, const svtb200_yuv_picture *rec, svtb200_sao_stats *out) { return sao_stats_impl(src, rec, out, 0); }
int port_sao_stats_picture16(const svtb200_yuv_picture *src, const svtb200_yuv_picture *rec, svtb200_sao_stats *out) { return sao_stats_impl(src, rec, out, 1); }

/* hbd: SAOApplyBO16bit / SAOApplyEO_*16bit (:431 ff): band = sample >> 5, results clipped to 1023 */
static int sao_apply_impl(const svtb200_yuv_picture *in, const svtb200_yuv_picture *outp, const svtb200_sao_params *prm, const uint8_t sao_flag[2], int hbd)
{
    const int bps = hbd ? 2 : 1, bsh = hbd ? 5 : 3, maxv = hbd ? 1023 : 255;
    static const int NB[5][4] = { { 0, 0, 0, 0 }, { -1, 0, 1, 0 }, { 0, -1, 0, 1 }, { -1, -1, 1, 1 }, { 1, -1, -1, 1 } };   /* dx0,dy0,dx1,dy1 per EO class */
    const int W = in->y.width, H = in->y.height, cols = (W + 63) / 64;
    for (int c = 0; c < 3; ++c) {
        const int sh = c ? 1 : 0, pw = W >> sh, ph = H >> sh;
        const svtb200_plane *ip = pl(in, c), *op = pl(outp, c);
        const uint8_t *I = ip->buf + ((size_t)ip->origin_y * ip->stride + ip->origin_x) * bps;
        uint8_t *O = op->buf + ((size_t)op->origin_y * op->stride + op->origin_x) * bps;
        for (int y = 0; y < ph; ++y)
            for (int x = 0; x < pw; ++x) {
                const svtb200_sao_params *p = &prm[(size_t)((y << sh) >> 6) * cols + ((x << sh) >> 6)];
                const int t = sao_flag[c ? 1 : 0] ? p->type[c ? 1 : 0] : 0;
                int v = LDS(I, (size_t)y * ip->stride + x);
                if (t == 5) {
                    const int band = v >> bsh, pos = p->band_position[c];
                    if (band >= pos && band <= pos + 3) { v += p->offset[c][band - pos]; v = v < 0 ? 0 : v > maxv ? maxv : v; }
                } else if (t >= 1 && t <= 4) {
                    const int *n = NB[t];
                    const int edge_x = (x == 0 || x == pw - 1), edge_y = (y == 0 || y == ph - 1);
                    const int skip = (t == 1) ? edge_x : (t == 2) ? edge_y : (edge_x || edge_y);
                    if (!skip) {
                        const int a = LDS(I, (size_t)(y + n[1]) * ip->stride + x + n[0]), b = LDS(I, (size_t)(y + n[3]) * ip->stride + x + n[2]);
                        const int idx = sgn(v, a) + sgn(v, b) + 2;
                        const int off = idx == 0 ? p->offset[c][0] : idx == 1 ? p->offset[c][1] : idx == 3 ? p->offset[c][2] : idx == 4 ? p->offset[c][3] : 0;
                        v += off; v = v < 0 ? 0 : v > maxv ? maxv : v;
                    }
                }
                if (hbd) ((uint16_t *)O)[(size_t)y * op->stride + x] = (uint16_t)v; else O[(size_t)y * op->stride + x] = (uint8_t)v;
            }
    }
    return 0;
}
int port_sao_apply_picture(const svtb200_yuv_picture *in, const svtb200_yuv_picture *outp, const svtb200_sao_params *prm, const uint8_t sao_flag[2])
{ return sao_apply_impl(in, outp, prm, sao_flag, 0); }
int port_sao_apply_picture16(const svtb200_yuv_picture *in, const svtb200_yuv_picture *outp, const svtb200_sao_params *prm, const uint8_t sao_flag[2])
{ return sao_apply_impl(in, outp, prm, sao_flag, 1); }
