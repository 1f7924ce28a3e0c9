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

/* SURVEY §7 H3 -- the deblocking state the encoder's SAO gather sees.  EncodePass deblocks LCU n and gathers its statistics before any later LCU is
 * even reconstructed (Codec/EbCodingLoop.c:4600-4760), and the per-LCU DLF drivers postpone everything that needs the right / bottom neighbour
 * (LCUInternalAreaDLFCore leaves the last 4 columns / rows alone, LCUBoundaryDLFCore of the NEXT LCUs finishes them, Codec/EbDeblockingFilter.c:2222,
 * 2828).  Checked sample by sample against the reference's own interleaving (oracle/ref_dlf_harness.c: svtref_dlf_sao_stats_picture) on every DLF test
 * picture: at gather time an LCU holds the FINAL deblocked samples everywhere except its last 4 columns (when a right neighbour exists) and its last
 * 4 rows (when a bottom neighbour exists) -- 4 samples in each plane's own units, luma and chroma alike -- which still hold the UNFILTERED
 * reconstruction.  The strips of different LCUs never overlap, so one composite picture serves every LCU: */
static int h3_compose_impl(const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered, const svtb200_yuv_picture *out, int bps)
{
    const int W = deblocked->y.width, H = deblocked->y.height;
    for (int c = 0; c < 3; ++c) {
        const int sh = c ? 1 : 0, pw = W >> sh, ph = H >> sh, L = 64 >> sh;
        const svtb200_plane *d = pl(deblocked, c), *u = pl(unfiltered, c), *o = pl(out, c);
        for (int y = 0; y < ph; ++y)
            for (int x = 0; x < pw; ++x) {
                const int x1 = ((x / L) + 1) * L < pw ? ((x / L) + 1) * L : pw, y1 = ((y / L) + 1) * L < ph ? ((y / L) + 1) * L : ph;   /* end of this LCU */
                const int strip = (x1 < pw && x >= x1 - 4) || (y1 < ph && y >= y1 - 4);
                const svtb200_plane *s = strip ? u : d;
                memcpy(o->buf + ((size_t)(o->origin_y + y) * o->stride + o->origin_x + x) * bps, s->buf + ((size_t)(s->origin_y + y) * s->stride + s->origin_x + x) * bps, (size_t)bps);
            }
    }
    return 0;
}
int port_sao_h3_compose(const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered, const svtb200_yuv_picture *out) { return h3_compose_impl(deblocked, unfiltered, out, 1); }
int port_sao_h3_compose16(const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered, const svtb200_yuv_picture *out) { return h3_compose_impl(deblocked, unfiltered, out, 2); }
/* statistics of every LCU in the state the encoder gathers them on a picture with deblocking: `scratch` (same geometry) receives the composite */
int port_sao_stats_picture_h3(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered,
                              const svtb200_yuv_picture *scratch, svtb200_sao_stats *out)
{
    port_sao_h3_compose(deblocked, unfiltered, scratch);
    return sao_stats_impl(src, scratch, out, 0);
}
int port_sao_stats_picture16_h3(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered,
                                const svtb200_yuv_picture *scratch, svtb200_sao_stats *out)
{
    port_sao_h3_compose16(deblocked, unfiltered, scratch);
    return sao_stats_impl(src, scratch, out, 1);
}

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
