/*
 * oracle/port/dlf_port.c -- TEST INFRASTRUCTURE ONLY (see me_port.c header for the rules).
 *
 * Plain-C restatement of the reference's in-loop deblocking, 8-bit 4:2:0:
 *   edge filters  Luma4SampleEdgeDLFCore / Chroma2SampleEdgeDLFCore   C_DEFAULT/EbDeblockingFilter_C.c:39,469
 *   per-edge parameter derivation (QP averaging, tc / beta lookup, chroma QP mapping)
 *                 Codec/EbDeblockingFilter.c:2303-2336, :2400-2440; tables :30-38; ChromaQpMap Codec/EbTransforms.c:174
 *   picture traversal: written in the HEVC order (all vertical edges of the picture, then all horizontal edges) --
 *                 tests/test_oracle_dlf.py checks that this equals the reference's per-LCU drivers
 *                 (LCUInternalAreaDLFCore / LCUBoundaryDLFCore / LCUPictureEdgeDLFCore, :2222,:2828,:3518).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "svt_hevc_b200.h"

static const uint8_t TC_TAB[54] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6,
                                    7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24 };
static const uint8_t BETA_TAB[52] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28,
                                      30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64 };
static const uint8_t CQP_TAB[58] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32,
                                     33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51 };

static inline int clampi(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }
static inline int iabs(int v) { return v < 0 ? -v : v; }

void port_luma_edge(uint8_t *e, uint32_t stride, uint8_t vertical, int32_t tc, int32_t beta)
{
    const int fs = vertical ? 1 : (int)stride, rs = vertical ? (int)stride : 1;   /* across / along the edge */
#define S(k, r) ((int)e[(k) * fs + (r) * rs])                                      /* k < 0: p side, k >= 0: q side */
    const int dp0 = iabs(S(-3, 0) - 2 * S(-2, 0) + S(-1, 0)), dp3 = iabs(S(-3, 3) - 2 * S(-2, 3) + S(-1, 3));
    const int dq0 = iabs(S(2, 0) - 2 * S(1, 0) + S(0, 0)), dq3 = iabs(S(2, 3) - 2 * S(1, 3) + S(0, 3));
    const int dp = dp0 + dp3, dq = dq0 + dq3, d0 = dp0 + dq0, d3 = dp3 + dq3;
    if (d0 + d3 >= beta) return;
    const int s0 = ((d0 << 1) < (beta >> 2)) && (beta >> 3) > iabs(S(-4, 0) - S(-1, 0)) + iabs(S(3, 0) - S(0, 0)) && ((5 * tc + 1) >> 1) > iabs(S(-1, 0) - S(0, 0));
    const int s3 = ((d3 << 1) < (beta >> 2)) && (beta >> 3) > iabs(S(-4, 3) - S(-1, 3)) + iabs(S(3, 3) - S(0, 3)) && ((5 * tc + 1) >> 1) > iabs(S(-1, 3) - S(0, 3));
    for (int r = 0; r < 4; ++r) {
        const int p3 = S(-4, r), p2 = S(-3, r), p1 = S(-2, r), p0 = S(-1, r), q0 = S(0, r), q1 = S(1, r), q2 = S(2, r), q3 = S(3, r);
        uint8_t *o = e + r * rs;
        if (s0 && s3) {
            o[0] = (uint8_t)clampi(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            o[-fs] = (uint8_t)clampi(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            o[fs] = (uint8_t)clampi(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
            o[-2 * fs] = (uint8_t)clampi(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
            o[2 * fs] = (uint8_t)clampi(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
            o[-3 * fs] = (uint8_t)clampi(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else {
            int delta = ((q0 - p0) * 9 - (q1 - p1) * 3 + 8) >> 4;
            if (iabs(delta) >= tc * 10) continue;
            delta = clampi(-tc, tc, delta);
            o[0] = (uint8_t)clampi(0, 255, q0 - delta);
            o[-fs] = (uint8_t)clampi(0, 255, p0 + delta);
            const int side = (beta + (beta >> 1)) >> 3, tc2 = tc >> 1;
            if (side > dp) o[-2 * fs] = (uint8_t)clampi(0, 255, p1 + clampi(-tc2, tc2, (((p0 + p2 + 1) >> 1) - p1 + delta) >> 1));
            if (side > dq) o[fs] = (uint8_t)clampi(0, 255, q1 + clampi(-tc2, tc2, (((q0 + q2 + 1) >> 1) - q1 - delta) >> 1));
        }
    }
#undef S
}

static void chroma_one(uint8_t *e, int fs, int tc)
{
    const int p1 = e[-2 * fs], p0 = e[-fs], q0 = e[0], q1 = e[fs];
    const int delta = clampi(-tc, tc, (((q0 - p0) << 2) + p1 - q1 + 4) >> 3);
    e[-fs] = (uint8_t)clampi(0, 255, p0 + delta);
    e[0] = (uint8_t)clampi(0, 255, q0 - delta);
}
void port_chroma_edge(uint8_t *cb, uint8_t *cr, uint32_t stride, uint8_t vertical, uint8_t cb_tc, uint8_t cr_tc)
{
    const int fs = vertical ? 1 : (int)stride, rs = vertical ? (int)stride : 1;
    chroma_one(cb, fs, cb_tc); chroma_one(cb + rs, fs, cb_tc);
    chroma_one(cr, fs, cr_tc); chroma_one(cr + rs, fs, cr_tc);
}

/* Luma4SampleEdgeDLFCore16bit, C_DEFAULT/EbDeblockingFilter_C.c:240: same decisions, samples clipped to 1023 */
void port_luma_edge16(uint16_t *e, uint32_t stride, uint8_t vertical, int32_t tc, int32_t beta)
{
    const int fs = vertical ? 1 : (int)stride, rs = vertical ? (int)stride : 1;   /* across / along the edge */
#define S(k, r) ((int)e[(k) * fs + (r) * rs])                                      /* k < 0: p side, k >= 0: q side */
    const int dp0 = iabs(S(-3, 0) - 2 * S(-2, 0) + S(-1, 0)), dp3 = iabs(S(-3, 3) - 2 * S(-2, 3) + S(-1, 3));
    const int dq0 = iabs(S(2, 0) - 2 * S(1, 0) + S(0, 0)), dq3 = iabs(S(2, 3) - 2 * S(1, 3) + S(0, 3));
    const int dp = dp0 + dp3, dq = dq0 + dq3, d0 = dp0 + dq0, d3 = dp3 + dq3;
    if (d0 + d3 >= beta) return;
    const int s0 = ((d0 << 1) < (beta >> 2)) && (beta >> 3) > iabs(S(-4, 0) - S(-1, 0)) + iabs(S(3, 0) - S(0, 0)) && ((5 * tc + 1) >> 1) > iabs(S(-1, 0) - S(0, 0));
    const int s3 = ((d3 << 1) < (beta >> 2)) && (beta >> 3) > iabs(S(-4, 3) - S(-1, 3)) + iabs(S(3, 3) - S(0, 3)) && ((5 * tc + 1) >> 1) > iabs(S(-1, 3) - S(0, 3));
    for (int r = 0; r < 4; ++r) {
        const int p3 = S(-4, r), p2 = S(-3, r), p1 = S(-2, r), p0 = S(-1, r), q0 = S(0, r), q1 = S(1, r), q2 = S(2, r), q3 = S(3, r);
        uint16_t *o = e + r * rs;
        if (s0 && s3) {
            o[0] = (uint16_t)clampi(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            o[-fs] = (uint16_t)clampi(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            o[fs] = (uint16_t)clampi(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
            o[-2 * fs] = (uint16_t)clampi(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
            o[2 * fs] = (uint16_t)clampi(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
            o[-3 * fs] = (uint16_t)clampi(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else {
            int delta = ((q0 - p0) * 9 - (q1 - p1) * 3 + 8) >> 4;
            if (iabs(delta) >= tc * 10) continue;
            delta = clampi(-tc, tc, delta);
            o[0] = (uint16_t)clampi(0, 1023, q0 - delta);
            o[-fs] = (uint16_t)clampi(0, 1023, p0 + delta);
            const int side = (beta + (beta >> 1)) >> 3, tc2 = tc >> 1;
            if (side > dp) o[-2 * fs] = (uint16_t)clampi(0, 1023, p1 + clampi(-tc2, tc2, (((p0 + p2 + 1) >> 1) - p1 + delta) >> 1));
            if (side > dq) o[fs] = (uint16_t)clampi(0, 1023, q1 + clampi(-tc2, tc2, (((q0 + q2 + 1) >> 1) - q1 - delta) >> 1));
        }
    }
#undef S
}


static void chroma_one16(uint16_t *e, int fs, int tc)
{
    const int p1 = e[-2 * fs], p0 = e[-fs], q0 = e[0], q1 = e[fs];
    const int delta = clampi(-tc, tc, (((q0 - p0) << 2) + p1 - q1 + 4) >> 3);
    e[-fs] = (uint16_t)clampi(0, 1023, p0 + delta);
    e[0] = (uint16_t)clampi(0, 1023, q0 - delta);
}
void port_chroma_edge16(uint16_t *cb, uint16_t *cr, uint32_t stride, uint8_t vertical, uint8_t cb_tc, uint8_t cr_tc)
{
    const int fs = vertical ? 1 : (int)stride, rs = vertical ? (int)stride : 1;
    chroma_one16(cb, fs, cb_tc); chroma_one16(cb + rs, fs, cb_tc);
    chroma_one16(cr, fs, cr_tc); chroma_one16(cr + rs, fs, cr_tc);
}


static int chroma_qp(int q) { return q < 0 ? q : q > 57 ? q - 6 : CQP_TAB[q]; }   /* convertToChromaQp, 4:2:0 branch */

/* hbd: LCU*DLFCore16bit (Codec/EbDeblockingFilter.c:2514-4300): tc and beta scaled by 4 (:1254), 16-bit sample planes */
static int dlf_picture_impl(const svtb200_yuv_picture *rec, const uint8_t *vbs, const uint8_t *hbs, const uint8_t *qp, const svtb200_dlf_params *p, int hbd)
{
    const int bps = hbd ? 2 : 1, sc = hbd ? 2 : 0;
    const int W = p->pic_width, H = p->pic_height, cols = (W + 63) / 64, qs = p->qp_array_stride;
    uint8_t *Y = rec->y.buf + ((size_t)rec->y.origin_y * rec->y.stride + rec->y.origin_x) * bps;
    uint8_t *Cb = rec->cb.buf + ((size_t)rec->cb.origin_y * rec->cb.stride + rec->cb.origin_x) * bps;
    uint8_t *Cr = rec->cr.buf + ((size_t)rec->cr.origin_y * rec->cr.stride + rec->cr.origin_x) * bps;
#define BS(arr, x, y) (arr)[((size_t)((y) >> 6) * cols + ((x) >> 6)) * 256 + (((x) & 63) >> 2) + ((((y) & 63) >> 2) << 4)]
#define QP(x, y) qp[((x) >> 3) + ((y) >> 3) * qs]
    for (int pass = 0; pass < 2; ++pass) {          /* 0: vertical edges, 1: horizontal edges */
        const uint8_t *bsarr = pass == 0 ? vbs : hbs;
        /* luma: edges on the 8-sample grid, 4-sample segments */
        for (int y = 0; y < H; y += (pass == 0 ? 4 : 8))
            for (int x = 0; x < W; x += (pass == 0 ? 8 : 4)) {
                if ((pass == 0 && x == 0) || (pass == 1 && y == 0)) continue;
                const int bs = BS(bsarr, x, y);
                if (!bs) continue;
                const int q = (QP(x, y) + (pass == 0 ? QP(x - 1, y) : QP(x, y - 1)) + 1) >> 1;
                const int tc = TC_TAB[clampi(0, 53, q + ((bs > 1) << 1) + p->tc_offset)], beta = BETA_TAB[clampi(0, 51, q + p->beta_offset)];
                if (hbd) port_luma_edge16((uint16_t *)Y + (size_t)y * rec->y.stride + x, rec->y.stride, pass == 0, tc << sc, beta << sc);
                else port_luma_edge(Y + (size_t)y * rec->y.stride + x, rec->y.stride, pass == 0, tc, beta);
            }
        /* chroma: edges on the 8-sample chroma grid, 2-sample segments, only strength 2 */
        for (int cy = 0; cy < H / 2; cy += (pass == 0 ? 2 : 8))
            for (int cx = 0; cx < W / 2; cx += (pass == 0 ? 8 : 2)) {
                if ((pass == 0 && cx == 0) || (pass == 1 && cy == 0)) continue;
                const int lx = 2 * cx, ly = 2 * cy;
                if (BS(bsarr, lx, ly) <= 1) continue;
                const int q = (QP(lx, ly) + (pass == 0 ? QP(lx - 2, ly) : QP(lx, ly - 2)) + 1) >> 1;
                const int cbq = chroma_qp(q + p->cb_qp_offset), crq = chroma_qp(q + p->cr_qp_offset);
                const int cbt = TC_TAB[clampi(0, 53, cbq + 2 + p->tc_offset)], crt = TC_TAB[clampi(0, 53, crq + 2 + p->tc_offset)];
                if (hbd) port_chroma_edge16((uint16_t *)Cb + (size_t)cy * rec->cb.stride + cx, (uint16_t *)Cr + (size_t)cy * rec->cr.stride + cx, rec->cb.stride, pass == 0, (uint8_t)(cbt << sc), (uint8_t)(crt << sc));
                else port_chroma_edge(Cb + (size_t)cy * rec->cb.stride + cx, Cr + (size_t)cy * rec->cr.stride + cx, rec->cb.stride, pass == 0, (uint8_t)cbt, (uint8_t)crt);
            }
    }
#undef BS
#undef QP
    return 0;
}
int port_dlf_picture(const svtb200_yuv_picture *rec, const uint8_t *vbs, const uint8_t *hbs, const uint8_t *qp, const svtb200_dlf_params *p)
{ return dlf_picture_impl(rec, vbs, hbs, qp, p, 0); }
int port_dlf_picture16(const svtb200_yuv_picture *rec, const uint8_t *vbs, const uint8_t *hbs, const uint8_t *qp, const svtb200_dlf_params *p)
{ return dlf_picture_impl(rec, vbs, hbs, qp, p, 1); }

/* ---- boundary-strength derivation --------------------------------------------------------------------------------
 * Restated CU by CU the way the encode pass runs it (Codec/EbCodingLoop.c:3554,4329,4364): for every CU of every LCU, SetBSArrayBasedOnPUBoundary
 * (Codec/EbDeblockingFilter.c:339: the CU's top row and left column against the neighbour arrays, skipped on the picture edge) and, for 64x64 CUs,
 * SetBSArrayBasedOnTUBoundary (:472) for the four 32x32 TUs; per edge CalculateBSForPUBoundary (:109) / CalculateBSForTUBoundaryInsidePU (:440).
 * The neighbour arrays are read from the final maps (the neighbours are coded before the CU, the CU's own cbf entries are written before the call). */
static int port_mv_far(const svtb200_cu_info *a, int la, const svtb200_cu_info *b, int lb)
{
    return abs(a->mv[la][0] - b->mv[lb][0]) >= 4 || abs(a->mv[la][1] - b->mv[lb][1]) >= 4;
}
static uint8_t port_bs_pu(const svtb200_cu_info *pu, const svtb200_cu_info *nb, int cbf_here, int cbf_nb, int slice_type, uint64_t poc0, uint64_t poc1)
{
    if (pu->intra || nb->intra) return 2;
    int c1;
    if (slice_type == 1) c1 = port_mv_far(pu, 0, nb, 0);
    else {
        switch (pu->pred_dir + nb->pred_dir * 3) {
        case 0: c1 = port_mv_far(pu, 0, nb, 0); break;
        case 1: c1 = poc1 != poc0 || port_mv_far(pu, 1, nb, 0); break;
        case 3: c1 = poc0 != poc1 || port_mv_far(pu, 0, nb, 1); break;
        case 4: c1 = port_mv_far(pu, 1, nb, 1); break;
        case 8:
            if (poc0 == poc1) c1 = (port_mv_far(pu, 0, nb, 0) || port_mv_far(pu, 1, nb, 1)) && (port_mv_far(pu, 0, nb, 1) || port_mv_far(pu, 1, nb, 0));
            else c1 = port_mv_far(pu, 0, nb, 0) || port_mv_far(pu, 1, nb, 1);
            break;
        default: c1 = 1; break;
        }
    }
    return (uint8_t)(c1 | (cbf_here || cbf_nb));
}

int port_dlf_bs_picture(const svtb200_cu_info *cu_map, uint32_t cu_stride, const uint8_t *cbf, uint32_t cbf_stride, uint16_t W, uint16_t H, uint8_t slice_type,
                        uint64_t poc0, uint64_t poc1, uint8_t *vbs, uint8_t *hbs)
{
    const int cols = (W + 63) / 64, rows = (H + 63) / 64;
    memset(vbs, 0, (size_t)cols * rows * 256);
    memset(hbs, 0, (size_t)cols * rows * 256);
    for (int ly = 0; ly < rows; ++ly)
        for (int lx = 0; lx < cols; ++lx) {
            uint8_t *v = vbs + (size_t)(ly * cols + lx) * 256, *h = hbs + (size_t)(ly * cols + lx) * 256;
            /* the CUs of this LCU: every 8x8 block whose position is a multiple of its CU size starts one */
            for (int by = 0; by < 64; by += 8)
                for (int bx = 0; bx < 64; bx += 8) {
                    const int x0 = lx * 64 + bx, y0 = ly * 64 + by;
                    if (x0 >= W || y0 >= H) continue;
                    const svtb200_cu_info *cu = &cu_map[(size_t)(y0 >> 3) * cu_stride + (x0 >> 3)];
                    const int size = 1 << cu->log2_cu_size;
                    if ((bx & (size - 1)) || (by & (size - 1))) continue;
                    if (y0 != 0)          /* horizontal PU boundary: the CU's top row (tileTopBoundary == picture top here) */
                        for (int x = x0; x < x0 + size && x < W; x += 4)
                            h[((x & 63) >> 2) + ((y0 & 63) >> 2) * 16] = port_bs_pu(cu, &cu_map[(size_t)((y0 >> 3) - 1) * cu_stride + (x >> 3)], cbf[(size_t)(y0 >> 2) * cbf_stride + (x >> 2)],
                                                                                   cbf[(size_t)((y0 >> 2) - 1) * cbf_stride + (x >> 2)], slice_type, poc0, poc1);
                    if (x0 != 0)
                        for (int y = y0; y < y0 + size && y < H; y += 4)
                            v[((x0 & 63) >> 2) + ((y & 63) >> 2) * 16] = port_bs_pu(cu, &cu_map[(size_t)(y >> 3) * cu_stride + (x0 >> 3) - 1], cbf[(size_t)(y >> 2) * cbf_stride + (x0 >> 2)],
                                                                                   cbf[(size_t)(y >> 2) * cbf_stride + (x0 >> 2) - 1], slice_type, poc0, poc1);
                    if (size == 64 && !cu->intra)      /* TU boundaries inside the CU: the 32-sample cross; only the inter path sets them (EbCodingLoop.c:4328) */
                        for (int t = 0; t < 64; t += 4) {
                            if (x0 + t < W && y0 + 32 < H)
                                h[(t >> 2) + 8 * 16] = (cbf[(size_t)((y0 + 32) >> 2) * cbf_stride + ((x0 + t) >> 2)] || cbf[(size_t)(((y0 + 32) >> 2) - 1) * cbf_stride + ((x0 + t) >> 2)]);
                            if (y0 + t < H && x0 + 32 < W)
                                v[8 + (t >> 2) * 16] = (cbf[(size_t)((y0 + t) >> 2) * cbf_stride + ((x0 + 32) >> 2)] || cbf[(size_t)((y0 + t) >> 2) * cbf_stride + ((x0 + 32) >> 2) - 1]);
                        }
                }
        }
    return 0;
}
