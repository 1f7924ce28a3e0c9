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
