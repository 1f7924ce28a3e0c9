/*
 * oracle/port/tq_port.c -- TEST INFRASTRUCTURE ONLY (see me_port.c header for the rules).
 *
 * Plain-C restatement of the reference's residual / transform / quantisation / reconstruction arithmetic:
 *   forward:  Transform4x4 / DstTransform4x4 / Transform8x8 / Transform16x16Estimate / Transform32x32Estimate
 *             (C_DEFAULT/EbTransforms_C.c:1634-1889, butterflies :492-1073) -- the functions every row of
 *             transformFunctionTableEstimate (Codec/EbTransforms.h:339) resolves to; EstimateTransform shapes
 *             (Codec/EbTransforms.c:3268): N2/N4 keep the low-frequency quadrant, ONLY_DC the single-rounding DC
 *   quant:    QuantizeInvQuantize (C_DEFAULT/EbTransforms_C.c:89), parameter derivation + sparse cleaning of
 *             UnifiedQuantizeInvQuantize (Codec/EbTransforms.c:2978-3261), UpdateQiQCoef (C_DEFAULT/EbTransforms_C.c:209)
 *   inverse:  InvTransform{4,8,16,32} / InvDstTransform4x4 (C_DEFAULT/EbTransforms_C.c:1910-2119)
 *   chain:    EncodeLoop / EncodeGenerateRecon (Codec/EbCodingLoop.c:651,1084)
 * Written as plain integer matrix products; the reference's partial butterflies factor the same matrices exactly.
 * One reference detail is kept: the 16/32-point "Estimate" butterflies hold their first add/sub stage in 16 bits.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "svt_hevc_b200.h"

/* HEVC core transform: M_N[k][n] = C((k * (2n+1) * (32/N)) mod 128), C = cosine table scaled to the standard's integers */
static const int COS_TAB[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
static int dct_coef(int N, int k, int n)
{
    const int m = (k * (2 * n + 1) * (32 / N)) & 127;
    if (m <= 32) return COS_TAB[m];
    if (m <= 64) return -COS_TAB[64 - m];
    if (m <= 96) return -COS_TAB[m - 64];
    return COS_TAB[128 - m];
}
static const int DST4[4][4] = { { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };

static inline int16_t clip16(int v) { return (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

/* one 1-D forward pass over `rows` rows: out[k * ostride + r] = (sum_n M[k][n] * in[r * istride + n] + off) >> shift, k < kmax */
static void fwd_pass(const int16_t *in, int istride, int16_t *out, int ostride, int N, int rows, int kmax, int shift, int dst, int narrow16)
{
    const int off = 1 << (shift - 1);
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < kmax; ++k) {
            int acc = 0;
            if (dst) for (int n = 0; n < 4; ++n) acc += DST4[k][n] * in[r * istride + n];
            else if (narrow16) {
                /* PartialButterfly{16,32}Estimate keep even/odd sums in EB_S16 (C_DEFAULT/EbTransforms_C.c:500-560, 835-860) */
                for (int n = 0; n < N / 2; ++n) {
                    const int16_t t = (k & 1) ? (int16_t)(in[r * istride + n] - in[r * istride + N - 1 - n]) : (int16_t)(in[r * istride + n] + in[r * istride + N - 1 - n]);
                    acc += dct_coef(N, k, n) * t;
                }
            } else for (int n = 0; n < N; ++n) acc += dct_coef(N, k, n) * in[r * istride + n];
            out[k * ostride + r] = (int16_t)((acc + off) >> shift);
        }
}

static int ilog2(uint32_t v) { int l = 0; while (v > 1) { v >>= 1; ++l; } return l; }

/* EstimateTransform, Codec/EbTransforms.c:3268 */
void port_estimate_transform(const int16_t *res, uint32_t rs, int16_t *coeff, uint32_t cs, uint32_t N, uint32_t bit_inc, uint8_t dst, uint8_t shape)
{
    const int lg = ilog2(N);
    if (shape == SVTB200_SHAPE_ONLY_DC) {
        int sum = 0;
        for (uint32_t y = 0; y < N; ++y) for (uint32_t x = 0; x < N; ++x) sum += res[y * rs + x];
        const int shift = (lg - 1 + (int)bit_inc) + (lg + 6);
        coeff[0] = (int16_t)((64 * 64 * sum + (1 << (shift - 1))) >> shift);
        return;
    }
    int s1, s2;
    if (N == 32) { s1 = 6 + (int)bit_inc; s2 = 9; } else if (N == 16) { s1 = 4 + (int)bit_inc; s2 = 9; }
    else { s1 = lg - 1 + (int)bit_inc; s2 = lg + 6; }
    const int act = (int)N >> shape;       /* N2 / N4: only the low-frequency act x act corner is defined */
    int16_t tmp[32 * 32];
    fwd_pass(res, (int)rs, tmp, (int)N, (int)N, (int)N, N == 4 ? 4 : act, s1, dst, N >= 16);
    /* second pass over the act kept horizontal frequencies */
    int16_t out[32 * 32];
    fwd_pass(tmp, (int)N, out, (int)N, (int)N, N == 4 ? 4 : act, N == 4 ? 4 : act, s2, dst, N >= 16);
    const int keep = N == 4 ? 4 : act;
    for (int v = 0; v < keep; ++v) for (int u = 0; u < keep; ++u) coeff[v * cs + u] = out[v * N + u];
}

/* QuantizeInvQuantize, C_DEFAULT/EbTransforms_C.c:89 */
void port_qiq(const int16_t *coeff, uint32_t cs, int16_t *q, int16_t *r, uint32_t qFunc, uint32_t q_offset, int32_t shiftedQBits,
              int32_t shiftedFFunc, int32_t iq_offset, int32_t shiftNum, uint32_t area, uint32_t *nz)
{
    *nz = 0;
    for (uint32_t y = 0; y < area; ++y)
        for (uint32_t x = 0; x < area; ++x) {
            const int c = coeff[y * cs + x];
            int32_t a = (int32_t)((uint32_t)abs(c) * qFunc + q_offset);   /* EB_S32 arithmetic like the reference */
            a >>= shiftedQBits;
            const int16_t qv = clip16(c < 0 ? -a : a);
            q[y * cs + x] = qv;
            *nz += (qv != 0);
            r[y * cs + x] = clip16((qv * shiftedFFunc + iq_offset) >> shiftNum);
        }
}

static void inv_pass(const int16_t *in, int istride, int16_t *out, int ostride, int N, int shift, int dst)
{
    const int off = 1 << (shift - 1);
    /* out[r * ostride + n] = clip16((sum_k M[k][n] * in[k * istride + r] + off) >> shift) */
    for (int r = 0; r < N; ++r)
        for (int n = 0; n < N; ++n) {
            int acc = 0;
            for (int k = 0; k < N; ++k) acc += (dst ? DST4[k][n] : dct_coef(N, k, n)) * in[k * istride + r];
            out[r * ostride + n] = clip16((acc + off) >> shift);
        }
}
/* EncodeInvTransform, Codec/EbTransforms.c:3502 (the isOnlyDc shortcut yields the same samples) */
void port_inv_transform(const int16_t *coeff, uint32_t cs, int16_t *res, uint32_t rs, uint32_t N, uint32_t bit_inc, uint8_t dst)
{
    int16_t tmp[32 * 32], out[32 * 32];
    inv_pass(coeff, (int)cs, tmp, (int)N, (int)N, 7, dst);             /* SHIFT_INV_1ST */
    inv_pass(tmp, (int)N, out, (int)N, (int)N, 12 - (int)bit_inc, dst);   /* SHIFT_INV_2ND - bitIncrement */
    for (uint32_t y = 0; y < N; ++y) for (uint32_t x = 0; x < N; ++x) res[y * rs + x] = out[y * N + x];
}

static const svtb200_plane *pl(const svtb200_yuv_picture *p, int k) { return k == 0 ? &p->y : k == 1 ? &p->cb : &p->cr; }

/* hbd: 16-bit sample planes of 10-bit video -- EncodeLoop16bit / EncodeGenerateRecon16bit (Codec/EbCodingLoop.c:1240,1640): EncodeTransform with
 * BIT_INCREMENT_10BIT (Codec/EbTransforms.c:3343; two-stage rounding in its ONLY_DC branch :3425-3447), quantiser at bitDepth 10 with
 * qp + QP_BD_OFFSET, inverse second shift 12 - 2, reconstruction clipped to 1023 (PictureAdditionKernel16bit). */
static int tq_batch_impl(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                         const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts, int hbd)
{
    const int bps = hbd ? 2 : 1, bit_inc = hbd ? 2 : 0, maxv = hbd ? 1023 : 255;
#define LD(ptr, i) (hbd ? (int)((const uint16_t *)(ptr))[i] : (int)((const uint8_t *)(ptr))[i])
    static const uint32_t QF[6] = { 26214, 23302, 20560, 18396, 16384, 14564 }, FF[6] = { 40, 45, 51, 57, 64, 72 };   /* EbTransforms.h:26-27 */
    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_tq_job *t = &jobs[j];
        const svtb200_plane *sp = pl(src, t->plane), *pp = pl(pred, t->plane), *rp = pl(recon, t->plane);
        const int N = t->size;
        const uint8_t *s = sp->buf + ((size_t)(sp->origin_y + t->y) * sp->stride + sp->origin_x + t->x) * bps;
        const uint8_t *p = pp->buf + ((size_t)(pp->origin_y + t->y) * pp->stride + pp->origin_x + t->x) * bps;
        uint8_t *r = rp->buf + ((size_t)(rp->origin_y + t->y) * rp->stride + rp->origin_x + t->x) * bps;
        int16_t res[32 * 32], tr[32 * 32], q[32 * 32], rc[32 * 32];
        memset(tr, 0, sizeof tr); memset(q, 0, sizeof q); memset(rc, 0, sizeof rc);
        for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) res[y * N + x] = (int16_t)(LD(s, (size_t)y * sp->stride + x) - LD(p, (size_t)y * pp->stride + x));
        if (hbd && t->shape == SVTB200_SHAPE_ONLY_DC) {
            int sum = 0;
            for (int i = 0; i < N * N; ++i) sum += res[i];
            const int lgn = ilog2((uint32_t)N), sh1 = lgn - 1 + bit_inc, sh2 = lgn + 6;
            const int d1 = (64 * sum + (1 << (sh1 - 1))) >> sh1;
            tr[0] = (int16_t)((64 * d1 + (1 << (sh2 - 1))) >> sh2);
        } else
            port_estimate_transform(res, (uint32_t)N, tr, (uint32_t)N, (uint32_t)N, (uint32_t)bit_inc, t->dst4x4, t->shape);
        /* UnifiedQuantizeInvQuantize parameter derivation, Codec/EbTransforms.c:3047-3060 / :3094-3110 */
        const int qp = t->qp + (hbd ? 12 : 0);     /* QP_BD_OFFSET */
        const int qpRem = qp % 6, qpPer = qp / 6;
        const int tshift = 15 - (hbd ? 10 : 8) - ilog2((uint32_t)N);
        const int shiftedQBits = 14 + qpPer + tshift;
        const uint32_t qoff = (uint32_t)((t->slice_type >= 2) ? 171 : 85) << (shiftedQBits - 9);
        const int shiftedFFunc = qpPer > 8 ? (int)FF[qpRem] << (qpPer - 2) : (int)FF[qpRem] << qpPer;
        const int shiftNum = qpPer > 8 ? 20 - 14 - tshift - 2 : 20 - 14 - tshift;
        const int iqoff = 1 << (shiftNum - 1);
        uint32_t nz = 0;
        int act;
        if (t->shape == SVTB200_SHAPE_ONLY_DC) {
            act = 1;
            port_qiq(tr, (uint32_t)N, q, rc, QF[qpRem], qoff, shiftedQBits, shiftedFFunc, iqoff, shiftNum, 1, &nz);
        } else {
            act = N >> t->shape;
            const uint32_t aoff = t->dz_offset ? (uint32_t)(t->dz_offset * (1u << shiftedQBits) / 20) : qoff;
            port_qiq(tr, (uint32_t)N, q, rc, QF[qpRem], aoff, shiftedQBits, shiftedFFunc, iqoff, shiftNum, (uint32_t)act, &nz);
            if ((t->flags & SVTB200_TQ_CLEAN_SPARSE) && N >= 8 && nz && t->slice_type != 2 && act >= 4) {
                /* isolated coefficient cleaning: a 4x4 group holding exactly one non-zero loses it, except the DC position */
                for (int by = 0; by < act / 4; ++by)
                    for (int bx = 0; bx < act / 4; ++bx) {
                        int cnt = 0;
                        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) cnt += q[(by * 4 + y) * N + bx * 4 + x] != 0;
                        if (cnt != 1) continue;
                        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) {
                            const int loc = (by * 4 + y) * N + bx * 4 + x;
                            if (q[loc] && loc) { q[loc] = 0; rc[loc] = 0; --nz; }
                        }
                    }
            }
            /* UpdateQiQCoef, C_DEFAULT/EbTransforms_C.c:209 -- note the (EB_S16) cast binds before the shift */
            if (nz < 10 && (t->flags & SVTB200_TQ_CONTOURING) && t->slice_type == 2 && t->temporal_layer == 0 && t->plane == 0) {
                const int loc = (act - 1) + (act - 1) * N;
                if (q[loc] == 0) { ++nz; q[loc] = 1; rc[loc] = (int16_t)((int16_t)(shiftedFFunc + iqoff) >> shiftNum); }
            }
            if (nz == 0 && (t->flags & SVTB200_TQ_ENABLE_CBF)) {
                const int loc = (act - 2) * N + (act - 1);
                nz = 1; q[loc] = 1; rc[loc] = (int16_t)((int16_t)(shiftedFFunc + iqoff) >> shiftNum);
            }
        }
        nz_counts[j] = nz;
        int16_t *co = coeff->buf[t->plane] + (size_t)t->y * coeff->stride[t->plane] + t->x;
        for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) co[(size_t)y * coeff->stride[t->plane] + x] = (y < act && x < act) ? q[y * N + x] : 0;
        if (t->flags & SVTB200_TQ_SKIP_RECON) continue;
        if (nz) {
            /* PfZeroOutUselessQuadrants: everything outside the active corner is zero for the inverse */
            for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) if (y >= act || x >= act) rc[y * N + x] = 0;
            int16_t rr[32 * 32];
            port_inv_transform(rc, (uint32_t)N, rr, (uint32_t)N, (uint32_t)N, (uint32_t)bit_inc, t->dst4x4);
            for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
                int v = LD(p, (size_t)y * pp->stride + x) + rr[y * N + x];
                v = v < 0 ? 0 : v > maxv ? maxv : v;
                if (hbd) ((uint16_t *)r)[(size_t)y * rp->stride + x] = (uint16_t)v; else r[(size_t)y * rp->stride + x] = (uint8_t)v;
            }
        } else if (r != p) {
            for (int y = 0; y < N; ++y) memcpy(r + (size_t)y * rp->stride * bps, p + (size_t)y * pp->stride * bps, (size_t)N * bps);
        }
    }
#undef LD
    return 0;
}

int port_tq_batch(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                  const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts)
{ return tq_batch_impl(src, pred, recon, coeff, jobs, n_jobs, nz_counts, 0); }
int port_tq_batch16(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                    const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts)
{ return tq_batch_impl(src, pred, recon, coeff, jobs, n_jobs, nz_counts, 1); }
