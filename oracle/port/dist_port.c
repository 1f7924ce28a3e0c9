/*
 * oracle/port/dist_port.c -- TEST INFRASTRUCTURE ONLY (see me_port.c header for the rules).
 *
 * Plain-C restatement of the distortion leaves mode decision calls, per block:
 *   port_spatial_distortion   SpatialFullDistortionKernel                   C_DEFAULT/EbPictureOperators_C.c:697
 *   port_full_distortion      FullDistortionKernel{,Intra,CbfZero}_32bit    C_DEFAULT/EbPictureOperators_C.c:385-475 (32-bit sums, zero-extended)
 *   port_satd8x8 / port_satd4x4  Compute8x8Satd (C_DEFAULT/EbPictureOperators_C.c:482) / Compute4x4Satd (Codec/EbHmCode.c:41) on an s16 block
 * and the batch drivers the CUDA entry points are compared with (one job = one block, svtb200_dist_job).
 * Pinned against the reference's own functions (C and SSE rows) by tests/test_oracle_dist.py.
 */
#include <stdint.h>
#include <stdlib.h>

#include "svt_hevc_b200.h"

uint64_t port_spatial_distortion(const uint8_t *in, uint32_t is, const uint8_t *rec, uint32_t rs, uint32_t w, uint32_t h)
{
    uint64_t s = 0;
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) { const int64_t d = (int64_t)in[y * is + x] - rec[y * rs + x]; s += (uint64_t)(d * d); }
    return s;
}

void port_full_distortion(const int16_t *c, uint32_t cs, const int16_t *r, uint32_t rs, uint64_t out[2], uint32_t w, uint32_t h, int kind)
{
    uint32_t res = 0, prd = 0;
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const int16_t cv = c[y * cs + x];
            if (kind != SVTB200_FULLDIST_CBF_ZERO) { const int16_t d = (int16_t)(cv - r[y * rs + x]); res += (uint32_t)((int32_t)d * d); }
            prd += (uint32_t)((int32_t)cv * cv);
        }
    out[0] = kind == SVTB200_FULLDIST_CBF_ZERO ? prd : res;
    out[1] = kind == SVTB200_FULLDIST_INTRA ? res : prd;
}

/* 2-D Hadamard of an NxN s16 block (N = 4, 8): rows then columns; every intermediate is stored in 16 bits by the reference, and since all steps are
 * sums the wrap can be taken once at the end */
static uint64_t hadamard_abs_sum(const int16_t *d, int n)
{
    int32_t m[8][8], t[8];
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) t[i] = d[j * n + i];
        for (int len = n >> 1; len >= 1; len >>= 1)
            for (int g = 0; g < n; g += 2 * len)
                for (int i = 0; i < len; ++i) { const int32_t a = t[g + i], b = t[g + i + len]; t[g + i] = a + b; t[g + i + len] = a - b; }
        for (int i = 0; i < n; ++i) m[j][i] = t[i];
    }
    uint64_t s = 0;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) t[j] = m[j][i];
        for (int len = n >> 1; len >= 1; len >>= 1)
            for (int g = 0; g < n; g += 2 * len)
                for (int k = 0; k < len; ++k) { const int32_t a = t[g + k], b = t[g + k + len]; t[g + k] = a + b; t[g + k + len] = a - b; }
        for (int j = 0; j < n; ++j) s += (uint64_t)abs((int)(int16_t)t[j]);
    }
    return s;
}
uint64_t port_satd8x8(const int16_t *diff) { return (hadamard_abs_sum(diff, 8) + 2) >> 2; }
uint64_t port_satd4x4(const int16_t *diff) { return (hadamard_abs_sum(diff, 4) + 1) >> 1; }

static const svtb200_plane *pl(const svtb200_yuv_picture *p, int k) { return k == 0 ? &p->y : k == 1 ? &p->cb : &p->cr; }
static const uint8_t *at(const svtb200_plane *p, int x, int y) { return p->buf + (size_t)(p->origin_y + y) * p->stride + p->origin_x + x; }

int port_spatial_distortion_batch(const svtb200_yuv_picture *in, const svtb200_yuv_picture *rec, const svtb200_dist_job *jobs, int n, uint64_t *out)
{
    for (int j = 0; j < n; ++j) {
        const svtb200_plane *a = pl(in, jobs[j].plane), *b = pl(rec, jobs[j].plane);
        out[j] = port_spatial_distortion(at(a, jobs[j].x, jobs[j].y), a->stride, at(b, jobs[j].x, jobs[j].y), b->stride, jobs[j].width, jobs[j].height);
    }
    return 0;
}
int port_full_distortion_batch(const svtb200_coeff_picture *c, const svtb200_coeff_picture *r, const svtb200_dist_job *jobs, int n, uint64_t *out)
{
    for (int j = 0; j < n; ++j) {
        const int p = jobs[j].plane;
        port_full_distortion(c->buf[p] + (size_t)jobs[j].y * c->stride[p] + jobs[j].x, c->stride[p], r->buf[p] + (size_t)jobs[j].y * r->stride[p] + jobs[j].x, r->stride[p],
                             out + 2 * j, jobs[j].width, jobs[j].height, jobs[j].kind);
    }
    return 0;
}
int port_satd_batch(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_dist_job *jobs, int n, uint64_t *out)
{
    for (int j = 0; j < n; ++j) {
        const svtb200_plane *a = pl(src, jobs[j].plane), *b = pl(pred, jobs[j].plane);
        const int u = jobs[j].kind == SVTB200_SATD_8X8 ? 8 : 4;
        uint64_t s = 0;
        for (int uy = 0; uy < jobs[j].height / u; ++uy)
            for (int ux = 0; ux < jobs[j].width / u; ++ux) {
                int16_t d[64];
                for (int y = 0; y < u; ++y)
                    for (int x = 0; x < u; ++x)
                        d[y * u + x] = (int16_t)(*at(a, jobs[j].x + ux * u + x, jobs[j].y + uy * u + y) - *at(b, jobs[j].x + ux * u + x, jobs[j].y + uy * u + y));
                s += u == 8 ? port_satd8x8(d) : port_satd4x4(d);
            }
        out[j] = s;
    }
    return 0;
}
