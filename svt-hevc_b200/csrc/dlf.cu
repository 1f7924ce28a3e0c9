// dlf.cu -- HEVC in-loop deblocking of a whole picture, 8-bit 4:2:0.
//
// Replaces:
//   Luma4SampleEdgeDLFCore / Chroma2SampleEdgeDLFCore             C_DEFAULT/EbDeblockingFilter_C.c:39,469 (+ _SSSE3 rows)
//   LCUInternalAreaDLFCore / LCUBoundaryDLFCore / LCUPictureEdgeDLFCore   Codec/EbDeblockingFilter.c:2222,2828,3518
//     (traversal, QP averaging, tc / beta lookup :2303-2336, chroma QP mapping :2430-2436, tables :30-38)
//
// Design: two launches per picture -- every vertical edge, then every horizontal edge on the result (the HEVC order; it
// equals the reference's per-LCU three-driver order once all LCUs are done, tests/test_gpu_dlf.py).  One thread per
// 4-sample luma (2-sample chroma) edge segment.  Threads are numbered along the picture row so that the 8-byte
// footprints of neighbouring vertical-edge segments tile each row contiguously (fully coalesced 32-bit accesses);
// horizontal-edge segments read one aligned word per row.  A segment with strength 0 costs one BS byte.  The kernel is
// HBM-bound: each pass reads and writes each touched sample once.
#include "common.cuh"

namespace svtb200 {

__constant__ uint8_t c_tc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6,
                                 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};                      // TcTable_8x8, Codec/EbDeblockingFilter.c:30
__constant__ uint8_t c_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28,
                                   30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};   // BetaTable_8x8, :35
__constant__ uint8_t c_cqp[58] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32,
                                  33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};   // ChromaQpMap, Codec/EbTransforms.c:174

struct DlfArgs {
    uint8_t *y, *cb, *cr;      // interior origins (top-left visible sample)
    int ys, cs;                // luma / chroma pitch
    const uint8_t *vbs, *hbs, *qp;
    int W, H, cols, qs;
    int tc_off, beta_off, cb_off, cr_off;
};

__device__ __forceinline__ int clampi(int lo, int hi, int v) { return min(max(v, lo), hi); }

// s[r][k]: r = position along the edge (0..3), k = 0..7 = p3,p2,p1,p0,q0,q1,q2,q3.  Returns false when nothing changes.
__device__ __forceinline__ bool luma_edge(int (&s)[4][8], int tc, int beta, int maxv = 255)
{
    const int dp0 = abs(s[0][1] - 2 * s[0][2] + s[0][3]), dp3 = abs(s[3][1] - 2 * s[3][2] + s[3][3]);
    const int dq0 = abs(s[0][6] - 2 * s[0][5] + s[0][4]), dq3 = abs(s[3][6] - 2 * s[3][5] + s[3][4]);
    const int dp = dp0 + dp3, dq = dq0 + dq3, d0 = dp0 + dq0, d3 = dp3 + dq3;
    if (d0 + d3 >= beta) return false;
    const bool s0 = ((d0 << 1) < (beta >> 2)) && (beta >> 3) > abs(s[0][0] - s[0][3]) + abs(s[0][7] - s[0][4]) && ((5 * tc + 1) >> 1) > abs(s[0][3] - s[0][4]);
    const bool s3 = ((d3 << 1) < (beta >> 2)) && (beta >> 3) > abs(s[3][0] - s[3][3]) + abs(s[3][7] - s[3][4]) && ((5 * tc + 1) >> 1) > abs(s[3][3] - s[3][4]);
    const bool strong = s0 && s3;
    const int side = (beta + (beta >> 1)) >> 3, tc2 = tc >> 1;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int p3 = s[r][0], p2 = s[r][1], p1 = s[r][2], p0 = s[r][3], q0 = s[r][4], q1 = s[r][5], q2 = s[r][6], q3 = s[r][7];
        if (strong) {
            s[r][4] = clampi(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            s[r][3] = clampi(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            s[r][5] = clampi(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
            s[r][2] = clampi(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
            s[r][6] = clampi(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
            s[r][1] = clampi(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else {
            int delta = ((q0 - p0) * 9 - (q1 - p1) * 3 + 8) >> 4;
            if (abs(delta) < tc * 10) {
                delta = clampi(-tc, tc, delta);
                s[r][4] = clampi(0, maxv, q0 - delta);
                s[r][3] = clampi(0, maxv, p0 + delta);
                if (side > dp) s[r][2] = clampi(0, maxv, p1 + clampi(-tc2, tc2, (((p0 + p2 + 1) >> 1) - p1 + delta) >> 1));
                if (side > dq) s[r][5] = clampi(0, maxv, q1 + clampi(-tc2, tc2, (((q0 + q2 + 1) >> 1) - q1 - delta) >> 1));
            }
        }
    }
    return true;
}

__device__ __forceinline__ void chroma_edge(int &p1, int &p0, int &q0, int &q1, int tc, int maxv = 255)
{
    const int delta = clampi(-tc, tc, (((q0 - p0) << 2) + p1 - q1 + 4) >> 3);
    p0 = clampi(0, maxv, p0 + delta);
    q0 = clampi(0, maxv, q0 - delta);
}

__device__ __forceinline__ int bs_at(const DlfArgs &A, const uint8_t *arr, int x, int y)
{
    return arr[((size_t)(y >> 6) * A.cols + (x >> 6)) * 256 + ((x & 63) >> 2) + (((y & 63) >> 2) << 4)];
}
__device__ __forceinline__ int qp_at(const DlfArgs &A, int x, int y) { return A.qp[(x >> 3) + (y >> 3) * A.qs]; }
__device__ __forceinline__ int chroma_qp(int q) { return q < 0 ? q : (q > 57 ? q - 6 : c_cqp[q]); }

// PASS 0: vertical edges, PASS 1: horizontal edges
// HBD: 16-bit sample planes (10-bit video): LCU*DLFCore16bit (Codec/EbDeblockingFilter.c:2514,3174,3922), Luma4SampleEdgeDLFCore16bit /
// Chroma2SampleEdgeDLFCore16bit (C_DEFAULT/EbDeblockingFilter_C.c:240,524): tc and beta scaled by 4 (:1254-1255), samples clipped to 1023
template <int PASS, bool HBD>
__global__ void __launch_bounds__(256) dlf_pass_kernel(const __grid_constant__ DlfArgs A)
{
    const int W = A.W, H = A.H;
    const int n_luma = PASS == 0 ? (W / 8 - 1) * (H / 4) : (W / 4) * (H / 8 - 1);
    const int ncv = (W / 2 + 7) / 8 - 1, nch = (H / 2 + 7) / 8 - 1;   // chroma edges on the 8-sample chroma grid
    const int n_chroma = PASS == 0 ? ncv * (H / 4) : (W / 4) * nch;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_luma) {
        int x, y;
        if (PASS == 0) { const int per = W / 8 - 1; y = (i / per) * 4; x = (i % per + 1) * 8; }
        else { const int per = W / 4; y = (i / per + 1) * 8; x = (i % per) * 4; }
        const int bs = bs_at(A, PASS == 0 ? A.vbs : A.hbs, x, y);
        if (!bs) return;
        const int q = (qp_at(A, x, y) + (PASS == 0 ? qp_at(A, x - 1, y) : qp_at(A, x, y - 1)) + 1) >> 1;
        const int tc = c_tc[clampi(0, 53, q + ((bs > 1) << 1) + A.tc_off)] << (HBD ? 2 : 0), beta = c_beta[clampi(0, 51, q + A.beta_off)] << (HBD ? 2 : 0);
        int s[4][8];
        if (HBD) {
            uint16_t *Y = (uint16_t *)A.y;
            const ptrdiff_t fs = PASS == 0 ? 1 : A.ys, rs = PASS == 0 ? A.ys : 1;     // across / along the edge
            uint16_t *e = Y + (ptrdiff_t)y * A.ys + x;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) s[r][k] = e[(k - 4) * fs + r * rs];
            if (!luma_edge(s, tc, beta, 1023)) return;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 1; k < 7; ++k) e[(k - 4) * fs + r * rs] = (uint16_t)s[r][k];
            return;
        }
        if (PASS == 0) {
            uint32_t *row = (uint32_t *)(A.y + (size_t)y * A.ys + x - 4);
            const int rs = A.ys >> 2;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t a = row[r * rs], b = row[r * rs + 1];
#pragma unroll
                for (int k = 0; k < 4; ++k) { s[r][k] = (a >> (8 * k)) & 255; s[r][4 + k] = (b >> (8 * k)) & 255; }
            }
            if (!luma_edge(s, tc, beta)) return;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                row[r * rs] = (uint32_t)s[r][0] | (s[r][1] << 8) | (s[r][2] << 16) | ((uint32_t)s[r][3] << 24);
                row[r * rs + 1] = (uint32_t)s[r][4] | (s[r][5] << 8) | (s[r][6] << 16) | ((uint32_t)s[r][7] << 24);
            }
        } else {
            uint32_t *col = (uint32_t *)(A.y + (size_t)(y - 4) * A.ys + x);
            const int rs = A.ys >> 2;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t a = col[k * rs];
#pragma unroll
                for (int r = 0; r < 4; ++r) s[r][k] = (a >> (8 * r)) & 255;
            }
            if (!luma_edge(s, tc, beta)) return;
#pragma unroll
            for (int k = 1; k < 7; ++k) col[k * rs] = (uint32_t)s[0][k] | (s[1][k] << 8) | (s[2][k] << 16) | ((uint32_t)s[3][k] << 24);
        }
        return;
    }
    i -= n_luma;
    if (i >= n_chroma) return;
    int cx, cy;
    if (PASS == 0) { const int per = ncv; cy = (i / per) * 2; cx = (i % per + 1) * 8; }
    else { const int per = W / 4; cy = (i / per + 1) * 8; cx = (i % per) * 2; }
    const int lx = 2 * cx, ly = 2 * cy;
    if (bs_at(A, PASS == 0 ? A.vbs : A.hbs, lx, ly) <= 1) return;
    const int q = (qp_at(A, lx, ly) + (PASS == 0 ? qp_at(A, lx - 2, ly) : qp_at(A, lx, ly - 2)) + 1) >> 1;
    const int tcv[2] = {c_tc[clampi(0, 53, chroma_qp(q + A.cb_off) + 2 + A.tc_off)] << (HBD ? 2 : 0), c_tc[clampi(0, 53, chroma_qp(q + A.cr_off) + 2 + A.tc_off)] << (HBD ? 2 : 0)};
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
        if (HBD) {
            uint16_t *e = (uint16_t *)(pl == 0 ? A.cb : A.cr) + (ptrdiff_t)cy * A.cs + cx;
            const ptrdiff_t fs = PASS == 0 ? 1 : A.cs, rs = PASS == 0 ? A.cs : 1;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                int p1 = e[-2 * fs + r * rs], p0 = e[-fs + r * rs], q0 = e[r * rs], q1 = e[fs + r * rs];
                chroma_edge(p1, p0, q0, q1, tcv[pl], 1023);
                e[-fs + r * rs] = (uint16_t)p0; e[r * rs] = (uint16_t)q0;
            }
            continue;
        }
        uint8_t *base = (pl == 0 ? A.cb : A.cr) + (size_t)cy * A.cs + cx;
        if (PASS == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint16_t *h = (uint16_t *)(base + (size_t)r * A.cs - 2);   // cx - 2 is even
                const uint32_t a = h[0], b = h[1];
                int p1 = a & 255, p0 = a >> 8, q0 = b & 255, q1 = b >> 8;
                chroma_edge(p1, p0, q0, q1, tcv[pl]);
                h[0] = (uint16_t)(p1 | (p0 << 8));
                h[1] = (uint16_t)(q0 | (q1 << 8));
            }
        } else {
            uint16_t *h = (uint16_t *)(base - 2 * (size_t)A.cs);
            const int hs = A.cs >> 1;
            const uint32_t a = h[0], b = h[hs], c = h[2 * hs], d = h[3 * hs];
            int p1 = a & 255, p0 = b & 255, q0 = c & 255, q1 = d & 255;
            int P1 = a >> 8, P0 = b >> 8, Q0 = c >> 8, Q1 = d >> 8;
            chroma_edge(p1, p0, q0, q1, tcv[pl]);
            chroma_edge(P1, P0, Q0, Q1, tcv[pl]);
            h[hs] = (uint16_t)(p0 | (P0 << 8));
            h[2 * hs] = (uint16_t)(q0 | (Q0 << 8));
        }
    }
}

__global__ void dlf_leaf_luma_kernel(uint8_t *blk /* 8x8, pitch 8, edge in the middle */, int vertical, int tc, int beta)
{
    int s[4][8];
    for (int r = 0; r < 4; ++r) for (int k = 0; k < 8; ++k) s[r][k] = vertical ? blk[r * 8 + k] : blk[k * 8 + r];
    if (!luma_edge(s, tc, beta)) return;
    for (int r = 0; r < 4; ++r) for (int k = 0; k < 8; ++k) { if (vertical) blk[r * 8 + k] = (uint8_t)s[r][k]; else blk[k * 8 + r] = (uint8_t)s[r][k]; }
}
__global__ void dlf_leaf_chroma_kernel(uint8_t *blk /* [plane][4 across][2 along] */, int cb_tc, int cr_tc)
{
    for (int pl = 0; pl < 2; ++pl)
        for (int r = 0; r < 2; ++r) {
            uint8_t *b = blk + pl * 8 + r;
            int p1 = b[0], p0 = b[2], q0 = b[4], q1 = b[6];
            chroma_edge(p1, p0, q0, q1, pl ? cr_tc : cb_tc);
            b[2] = (uint8_t)p0; b[4] = (uint8_t)q0;
        }
}

static int dlf_launch(const DlfArgs &A, cudaStream_t st, bool hbd = false)
{
    const int W = A.W, H = A.H;
    const int ncv = (W / 2 + 7) / 8 - 1, nch = (H / 2 + 7) / 8 - 1;
    const int n0 = (W / 8 - 1) * (H / 4) + ncv * (H / 4), n1 = (W / 4) * (H / 8 - 1) + (W / 4) * nch;
    if (n0 > 0) { if (hbd) dlf_pass_kernel<0, true><<<(n0 + 255) / 256, 256, 0, st>>>(A); else dlf_pass_kernel<0, false><<<(n0 + 255) / 256, 256, 0, st>>>(A); SVTB200_COUNT_LAUNCH(); }
    if (n1 > 0) { if (hbd) dlf_pass_kernel<1, true><<<(n1 + 255) / 256, 256, 0, st>>>(A); else dlf_pass_kernel<1, false><<<(n1 + 255) / 256, 256, 0, st>>>(A); SVTB200_COUNT_LAUNCH(); }
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static int dlf_check(const svtb200_yuv_picture *rec, const svtb200_dlf_params *p)
{
    if (!rec || !p || !rec->y.buf || !rec->cb.buf || !rec->cr.buf || p->pic_width < 16 || p->pic_height < 16 || (p->pic_width & 7) || (p->pic_height & 7) ||
        p->qp_array_stride * 8 < p->pic_width || (rec->y.stride & 3) || (rec->cb.stride & 1) || (rec->y.origin_x & 3) || (rec->cb.origin_x & 1) ||
        rec->y.origin_x < 4 || rec->y.origin_y < 4 || rec->cb.stride != rec->cr.stride) {
        set_error("dlf_picture: bad argument (sizes must be multiples of 8, luma pitch/origin multiples of 4, chroma even, padding >= 4)");
        return SVTB200_ERR_BAD_ARG;
    }
    return SVTB200_OK;
}

static DlfArgs dlf_args(const svtb200_yuv_picture &rec, uint8_t *y, uint8_t *cb, uint8_t *cr, const uint8_t *vbs, const uint8_t *hbs, const uint8_t *qp,
                        const svtb200_dlf_params &p, size_t bps = 1)
{
    DlfArgs A;
    A.y = y + ((size_t)rec.y.origin_y * rec.y.stride + rec.y.origin_x) * bps;
    A.cb = cb + ((size_t)rec.cb.origin_y * rec.cb.stride + rec.cb.origin_x) * bps;
    A.cr = cr + ((size_t)rec.cr.origin_y * rec.cr.stride + rec.cr.origin_x) * bps;
    A.ys = (int)rec.y.stride; A.cs = (int)rec.cb.stride;
    A.vbs = vbs; A.hbs = hbs; A.qp = qp;
    A.W = p.pic_width; A.H = p.pic_height; A.cols = (p.pic_width + 63) / 64; A.qs = p.qp_array_stride;
    A.tc_off = p.tc_offset; A.beta_off = p.beta_offset; A.cb_off = p.cb_qp_offset; A.cr_off = p.cr_qp_offset;
    return A;
}

static size_t pbytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }

}  // namespace svtb200
using namespace svtb200;

extern "C" {

static int dlf_picture_dev_impl(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs, const uint8_t *qp_array,
                                const svtb200_dlf_params *prm, void *cuda_stream, bool hbd)
{
    int rc = dlf_check(recon, prm);
    if (rc) return rc;
    if (!vertical_bs || !horizontal_bs || !qp_array) { set_error("dlf_picture_dev: null array"); return SVTB200_ERR_BAD_ARG; }
    if (((uintptr_t)recon->y.buf & 3) || ((uintptr_t)recon->cb.buf & 1) || ((uintptr_t)recon->cr.buf & 1)) { set_error("dlf_picture_dev: planes must be 4-byte (luma) / 2-byte (chroma) aligned"); return SVTB200_ERR_BAD_ARG; }
    if ((rc = ensure_device())) return rc;
    return dlf_launch(dlf_args(*recon, recon->y.buf, recon->cb.buf, recon->cr.buf, vertical_bs, horizontal_bs, qp_array, *prm, hbd ? 2 : 1), (cudaStream_t)cuda_stream, hbd);
}

static int dlf_picture_impl(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs, const uint8_t *qp_array,
                            const svtb200_dlf_params *prm, bool hbd)
{
    const size_t bps = hbd ? 2 : 1;
    int rc = dlf_check(recon, prm);
    if (rc) return rc;
    if (!vertical_bs || !horizontal_bs || !qp_array) { set_error("dlf_picture: null array"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t n_lcu = (size_t)((prm->pic_width + 63) / 64) * ((prm->pic_height + 63) / 64);
    const size_t bs_bytes = n_lcu * 256, qp_bytes = (size_t)prm->qp_array_stride * ((prm->pic_height + 7) / 8);
    const svtb200_plane *pl[3] = {&recon->y, &recon->cb, &recon->cr};
    size_t need = 8192 + 2 * align_up(bs_bytes, 256) + align_up(qp_bytes, 256);
    for (int k = 0; k < 3; ++k) need += align_up(bps * pbytes(*pl[k]), 256);
    if ((rc = c->dev_reserve(need))) return rc;
    void *d[3], *d_v, *d_h, *d_q;
    for (int k = 0; k < 3; ++k) {
        if ((rc = c->dev_alloc(bps * pbytes(*pl[k]), &d[k]))) return rc;
        SVTB200_CUDA_TRY(cudaMemcpyAsync(d[k], pl[k]->buf, bps * pbytes(*pl[k]), cudaMemcpyHostToDevice, c->stream));
    }
    if ((rc = c->dev_alloc(bs_bytes, &d_v)) || (rc = c->dev_alloc(bs_bytes, &d_h)) || (rc = c->dev_alloc(qp_bytes, &d_q))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_v, vertical_bs, bs_bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_h, horizontal_bs, bs_bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_q, qp_array, qp_bytes, cudaMemcpyHostToDevice, c->stream));
    if ((rc = dlf_launch(dlf_args(*recon, (uint8_t *)d[0], (uint8_t *)d[1], (uint8_t *)d[2], (const uint8_t *)d_v, (const uint8_t *)d_h, (const uint8_t *)d_q, *prm, bps), c->stream, hbd))) return rc;
    for (int k = 0; k < 3; ++k) SVTB200_CUDA_TRY(cudaMemcpyAsync(pl[k]->buf, d[k], bps * pbytes(*pl[k]), cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_dlf_picture_dev(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs, const uint8_t *qp_array,
                            const svtb200_dlf_params *prm, void *cuda_stream)
{ return dlf_picture_dev_impl(recon, vertical_bs, horizontal_bs, qp_array, prm, cuda_stream, false); }
int svtb200_dlf_picture(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs, const uint8_t *qp_array, const svtb200_dlf_params *prm)
{ return dlf_picture_impl(recon, vertical_bs, horizontal_bs, qp_array, prm, false); }
int svtb200_dlf_picture16_dev(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs, const uint8_t *qp_array,
                              const svtb200_dlf_params *prm, void *cuda_stream)
{ return dlf_picture_dev_impl(recon, vertical_bs, horizontal_bs, qp_array, prm, cuda_stream, true); }
int svtb200_dlf_picture16(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs, const uint8_t *qp_array, const svtb200_dlf_params *prm)
{ return dlf_picture_impl(recon, vertical_bs, horizontal_bs, qp_array, prm, true); }

int svtb200_Luma4SampleEdgeDLFCore(uint8_t *e, uint32_t stride, uint8_t isVerticalEdge, int32_t tc, int32_t beta)
{
    if (!e || stride == 0) { set_error("Luma4SampleEdgeDLFCore: bad argument"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    void *d, *h;
    if ((rc = c->dev_alloc(64, &d)) || (rc = c->pin_alloc(64, &h))) return rc;
    uint8_t *hb = (uint8_t *)h;
    // gather the 8 (across) x 4 (along) footprint into an 8x8 staging tile with the edge in the middle
    const ptrdiff_t fs = isVerticalEdge ? 1 : (ptrdiff_t)stride, rs = isVerticalEdge ? (ptrdiff_t)stride : 1;
    memset(hb, 0, 64);
    for (int r = 0; r < 4; ++r) for (int k = -4; k < 4; ++k) { if (isVerticalEdge) hb[r * 8 + k + 4] = e[k * fs + r * rs]; else hb[(k + 4) * 8 + r] = e[k * fs + r * rs]; }
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d, hb, 64, cudaMemcpyHostToDevice, c->stream));
    dlf_leaf_luma_kernel<<<1, 1, 0, c->stream>>>((uint8_t *)d, isVerticalEdge ? 1 : 0, tc, beta);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    SVTB200_CUDA_TRY(cudaMemcpyAsync(hb, d, 64, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < 4; ++r) for (int k = -3; k < 3; ++k) e[k * fs + r * rs] = isVerticalEdge ? hb[r * 8 + k + 4] : hb[(k + 4) * 8 + r];
    return SVTB200_OK;
}

int svtb200_Chroma2SampleEdgeDLFCore(uint8_t *cb, uint8_t *cr, uint32_t stride, uint8_t isVerticalEdge, uint8_t cbTc, uint8_t crTc)
{
    if (!cb || !cr || stride == 0) { set_error("Chroma2SampleEdgeDLFCore: bad argument"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    void *d, *h;
    if ((rc = c->dev_alloc(16, &d)) || (rc = c->pin_alloc(16, &h))) return rc;
    uint8_t *hb = (uint8_t *)h;
    const ptrdiff_t fs = isVerticalEdge ? 1 : (ptrdiff_t)stride, rs = isVerticalEdge ? (ptrdiff_t)stride : 1;
    uint8_t *pl[2] = {cb, cr};
    for (int p = 0; p < 2; ++p) for (int r = 0; r < 2; ++r) for (int k = -2; k < 2; ++k) hb[p * 8 + (k + 2) * 2 + r] = pl[p][k * fs + r * rs];
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d, hb, 16, cudaMemcpyHostToDevice, c->stream));
    dlf_leaf_chroma_kernel<<<1, 1, 0, c->stream>>>((uint8_t *)d, cbTc, crTc);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    SVTB200_CUDA_TRY(cudaMemcpyAsync(hb, d, 16, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int p = 0; p < 2; ++p) for (int r = 0; r < 2; ++r) for (int k = -1; k < 1; ++k) pl[p][k * fs + r * rs] = hb[p * 8 + (k + 2) * 2 + r];
    return SVTB200_OK;
}

}  // extern "C"
