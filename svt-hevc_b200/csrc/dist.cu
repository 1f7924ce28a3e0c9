// dist.cu -- the distortion leaves mode decision calls between the kernels of the path (the first bricks of N2).
//
// Replaces:
//   SpatialFullDistortionKernel + ..4x4/8x8/16MxN_SSSE3_INTRIN      (C_DEFAULT/EbPictureOperators_C.c:697, table Codec/EbPictureOperators.h:633-660)
//       callers: the spatial-SSE full loop (Codec/EbFullLoop.c:384-392, 950-975), the SSD half-sample refinement (Codec/EbMotionEstimation.c:799-979)
//   FullDistortionKernel_32bit / ..Intra_32bit / ..CbfZero_32bit    (C_DEFAULT/EbPictureOperators_C.c:385-475, SSE2 rows ASM_SSE2/EbPictureOperators_Intrinsic_SSE2.c:16-330,
//       table Codec/EbPictureOperators.h:520-606) -- the coefficient-domain distortion of the full loop
//   Compute8x8Satd / Compute4x4Satd over NxM blocks                 (C_DEFAULT/EbPictureOperators_C.c:482, Codec/EbHmCode.c:41, drivers Codec/EbProductCodingLoop.c:2642)
// (sumResidual / memset16bitBlock, Codec/EbPictureOperators.c:668-700, are the ONLY_DC shortcut of the transform and live fused in tq.cu.)
//
// Design.  The callers evaluate these per candidate block; on the device they are batched: one WARP per job, 8 jobs per CTA, lanes stride over the
// block's 4-sample words (SSD: VABSDIFF4 + DP4A of the difference word with itself = 2 instructions per 4 samples), s16 pairs (coefficient domain:
// the reference sums in 32 bits, both rows -- the sums here wrap at 2^32 exactly like it) or 8x8 / 4x4 Hadamard units (one unit per lane, butterflies
// in registers, 16-bit wrap applied where the reference stores EB_S16).  All of them read each input byte once: HBM / L2 bound, 1 byte-op per byte.
#include "common.cuh"

namespace svtb200 {

struct DistArgs {
    const uint8_t *a[3], *b[3];      // sample planes at the interior origin (spatial / SATD)
    int sa[3], sb[3];
    const int16_t *ca[3], *cb[3];    // coefficient planes
    int csa[3], csb[3];
    const svtb200_dist_job *jobs;
    int n;
    unsigned long long *out;
};
constexpr int DIST_NT = 256;

__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__global__ void __launch_bounds__(DIST_NT) spatial_dist_kernel(const __grid_constant__ DistArgs A)
{
    const int j = blockIdx.x * (DIST_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= A.n) return;
    const svtb200_dist_job t = A.jobs[j];
    const int pl = t.plane, w4 = t.width >> 2, n = w4 * t.height;
    const uint8_t *a = A.a[pl] + (size_t)t.y * A.sa[pl] + t.x, *b = A.b[pl] + (size_t)t.y * A.sb[pl] + t.x;
    const SmallDiv dv = mk_div(w4);
    uint32_t acc = 0;      // <= 64 * 64 * 255^2 < 2^32
    for (int i = lane; i < n; i += 32) {
        const int r = udiv(i, dv), c = i - r * w4;
        const uint32_t x = ld_u32_unaligned(a + (size_t)r * A.sa[pl] + 4 * c), y = ld_u32_unaligned(b + (size_t)r * A.sb[pl] + 4 * c);
        const uint32_t d = __vabsdiffu4(x, y);
        acc = dp4a_uu(d, d, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) A.out[j] = acc;
}

__global__ void __launch_bounds__(DIST_NT) full_dist_kernel(const __grid_constant__ DistArgs A)
{
    const int j = blockIdx.x * (DIST_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= A.n) return;
    const svtb200_dist_job t = A.jobs[j];
    const int pl = t.plane, w2 = t.width >> 1, n = w2 * t.height;
    const int16_t *c = A.ca[pl] + (size_t)t.y * A.csa[pl] + t.x, *r = A.cb[pl] + (size_t)t.y * A.csb[pl] + t.x;
    const SmallDiv dv = mk_div(w2);
    uint32_t res = 0, prd = 0;      // 32-bit sums, wrapping: EB_U32 residualDistortion / _mm_add_epi32
    for (int i = lane; i < n; i += 32) {
        const int y = udiv(i, dv), x = i - y * w2;
        const uint32_t cw = *(const uint32_t *)(c + (size_t)y * A.csa[pl] + 2 * x);      // origins and strides are even: 4-byte aligned pairs
        const uint32_t rw = t.kind == SVTB200_FULLDIST_CBF_ZERO ? 0u : *(const uint32_t *)(r + (size_t)y * A.csb[pl] + 2 * x);
        const uint32_t dw = __vsub2(cw, rw);                                             // (EB_S16)(coeff - reconCoeff) / _mm_sub_epi16
        const int c0 = (short)(cw & 0xffff), c1 = (short)(cw >> 16), d0 = (short)(dw & 0xffff), d1 = (short)(dw >> 16);
        prd += (uint32_t)(c0 * c0) + (uint32_t)(c1 * c1);
        res += (uint32_t)(d0 * d0) + (uint32_t)(d1 * d1);
    }
    res = warp_sum(res); prd = warp_sum(prd);
    if (lane == 0) {
        A.out[2 * j] = t.kind == SVTB200_FULLDIST_CBF_ZERO ? prd : res;
        A.out[2 * j + 1] = t.kind == SVTB200_FULLDIST_INTRA ? res : prd;
    }
}

// one butterfly level over 8 values at distance D (the reference's three levels run D = 4, 2, 1 with the sums first)
template <int D>
__device__ __forceinline__ void had8_level(int (&v)[8])
{
    int o[8];
#pragma unroll
    for (int g = 0; g < 8; g += 2 * D)
#pragma unroll
        for (int i = 0; i < D; ++i) { o[g + i] = v[g + i] + v[g + i + D]; o[g + i + D] = v[g + i] - v[g + i + D]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = o[i];
}
// Compute8x8Satd: the sum of |H8 d H8'| over the unit, (sum + 2) >> 2.  The reference's three stages order the outputs differently from level to level
// (m2[j][4..7] = differences, then pairs ...); the multiset of final magnitudes is the same as the plain in-place butterflies' -- only a sum is taken.
__device__ __forceinline__ uint32_t satd8x8(const int (&d)[8][8])
{
    int m[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = d[j][i];
        had8_level<4>(v); had8_level<2>(v); had8_level<1>(v);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[j][i] = v[i];
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = m[j][i];
        had8_level<4>(v); had8_level<2>(v); had8_level<1>(v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += (uint32_t)abs((int)(short)v[j]);      // EB_S16 m2[][]: every intermediate is linear, so one wrap at the end equals the wraps along the way
    }
    return (s + 2) >> 2;
}
__device__ __forceinline__ uint32_t satd4x4(const int (&d)[4][4])
{
    int m[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int a = d[j][0] + d[j][2], b = d[j][1] + d[j][3], c = d[j][0] - d[j][2], e = d[j][1] - d[j][3];
        m[j][0] = a + b; m[j][1] = a - b; m[j][2] = c + e; m[j][3] = c - e;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = m[0][i] + m[2][i], b = m[1][i] + m[3][i], c = m[0][i] - m[2][i], e = m[1][i] - m[3][i];
        s += (uint32_t)abs((int)(short)(a + b)) + (uint32_t)abs((int)(short)(a - b)) + (uint32_t)abs((int)(short)(c + e)) + (uint32_t)abs((int)(short)(c - e));
    }
    return (s + 1) >> 1;
}

__global__ void __launch_bounds__(DIST_NT) satd_kernel(const __grid_constant__ DistArgs A)
{
    const int j = blockIdx.x * (DIST_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= A.n) return;
    const svtb200_dist_job t = A.jobs[j];
    const int pl = t.plane;
    const uint8_t *a = A.a[pl] + (size_t)t.y * A.sa[pl] + t.x, *b = A.b[pl] + (size_t)t.y * A.sb[pl] + t.x;
    unsigned long long acc = 0;
    if (t.kind == SVTB200_SATD_8X8) {
        const int uw = t.width >> 3, n = uw * (t.height >> 3);
        for (int u = lane; u < n; u += 32) {
            const int uy = u / uw, ux = u - uy * uw;
            int d[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint8_t *pa = a + (size_t)(8 * uy + r) * A.sa[pl] + 8 * ux, *pb = b + (size_t)(8 * uy + r) * A.sb[pl] + 8 * ux;
                const uint32_t x0 = ld_u32_unaligned(pa), x1 = ld_u32_unaligned(pa + 4), y0 = ld_u32_unaligned(pb), y1 = ld_u32_unaligned(pb + 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) { d[r][i] = (int)((x0 >> (8 * i)) & 255) - (int)((y0 >> (8 * i)) & 255); d[r][4 + i] = (int)((x1 >> (8 * i)) & 255) - (int)((y1 >> (8 * i)) & 255); }
            }
            acc += satd8x8(d);
        }
    } else {
        const int uw = t.width >> 2, n = uw * (t.height >> 2);
        for (int u = lane; u < n; u += 32) {
            const int uy = u / uw, ux = u - uy * uw;
            int d[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t x0 = ld_u32_unaligned(a + (size_t)(4 * uy + r) * A.sa[pl] + 4 * ux), y0 = ld_u32_unaligned(b + (size_t)(4 * uy + r) * A.sb[pl] + 4 * ux);
#pragma unroll
                for (int i = 0; i < 4; ++i) d[r][i] = (int)((x0 >> (8 * i)) & 255) - (int)((y0 >> (8 * i)) & 255);
            }
            acc += satd4x4(d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) A.out[j] = acc;
}

static const svtb200_plane *plane_of(const svtb200_yuv_picture *p, int k) { return k == 0 ? &p->y : k == 1 ? &p->cb : &p->cr; }
static size_t pbytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }

static int check_pics(const svtb200_yuv_picture *a, const svtb200_yuv_picture *b, const char *who)
{
    if (!a || !b) { set_error("%s: bad argument", who); return SVTB200_ERR_BAD_ARG; }
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane *pa = plane_of(a, k), *pb = plane_of(b, k);
        if (!pa->buf || !pb->buf || pa->width != pb->width || pa->height != pb->height) { set_error("%s: plane %d: null buffer or the two pictures differ in size", who, k); return SVTB200_ERR_BAD_ARG; }
    }
    return SVTB200_OK;
}
// kind_max: highest legal job.kind; unit: the block dimensions must be multiples of it (8 for SATD_8X8 jobs, checked per job)
static int check_jobs(const svtb200_dist_job *jobs, int n, const uint32_t w[3], const uint32_t h[3], int kind_max, bool satd, const char *who)
{
    for (int j = 0; j < n; ++j) {
        const svtb200_dist_job &t = jobs[j];
        const int unit = satd && t.kind == SVTB200_SATD_8X8 ? 8 : 4;
        const bool ok = t.plane <= 2 && t.kind <= kind_max && t.width >= unit && t.height >= unit && t.width <= 64 && t.height <= 64 && t.width % unit == 0 && t.height % unit == 0 &&
                        t.x + (uint32_t)t.width <= w[t.plane] && t.y + (uint32_t)t.height <= h[t.plane];
        if (!ok) { set_error("%s: job %d malformed (plane / kind / dimensions out of range or block outside the picture)", who, j); return SVTB200_ERR_BAD_ARG; }
    }
    return SVTB200_OK;
}

static void fill_samples(DistArgs &A, const svtb200_yuv_picture *a, const svtb200_yuv_picture *b)
{
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane *pa = plane_of(a, k), *pb = plane_of(b, k);
        A.a[k] = pa->buf + (size_t)pa->origin_y * pa->stride + pa->origin_x; A.sa[k] = (int)pa->stride;
        A.b[k] = pb->buf + (size_t)pb->origin_y * pb->stride + pb->origin_x; A.sb[k] = (int)pb->stride;
    }
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

static int sample_dist_dev(const svtb200_yuv_picture *a, const svtb200_yuv_picture *b, const svtb200_dist_job *jobs_dev, int n_jobs, uint64_t *out_dev, void *cuda_stream, bool satd)
{
    const char *who = satd ? "satd_batch_dev" : "spatial_distortion_batch_dev";
    int rc = check_pics(a, b, who);
    if (rc) return rc;
    if ((!jobs_dev && n_jobs) || n_jobs < 0 || !out_dev) { set_error("%s: bad argument", who); return SVTB200_ERR_BAD_ARG; }
    if ((rc = ensure_device())) return rc;
    if (n_jobs == 0) return SVTB200_OK;
    DistArgs A = {};
    fill_samples(A, a, b);
    A.jobs = jobs_dev; A.n = n_jobs; A.out = (unsigned long long *)out_dev;
    const int grid = (n_jobs + DIST_NT / 32 - 1) / (DIST_NT / 32);
    if (satd) satd_kernel<<<grid, DIST_NT, 0, (cudaStream_t)cuda_stream>>>(A); else spatial_dist_kernel<<<grid, DIST_NT, 0, (cudaStream_t)cuda_stream>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static int sample_dist_host(const svtb200_yuv_picture *a, const svtb200_yuv_picture *b, const svtb200_dist_job *jobs, int n_jobs, uint64_t *out, bool satd)
{
    const char *who = satd ? "satd_batch" : "spatial_distortion_batch";
    int rc = check_pics(a, b, who);
    if (rc) return rc;
    if ((!jobs && n_jobs) || n_jobs < 0 || (!out && n_jobs)) { set_error("%s: bad argument", who); return SVTB200_ERR_BAD_ARG; }
    const uint32_t w[3] = {a->y.width, a->cb.width, a->cr.width}, h[3] = {a->y.height, a->cb.height, a->cr.height};
    if ((rc = check_jobs(jobs, n_jobs, w, h, satd ? SVTB200_SATD_4X4 : 255, satd, who))) return rc;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    if (n_jobs == 0) return SVTB200_OK;
    c->reset();
    size_t need = (size_t)n_jobs * (sizeof(svtb200_dist_job) + 8) + 4096;
    for (int k = 0; k < 3; ++k) need += align_up(pbytes(*plane_of(a, k)), 256) + align_up(pbytes(*plane_of(b, k)), 256);
    if ((rc = c->dev_reserve(need))) return rc;
    svtb200_yuv_picture da = *a, db = *b;
    svtb200_plane *pa[3] = {&da.y, &da.cb, &da.cr}, *pb[3] = {&db.y, &db.cb, &db.cr};
    for (int k = 0; k < 3; ++k) {
        void *p;
        if ((rc = c->dev_alloc(pbytes(*pa[k]), &p))) return rc;
        SVTB200_CUDA_TRY(cudaMemcpyAsync(p, pa[k]->buf, pbytes(*pa[k]), cudaMemcpyHostToDevice, c->stream));
        pa[k]->buf = (uint8_t *)p;
        if ((rc = c->dev_alloc(pbytes(*pb[k]), &p))) return rc;
        SVTB200_CUDA_TRY(cudaMemcpyAsync(p, pb[k]->buf, pbytes(*pb[k]), cudaMemcpyHostToDevice, c->stream));
        pb[k]->buf = (uint8_t *)p;
    }
    void *dj, *dout;
    if ((rc = c->dev_alloc((size_t)n_jobs * sizeof(svtb200_dist_job), &dj)) || (rc = c->dev_alloc((size_t)n_jobs * 8, &dout))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(dj, jobs, (size_t)n_jobs * sizeof(svtb200_dist_job), cudaMemcpyHostToDevice, c->stream));
    if ((rc = sample_dist_dev(&da, &db, (const svtb200_dist_job *)dj, n_jobs, (uint64_t *)dout, c->stream, satd))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)n_jobs * 8, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_spatial_distortion_batch(const svtb200_yuv_picture *input, const svtb200_yuv_picture *recon, const svtb200_dist_job *jobs, int n_jobs, uint64_t *out)
{ return sample_dist_host(input, recon, jobs, n_jobs, out, false); }
int svtb200_spatial_distortion_batch_dev(const svtb200_yuv_picture *input, const svtb200_yuv_picture *recon, const svtb200_dist_job *jobs_dev, int n_jobs, uint64_t *out_dev, void *cuda_stream)
{ return sample_dist_dev(input, recon, jobs_dev, n_jobs, out_dev, cuda_stream, false); }
int svtb200_satd_batch(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_dist_job *jobs, int n_jobs, uint64_t *out)
{ return sample_dist_host(src, pred, jobs, n_jobs, out, true); }
int svtb200_satd_batch_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_dist_job *jobs_dev, int n_jobs, uint64_t *out_dev, void *cuda_stream)
{ return sample_dist_dev(src, pred, jobs_dev, n_jobs, out_dev, cuda_stream, true); }

int svtb200_full_distortion_batch_dev(const svtb200_coeff_picture *coeff, const svtb200_coeff_picture *recon_coeff, const svtb200_dist_job *jobs_dev, int n_jobs,
                                      uint64_t *out_dev, void *cuda_stream)
{
    if (!coeff || !recon_coeff || (!jobs_dev && n_jobs) || n_jobs < 0 || !out_dev) { set_error("full_distortion_batch_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    for (int k = 0; k < 3; ++k)
        if (!coeff->buf[k] || !recon_coeff->buf[k] || (coeff->stride[k] & 1) || (recon_coeff->stride[k] & 1) || ((uintptr_t)coeff->buf[k] & 3) || ((uintptr_t)recon_coeff->buf[k] & 3)) {
            set_error("full_distortion_batch_dev: plane %d: null buffer, odd stride or a base that is not 4-byte aligned", k); return SVTB200_ERR_BAD_ARG;
        }
    int rc = ensure_device();
    if (rc) return rc;
    if (n_jobs == 0) return SVTB200_OK;
    DistArgs A = {};
    for (int k = 0; k < 3; ++k) { A.ca[k] = coeff->buf[k]; A.csa[k] = (int)coeff->stride[k]; A.cb[k] = recon_coeff->buf[k]; A.csb[k] = (int)recon_coeff->stride[k]; }
    A.jobs = jobs_dev; A.n = n_jobs; A.out = (unsigned long long *)out_dev;
    full_dist_kernel<<<(n_jobs + DIST_NT / 32 - 1) / (DIST_NT / 32), DIST_NT, 0, (cudaStream_t)cuda_stream>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_full_distortion_batch(const svtb200_coeff_picture *coeff, const svtb200_coeff_picture *recon_coeff, uint32_t width, uint32_t height,
                                  const svtb200_dist_job *jobs, int n_jobs, uint64_t *out)
{
    if (!coeff || !recon_coeff || (!jobs && n_jobs) || n_jobs < 0 || (!out && n_jobs) || !width || !height) { set_error("full_distortion_batch: bad argument"); return SVTB200_ERR_BAD_ARG; }
    const uint32_t w[3] = {width, width / 2, width / 2}, h[3] = {height, height / 2, height / 2};
    for (int k = 0; k < 3; ++k)
        if (!coeff->buf[k] || !recon_coeff->buf[k] || coeff->stride[k] < w[k] || recon_coeff->stride[k] < w[k]) { set_error("full_distortion_batch: plane %d: null buffer or stride below the width", k); return SVTB200_ERR_BAD_ARG; }
    int rc = check_jobs(jobs, n_jobs, w, h, SVTB200_FULLDIST_CBF_ZERO, false, "full_distortion_batch");
    if (rc) return rc;
    for (int j = 0; j < n_jobs; ++j)
        if ((jobs[j].x | jobs[j].width) & 1) { set_error("full_distortion_batch: job %d: odd origin or width", j); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    if (n_jobs == 0) return SVTB200_OK;
    c->reset();
    size_t need = (size_t)n_jobs * (sizeof(svtb200_dist_job) + 16) + 4096;
    for (int k = 0; k < 3; ++k) need += align_up((size_t)coeff->stride[k] * h[k] * 2, 256) + align_up((size_t)recon_coeff->stride[k] * h[k] * 2, 256);
    if ((rc = c->dev_reserve(need))) return rc;
    svtb200_coeff_picture dc = *coeff, dr = *recon_coeff;
    for (int k = 0; k < 3; ++k) {
        void *p;
        if ((rc = c->dev_alloc((size_t)coeff->stride[k] * h[k] * 2, &p))) return rc;
        SVTB200_CUDA_TRY(cudaMemcpyAsync(p, coeff->buf[k], (size_t)coeff->stride[k] * h[k] * 2, cudaMemcpyHostToDevice, c->stream));
        dc.buf[k] = (int16_t *)p;
        if ((rc = c->dev_alloc((size_t)recon_coeff->stride[k] * h[k] * 2, &p))) return rc;
        SVTB200_CUDA_TRY(cudaMemcpyAsync(p, recon_coeff->buf[k], (size_t)recon_coeff->stride[k] * h[k] * 2, cudaMemcpyHostToDevice, c->stream));
        dr.buf[k] = (int16_t *)p;
    }
    void *dj, *dout;
    if ((rc = c->dev_alloc((size_t)n_jobs * sizeof(svtb200_dist_job), &dj)) || (rc = c->dev_alloc((size_t)n_jobs * 16, &dout))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(dj, jobs, (size_t)n_jobs * sizeof(svtb200_dist_job), cudaMemcpyHostToDevice, c->stream));
    if ((rc = svtb200_full_distortion_batch_dev(&dc, &dr, (const svtb200_dist_job *)dj, n_jobs, (uint64_t *)dout, c->stream))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)n_jobs * 16, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

}  // extern "C"
