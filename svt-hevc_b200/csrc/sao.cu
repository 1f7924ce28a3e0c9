// sao.cu -- sample adaptive offset: per-LCU statistics and picture-level application, 8-bit 4:2:0.
//
// Replaces:
//   GatherSaoStatisticsLcuLossy_62x62 (C_DEFAULT/EbSampleAdaptiveOffset_C.c:125) and GatherSaoStatisticsLcu_BT_SSE2
//   SAOApplyBO / SAOApplyEO_{0,90,135,45} (C_DEFAULT/EbSampleAdaptiveOffset_C.c:394-861) and their SSSE3 rows,
//   ApplySaoOffsetsLcu / ApplySaoOffsetsPicture (Codec/EbEncDecProcess.c:215,522)
// The RD decision (SaoGenerationDecision, Codec/EbSampleAdaptiveOffsetGenerationDecision.c:647) stays on the host: it
// needs the CABAC rate tables and is a few hundred scalar operations per LCU.
//
// Design.  Statistics: one CTA per (LCU, component); the reconstructed block is staged once in shared memory, every
// thread keeps the 4 x 5 edge-class accumulators in registers (predicated adds, no atomics) and band sums are
// warp-aggregated with __match_any_sync before touching a shared 32-bin histogram.  Application: out = SAO(in), one
// thread per 4 samples, neighbours always read from the unmodified input picture, so the reference's line buffers
// (saoUpBuffer / saoLeftBuffer) have no counterpart.  Both kernels are HBM-bound (18 432 B per CTU).
#include "common.cuh"

namespace svtb200 {

struct PlaneRO { const uint8_t *p; int stride; };
struct StatsArgs {
    PlaneRO src[3], rec[3];      // pointers at the interior origin of each plane
    svtb200_sao_stats *out;
    int W, H, cols;
};

__device__ __forceinline__ int sgn(int a, int b) { return (a > b) - (a < b); }

__global__ void __launch_bounds__(256) sao_stats_kernel(const __grid_constant__ StatsArgs A)
{
    __shared__ __align__(16) uint8_t s_rec[64 * 68];
    __shared__ int s_bo_diff[32];
    __shared__ unsigned s_bo_cnt[32];
    __shared__ int s_eo[8][40];      // per-warp [type][idx] diff (20) + count (20)
    const int job = blockIdx.x, lcu = job / 3, c = job - lcu * 3, sh = c ? 1 : 0;
    const int lx = lcu % A.cols, ly = lcu / A.cols;
    const int ox = (lx * 64) >> sh, oy = (ly * 64) >> sh;
    const int lw = min(64, A.W - lx * 64) >> sh, lh = min(64, A.H - ly * 64) >> sh;
    const uint8_t *r = A.rec[c].p + (size_t)oy * A.rec[c].stride + ox, *s = A.src[c].p + (size_t)oy * A.src[c].stride + ox;
    for (int i = threadIdx.x; i < lh * (lw >> 2); i += 256) {
        const int y = i / (lw >> 2), x4 = i - y * (lw >> 2);
        ((uint32_t *)(s_rec + y * 68))[x4] = *(const uint32_t *)(r + (size_t)y * A.rec[c].stride + 4 * x4);   // block origins are 4-aligned
    }
    if (threadIdx.x < 32) { s_bo_diff[threadIdx.x] = 0; s_bo_cnt[threadIdx.x] = 0; }
    __syncthreads();
    int ed[4][5], ec[4][5];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int k = 0; k < 5; ++k) { ed[t][k] = 0; ec[t][k] = 0; }
    const int iw = lw - 2, n = iw * (lh - 2);
    const int iters = (n + 255) / 256;
    for (int it = 0; it < iters; ++it) {
        const int i = it * 256 + threadIdx.x;
        const bool act = i < n;
        int band = 32, d = 0;   // band 32 = inactive lane: never matches a real band
        if (act) {
            const int y = 1 + i / iw, x = 1 + i % iw;
            const uint8_t *q = s_rec + y * 68 + x;
            const int v = q[0];
            d = (int)s[(size_t)y * A.src[c].stride + x] - v;
            d = min(max(d, -128), 127);                         // clipToSigned8Bit
            band = v >> 3;
            const int e[4] = {sgn(v, q[-1]) + sgn(v, q[1]) + 2, sgn(v, q[-68]) + sgn(v, q[68]) + 2, sgn(v, q[-69]) + sgn(v, q[69]) + 2,
                              sgn(v, q[-67]) + sgn(v, q[67]) + 2};
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int k = 0; k < 5; ++k) { const bool hit = e[t] == k; ed[t][k] += hit ? d : 0; ec[t][k] += hit; }
        }
        // warp-aggregated band histogram: one shared-memory atomic per distinct band per warp
        const unsigned peers = __match_any_sync(0xffffffffu, band);
        const int sum = __reduce_add_sync(peers, d);
        if (act && (threadIdx.x & 31) == __ffs(peers) - 1) { atomicAdd(&s_bo_diff[band], sum); atomicAdd(&s_bo_cnt[band], (unsigned)__popc(peers)); }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int dsum = __reduce_add_sync(0xffffffffu, ed[t][k]), csum = __reduce_add_sync(0xffffffffu, ec[t][k]);
            if (lane == 0) { s_eo[wid][t * 5 + k] = dsum; s_eo[wid][20 + t * 5 + k] = csum; }
        }
    __syncthreads();
    svtb200_sao_stats *o = A.out + job;
    if (threadIdx.x < 32) { o->bo_diff[threadIdx.x] = s_bo_diff[threadIdx.x]; o->bo_count[threadIdx.x] = (uint16_t)s_bo_cnt[threadIdx.x]; }
    if (threadIdx.x < 20) {
        // categories compacted like the reference: slot 2 <- raw 3, slot 3 <- raw 4, slot 4 scratch (we write 0)
        const int t = threadIdx.x / 5, k = threadIdx.x % 5;
        const int rawk = k < 2 ? k : k + 1;
        int dsum = 0, csum = 0;
        if (k < 4)
            for (int w = 0; w < 8; ++w) { dsum += s_eo[w][t * 5 + rawk]; csum += s_eo[w][20 + t * 5 + rawk]; }
        o->eo_diff[t][k] = dsum;
        o->eo_count[t][k] = (uint16_t)csum;
    }
}

struct ApplyArgs {
    PlaneRO in[3];
    uint8_t *out[3];
    int ostride[3];
    const svtb200_sao_params *prm;
    int W, H, cols;
    int flag[2];
};

// one thread per 4 horizontally adjacent samples of one plane (blockIdx.z = plane)
__global__ void __launch_bounds__(256) sao_apply_kernel(const __grid_constant__ ApplyArgs A)
{
    const int c = blockIdx.z, sh = c ? 1 : 0, pw = A.W >> sh, ph = A.H >> sh;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= pw || y >= ph) return;
    const svtb200_sao_params p = A.prm[(size_t)((y << sh) >> 6) * A.cols + ((x0 << sh) >> 6)];   // 4 samples never straddle an LCU
    const int t = A.flag[c ? 1 : 0] ? p.type[c ? 1 : 0] : 0;
    const int is = A.in[c].stride;
    const uint8_t *row = A.in[c].p + (size_t)y * is + x0;
    const uint32_t cur = *(const uint32_t *)row;
    uint32_t res = cur;
    if (t == 5) {
        const int pos = p.band_position[c];
        res = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int v = (cur >> (8 * k)) & 255;
            const int b = (v >> 3) - pos;
            if (b >= 0 && b <= 3) v = min(max(v + p.offset[c][b], 0), 255);
            res |= (uint32_t)v << (8 * k);
        }
    } else if (t >= 1 && t <= 4) {
        const int dx = (t == 2) ? 0 : (t == 4 ? 1 : -1);      // first neighbour: EO-0 (-1,0) EO-90 (0,-1) EO-135 (-1,-1) EO-45 (+1,-1)
        const int dy = (t == 1) ? 0 : -1;
        const bool skip_row = (t != 1) && (y == 0 || y == ph - 1);
        if (!skip_row) {
            const uint8_t *ra = row + (ptrdiff_t)dy * is, *rb = row - (ptrdiff_t)dy * is;
            // bytes x0-1 .. x0+4 of the two neighbour rows
            const uint32_t a0 = *(const uint32_t *)(ra - 4), a1 = *(const uint32_t *)ra, a2 = *(const uint32_t *)(ra + 4);
            const uint32_t b0 = *(const uint32_t *)(rb - 4), b1 = *(const uint32_t *)rb, b2 = *(const uint32_t *)(rb + 4);
            const uint32_t na = dx == 0 ? a1 : (dx < 0 ? __funnelshift_r(a0, a1, 24) : __funnelshift_r(a1, a2, 8));
            const uint32_t nb = dx == 0 ? b1 : (dx < 0 ? __funnelshift_r(b1, b2, 8) : __funnelshift_r(b0, b1, 24));
            res = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int v = (cur >> (8 * k)) & 255;
                const int x = x0 + k;
                if (!((t != 2) && (x == 0 || x == pw - 1))) {
                    const int idx = sgn(v, (int)((na >> (8 * k)) & 255)) + sgn(v, (int)((nb >> (8 * k)) & 255)) + 2;
                    const int off = idx == 0 ? p.offset[c][0] : idx == 1 ? p.offset[c][1] : idx == 3 ? p.offset[c][2] : idx == 4 ? p.offset[c][3] : 0;
                    v = min(max(v + off, 0), 255);
                }
                res |= (uint32_t)v << (8 * k);
            }
        }
    }
    *(uint32_t *)(A.out[c] + (size_t)y * A.ostride[c] + x0) = res;
}

static const svtb200_plane *plane_of(const svtb200_yuv_picture *p, int k) { return k == 0 ? &p->y : k == 1 ? &p->cb : &p->cr; }
static size_t pbytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }

static int sao_check(const svtb200_yuv_picture *a, const svtb200_yuv_picture *b, const char *what)
{
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane *pa = plane_of(a, k), *pb = plane_of(b, k);
        if (!pa->buf || !pb->buf || pa->width != pb->width || pa->height != pb->height || (pa->stride & 3) || (pb->stride & 3) || (pa->origin_x & 3) || (pb->origin_x & 3) ||
            pa->origin_x < 4 || pb->origin_x < 4 || pa->origin_y < 1 || pb->origin_y < 1) {
            set_error("%s: planes must share their size, have pitch and x-origin multiples of 4 and >= 4 samples of padding", what);
            return SVTB200_ERR_BAD_ARG;
        }
    }
    if ((a->y.width & 7) || (a->y.height & 7) || a->y.width < 16 || a->y.height < 16) { set_error("%s: picture size must be a multiple of 8", what); return SVTB200_ERR_BAD_ARG; }
    return SVTB200_OK;
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

int svtb200_sao_stats_picture_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out_dev, void *cuda_stream)
{
    if (!src || !recon || !out_dev) { set_error("sao_stats_picture_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = sao_check(src, recon, "sao_stats_picture");
    if (rc || (rc = ensure_device())) return rc;
    StatsArgs A;
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane *s = plane_of(src, k), *r = plane_of(recon, k);
        A.src[k] = {s->buf + (size_t)s->origin_y * s->stride + s->origin_x, (int)s->stride};
        A.rec[k] = {r->buf + (size_t)r->origin_y * r->stride + r->origin_x, (int)r->stride};
        if (((uintptr_t)A.rec[k].p) & 3) { set_error("sao_stats_picture: reconstruction interior must be 4-byte aligned"); return SVTB200_ERR_BAD_ARG; }
    }
    A.out = out_dev; A.W = recon->y.width; A.H = recon->y.height; A.cols = (A.W + 63) / 64;
    const int n = A.cols * ((A.H + 63) / 64) * 3;
    sao_stats_kernel<<<n, 256, 0, (cudaStream_t)cuda_stream>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_sao_apply_picture_dev(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params_dev,
                                  const uint8_t sao_flag[2], void *cuda_stream)
{
    if (!in || !out || !lcu_params_dev || !sao_flag) { set_error("sao_apply_picture_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = sao_check(in, out, "sao_apply_picture");
    if (rc || (rc = ensure_device())) return rc;
    if (in->y.buf == out->y.buf || in->cb.buf == out->cb.buf || in->cr.buf == out->cr.buf) { set_error("sao_apply_picture: in and out must be different pictures"); return SVTB200_ERR_BAD_ARG; }
    ApplyArgs A;
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane *i = plane_of(in, k), *o = plane_of(out, k);
        A.in[k] = {i->buf + (size_t)i->origin_y * i->stride + i->origin_x, (int)i->stride};
        A.out[k] = o->buf + (size_t)o->origin_y * o->stride + o->origin_x;
        A.ostride[k] = (int)o->stride;
        if ((((uintptr_t)A.in[k].p) & 3) || (((uintptr_t)A.out[k]) & 3)) { set_error("sao_apply_picture: plane interiors must be 4-byte aligned"); return SVTB200_ERR_BAD_ARG; }
    }
    A.prm = lcu_params_dev; A.W = in->y.width; A.H = in->y.height; A.cols = (A.W + 63) / 64; A.flag[0] = sao_flag[0]; A.flag[1] = sao_flag[1];
    dim3 grid((A.W / 4 + 255) / 256, A.H, 3);
    sao_apply_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static int upload_pic(ThreadCtx *c, const svtb200_yuv_picture *h, svtb200_yuv_picture *d, bool copy)
{
    *d = *h;
    svtb200_plane *dp[3] = {&d->y, &d->cb, &d->cr};
    for (int k = 0; k < 3; ++k) {
        void *p;
        int rc = c->dev_alloc(pbytes(*dp[k]), &p);
        if (rc) return rc;
        if (copy) SVTB200_CUDA_TRY(cudaMemcpyAsync(p, plane_of(h, k)->buf, pbytes(*dp[k]), cudaMemcpyHostToDevice, c->stream));
        dp[k]->buf = (uint8_t *)p;
    }
    return SVTB200_OK;
}
static size_t pic_bytes(const svtb200_yuv_picture *p) { return align_up(pbytes(p->y), 256) + align_up(pbytes(p->cb), 256) + align_up(pbytes(p->cr), 256); }

int svtb200_sao_stats_picture(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out)
{
    if (!src || !recon || !out) { set_error("sao_stats_picture: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = sao_check(src, recon, "sao_stats_picture");
    if (rc) return rc;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t n = (size_t)((recon->y.width + 63) / 64) * ((recon->y.height + 63) / 64) * 3, ob = n * sizeof(svtb200_sao_stats);
    if ((rc = c->dev_reserve(pic_bytes(src) + pic_bytes(recon) + ob + 8192))) return rc;
    svtb200_yuv_picture ds, dr;
    void *d_out;
    if ((rc = upload_pic(c, src, &ds, true)) || (rc = upload_pic(c, recon, &dr, true)) || (rc = c->dev_alloc(ob, &d_out))) return rc;
    if ((rc = svtb200_sao_stats_picture_dev(&ds, &dr, (svtb200_sao_stats *)d_out, c->stream))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_sao_apply_picture(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params, const uint8_t sao_flag[2])
{
    if (!in || !out || !lcu_params || !sao_flag) { set_error("sao_apply_picture: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = sao_check(in, out, "sao_apply_picture");
    if (rc) return rc;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t n = (size_t)((in->y.width + 63) / 64) * ((in->y.height + 63) / 64), pb = n * sizeof(svtb200_sao_params);
    if ((rc = c->dev_reserve(pic_bytes(in) + pic_bytes(out) + pb + 8192))) return rc;
    svtb200_yuv_picture di, dout;
    void *d_prm;
    // the output keeps whatever the caller had in its padding: bring it over, the kernel rewrites the interior only
    if ((rc = upload_pic(c, in, &di, true)) || (rc = upload_pic(c, out, &dout, true)) || (rc = c->dev_alloc(pb, &d_prm))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_prm, lcu_params, pb, cudaMemcpyHostToDevice, c->stream));
    if ((rc = svtb200_sao_apply_picture_dev(&di, &dout, (const svtb200_sao_params *)d_prm, sao_flag, c->stream))) return rc;
    const svtb200_plane *hp[3] = {&out->y, &out->cb, &out->cr}, *dp[3] = {&dout.y, &dout.cb, &dout.cr};
    for (int k = 0; k < 3; ++k) SVTB200_CUDA_TRY(cudaMemcpyAsync(hp[k]->buf, dp[k]->buf, pbytes(*hp[k]), cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

}  // extern "C"
