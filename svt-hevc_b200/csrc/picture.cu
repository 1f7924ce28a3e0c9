// picture.cu -- device-resident source pictures: edge padding + point decimation on the GPU.
// Replaces the host-side feeders of ME: GeneratePadding (Codec/EbMcp.c:1017),
// Decimation2D (Codec/EbPictureAnalysisProcess.c:173), DecimateInputPicture (:4139).
#include "common.cuh"

namespace svtb200 {

// out(x,y) over the PADDED extent of a plane whose interior is (w x h) at pad; value = full-res interior sample
// at (clamp(x - pad) * step, clamp(y - pad) * step): decimation followed by edge replication in one pass.
__global__ void __launch_bounds__(256) decim_pad_kernel(const uint8_t *full, int fstride, int fpad, uint8_t *out, int ostride,
                                                        int opad, int w, int h, int step, int skip_interior)
{
    const int tw = w + 2 * opad, th = h + 2 * opad;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= tw || y >= th) return;
    const int sy = min(max(y - opad, 0), h - 1) * step;
    const uint8_t *srow = full + (size_t)(fpad + sy) * fstride + fpad;
    const bool row_in = (y >= opad && y < opad + h);
    uint8_t *o = out + (size_t)y * ostride + x4;
    if (skip_interior && row_in && x4 >= opad && x4 + 3 < opad + w) return;   // in-place full plane: interior untouched
    if (step == 1 && x4 >= opad && x4 + 3 < opad + w && ((((uintptr_t)(srow + x4 - opad)) | (uintptr_t)o) & 3) == 0) {
        *(uint32_t *)o = *(const uint32_t *)(srow + x4 - opad);              // four interior samples of a full-resolution plane: one word
        return;
    }
    uint8_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        const int sx = min(max(x - opad, 0), w - 1) * step;
        v[k] = srow[sx];
    }
    if (x4 + 3 < tw && (((uintptr_t)o) & 3) == 0 && !(skip_interior && row_in)) {
        *(uint32_t *)o = v[0] | (v[1] << 8) | (v[2] << 16) | ((uint32_t)v[3] << 24);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = x4 + k;
            if (x >= tw) break;
            if (skip_interior && row_in && x >= opad && x < opad + w) continue;
            o[k] = v[k];
        }
    }
}

// src: full-resolution luma the planes are derived from (interior origin at src + spad * (sstride + 1)).  in_place: src is d->full itself,
// whose interior is already there (only its border is written).
static int build_planes(const svtb200_pa_picture *d, cudaStream_t st, const uint8_t *src, int sstride, int spad, bool in_place)
{
    const svtb200_plane *pl[3] = {&d->full, &d->quarter, &d->sixteenth};
    const int steps[3] = {1, 2, 4};
    // decimated planes first (they read the full-plane interior only), then the full plane's own border
    for (int k = 2; k >= 0; --k) {
        const svtb200_plane &p = *pl[k];
        if (!p.buf) continue;
        const int tw = p.width + 2 * p.origin_x, th = p.height + 2 * p.origin_y;
        const dim3 grid(((tw + 3) / 4 + 255) / 256, th);      // one thread per 4 output bytes of a padded row
        decim_pad_kernel<<<grid, 256, 0, st>>>(src, sstride, spad, p.buf, (int)p.stride, p.origin_x, p.width, p.height, steps[k], k == 0 && in_place);
        SVTB200_COUNT_LAUNCH();
    }
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

int svtb200_pa_picture_create(uint16_t width, uint16_t height, svtb200_pa_picture *out)
{
    if (!out || width < 8 || height < 8) { set_error("pa_picture_create: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    svtb200_plane *pl[3] = {&out->full, &out->quarter, &out->sixteenth};
    const int pads[3] = {68, 32, 16}, sh[3] = {0, 1, 2};
    for (int k = 0; k < 3; ++k) {
        svtb200_plane &p = *pl[k];
        p.width = (uint16_t)((width + (1 << sh[k]) - 1) >> sh[k]);
        p.height = (uint16_t)((height + (1 << sh[k]) - 1) >> sh[k]);
        p.origin_x = p.origin_y = (uint16_t)pads[k];
        p.stride = (uint32_t)align_up((size_t)p.width + 2 * pads[k], 16);   // 16-byte pitch: windows are staged with 128-bit loads
        void *d = nullptr;
        SVTB200_CUDA_TRY(cudaMalloc(&d, (size_t)p.stride * (p.height + 2 * pads[k]) + 64));
        p.buf = (uint8_t *)d;
    }
    return SVTB200_OK;
}

int svtb200_pa_picture_build_dev(const svtb200_pa_picture *dev, const uint8_t *dev_luma, uint32_t dev_stride, void *cuda_stream)
{
    if (!dev || !dev->full.buf || !dev_luma) { set_error("pa_picture_build_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const svtb200_plane &f = dev->full;
    (void)f;
    return build_planes(dev, st, dev_luma, (int)dev_stride, 0, false);      // all three planes straight from the caller's luma: no intermediate copy
}

int svtb200_pa_picture_upload(const svtb200_pa_picture *dev, const uint8_t *host_luma, uint32_t host_stride, void *cuda_stream)
{
    if (!dev || !dev->full.buf || !host_luma) { set_error("pa_picture_upload: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const svtb200_plane &f = dev->full;
    SVTB200_CUDA_TRY(cudaMemcpy2DAsync(f.buf + (size_t)f.origin_y * f.stride + f.origin_x, f.stride, host_luma, host_stride, f.width, f.height,
                                       cudaMemcpyHostToDevice, st));
    return build_planes(dev, st, f.buf, (int)f.stride, f.origin_x, true);
}

// Device picture -> device picture (same geometry), whole padded planes: the reconstruction a later stage needs unmodified (H3: the SAO statistics of a
// deblocked picture read the strips that were still unfiltered at gather time from the picture before svtb200_dlf_picture_dev).
int svtb200_yuv_picture_copy_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *dst, uint32_t bytes_per_sample, void *cuda_stream)
{
    if (!src || !dst || (bytes_per_sample != 1 && bytes_per_sample != 2)) { set_error("yuv_picture_copy_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    const svtb200_plane *sp[3] = {&src->y, &src->cb, &src->cr}, *dp[3] = {&dst->y, &dst->cb, &dst->cr};
    for (int k = 0; k < 3; ++k) {
        if (!sp[k]->buf || !dp[k]->buf || sp[k]->width != dp[k]->width || sp[k]->height != dp[k]->height || sp[k]->origin_x != dp[k]->origin_x || sp[k]->origin_y != dp[k]->origin_y) {
            set_error("yuv_picture_copy_dev: plane %d: null buffer or different geometry", k); return SVTB200_ERR_BAD_ARG;
        }
        const size_t row = (size_t)(sp[k]->width + 2 * sp[k]->origin_x) * bytes_per_sample, rows = (size_t)sp[k]->height + 2 * sp[k]->origin_y;
        SVTB200_CUDA_TRY(cudaMemcpy2DAsync(dp[k]->buf, (size_t)dp[k]->stride * bytes_per_sample, sp[k]->buf, (size_t)sp[k]->stride * bytes_per_sample, row, rows,
                                           cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
    }
    return SVTB200_OK;
}

int svtb200_yuv_picture_upload(const svtb200_yuv_picture *dev, const void *host_y, uint32_t host_y_stride, const void *host_cb, uint32_t host_cb_stride,
                               const void *host_cr, uint32_t host_cr_stride, uint32_t bytes_per_sample, void *cuda_stream)
{
    if (!dev || !host_y || !host_cb || !host_cr || (bytes_per_sample != 1 && bytes_per_sample != 2)) { set_error("yuv_picture_upload: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const svtb200_plane *pl[3] = {&dev->y, &dev->cb, &dev->cr};
    const void *h[3] = {host_y, host_cb, host_cr};
    const uint32_t hs[3] = {host_y_stride, host_cb_stride, host_cr_stride};
    const size_t b = bytes_per_sample;
    for (int k = 0; k < 3; ++k) {
        if (!pl[k]->buf) { set_error("yuv_picture_upload: null device plane"); return SVTB200_ERR_BAD_ARG; }
        SVTB200_CUDA_TRY(cudaMemcpy2DAsync(pl[k]->buf + ((size_t)pl[k]->origin_y * pl[k]->stride + pl[k]->origin_x) * b, pl[k]->stride * b, h[k], hs[k] * b,
                                           pl[k]->width * b, pl[k]->height, cudaMemcpyHostToDevice, st));
    }
    return SVTB200_OK;
}

// Edge replication of a device-resident Y/Cb/Cr picture in place (GeneratePadding, Codec/EbMcp.c:1017, as the reference runs it on a reconstructed
// picture before it serves as a reference: PadRefAndSetFlags, Codec/EbEncDecProcess.c): the interior stays, every padding sample takes the nearest
// interior sample.  8-bit planes.
int svtb200_yuv_picture_pad_dev(const svtb200_yuv_picture *dev, void *cuda_stream)
{
    if (!dev) { set_error("yuv_picture_pad_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const svtb200_plane *pl[3] = {&dev->y, &dev->cb, &dev->cr};
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane &p = *pl[k];
        if (!p.buf || p.origin_x != p.origin_y) { set_error("yuv_picture_pad_dev: null plane or unequal horizontal / vertical padding"); return SVTB200_ERR_BAD_ARG; }
        const int tw = p.width + 2 * p.origin_x, th = p.height + 2 * p.origin_y;
        if ((uint32_t)tw > p.stride) { set_error("yuv_picture_pad_dev: pitch smaller than the padded width"); return SVTB200_ERR_BAD_ARG; }
        const dim3 grid(((tw + 3) / 4 + 255) / 256, th);
        decim_pad_kernel<<<grid, 256, 0, st>>>(p.buf, (int)p.stride, p.origin_x, p.buf, (int)p.stride, p.origin_x, p.width, p.height, 1, 1);
        SVTB200_COUNT_LAUNCH();
    }
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_copy_async(void *dst, const void *src, size_t bytes, int device_to_host, void *cuda_stream)
{
    if ((!dst || !src) && bytes) { set_error("copy_async: null pointer"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    if (bytes) SVTB200_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, device_to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream));
    return SVTB200_OK;
}

// ---- PictureAnalysis feeder: the per-LCU mean / variance pyramid -------------------------------------------------------------
// ComputeBlockMeanComputeVariance (Codec/EbPictureAnalysisProcess.c:1646-2400) with the 8x8 figures of ComputeSubMean8x8_SSE2_INTRIN /
// ComputeSubdMeanOfSquaredValues8x8_SSE2_INTRIN (ASM_SSE2/EbComputeMean_Intrinsic_SSE2.c:53,10: rows 0, 2, 4, 6 of the block; sum << 3, sum of squares
// << 11) -- the AVX2 row (ComputeIntermVarFour8x8_AVX2_INTRIN) is arithmetically the same.  One CTA of 64 threads per LCU, a thread per 8x8 block
// (VABSDIFF4 against 0 for the sums, DP4A(w, w) for the squares), the 16x16 / 32x32 / 64x64 levels ((sum of four children) >> 2) through shared memory.
// Outputs in the order of ME_TIER_ZERO_PU_*: 64x64, 4 x 32x32, 16 x 16x16, 64 x 8x8 (raster each): yMean = mean >> 8 (EB_U8), variance =
// (mean of squares - mean^2) >> 16 (EB_U16).  A partial last LCU reads the picture's padding like the reference.
__global__ void __launch_bounds__(64) variance_kernel(const uint8_t *luma00, int stride, int cols, uint8_t *y_mean, uint16_t *variance)
{
    __shared__ unsigned long long s_m[64 + 16 + 4 + 1], s_s[64 + 16 + 4 + 1];
    const int lcu = blockIdx.x, b = threadIdx.x;
    const uint8_t *q = luma00 + (size_t)((lcu / cols) * 64 + (b >> 3) * 8) * stride + (lcu % cols) * 64 + (b & 7) * 8;
    uint32_t sum = 0, sq = 0;
#pragma unroll
    for (int r = 0; r < 8; r += 2) {
        const uint32_t w0 = ld_u32_unaligned(q + (size_t)r * stride), w1 = ld_u32_unaligned(q + (size_t)r * stride + 4);
        sum = vsad4_acc(w0, 0u, sum); sum = vsad4_acc(w1, 0u, sum);
        asm("dp4a.u32.u32 %0, %1, %1, %0;" : "+r"(sq) : "r"(w0));
        asm("dp4a.u32.u32 %0, %1, %1, %0;" : "+r"(sq) : "r"(w1));
    }
    s_m[21 + b] = (unsigned long long)sum << 3;
    s_s[21 + b] = (unsigned long long)sq << 11;
    __syncthreads();
    if (b < 16) {
        const int i = 21 + (b >> 2) * 16 + (b & 3) * 2;
        s_m[5 + b] = (s_m[i] + s_m[i + 1] + s_m[i + 8] + s_m[i + 9]) >> 2;
        s_s[5 + b] = (s_s[i] + s_s[i + 1] + s_s[i + 8] + s_s[i + 9]) >> 2;
    }
    __syncthreads();
    if (b < 4) {
        const int i = 5 + (b >> 1) * 8 + (b & 1) * 2;
        s_m[1 + b] = (s_m[i] + s_m[i + 1] + s_m[i + 4] + s_m[i + 5]) >> 2;
        s_s[1 + b] = (s_s[i] + s_s[i + 1] + s_s[i + 4] + s_s[i + 5]) >> 2;
    }
    __syncthreads();
    if (b == 0) { s_m[0] = (s_m[1] + s_m[2] + s_m[3] + s_m[4]) >> 2; s_s[0] = (s_s[1] + s_s[2] + s_s[3] + s_s[4]) >> 2; }
    __syncthreads();
    for (int k = b; k < 85; k += 64) {
        const unsigned long long m = s_m[k];
        y_mean[(size_t)lcu * 85 + k] = (uint8_t)(m >> 8);
        variance[(size_t)lcu * 85 + k] = (uint16_t)((s_s[k] - m * m) >> 16);
    }
}

static int variance_check(const svtb200_plane *p, const void *a, const void *b, const char *who)
{
    if (!p || !p->buf || !a || !b || !p->width || !p->height) { set_error("%s: bad argument", who); return SVTB200_ERR_BAD_ARG; }
    const uint32_t cols = (p->width + 63) / 64, rows = (p->height + 63) / 64;
    // the last LCU row / column is read in full: the padding has to hold it (it does for the encoder's 68-sample padding and any picture size it accepts)
    if (p->stride < p->origin_x + cols * 64 || (uint32_t)p->height + 2 * p->origin_y < p->origin_y + rows * 64) {
        set_error("%s: the plane's padding does not cover the last (partial) LCU row / column", who); return SVTB200_ERR_BAD_ARG;
    }
    return SVTB200_OK;
}

int svtb200_variance_picture_dev(const svtb200_plane *luma_dev, uint8_t *y_mean_dev, uint16_t *variance_dev, void *cuda_stream)
{
    int rc = variance_check(luma_dev, y_mean_dev, variance_dev, "variance_picture_dev");
    if (rc || (rc = ensure_device())) return rc;
    const int cols = (luma_dev->width + 63) / 64, rows = (luma_dev->height + 63) / 64;
    variance_kernel<<<cols * rows, 64, 0, (cudaStream_t)cuda_stream>>>(luma_dev->buf + (size_t)luma_dev->origin_y * luma_dev->stride + luma_dev->origin_x, (int)luma_dev->stride, cols,
                                                                       y_mean_dev, variance_dev);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_variance_picture(const svtb200_plane *luma, uint8_t *y_mean, uint16_t *variance)
{
    int rc = variance_check(luma, y_mean, variance, "variance_picture");
    if (rc) return rc;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t pb = (size_t)luma->stride * (luma->height + 2 * luma->origin_y), n = (size_t)((luma->width + 63) / 64) * ((luma->height + 63) / 64);
    void *d_p, *d_m, *d_v;
    if ((rc = c->dev_reserve(pb + n * 85 * 3 + 4096)) || (rc = c->dev_alloc(pb, &d_p)) || (rc = c->dev_alloc(n * 85, &d_m)) || (rc = c->dev_alloc(n * 85 * 2, &d_v))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_p, luma->buf, pb, cudaMemcpyHostToDevice, c->stream));
    svtb200_plane dp = *luma;
    dp.buf = (uint8_t *)d_p;
    if ((rc = svtb200_variance_picture_dev(&dp, (uint8_t *)d_m, (uint16_t *)d_v, c->stream))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(y_mean, d_m, n * 85, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(variance, d_v, n * 85 * 2, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

// ---- peer memory for the GOP-boundary hand-off: one process per GPU, the entropy-coding rank's receive buffers mapped into every
// producer's address space (CUDA IPC) so that a finished picture's outputs move GPU to GPU over NVLink by the copy engines -- no SM, no staging, no
// collective in the data path.  Exportable memory comes from cudaMalloc directly (pool allocators hand out sub-ranges an IPC handle cannot name).
int svtb200_peer_alloc(size_t bytes, void **dev_ptr)
{
    if (!dev_ptr || !bytes) { set_error("peer_alloc: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    SVTB200_CUDA_TRY(cudaMalloc(dev_ptr, bytes));
    SVTB200_CUDA_TRY(cudaMemset(*dev_ptr, 0, bytes));
    return SVTB200_OK;
}
int svtb200_peer_free(void *dev_ptr)
{
    if (dev_ptr) SVTB200_CUDA_TRY(cudaFree(dev_ptr));
    return SVTB200_OK;
}
int svtb200_peer_export(const void *dev_ptr, uint8_t handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!dev_ptr || !handle) { set_error("peer_export: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    SVTB200_CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle, &h, 64);
    return SVTB200_OK;
}
int svtb200_peer_open(const uint8_t handle[64], void **dev_ptr)
{
    if (!dev_ptr || !handle) { set_error("peer_open: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SVTB200_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SVTB200_OK;
}
int svtb200_peer_close(void *dev_ptr)
{
    if (dev_ptr) SVTB200_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return SVTB200_OK;
}
int svtb200_peer_copy_async(void *dst_dev, const void *src_dev, size_t bytes, void *cuda_stream)
{
    if ((!dst_dev || !src_dev) && bytes) { set_error("peer_copy_async: null pointer"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    if (bytes) SVTB200_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));      // unified addressing: the peer mapping routes it over NVLink
    return SVTB200_OK;
}

int svtb200_pa_picture_download(const svtb200_pa_picture *dev, const svtb200_pa_picture *host)
{
    if (!dev || !host) { set_error("pa_picture_download: bad argument"); return SVTB200_ERR_BAD_ARG; }
    const svtb200_plane *d[3] = {&dev->full, &dev->quarter, &dev->sixteenth}, *h[3] = {&host->full, &host->quarter, &host->sixteenth};
    SVTB200_CUDA_TRY(cudaDeviceSynchronize());
    for (int k = 0; k < 3; ++k) {
        if (!h[k]->buf || !d[k]->buf) continue;
        const int tw = d[k]->width + 2 * d[k]->origin_x, th = d[k]->height + 2 * d[k]->origin_y;
        if (h[k]->width != d[k]->width || h[k]->height != d[k]->height || h[k]->origin_x != d[k]->origin_x) { set_error("pa_picture_download: geometry mismatch"); return SVTB200_ERR_BAD_ARG; }
        SVTB200_CUDA_TRY(cudaMemcpy2D(h[k]->buf, h[k]->stride, d[k]->buf, d[k]->stride, tw, th, cudaMemcpyDeviceToHost));
    }
    return SVTB200_OK;
}

void svtb200_pa_picture_destroy(svtb200_pa_picture *dev)
{
    if (!dev) return;
    svtb200_plane *pl[3] = {&dev->full, &dev->quarter, &dev->sixteenth};
    for (int k = 0; k < 3; ++k) if (pl[k]->buf) { cudaFree(pl[k]->buf); pl[k]->buf = nullptr; }
}

}  // extern "C"
