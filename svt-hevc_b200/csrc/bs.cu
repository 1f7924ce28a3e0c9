// bs.cu -- boundary-strength derivation of the deblocking filter on the device.
//
// Replaces CalculateBSForPUBoundary (Codec/EbDeblockingFilter.c:109), SetBSArrayBasedOnPUBoundary (:339), CalculateBSForTUBoundaryInsidePU /
// SetBSArrayBasedOnTUBoundary (:440,472) and the Intra4x4 variants (:527 ff), which the encode pass calls CU by CU (Codec/EbCodingLoop.c:3554,
// 4329,4364) against its neighbour arrays.  By the time a CU's strengths are set, its own cbf entries and everything left of / above it are final
// (EbCodingLoop.c:3547-3552,4320-4326), so the strengths are a pure function of the picture's FINAL mode / motion / cbf maps -- which is what the
// kernel reads: one thread per 4x4 block writes the strength of the edge on its left and on its top, in the [n_lcu][256] layout svtb200_dlf_picture
// consumes (verticalEdgeBSArray / horizontalEdgeBSArray).  Edges exist on the 8-sample grid only: a CU boundary takes the PU rule (intra on either
// side 2; else motion / reference / cbf tests by slice type), the 32-sample split inside an inter 64x64 CU the TU rule (cbf of either side; only the
// inter path of 64x64 CUs sets TU-boundary strengths, EbCodingLoop.c:4328); picture edges stay 0 (tileLeftBoundary / tileTopBoundary, one tile).
#include "common.cuh"

namespace svtb200 {

struct BsArgs {
    const svtb200_cu_info *cu;      // one entry per 8x8 luma block
    int cu_stride;
    const uint8_t *cbf;             // one entry per 4x4 luma block (cbfMapArray)
    int cbf_stride;
    int W, H, cols;
    int slice_type;                 // 0 = B, 1 = P, 2 = I
    int same_poc;                   // the two lists point at the same picture
    uint8_t *vbs, *hbs;
};

__device__ __forceinline__ bool mv_far(const svtb200_cu_info &a, int la, const svtb200_cu_info &b, int lb)
{
    return abs((int)a.mv[la][0] - (int)b.mv[lb][0]) >= 4 || abs((int)a.mv[la][1] - (int)b.mv[lb][1]) >= 4;
}
// the PU-boundary rule for one 4x4 edge segment: a = the block's PU, b = the neighbour's
__device__ __forceinline__ int pu_rule(const BsArgs &A, const svtb200_cu_info &a, const svtb200_cu_info &b, bool cbf_any)
{
    if (a.intra || b.intra) return 2;
    bool c1;
    if (A.slice_type == 1) c1 = mv_far(a, 0, b, 0);             // one reference per list: the POC test of a list against itself never fires
    else {
        switch (a.pred_dir + 3 * b.pred_dir) {
        case 0: c1 = mv_far(a, 0, b, 0); break;
        case 1: c1 = !A.same_poc || mv_far(a, 1, b, 0); break;
        case 3: c1 = !A.same_poc || mv_far(a, 0, b, 1); break;
        case 4: c1 = mv_far(a, 1, b, 1); break;
        case 8:
            c1 = A.same_poc ? (mv_far(a, 0, b, 0) || mv_far(a, 1, b, 1)) && (mv_far(a, 0, b, 1) || mv_far(a, 1, b, 0)) : (mv_far(a, 0, b, 0) || mv_far(a, 1, b, 1));
            break;
        default: c1 = true; break;      // one side bi-predicted, the other not
        }
    }
    return (c1 || cbf_any) ? 1 : 0;
}

__global__ void __launch_bounds__(256) dlf_bs_kernel(const __grid_constant__ BsArgs A)
{
    const int lcu = blockIdx.x, lx = lcu % A.cols, ly = lcu / A.cols;
    const int bx = threadIdx.x & 15, by = threadIdx.x >> 4;
    const int x = lx * 64 + 4 * bx, y = ly * 64 + 4 * by;
    int v = 0, h = 0;
    if (x < A.W && y < A.H) {
        const svtb200_cu_info me = A.cu[(y >> 3) * A.cu_stride + (x >> 3)];
        const int cu = 1 << me.log2_cu_size;
        const bool my_cbf = A.cbf[(y >> 2) * A.cbf_stride + (x >> 2)] != 0;
        if ((x & 7) == 0 && x != 0) {
            const bool cbf_any = my_cbf || A.cbf[(y >> 2) * A.cbf_stride + (x >> 2) - 1] != 0;
            if ((x & (cu - 1)) == 0) v = pu_rule(A, me, A.cu[(y >> 3) * A.cu_stride + (x >> 3) - 1], cbf_any);
            else if (cu == 64 && !me.intra && (x & 31) == 0) v = cbf_any ? 1 : 0;
        }
        if ((y & 7) == 0 && y != 0) {
            const bool cbf_any = my_cbf || A.cbf[((y >> 2) - 1) * A.cbf_stride + (x >> 2)] != 0;
            if ((y & (cu - 1)) == 0) h = pu_rule(A, me, A.cu[((y >> 3) - 1) * A.cu_stride + (x >> 3)], cbf_any);
            else if (cu == 64 && !me.intra && (y & 31) == 0) h = cbf_any ? 1 : 0;
        }
    }
    A.vbs[(size_t)lcu * 256 + threadIdx.x] = (uint8_t)v;
    A.hbs[(size_t)lcu * 256 + threadIdx.x] = (uint8_t)h;
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

static int bs_check(const void *cu, uint32_t cu_stride, const void *cbf, uint32_t cbf_stride, uint16_t W, uint16_t H, uint8_t slice_type, const void *v, const void *h, const char *who)
{
    if (!cu || !cbf || !v || !h || !W || !H || (W & 7) || (H & 7) || cu_stride < (uint32_t)(W >> 3) || cbf_stride < (uint32_t)(W >> 2) || slice_type > 2) {
        set_error("%s: bad argument (null pointer, picture size not a multiple of 8, strides below the picture width, or slice type > 2)", who);
        return SVTB200_ERR_BAD_ARG;
    }
    return SVTB200_OK;
}

int svtb200_dlf_bs_picture_dev(const svtb200_cu_info *cu_map_dev, uint32_t cu_stride, const uint8_t *cbf_map_dev, uint32_t cbf_stride, uint16_t pic_width,
                               uint16_t pic_height, uint8_t slice_type, uint64_t ref_poc_l0, uint64_t ref_poc_l1, uint8_t *vertical_bs_dev, uint8_t *horizontal_bs_dev,
                               void *cuda_stream)
{
    int rc = bs_check(cu_map_dev, cu_stride, cbf_map_dev, cbf_stride, pic_width, pic_height, slice_type, vertical_bs_dev, horizontal_bs_dev, "dlf_bs_picture_dev");
    if (rc || (rc = ensure_device())) return rc;
    BsArgs A;
    A.cu = cu_map_dev; A.cu_stride = (int)cu_stride; A.cbf = cbf_map_dev; A.cbf_stride = (int)cbf_stride;
    A.W = pic_width; A.H = pic_height; A.cols = (pic_width + 63) / 64;
    A.slice_type = slice_type; A.same_poc = ref_poc_l0 == ref_poc_l1;
    A.vbs = vertical_bs_dev; A.hbs = horizontal_bs_dev;
    dlf_bs_kernel<<<A.cols * ((pic_height + 63) / 64), 256, 0, (cudaStream_t)cuda_stream>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_dlf_bs_picture(const svtb200_cu_info *cu_map, uint32_t cu_stride, const uint8_t *cbf_map, uint32_t cbf_stride, uint16_t pic_width, uint16_t pic_height,
                           uint8_t slice_type, uint64_t ref_poc_l0, uint64_t ref_poc_l1, uint8_t *vertical_bs, uint8_t *horizontal_bs)
{
    int rc = bs_check(cu_map, cu_stride, cbf_map, cbf_stride, pic_width, pic_height, slice_type, vertical_bs, horizontal_bs, "dlf_bs_picture");
    if (rc) return rc;
    for (int y = 0; y < pic_height >> 3; ++y)
        for (int x = 0; x < pic_width >> 3; ++x) {
            const svtb200_cu_info &c = cu_map[(size_t)y * cu_stride + x];
            if (c.log2_cu_size < 3 || c.log2_cu_size > 6 || c.pred_dir > 2) { set_error("dlf_bs_picture: block (%d, %d): CU size or prediction direction out of range", x, y); return SVTB200_ERR_BAD_ARG; }
        }
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t cb = (size_t)cu_stride * (pic_height >> 3) * sizeof(svtb200_cu_info), fb = (size_t)cbf_stride * (pic_height >> 2);
    const size_t ob = (size_t)((pic_width + 63) / 64) * ((pic_height + 63) / 64) * 256;
    void *d_cu, *d_cbf, *d_v, *d_h;
    if ((rc = c->dev_reserve(cb + fb + 2 * ob + 4096)) || (rc = c->dev_alloc(cb, &d_cu)) || (rc = c->dev_alloc(fb, &d_cbf)) || (rc = c->dev_alloc(ob, &d_v)) || (rc = c->dev_alloc(ob, &d_h))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_cu, cu_map, cb, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_cbf, cbf_map, fb, cudaMemcpyHostToDevice, c->stream));
    if ((rc = svtb200_dlf_bs_picture_dev((const svtb200_cu_info *)d_cu, cu_stride, (const uint8_t *)d_cbf, cbf_stride, pic_width, pic_height, slice_type, ref_poc_l0, ref_poc_l1,
                                         (uint8_t *)d_v, (uint8_t *)d_h, c->stream))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(vertical_bs, d_v, ob, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(horizontal_bs, d_h, ob, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

}  // extern "C"
