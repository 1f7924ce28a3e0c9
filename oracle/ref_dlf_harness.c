/*
 * oracle/ref_dlf_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Path-level oracle for deblocking: runs the reference's OWN LCU drivers in the encoder's order -- for each LCU in
 * raster order LCUInternalAreaDLFCore, LCUBoundaryDLFCore, LCUPictureEdgeDLFCore
 * (Source/Lib/Codec/EbDeblockingFilter.c:2222,2828,3518; call site Codec/EbCodingLoop.c:4604-4635) -- on a picture
 * described with the product's ABI structs.  Single tile, 8-bit 4:2:0.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "EbDefinitions.h"
#include "EbPictureControlSet.h"
#include "EbSequenceControlSet.h"
#include "EbEncodeContext.h"
#include "EbCodingUnit.h"
#include "EbDeblockingFilter.h"
#include "EbPictureBufferDesc.h"
#include "EbSystemResourceManager.h"
#include "EbSampleAdaptiveOffset.h"

#include "svt_hevc_b200.h"

extern EB_U32 EbHevcGetCpuAsmType(void);

static void on_error(EB_PTR handle, EB_U32 code)
{
    (void)handle;
    fprintf(stderr, "ref_dlf_harness: reference raised error 0x%x\n", code);
    abort();
}

int svtref_dlf_picture(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs,
                       const uint8_t *qp_array, const svtb200_dlf_params *prm, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = prm->pic_width, H = prm->pic_height;
    const EB_U32 cols = (W + 63) / 64, rows = (H + 63) / 64, nLcu = cols * rows;

    EbPictureBufferDesc_t rd; memset(&rd, 0, sizeof rd);
    rd.bufferY = recon->y.buf; rd.bufferCb = recon->cb.buf; rd.bufferCr = recon->cr.buf;
    rd.strideY = (EB_U16)recon->y.stride; rd.strideCb = (EB_U16)recon->cb.stride; rd.strideCr = (EB_U16)recon->cr.stride;
    rd.originX = recon->y.origin_x; rd.originY = recon->y.origin_y; rd.width = (EB_U16)W; rd.height = (EB_U16)H;
    rd.maxWidth = (EB_U16)W; rd.maxHeight = (EB_U16)H; rd.bitDepth = EB_8BIT; rd.colorFormat = EB_YUV420;

    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec)); ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec; scs->lumaWidth = (EB_U16)W; scs->lumaHeight = (EB_U16)H; scs->chromaWidth = (EB_U16)(W / 2); scs->chromaHeight = (EB_U16)(H / 2);
    scs->lcuSize = 64;
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->sequenceControlSetWrapperPtr = &scsWrap;
    pcs->qpArray = (EB_U8 *)qp_array; pcs->qpArrayStride = prm->qp_array_stride;
    pcs->tcOffset = prm->tc_offset; pcs->betaOffset = prm->beta_offset; pcs->cbQpOffset = prm->cb_qp_offset; pcs->crQpOffset = prm->cr_qp_offset;
    pcs->verticalEdgeBSArray = (EB_U8 **)calloc(nLcu, sizeof(EB_U8 *));
    pcs->horizontalEdgeBSArray = (EB_U8 **)calloc(nLcu, sizeof(EB_U8 *));
    pcs->lcuPtrArray = (LargestCodingUnit_t **)calloc(nLcu, sizeof(LargestCodingUnit_t *));
    TileInfo_t tile; memset(&tile, 0, sizeof tile);
    tile.tileLcuEndX = (EB_U16)cols; tile.tileLcuEndY = (EB_U16)rows; tile.tilePxlEndX = (EB_U16)W; tile.tilePxlEndY = (EB_U16)H;
    LargestCodingUnit_t *lcus = (LargestCodingUnit_t *)calloc(nLcu, sizeof(LargestCodingUnit_t));
    for (EB_U32 i = 0; i < nLcu; ++i) {
        pcs->verticalEdgeBSArray[i] = (EB_U8 *)vertical_bs + (size_t)i * 256;
        pcs->horizontalEdgeBSArray[i] = (EB_U8 *)horizontal_bs + (size_t)i * 256;
        lcus[i].tileInfoPtr = &tile; lcus[i].sizeLog2 = 6;
        pcs->lcuPtrArray[i] = &lcus[i];
    }
    for (EB_U32 ly = 0; ly < rows; ++ly)
        for (EB_U32 lx = 0; lx < cols; ++lx) {
            const EB_U32 i = ly * cols + lx, ox = lx * 64, oy = ly * 64;
            const EB_U32 lw = W - ox < 64 ? W - ox : 64, lh = H - oy < 64 ? H - oy : 64;
            LCUInternalAreaDLFCore(&rd, ox, oy, lw, lh, pcs->verticalEdgeBSArray[i], pcs->horizontalEdgeBSArray[i], pcs);
            LCUBoundaryDLFCore(&rd, ox, oy, lw, lh, pcs->verticalEdgeBSArray[i], pcs->horizontalEdgeBSArray[i],
                               ly == 0 ? (EB_U8 *)EB_NULL : pcs->verticalEdgeBSArray[i - cols], lx == 0 ? (EB_U8 *)EB_NULL : pcs->horizontalEdgeBSArray[i - 1], pcs);
            LCUPictureEdgeDLFCore(&rd, i, ox, oy, lw, lh, pcs);
        }
    free(lcus); free(pcs->lcuPtrArray); free(pcs->verticalEdgeBSArray); free(pcs->horizontalEdgeBSArray); free(pcs); free(scs); free(ec);
    return 0;
}

/* 16-bit samples (10-bit video): the reference's LCU*DLFCore16bit drivers on EB_U16 planes (strides / origins in samples) */
int svtref_dlf_picture16(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs,
                       const uint8_t *qp_array, const svtb200_dlf_params *prm, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = prm->pic_width, H = prm->pic_height;
    const EB_U32 cols = (W + 63) / 64, rows = (H + 63) / 64, nLcu = cols * rows;

    EbPictureBufferDesc_t rd; memset(&rd, 0, sizeof rd);
    rd.bufferY = recon->y.buf; rd.bufferCb = recon->cb.buf; rd.bufferCr = recon->cr.buf;
    rd.strideY = (EB_U16)recon->y.stride; rd.strideCb = (EB_U16)recon->cb.stride; rd.strideCr = (EB_U16)recon->cr.stride;
    rd.originX = recon->y.origin_x; rd.originY = recon->y.origin_y; rd.width = (EB_U16)W; rd.height = (EB_U16)H;
    rd.maxWidth = (EB_U16)W; rd.maxHeight = (EB_U16)H; rd.bitDepth = EB_16BIT; rd.colorFormat = EB_YUV420;

    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec)); ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec; scs->lumaWidth = (EB_U16)W; scs->lumaHeight = (EB_U16)H; scs->chromaWidth = (EB_U16)(W / 2); scs->chromaHeight = (EB_U16)(H / 2);
    scs->lcuSize = 64;
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->sequenceControlSetWrapperPtr = &scsWrap;
    pcs->qpArray = (EB_U8 *)qp_array; pcs->qpArrayStride = prm->qp_array_stride;
    pcs->tcOffset = prm->tc_offset; pcs->betaOffset = prm->beta_offset; pcs->cbQpOffset = prm->cb_qp_offset; pcs->crQpOffset = prm->cr_qp_offset;
    pcs->verticalEdgeBSArray = (EB_U8 **)calloc(nLcu, sizeof(EB_U8 *));
    pcs->horizontalEdgeBSArray = (EB_U8 **)calloc(nLcu, sizeof(EB_U8 *));
    pcs->lcuPtrArray = (LargestCodingUnit_t **)calloc(nLcu, sizeof(LargestCodingUnit_t *));
    TileInfo_t tile; memset(&tile, 0, sizeof tile);
    tile.tileLcuEndX = (EB_U16)cols; tile.tileLcuEndY = (EB_U16)rows; tile.tilePxlEndX = (EB_U16)W; tile.tilePxlEndY = (EB_U16)H;
    LargestCodingUnit_t *lcus = (LargestCodingUnit_t *)calloc(nLcu, sizeof(LargestCodingUnit_t));
    for (EB_U32 i = 0; i < nLcu; ++i) {
        pcs->verticalEdgeBSArray[i] = (EB_U8 *)vertical_bs + (size_t)i * 256;
        pcs->horizontalEdgeBSArray[i] = (EB_U8 *)horizontal_bs + (size_t)i * 256;
        lcus[i].tileInfoPtr = &tile; lcus[i].sizeLog2 = 6;
        pcs->lcuPtrArray[i] = &lcus[i];
    }
    for (EB_U32 ly = 0; ly < rows; ++ly)
        for (EB_U32 lx = 0; lx < cols; ++lx) {
            const EB_U32 i = ly * cols + lx, ox = lx * 64, oy = ly * 64;
            const EB_U32 lw = W - ox < 64 ? W - ox : 64, lh = H - oy < 64 ? H - oy : 64;
            LCUInternalAreaDLFCore16bit(&rd, ox, oy, lw, lh, pcs->verticalEdgeBSArray[i], pcs->horizontalEdgeBSArray[i], pcs);
            LCUBoundaryDLFCore16bit(&rd, ox, oy, lw, lh, pcs->verticalEdgeBSArray[i], pcs->horizontalEdgeBSArray[i],
                               ly == 0 ? (EB_U8 *)EB_NULL : pcs->verticalEdgeBSArray[i - cols], lx == 0 ? (EB_U8 *)EB_NULL : pcs->horizontalEdgeBSArray[i - 1], pcs);
            LCUPictureEdgeDLFCore16bit(&rd, i, ox, oy, lw, lh, pcs);
        }
    free(lcus); free(pcs->lcuPtrArray); free(pcs->verticalEdgeBSArray); free(pcs->horizontalEdgeBSArray); free(pcs); free(scs); free(ec);
    return 0;
}

/* the encoder gathers the SAO statistics of LCU n right after LCU n's own deblocking step (Codec/EbCodingLoop.c:4600-4760), i.e. on a
 * picture whose later LCUs are not deblocked yet and whose own right / bottom strips still wait for their neighbours' boundary step.  This driver
 * reproduces exactly that interleaving on `recon` (deblocked in place, final result = svtref_dlf_picture) and returns the statistics the encoder
 * would see; `snapshot` (optional, same geometry as recon) receives every LCU's samples as they were at its gather time. */
static const svtb200_plane *plq2(const svtb200_yuv_picture *p, int c) { return c == 0 ? &p->y : c == 1 ? &p->cb : &p->cr; }
int svtref_dlf_sao_stats_picture(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *src, const uint8_t *vertical_bs, const uint8_t *horizontal_bs,
                                 const uint8_t *qp_array, const svtb200_dlf_params *prm, svtb200_sao_stats *stats, const svtb200_yuv_picture *snapshot, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = prm->pic_width, H = prm->pic_height;
    const EB_U32 cols = (W + 63) / 64, rows = (H + 63) / 64, nLcu = cols * rows;

    EbPictureBufferDesc_t rd; memset(&rd, 0, sizeof rd);
    rd.bufferY = recon->y.buf; rd.bufferCb = recon->cb.buf; rd.bufferCr = recon->cr.buf;
    rd.strideY = (EB_U16)recon->y.stride; rd.strideCb = (EB_U16)recon->cb.stride; rd.strideCr = (EB_U16)recon->cr.stride;
    rd.originX = recon->y.origin_x; rd.originY = recon->y.origin_y; rd.width = (EB_U16)W; rd.height = (EB_U16)H;
    rd.maxWidth = (EB_U16)W; rd.maxHeight = (EB_U16)H; rd.bitDepth = EB_8BIT; rd.colorFormat = EB_YUV420;

    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec)); ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec; scs->lumaWidth = (EB_U16)W; scs->lumaHeight = (EB_U16)H; scs->chromaWidth = (EB_U16)(W / 2); scs->chromaHeight = (EB_U16)(H / 2);
    scs->lcuSize = 64;
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->sequenceControlSetWrapperPtr = &scsWrap;
    pcs->qpArray = (EB_U8 *)qp_array; pcs->qpArrayStride = prm->qp_array_stride;
    pcs->tcOffset = prm->tc_offset; pcs->betaOffset = prm->beta_offset; pcs->cbQpOffset = prm->cb_qp_offset; pcs->crQpOffset = prm->cr_qp_offset;
    pcs->verticalEdgeBSArray = (EB_U8 **)calloc(nLcu, sizeof(EB_U8 *));
    pcs->horizontalEdgeBSArray = (EB_U8 **)calloc(nLcu, sizeof(EB_U8 *));
    pcs->lcuPtrArray = (LargestCodingUnit_t **)calloc(nLcu, sizeof(LargestCodingUnit_t *));
    TileInfo_t tile; memset(&tile, 0, sizeof tile);
    tile.tileLcuEndX = (EB_U16)cols; tile.tileLcuEndY = (EB_U16)rows; tile.tilePxlEndX = (EB_U16)W; tile.tilePxlEndY = (EB_U16)H;
    LargestCodingUnit_t *lcus = (LargestCodingUnit_t *)calloc(nLcu, sizeof(LargestCodingUnit_t));
    for (EB_U32 i = 0; i < nLcu; ++i) {
        pcs->verticalEdgeBSArray[i] = (EB_U8 *)vertical_bs + (size_t)i * 256;
        pcs->horizontalEdgeBSArray[i] = (EB_U8 *)horizontal_bs + (size_t)i * 256;
        lcus[i].tileInfoPtr = &tile; lcus[i].sizeLog2 = 6;
        pcs->lcuPtrArray[i] = &lcus[i];
    }
    for (EB_U32 ly = 0; ly < rows; ++ly)
        for (EB_U32 lx = 0; lx < cols; ++lx) {
            const EB_U32 i = ly * cols + lx, ox = lx * 64, oy = ly * 64;
            const EB_U32 lw = W - ox < 64 ? W - ox : 64, lh = H - oy < 64 ? H - oy : 64;
            LCUInternalAreaDLFCore(&rd, ox, oy, lw, lh, pcs->verticalEdgeBSArray[i], pcs->horizontalEdgeBSArray[i], pcs);
            LCUBoundaryDLFCore(&rd, ox, oy, lw, lh, pcs->verticalEdgeBSArray[i], pcs->horizontalEdgeBSArray[i],
                               ly == 0 ? (EB_U8 *)EB_NULL : pcs->verticalEdgeBSArray[i - cols], lx == 0 ? (EB_U8 *)EB_NULL : pcs->horizontalEdgeBSArray[i - 1], pcs);
            LCUPictureEdgeDLFCore(&rd, i, ox, oy, lw, lh, pcs);
            /* SaoGenerationDecision :700-735: gather on the LCU as it is now */
            for (int c = 0; c < 3; ++c) {
                const int sh = c ? 1 : 0;
                const svtb200_plane *sp = plq2(src, c), *rp = plq2(recon, c);
                const EB_U32 cx = ox >> sh, cy = oy >> sh, cw = lw >> sh, ch = lh >> sh;
                svtb200_sao_stats *o = &stats[(size_t)i * 3 + c];
                SaoGatherFunctionTableLossy[!!(ASM_TYPES & PREAVX2_MASK)](
                    sp->buf + (size_t)(sp->origin_y + cy) * sp->stride + sp->origin_x + cx, sp->stride,
                    rp->buf + (size_t)(rp->origin_y + cy) * rp->stride + rp->origin_x + cx, rp->stride, cw, ch,
                    o->bo_diff, o->bo_count, o->eo_diff, o->eo_count);
                if (snapshot) {
                    const svtb200_plane *qp_ = plq2(snapshot, c);
                    for (EB_U32 y = 0; y < ch; ++y)
                        memcpy(qp_->buf + (size_t)(qp_->origin_y + cy + y) * qp_->stride + qp_->origin_x + cx,
                               rp->buf + (size_t)(rp->origin_y + cy + y) * rp->stride + rp->origin_x + cx, cw);
                }
            }
        }
    free(lcus); free(pcs->lcuPtrArray); free(pcs->verticalEdgeBSArray); free(pcs->horizontalEdgeBSArray); free(pcs); free(scs); free(ec);
    return 0;
}

/* the same interleaving on 16-bit sample planes (10-bit video): LCU*DLFCore16bit + GatherSaoStatisticsLcu_62x62_16bit */
int svtref_dlf_sao_stats_picture16(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *src, const uint8_t *vertical_bs, const uint8_t *horizontal_bs,
                                 const uint8_t *qp_array, const svtb200_dlf_params *prm, svtb200_sao_stats *stats, const svtb200_yuv_picture *snapshot, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = prm->pic_width, H = prm->pic_height;
    const EB_U32 cols = (W + 63) / 64, rows = (H + 63) / 64, nLcu = cols * rows;

    EbPictureBufferDesc_t rd; memset(&rd, 0, sizeof rd);
    rd.bufferY = recon->y.buf; rd.bufferCb = recon->cb.buf; rd.bufferCr = recon->cr.buf;
    rd.strideY = (EB_U16)recon->y.stride; rd.strideCb = (EB_U16)recon->cb.stride; rd.strideCr = (EB_U16)recon->cr.stride;
    rd.originX = recon->y.origin_x; rd.originY = recon->y.origin_y; rd.width = (EB_U16)W; rd.height = (EB_U16)H;
    rd.maxWidth = (EB_U16)W; rd.maxHeight = (EB_U16)H; rd.bitDepth = EB_16BIT; rd.colorFormat = EB_YUV420;

    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec)); ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec; scs->lumaWidth = (EB_U16)W; scs->lumaHeight = (EB_U16)H; scs->chromaWidth = (EB_U16)(W / 2); scs->chromaHeight = (EB_U16)(H / 2);
    scs->lcuSize = 64;
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->sequenceControlSetWrapperPtr = &scsWrap;
    pcs->qpArray = (EB_U8 *)qp_array; pcs->qpArrayStride = prm->qp_array_stride;
    pcs->tcOffset = prm->tc_offset; pcs->betaOffset = prm->beta_offset; pcs->cbQpOffset = prm->cb_qp_offset; pcs->crQpOffset = prm->cr_qp_offset;
    pcs->verticalEdgeBSArray = (EB_U8 **)calloc(nLcu, sizeof(EB_U8 *));
    pcs->horizontalEdgeBSArray = (EB_U8 **)calloc(nLcu, sizeof(EB_U8 *));
    pcs->lcuPtrArray = (LargestCodingUnit_t **)calloc(nLcu, sizeof(LargestCodingUnit_t *));
    TileInfo_t tile; memset(&tile, 0, sizeof tile);
    tile.tileLcuEndX = (EB_U16)cols; tile.tileLcuEndY = (EB_U16)rows; tile.tilePxlEndX = (EB_U16)W; tile.tilePxlEndY = (EB_U16)H;
    LargestCodingUnit_t *lcus = (LargestCodingUnit_t *)calloc(nLcu, sizeof(LargestCodingUnit_t));
    for (EB_U32 i = 0; i < nLcu; ++i) {
        pcs->verticalEdgeBSArray[i] = (EB_U8 *)vertical_bs + (size_t)i * 256;
        pcs->horizontalEdgeBSArray[i] = (EB_U8 *)horizontal_bs + (size_t)i * 256;
        lcus[i].tileInfoPtr = &tile; lcus[i].sizeLog2 = 6;
        pcs->lcuPtrArray[i] = &lcus[i];
    }
    for (EB_U32 ly = 0; ly < rows; ++ly)
        for (EB_U32 lx = 0; lx < cols; ++lx) {
            const EB_U32 i = ly * cols + lx, ox = lx * 64, oy = ly * 64;
            const EB_U32 lw = W - ox < 64 ? W - ox : 64, lh = H - oy < 64 ? H - oy : 64;
            LCUInternalAreaDLFCore16bit(&rd, ox, oy, lw, lh, pcs->verticalEdgeBSArray[i], pcs->horizontalEdgeBSArray[i], pcs);
            LCUBoundaryDLFCore16bit(&rd, ox, oy, lw, lh, pcs->verticalEdgeBSArray[i], pcs->horizontalEdgeBSArray[i],
                               ly == 0 ? (EB_U8 *)EB_NULL : pcs->verticalEdgeBSArray[i - cols], lx == 0 ? (EB_U8 *)EB_NULL : pcs->horizontalEdgeBSArray[i - 1], pcs);
            LCUPictureEdgeDLFCore16bit(&rd, i, ox, oy, lw, lh, pcs);
            /* SaoGenerationDecision :700-735: gather on the LCU as it is now */
            for (int c = 0; c < 3; ++c) {
                const int sh = c ? 1 : 0;
                const svtb200_plane *sp = plq2(src, c), *rp = plq2(recon, c);
                const EB_U32 cx = ox >> sh, cy = oy >> sh, cw = lw >> sh, ch = lh >> sh;
                svtb200_sao_stats *o = &stats[(size_t)i * 3 + c];
                SaoGatherFunctionTabl_16bit[!!(ASM_TYPES & PREAVX2_MASK)](
                    (EB_U16 *)sp->buf + (size_t)(sp->origin_y + cy) * sp->stride + sp->origin_x + cx, sp->stride,
                    (EB_U16 *)rp->buf + (size_t)(rp->origin_y + cy) * rp->stride + rp->origin_x + cx, rp->stride, cw, ch,
                    o->bo_diff, o->bo_count, o->eo_diff, o->eo_count);
                if (snapshot) {
                    const svtb200_plane *qp_ = plq2(snapshot, c);
                    for (EB_U32 y = 0; y < ch; ++y)
                        memcpy((EB_U16 *)qp_->buf + (size_t)(qp_->origin_y + cy + y) * qp_->stride + qp_->origin_x + cx,
                               (EB_U16 *)rp->buf + (size_t)(rp->origin_y + cy + y) * rp->stride + rp->origin_x + cx, cw * 2);
                }
            }
        }
    free(lcus); free(pcs->lcuPtrArray); free(pcs->verticalEdgeBSArray); free(pcs->horizontalEdgeBSArray); free(pcs); free(scs); free(ec);
    return 0;
}
