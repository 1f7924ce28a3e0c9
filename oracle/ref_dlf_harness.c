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

/* ---- boundary strengths: the reference's OWN SetBSArrayBasedOnPUBoundary / SetBSArrayBasedOnTUBoundary (Codec/EbDeblockingFilter.c:339,472), called CU by
 * CU like the encode pass does (Codec/EbCodingLoop.c:3554,4329,4364), with the neighbour arrays of each call filled from the picture's final maps (the
 * blocks left of / above a CU are coded before it; its own cbf entries are written before the call, :4320-4326). */
#include "EbNeighborArrays.h"
#include "EbUtility.h"
#include "EbReferenceObject.h"
int svtref_dlf_bs_picture(const svtb200_cu_info *cu_map, uint32_t cu_stride, const uint8_t *cbf, uint32_t cbf_stride, uint16_t W, uint16_t H, uint8_t slice_type,
                          uint64_t poc0, uint64_t poc1, uint8_t *vbs, uint8_t *hbs)
{
    const EB_U32 cols = (W + 63) / 64, rows = (H + 63) / 64, nLcu = cols * rows;
    memset(vbs, 0, (size_t)nLcu * 256); memset(hbs, 0, (size_t)nLcu * 256);
    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec)); ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec; scs->lumaWidth = W; scs->lumaHeight = H; scs->lcuSize = 64;
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->sequenceControlSetWrapperPtr = &scsWrap;
    pcs->sliceType = slice_type == 0 ? EB_B_PICTURE : slice_type == 1 ? EB_P_PICTURE : EB_I_PICTURE;
    /* the reference indexes cbfMapArray with lumaWidth >> 2 entries per row */
    pcs->cbfMapArray = (EB_U8 *)calloc((size_t)(W >> 2) * (H >> 2), 1);
    for (int y = 0; y < H >> 2; ++y) memcpy(pcs->cbfMapArray + (size_t)y * (W >> 2), cbf + (size_t)y * cbf_stride, W >> 2);
    EbReferenceObject_t ro[2]; memset(ro, 0, sizeof ro); ro[0].refPOC = poc0; ro[1].refPOC = poc1;
    EbObjectWrapper_t rw[2]; memset(rw, 0, sizeof rw); rw[0].objectPtr = &ro[0]; rw[1].objectPtr = &ro[1];
    pcs->refPicPtrArray[0] = &rw[0]; pcs->refPicPtrArray[1] = &rw[1];
    /* neighbour arrays at the 4-sample granularity the callee indexes them with (LOG_MIN_PU_SIZE) */
    NeighborArrayUnit_t modeNA, mvNA; memset(&modeNA, 0, sizeof modeNA); memset(&mvNA, 0, sizeof mvNA);
    modeNA.topArray = (EB_U8 *)calloc(W / 4 + 16, 1); modeNA.leftArray = (EB_U8 *)calloc(H / 4 + 16, 1);
    mvNA.topArray = (EB_U8 *)calloc(W / 4 + 16, sizeof(MvUnit_t)); mvNA.leftArray = (EB_U8 *)calloc(H / 4 + 16, sizeof(MvUnit_t));
    for (EB_U32 lcu = 0; lcu < nLcu; ++lcu) {
        const EB_U32 ox = (lcu % cols) * 64, oy = (lcu / cols) * 64;
        for (EB_U32 ci = 0; ci < CU_MAX_COUNT; ++ci) {
            const CodedUnitStats_t *st = GetCodedUnitStats(ci);
            const EB_U32 x0 = ox + st->originX, y0 = oy + st->originY;
            if (x0 >= W || y0 >= H) continue;
            const svtb200_cu_info *me = &cu_map[(size_t)(y0 >> 3) * cu_stride + (x0 >> 3)];
            if ((1u << me->log2_cu_size) != st->size) continue;       /* not a CU of this picture's partitioning */
            for (EB_U32 k = 0; k < st->size && x0 + k < W; k += 4) {  /* row above */
                if (y0 == 0) break;
                const svtb200_cu_info *n = &cu_map[(size_t)((y0 >> 3) - 1) * cu_stride + ((x0 + k) >> 3)];
                MvUnit_t *m = (MvUnit_t *)mvNA.topArray + ((x0 + k) >> 2);
                modeNA.topArray[(x0 + k) >> 2] = n->intra ? INTRA_MODE : INTER_MODE;
                m->predDirection = n->pred_dir; m->mv[0].x = n->mv[0][0]; m->mv[0].y = n->mv[0][1]; m->mv[1].x = n->mv[1][0]; m->mv[1].y = n->mv[1][1];
            }
            for (EB_U32 k = 0; k < st->size && y0 + k < H; k += 4) {  /* column to the left */
                if (x0 == 0) break;
                const svtb200_cu_info *n = &cu_map[(size_t)((y0 + k) >> 3) * cu_stride + (x0 >> 3) - 1];
                MvUnit_t *m = (MvUnit_t *)mvNA.leftArray + ((y0 + k) >> 2);
                modeNA.leftArray[(y0 + k) >> 2] = n->intra ? INTRA_MODE : INTER_MODE;
                m->predDirection = n->pred_dir; m->mv[0].x = n->mv[0][0]; m->mv[0].y = n->mv[0][1]; m->mv[1].x = n->mv[1][0]; m->mv[1].y = n->mv[1][1];
            }
            PredictionUnit_t pu; memset(&pu, 0, sizeof pu);
            pu.interPredDirectionIndex = me->pred_dir; pu.mv[0].x = me->mv[0][0]; pu.mv[0].y = me->mv[0][1]; pu.mv[1].x = me->mv[1][0]; pu.mv[1].y = me->mv[1][1];
            CodingUnit_t cu; memset(&cu, 0, sizeof cu);
            cu.predictionModeFlag = me->intra ? INTRA_MODE : INTER_MODE;
            /* the picture is partial in its last LCU row / column: the callee walks the whole CU edge, so clip by giving it only CUs that lie inside (the
             * encoder never codes a CU crossing the picture edge) */
            if (x0 + st->size > W || y0 + st->size > H) continue;
            if (st->size == 64 && !me->intra) {      /* inter 64x64: four 32x32 TUs (EbCodingLoop.c:4328) */
                for (int t = 0; t < 4; ++t)
                    SetBSArrayBasedOnTUBoundary(x0 + (t & 1) * 32, y0 + (t >> 1) * 32, 32, 32, st, (EB_MODETYPE)cu.predictionModeFlag, ox, oy, pcs, hbs + (size_t)lcu * 256, vbs + (size_t)lcu * 256);
            }
            SetBSArrayBasedOnPUBoundary(&modeNA, &mvNA, &pu, &cu, st, ox, oy, x0 == 0 ? EB_TRUE : EB_FALSE, y0 == 0 ? EB_TRUE : EB_FALSE, pcs, hbs + (size_t)lcu * 256, vbs + (size_t)lcu * 256);
        }
    }
    free(modeNA.topArray); free(modeNA.leftArray); free(mvNA.topArray); free(mvNA.leftArray); free(pcs->cbfMapArray); free(pcs); free(scs); free(ec);
    return 0;
}
