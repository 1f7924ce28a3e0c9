/*
 * oracle/refinc_sao_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Path-level oracle for SAO.  ApplySaoOffsetsPicture / ApplySaoOffsetsLcu are `static` in the reference's
 * Source/Lib/Codec/EbEncDecProcess.c (:522, :215), so this translation unit #includes that file WHERE IT LIES (no copy)
 * and calls the static functions directly; it is built into its own shared object (oracle/_ref/libsvtref_sao_*.so) that
 * resolves every other symbol from libsvtref_*.so.  Statistics use the exported gather kernels through the reference's
 * own dispatch table (SaoGatherFunctionTableLossy, Codec/EbSampleAdaptiveOffset.h:178).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "EbEncDecProcess.c"        /* the reference source, compiled in place */
#include "EbSampleAdaptiveOffset.h"

#include "svt_hevc_b200.h"

extern EB_U32 EbHevcGetCpuAsmType(void);

static void sao_on_error(EB_PTR handle, EB_U32 code)
{
    (void)handle;
    fprintf(stderr, "refinc_sao_harness: reference raised error 0x%x\n", code);
    abort();
}

static void fill_rec(EbPictureBufferDesc_t *d, const svtb200_yuv_picture *p)
{
    memset(d, 0, sizeof(*d));
    d->bufferY = p->y.buf; d->bufferCb = p->cb.buf; d->bufferCr = p->cr.buf;
    d->strideY = (EB_U16)p->y.stride; d->strideCb = (EB_U16)p->cb.stride; d->strideCr = (EB_U16)p->cr.stride;
    d->originX = p->y.origin_x; d->originY = p->y.origin_y; d->width = p->y.width; d->height = p->y.height;
    d->maxWidth = p->y.width; d->maxHeight = p->y.height; d->bitDepth = EB_8BIT; d->colorFormat = EB_YUV420;
}

/* in-place on `rec`, exactly like the encoder */
int svtref_sao_apply_picture(const svtb200_yuv_picture *rec, const svtb200_sao_params *prm, const uint8_t sao_flag[2], int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = rec->y.width, H = rec->y.height, cols = (W + 63) / 64, rows = (H + 63) / 64, n = cols * rows;
    EbPictureBufferDesc_t rd, enh; fill_rec(&rd, rec); memset(&enh, 0, sizeof enh); enh.bitDepth = EB_8BIT;
    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = sao_on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec)); ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec; scs->lumaWidth = (EB_U16)W; scs->lumaHeight = (EB_U16)H; scs->chromaWidth = (EB_U16)(W / 2); scs->chromaHeight = (EB_U16)(H / 2); scs->lcuSize = 64;
    scs->lcuParamsArray = (LcuParams_t *)calloc(n, sizeof(LcuParams_t));
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureParentControlSet_t *ppcs = (PictureParentControlSet_t *)calloc(1, sizeof(*ppcs));
    ppcs->isUsedAsReferenceFlag = EB_FALSE; ppcs->enhancedPicturePtr = &enh; ppcs->sequenceControlSetWrapperPtr = &scsWrap;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->ParentPcsPtr = ppcs; pcs->sequenceControlSetWrapperPtr = &scsWrap; pcs->reconPicturePtr = &rd; pcs->colorFormat = EB_YUV420;
    pcs->saoFlag[0] = sao_flag[0]; pcs->saoFlag[1] = sao_flag[1];
    pcs->lcuPtrArray = (LargestCodingUnit_t **)calloc(n, sizeof(LargestCodingUnit_t *));
    LargestCodingUnit_t *lcus = (LargestCodingUnit_t *)calloc(n, sizeof(*lcus));
    LcuEdgeInfo_t *edges = (LcuEdgeInfo_t *)calloc(n, sizeof(*edges));
    TileInfo_t tile; memset(&tile, 0, sizeof tile);
    tile.tileLcuEndX = (EB_U16)cols; tile.tileLcuEndY = (EB_U16)rows; tile.tilePxlEndX = (EB_U16)W; tile.tilePxlEndY = (EB_U16)H;
    for (EB_U32 ly = 0; ly < rows; ++ly)
        for (EB_U32 lx = 0; lx < cols; ++lx) {
            const EB_U32 i = ly * cols + lx;
            LcuParams_t *lp = &scs->lcuParamsArray[i];
            lp->originX = (EB_U16)(lx * 64); lp->originY = (EB_U16)(ly * 64);
            lp->width = (EB_U8)(W - lx * 64 < 64 ? W - lx * 64 : 64); lp->height = (EB_U8)(H - ly * 64 < 64 ? H - ly * 64 : 64);
            edges[i].tileLeftEdgeFlag = lx == 0; edges[i].tileRightEdgeFlag = lx == cols - 1; edges[i].tileTopEdgeFlag = ly == 0;
            edges[i].pictureLeftEdgeFlag = lx == 0; edges[i].pictureRightEdgeFlag = lx == cols - 1; edges[i].pictureTopEdgeFlag = ly == 0;
            lcus[i].lcuEdgeInfoPtr = &edges[i]; lcus[i].tileInfoPtr = &tile; lcus[i].originX = lx * 64; lcus[i].originY = ly * 64; lcus[i].sizeLog2 = 6;
            lcus[i].pictureControlSetPtr = pcs;
            SaoParameters_t *sp = &lcus[i].saoParams;
            sp->saoTypeIndex[0] = prm[i].type[0]; sp->saoTypeIndex[1] = prm[i].type[1];
            for (int c = 0; c < 3; ++c) { sp->saoBandPosition[c] = prm[i].band_position[c]; for (int k = 0; k < 4; ++k) sp->saoOffset[c][k] = prm[i].offset[c][k]; }
            pcs->lcuPtrArray[i] = &lcus[i];
        }
    EncDecContext_t *ctx = (EncDecContext_t *)calloc(1, sizeof(*ctx));
    for (int k = 0; k < 2; ++k) { ctx->saoUpBuffer[k] = (EB_U8 *)calloc(W + 64, 1); ctx->saoLeftBuffer[k] = (EB_U8 *)calloc(64 + 8, 1); }
    ApplySaoOffsetsPicture(ctx, scs, pcs);
    for (int k = 0; k < 2; ++k) { free(ctx->saoUpBuffer[k]); free(ctx->saoLeftBuffer[k]); }
    free(ctx); free(edges); free(lcus); free(pcs->lcuPtrArray); free(pcs); free(ppcs); free(scs->lcuParamsArray); free(scs); free(ec);
    return 0;
}

static const svtb200_plane *plq(const svtb200_yuv_picture *p, int k) { return k == 0 ? &p->y : k == 1 ? &p->cb : &p->cr; }

int svtref_sao_stats_picture(const svtb200_yuv_picture *src, const svtb200_yuv_picture *rec, svtb200_sao_stats *out, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = rec->y.width, H = rec->y.height, cols = (W + 63) / 64, rows = (H + 63) / 64;
    for (EB_U32 ly = 0; ly < rows; ++ly)
        for (EB_U32 lx = 0; lx < cols; ++lx)
            for (int c = 0; c < 3; ++c) {
                const int sh = c ? 1 : 0;
                const svtb200_plane *sp = plq(src, c), *rp = plq(rec, c);
                const EB_U32 ox = (lx * 64) >> sh, oy = (ly * 64) >> sh;
                const EB_U32 lw = ((W - lx * 64 < 64 ? W - lx * 64 : 64)) >> sh, lh = ((H - ly * 64 < 64 ? H - ly * 64 : 64)) >> sh;
                svtb200_sao_stats *o = &out[(size_t)(ly * cols + lx) * 3 + c];
                SaoGatherFunctionTableLossy[!!(ASM_TYPES & PREAVX2_MASK)](
                    sp->buf + (size_t)(sp->origin_y + oy) * sp->stride + sp->origin_x + ox, sp->stride,
                    rp->buf + (size_t)(rp->origin_y + oy) * rp->stride + rp->origin_x + ox, rp->stride, lw, lh,
                    o->bo_diff, o->bo_count, o->eo_diff, o->eo_count);
            }
    return 0;
}

/* 16-bit samples (10-bit video) */
/* in-place on `rec`, exactly like the encoder */
int svtref_sao_apply_picture16(const svtb200_yuv_picture *rec, const svtb200_sao_params *prm, const uint8_t sao_flag[2], int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = rec->y.width, H = rec->y.height, cols = (W + 63) / 64, rows = (H + 63) / 64, n = cols * rows;
    EbPictureBufferDesc_t rd, enh; fill_rec(&rd, rec); rd.bitDepth = EB_16BIT; memset(&enh, 0, sizeof enh); enh.bitDepth = EB_8BIT;
    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = sao_on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec)); ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec; scs->lumaWidth = (EB_U16)W; scs->lumaHeight = (EB_U16)H; scs->chromaWidth = (EB_U16)(W / 2); scs->chromaHeight = (EB_U16)(H / 2); scs->lcuSize = 64;
    scs->lcuParamsArray = (LcuParams_t *)calloc(n, sizeof(LcuParams_t));
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureParentControlSet_t *ppcs = (PictureParentControlSet_t *)calloc(1, sizeof(*ppcs));
    ppcs->isUsedAsReferenceFlag = EB_FALSE; ppcs->enhancedPicturePtr = &enh; ppcs->sequenceControlSetWrapperPtr = &scsWrap;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->ParentPcsPtr = ppcs; pcs->sequenceControlSetWrapperPtr = &scsWrap; pcs->reconPicturePtr = &rd; pcs->reconPicture16bitPtr = &rd; pcs->colorFormat = EB_YUV420;
    pcs->saoFlag[0] = sao_flag[0]; pcs->saoFlag[1] = sao_flag[1];
    pcs->lcuPtrArray = (LargestCodingUnit_t **)calloc(n, sizeof(LargestCodingUnit_t *));
    LargestCodingUnit_t *lcus = (LargestCodingUnit_t *)calloc(n, sizeof(*lcus));
    LcuEdgeInfo_t *edges = (LcuEdgeInfo_t *)calloc(n, sizeof(*edges));
    TileInfo_t tile; memset(&tile, 0, sizeof tile);
    tile.tileLcuEndX = (EB_U16)cols; tile.tileLcuEndY = (EB_U16)rows; tile.tilePxlEndX = (EB_U16)W; tile.tilePxlEndY = (EB_U16)H;
    for (EB_U32 ly = 0; ly < rows; ++ly)
        for (EB_U32 lx = 0; lx < cols; ++lx) {
            const EB_U32 i = ly * cols + lx;
            LcuParams_t *lp = &scs->lcuParamsArray[i];
            lp->originX = (EB_U16)(lx * 64); lp->originY = (EB_U16)(ly * 64);
            lp->width = (EB_U8)(W - lx * 64 < 64 ? W - lx * 64 : 64); lp->height = (EB_U8)(H - ly * 64 < 64 ? H - ly * 64 : 64);
            edges[i].tileLeftEdgeFlag = lx == 0; edges[i].tileRightEdgeFlag = lx == cols - 1; edges[i].tileTopEdgeFlag = ly == 0;
            edges[i].pictureLeftEdgeFlag = lx == 0; edges[i].pictureRightEdgeFlag = lx == cols - 1; edges[i].pictureTopEdgeFlag = ly == 0;
            lcus[i].lcuEdgeInfoPtr = &edges[i]; lcus[i].tileInfoPtr = &tile; lcus[i].originX = lx * 64; lcus[i].originY = ly * 64; lcus[i].sizeLog2 = 6;
            lcus[i].pictureControlSetPtr = pcs;
            SaoParameters_t *sp = &lcus[i].saoParams;
            sp->saoTypeIndex[0] = prm[i].type[0]; sp->saoTypeIndex[1] = prm[i].type[1];
            for (int c = 0; c < 3; ++c) { sp->saoBandPosition[c] = prm[i].band_position[c]; for (int k = 0; k < 4; ++k) sp->saoOffset[c][k] = prm[i].offset[c][k]; }
            pcs->lcuPtrArray[i] = &lcus[i];
        }
    EncDecContext_t *ctx = (EncDecContext_t *)calloc(1, sizeof(*ctx));
    for (int k = 0; k < 2; ++k) { ctx->saoUpBuffer16[k] = (EB_U16 *)calloc(W + 64, 2); ctx->saoLeftBuffer16[k] = (EB_U16 *)calloc(64 + 8, 2); }
    ApplySaoOffsetsPicture16bit(ctx, scs, pcs);
    for (int k = 0; k < 2; ++k) { free(ctx->saoUpBuffer16[k]); free(ctx->saoLeftBuffer16[k]); }
    free(ctx); free(edges); free(lcus); free(pcs->lcuPtrArray); free(pcs); free(ppcs); free(scs->lcuParamsArray); free(scs); free(ec);
    return 0;
}


int svtref_sao_stats_picture16(const svtb200_yuv_picture *src, const svtb200_yuv_picture *rec, svtb200_sao_stats *out, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = rec->y.width, H = rec->y.height, cols = (W + 63) / 64, rows = (H + 63) / 64;
    for (EB_U32 ly = 0; ly < rows; ++ly)
        for (EB_U32 lx = 0; lx < cols; ++lx)
            for (int c = 0; c < 3; ++c) {
                const int sh = c ? 1 : 0;
                const svtb200_plane *sp = plq(src, c), *rp = plq(rec, c);
                const EB_U32 ox = (lx * 64) >> sh, oy = (ly * 64) >> sh;
                const EB_U32 lw = ((W - lx * 64 < 64 ? W - lx * 64 : 64)) >> sh, lh = ((H - ly * 64 < 64 ? H - ly * 64 : 64)) >> sh;
                svtb200_sao_stats *o = &out[(size_t)(ly * cols + lx) * 3 + c];
                SaoGatherFunctionTabl_16bit[!!(ASM_TYPES & PREAVX2_MASK)](
                    (EB_U16 *)sp->buf + (size_t)(sp->origin_y + oy) * sp->stride + sp->origin_x + ox, sp->stride,
                    (EB_U16 *)rp->buf + (size_t)(rp->origin_y + oy) * rp->stride + rp->origin_x + ox, rp->stride, lw, lh,
                    o->bo_diff, o->bo_count, o->eo_diff, o->eo_count);
            }
    return 0;
}
