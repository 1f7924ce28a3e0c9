/*
 * oracle/ref_mc_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Path-level oracle for inter prediction: runs the reference's OWN EncodePassInterPrediction
 * (Source/Lib/Codec/EbInterPrediction.c:761) -- and through it EncodeUniPredInterpolation /
 * EncodeBiPredInterpolation (Codec/EbMcp.c:175,562) and the table-dispatched interpolation kernels -- for a
 * list of PUs described with the product's ABI structs.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "EbDefinitions.h"
#include "EbPictureControlSet.h"
#include "EbSequenceControlSet.h"
#include "EbEncodeContext.h"
#include "EbInterPrediction.h"
#include "EbMcp.h"
#include "EbReferenceObject.h"
#include "EbSystemResourceManager.h"
#include "EbPictureBufferDesc.h"

#include "svt_hevc_b200.h"

extern EB_U32 EbHevcGetCpuAsmType(void);

static void fill_yuv(EbPictureBufferDesc_t *d, const svtb200_yuv_picture *p)
{
    memset(d, 0, sizeof(*d));
    d->bufferY = p->y.buf; d->bufferCb = p->cb.buf; d->bufferCr = p->cr.buf;
    d->strideY = (EB_U16)p->y.stride; d->strideCb = (EB_U16)p->cb.stride; d->strideCr = (EB_U16)p->cr.stride;
    d->originX = p->y.origin_x; d->originY = p->y.origin_y;
    d->width = p->y.width; d->height = p->y.height;
    d->maxWidth = p->y.width; d->maxHeight = p->y.height;
    d->bitDepth = EB_8BIT;
    d->colorFormat = EB_YUV420;
}

static void on_error(EB_PTR handle, EB_U32 code)
{
    (void)handle;
    fprintf(stderr, "ref_mc_harness: reference raised error 0x%x\n", code);
    abort();
}

int svtref_mc_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs,
                    const svtb200_yuv_picture *dst, int asm_mode)
{
    /* note: ASM_TYPES is a process global of the reference; concurrent callers must agree on asm_mode */
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    EbPictureBufferDesc_t *refDesc = (EbPictureBufferDesc_t *)calloc((size_t)n_refs, sizeof(*refDesc)), dstDesc;
    EbReferenceObject_t *refObj = (EbReferenceObject_t *)calloc((size_t)n_refs, sizeof(*refObj));
    EbObjectWrapper_t *refWrap = (EbObjectWrapper_t *)calloc((size_t)n_refs, sizeof(*refWrap));
    for (int i = 0; i < n_refs; ++i) {
        fill_yuv(&refDesc[i], &refs[i]);
        refObj[i].referencePicture = &refDesc[i];
        refWrap[i].objectPtr = &refObj[i];
    }
    fill_yuv(&dstDesc, dst);

    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec));
    ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec;
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->sequenceControlSetWrapperPtr = &scsWrap;

    MotionCompensationPredictionContext_t *mcp = (MotionCompensationPredictionContext_t *)calloc(1, sizeof(*mcp));
    if (MotionCompensationPredictionContextCtor(mcp, 64, 64, EB_FALSE) != EB_ErrorNone) return -1;

    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_mc_job *q = &jobs[j];
        MvUnit_t mvu; memset(&mvu, 0, sizeof mvu);
        mvu.predDirection = q->dir;
        for (int l = 0; l < 2; ++l) { mvu.mv[l].x = q->mv[l][0]; mvu.mv[l].y = q->mv[l][1]; }
        pcs->refPicPtrArray[0] = &refWrap[q->ref_idx[0]];
        pcs->refPicPtrArray[1] = &refWrap[q->ref_idx[1]];
        EncodePassInterPrediction(&mvu, q->x, q->y, q->w, q->h, pcs, &dstDesc, mcp);
    }
    if (mcp->dctor) mcp->dctor(mcp);
    free(mcp); free(pcs); free(scs); free(ec); free(refWrap); free(refObj); free(refDesc);
    return 0;
}

/* 16-bit samples: EncodePassInterPrediction16bit (Codec/EbInterPrediction.c:928) on EbReferenceObject_t::referencePicture16bit.
 * Planes hold EB_U16 samples (svtb200_plane::buf points at them; strides / origins in samples). */
int svtref_mc_batch16(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    EbPictureBufferDesc_t *refDesc = (EbPictureBufferDesc_t *)calloc((size_t)n_refs, sizeof(*refDesc)), dstDesc;
    EbReferenceObject_t *refObj = (EbReferenceObject_t *)calloc((size_t)n_refs, sizeof(*refObj));
    EbObjectWrapper_t *refWrap = (EbObjectWrapper_t *)calloc((size_t)n_refs, sizeof(*refWrap));
    for (int i = 0; i < n_refs; ++i) {
        fill_yuv(&refDesc[i], &refs[i]);
        refDesc[i].bitDepth = EB_16BIT;
        refObj[i].referencePicture16bit = &refDesc[i];
        refWrap[i].objectPtr = &refObj[i];
    }
    fill_yuv(&dstDesc, dst);
    dstDesc.bitDepth = EB_16BIT;

    EbCallback_t cb; memset(&cb, 0, sizeof cb); cb.ErrorHandler = on_error;
    EncodeContext_t *ec = (EncodeContext_t *)calloc(1, sizeof(*ec));
    ec->appCallbackPtr = &cb;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    scs->encodeContextPtr = ec;
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->sequenceControlSetWrapperPtr = &scsWrap;

    MotionCompensationPredictionContext_t *mcp = (MotionCompensationPredictionContext_t *)calloc(1, sizeof(*mcp));
    if (MotionCompensationPredictionContextCtor(mcp, 64, 64, EB_TRUE) != EB_ErrorNone) return -1;
    /* the 16-bit path re-points the local reference block descriptors into the reference pictures: keep the owned pointers for the dtor */
    EB_U8 *keep[2][3] = { { mcp->localReferenceBlockL0->bufferY, mcp->localReferenceBlockL0->bufferCb, mcp->localReferenceBlockL0->bufferCr },
                          { mcp->localReferenceBlockL1->bufferY, mcp->localReferenceBlockL1->bufferCb, mcp->localReferenceBlockL1->bufferCr } };

    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_mc_job *q = &jobs[j];
        MvUnit_t mvu; memset(&mvu, 0, sizeof mvu);
        mvu.predDirection = q->dir;
        for (int l = 0; l < 2; ++l) { mvu.mv[l].x = q->mv[l][0]; mvu.mv[l].y = q->mv[l][1]; }
        pcs->refPicPtrArray[0] = &refWrap[q->ref_idx[0]];
        pcs->refPicPtrArray[1] = &refWrap[q->ref_idx[1]];
        EncodePassInterPrediction16bit(&mvu, q->x, q->y, q->w, q->h, pcs, &dstDesc, mcp);
    }
    mcp->localReferenceBlockL0->bufferY = keep[0][0]; mcp->localReferenceBlockL0->bufferCb = keep[0][1]; mcp->localReferenceBlockL0->bufferCr = keep[0][2];
    mcp->localReferenceBlockL1->bufferY = keep[1][0]; mcp->localReferenceBlockL1->bufferCb = keep[1][1]; mcp->localReferenceBlockL1->bufferCr = keep[1][2];
    if (mcp->dctor) mcp->dctor(mcp);
    free(mcp); free(pcs); free(scs); free(ec); free(refWrap); free(refObj); free(refDesc);
    return 0;
}

/* Interpolation-free prediction of the mode-decision fast loop: the reference's OWN UniPredIFreeRef8Bit / BiPredIFreeRef8Bit
 * (Source/Lib/Codec/EbAvcStyleMcp.c:172,321) with the position arithmetic of their caller Inter2Nx2NPuPredictionInterpolationFree
 * (Codec/EbInterPrediction.c:100-115,340-400; reference pictures carry the 80-sample padding REFPADD_QPEL assumes).  Every PU is written at its
 * own position of `dst` (the fast loop writes the candidate's prediction buffer at puOriginIndex). */
#include "EbAvcStyleMcp.h"
int svtref_mc_ifree_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    EbPictureBufferDesc_t *refDesc = (EbPictureBufferDesc_t *)calloc((size_t)n_refs, sizeof(*refDesc)), dstDesc;
    for (int i = 0; i < n_refs; ++i) {
        if (refs[i].y.origin_x != 80 || refs[i].y.origin_y != 80) { free(refDesc); return -2; }
        fill_yuv(&refDesc[i], &refs[i]);
    }
    fill_yuv(&dstDesc, dst);
    EB_U8 *t0 = (EB_U8 *)malloc(64 * 64 * 4), *t1 = (EB_U8 *)malloc(64 * 64 * 4), *t2 = (EB_U8 *)malloc(64 * 64 * 8);
    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_mc_job *q = &jobs[j];
        EB_U32 pos[2][2];
        EbPictureBufferDesc_t *rp[2];
        for (int l = 0; l < 2; ++l) {
            rp[l] = &refDesc[q->ref_idx[l] < n_refs ? q->ref_idx[l] : 0];
            pos[l][0] = (EB_U32)CLIP3((EB_S32)MVBOUNDLOW, (EB_S32)((rp[l]->width << 2) + MVBOUNDHIGH), (EB_S32)((q->x << 2) + REFPADD_QPEL + q->mv[l][0]));
            pos[l][1] = (EB_U32)CLIP3((EB_S32)MVBOUNDLOW, (EB_S32)((rp[l]->height << 2) + MVBOUNDHIGH), (EB_S32)((q->y << 2) + REFPADD_QPEL + q->mv[l][1]));
        }
        const EB_U32 li = (EB_U32)(dstDesc.originX + q->x + (dstDesc.originY + q->y) * dstDesc.strideY);
        const EB_U32 ci = (EB_U32)(((dstDesc.originX + q->x) >> 1) + ((dstDesc.originY + q->y) >> 1) * dstDesc.strideCb);
        if (q->dir == 2)
            BiPredIFreeRef8Bit(rp[0], rp[1], pos[0][0], pos[0][1], pos[1][0], pos[1][1], q->w, q->h, &dstDesc, li, ci, PICTURE_BUFFER_DESC_FULL_MASK, t0, t1, t2);
        else
            UniPredIFreeRef8Bit(rp[q->dir], pos[q->dir][0], pos[q->dir][1], q->w, q->h, &dstDesc, li, ci, PICTURE_BUFFER_DESC_FULL_MASK, t2);
    }
    free(t0); free(t1); free(t2); free(refDesc);
    return 0;
}

/* The mode-decision side of HEVC interpolation: the reference's OWN UniPredHevcInterpolationMd / BiPredHevcInterpolationMd (Source/Lib/Codec/EbMcp.c:99,321)
 * with the position arithmetic of their caller Inter2Nx2NPuPredictionHevc (Codec/EbInterPrediction.c:469-760; 8-bit path).  They are the same table-dispatched
 * kernels the encode pass runs, writing the candidate's prediction buffer; here every PU lands at its own position of `dst`, so the result must equal
 * svtref_mc_batch's (and through it svtb200_mc_batch's). */
int svtref_mc_md_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    EbPictureBufferDesc_t *refDesc = (EbPictureBufferDesc_t *)calloc((size_t)n_refs, sizeof(*refDesc)), dstDesc;
    for (int i = 0; i < n_refs; ++i) {
        if (refs[i].y.origin_x != 80 || refs[i].y.origin_y != 80) { free(refDesc); return -2; }
        fill_yuv(&refDesc[i], &refs[i]);
    }
    fill_yuv(&dstDesc, dst);
    EB_S16 *t0 = (EB_S16 *)malloc(64 * 64 * 8), *t1 = (EB_S16 *)malloc(64 * 64 * 8), *t2 = (EB_S16 *)malloc(72 * 72 * 8);
    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_mc_job *q = &jobs[j];
        EB_U32 pos[2][2];
        EbPictureBufferDesc_t *rp[2];
        for (int l = 0; l < 2; ++l) {
            rp[l] = &refDesc[q->ref_idx[l] < n_refs ? q->ref_idx[l] : 0];
            pos[l][0] = (EB_U32)CLIP3((EB_S32)MVBOUNDLOW, (EB_S32)((rp[l]->width << 2) + MVBOUNDHIGH), (EB_S32)((q->x << 2) + REFPADD_QPEL + q->mv[l][0]));
            pos[l][1] = (EB_U32)CLIP3((EB_S32)MVBOUNDLOW, (EB_S32)((rp[l]->height << 2) + MVBOUNDHIGH), (EB_S32)((q->y << 2) + REFPADD_QPEL + q->mv[l][1]));
        }
        const EB_U32 li = (EB_U32)(dstDesc.originX + q->x + (dstDesc.originY + q->y) * dstDesc.strideY);
        const EB_U32 ci = (EB_U32)(((dstDesc.originX + q->x) >> 1) + ((dstDesc.originY + q->y) >> 1) * dstDesc.strideCb);
        if (q->dir == 2)
            BiPredHevcInterpolationMd(rp[0], rp[1], pos[0][0], pos[0][1], pos[1][0], pos[1][1], q->w, q->h, &dstDesc, li, ci, t0, t1, t2, EB_FALSE, PICTURE_BUFFER_DESC_FULL_MASK);
        else
            UniPredHevcInterpolationMd(rp[q->dir], pos[q->dir][0], pos[q->dir][1], q->w, q->h, &dstDesc, li, ci, t0, t2, EB_FALSE, PICTURE_BUFFER_DESC_FULL_MASK);
    }
    free(t0); free(t1); free(t2); free(refDesc);
    return 0;
}
