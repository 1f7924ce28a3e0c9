/*
 * oracle/ref_intra_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Path-level oracle for intra prediction: drives the reference's OWN, unmodified
 *   EncodePassIntraPrediction               Source/Lib/Codec/EbIntraPrediction.c:4395   (mode dispatch, filtered-line choice)
 *   GenerateIntraReferenceSamplesEncodePass :212   (neighbour-array walk, substitution, smoothing)
 *   OpenLoopIntraSearchLcu                  Source/Lib/Codec/EbMotionEstimation.c:5053
 * with hand-built neighbour arrays / control sets, the way EncodePass (EbCodingLoop.c:3292-3350) and MotionEstimationKernel
 * (EbMotionEstimationProcess.c:790-800) call them.  Compiled against the reference headers where they lie and linked into
 * oracle/_ref/libsvtref_*.so by oracle/build_ref.sh.  Arguments use the product's ABI structs (include/svt_hevc_b200.h).
 */
#include <stdlib.h>
#include <string.h>

#include "EbDefinitions.h"
#include "EbPictureControlSet.h"
#include "EbSequenceControlSet.h"
#include "EbMotionEstimation.h"
#include "EbMotionEstimationProcess.h"
#include "EbMotionEstimationContext.h"
#include "EbIntraPrediction.h"
#include "EbNeighborArrays.h"
#include "EbPictureBufferDesc.h"
#include "EbCodingUnit.h"

#include "svt_hevc_b200.h"

extern EB_U32 EbHevcGetCpuAsmType(void);
extern EB_ERRORTYPE OpenLoopIntraSearchLcu(PictureParentControlSet_t *pictureControlSetPtr, EB_U32 lcuIndex, MotionEstimationContext_t *contextPtr,
                                           EbPictureBufferDesc_t *inputPtr);

static void yuv_desc(EbPictureBufferDesc_t *d, const svtb200_yuv_picture *p)
{
    memset(d, 0, sizeof(*d));
    d->bufferY = p->y.buf; d->bufferCb = p->cb.buf; d->bufferCr = p->cr.buf;
    d->strideY = (EB_U16)p->y.stride; d->strideCb = (EB_U16)p->cb.stride; d->strideCr = (EB_U16)p->cr.stride;
    d->originX = p->y.origin_x; d->originY = p->y.origin_y;
    d->width = p->y.width; d->height = p->y.height; d->maxWidth = p->y.width; d->maxHeight = p->y.height;
    d->bitDepth = EB_8BIT; d->colorFormat = EB_YUV420;
}

/* one mode on an explicit reference line: the line is installed as both the filtered and the unfiltered array, so whichever
 * EncodePassIntraPrediction selects is the caller's line */
int svtref_intra_predict(uint32_t mode, uint32_t size, int is_chroma, const uint8_t *line, uint8_t *pred, uint32_t stride, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    IntraReferenceSamples_t *r = (IntraReferenceSamples_t *)calloc(1, sizeof(*r));
    if (IntraReferenceSamplesCtor(r, EB_YUV420) != EB_ErrorNone) return -1;
    const uint32_t N = size;
    EB_U8 *fwd[4] = { r->yIntraReferenceArray, r->yIntraFilteredReferenceArray, r->cbIntraReferenceArray, r->crIntraReferenceArray };
    EB_U8 *rev[4] = { r->yIntraReferenceArrayReverse, r->yIntraFilteredReferenceArrayReverse, r->cbIntraReferenceArrayReverse, r->crIntraReferenceArrayReverse };
    for (int k = 0; k < 4; ++k) {
        memcpy(fwd[k], line, 4 * N + 1);
        memcpy(rev[k] + 2 * N, line + 2 * N, 2 * N + 1);
        for (uint32_t i = 0; i < 2 * N; ++i) rev[k][i] = line[2 * N - 1 - i];
    }
    /* prediction target: a scratch 64x64 (stride 64) block in each plane */
    EbPictureBufferDesc_t d;
    memset(&d, 0, sizeof(d));
    EB_U8 *tmp = (EB_U8 *)calloc(3, 64 * 64);
    d.bufferY = tmp; d.bufferCb = tmp + 64 * 64; d.bufferCr = tmp + 2 * 64 * 64;
    d.strideY = d.strideCb = d.strideCr = 64;
    d.bitDepth = EB_8BIT; d.colorFormat = EB_YUV420;
    if (!is_chroma) {
        EncodePassIntraPrediction(r, 0, 0, N, N >> 1, &d, EB_YUV420, EB_FALSE, mode, EB_INTRA_CHROMA_DM, PICTURE_BUFFER_DESC_LUMA_MASK);
        for (uint32_t y = 0; y < N; ++y) memcpy(pred + (size_t)y * stride, tmp + y * 64, N);
    } else {
        /* chroma origin = luma origin >> 1; the chroma block size is passed explicitly */
        EncodePassIntraPrediction(r, 0, 0, N << 1, N, &d, EB_YUV420, EB_FALSE, mode, EB_INTRA_CHROMA_DM, PICTURE_BUFFER_DESC_CHROMA_MASK);
        if (memcmp(d.bufferCb, d.bufferCr, 64 * 64)) return -2;
        for (uint32_t y = 0; y < N; ++y) memcpy(pred + (size_t)y * stride, d.bufferCb + y * 64, N);
    }
    free(tmp);
    if (r->dctor) r->dctor(r);
    free(r);
    return 0;
}

int svtref_intra_batch(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    NeighborArrayUnit_t *na[4];   /* mode type, luma, cb, cr */
    for (int k = 0; k < 4; ++k) {
        na[k] = (NeighborArrayUnit_t *)calloc(1, sizeof(NeighborArrayUnit_t));
        const EB_U32 w = k >= 2 ? MAX_PICTURE_WIDTH_SIZE >> 1 : MAX_PICTURE_WIDTH_SIZE, h = k >= 2 ? MAX_PICTURE_HEIGHT_SIZE >> 1 : MAX_PICTURE_HEIGHT_SIZE;
        const EB_U32 gran = k == 0 ? PU_NEIGHBOR_ARRAY_GRANULARITY : SAMPLE_NEIGHBOR_ARRAY_GRANULARITY;
        if (NeighborArrayUnitCtor(na[k], w, h, sizeof(EB_U8), gran, gran, NEIGHBOR_ARRAY_UNIT_FULL_MASK) != EB_ErrorNone) return -1;
    }
    IntraReferenceSamples_t *r = (IntraReferenceSamples_t *)calloc(1, sizeof(*r));
    if (IntraReferenceSamplesCtor(r, EB_YUV420) != EB_ErrorNone) return -1;
    EbPictureBufferDesc_t pd;
    yuv_desc(&pd, pred);
    const uint8_t *ry = recon->y.buf + (size_t)recon->y.origin_y * recon->y.stride + recon->y.origin_x;
    const uint8_t *rcb = recon->cb.buf + (size_t)recon->cb.origin_y * recon->cb.stride + recon->cb.origin_x;
    const uint8_t *rcr = recon->cr.buf + (size_t)recon->cr.origin_y * recon->cr.stride + recon->cr.origin_x;
    /* scratch copies: prediction may be in place, and each job must see the neighbours as they are before it */
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const svtb200_intra_job *t = &jobs[j];
        const int x = t->x, y = t->y, N = t->size;
        NeighborArrayUnitReset(na[0]);
        for (int i = 0; i < 2 * N; ++i) {
            na[1]->topArray[x + i] = ry[(ptrdiff_t)(y - 1) * (ptrdiff_t)recon->y.stride + x + i];
            na[1]->leftArray[y + i] = ry[(ptrdiff_t)(y + i) * (ptrdiff_t)recon->y.stride + x - 1];
        }
        na[1]->topLeftArray[MAX_PICTURE_HEIGHT_SIZE + x - y] = ry[(ptrdiff_t)(y - 1) * (ptrdiff_t)recon->y.stride + x - 1];
        const int xc = x >> 1, yc = y >> 1;
        for (int i = 0; i < N; ++i) {
            na[2]->topArray[xc + i] = rcb[(ptrdiff_t)(yc - 1) * (ptrdiff_t)recon->cb.stride + xc + i];
            na[2]->leftArray[yc + i] = rcb[(ptrdiff_t)(yc + i) * (ptrdiff_t)recon->cb.stride + xc - 1];
            na[3]->topArray[xc + i] = rcr[(ptrdiff_t)(yc - 1) * (ptrdiff_t)recon->cr.stride + xc + i];
            na[3]->leftArray[yc + i] = rcr[(ptrdiff_t)(yc + i) * (ptrdiff_t)recon->cr.stride + xc - 1];
        }
        na[2]->topLeftArray[((MAX_PICTURE_HEIGHT_SIZE - y) >> 1) + xc] = rcb[(ptrdiff_t)(yc - 1) * (ptrdiff_t)recon->cb.stride + xc - 1];
        na[3]->topLeftArray[((MAX_PICTURE_HEIGHT_SIZE - y) >> 1) + xc] = rcr[(ptrdiff_t)(yc - 1) * (ptrdiff_t)recon->cr.stride + xc - 1];
        const int nu = 2 * N / 4;
        for (int u = 0; u < nu; ++u) {
            if ((t->left_avail >> u) & 1) na[0]->leftArray[(y + 2 * N - 4 - 4 * u) >> 2] = INTRA_MODE;
            if ((t->top_avail >> u) & 1) na[0]->topArray[(x + 4 * u) >> 2] = INTRA_MODE;
        }
        if (t->flags & SVTB200_INTRA_TOPLEFT_AVAIL) na[0]->topLeftArray[GetNeighborArrayUnitTopLeftIndex(na[0], x, y)] = INTRA_MODE;
        EB_U32 depth = 0;
        while ((64u >> depth) > (EB_U32)N) ++depth;
        GenerateIntraReferenceSamplesEncodePass(EB_FALSE, (t->flags & SVTB200_INTRA_STRONG_SMOOTHING) ? EB_TRUE : EB_FALSE, (EB_U32)x, (EB_U32)y, (EB_U32)N, 64, depth,
                                                na[0], na[1], na[2], na[3], r, EB_YUV420, EB_FALSE, EB_FALSE, EB_FALSE);
        EB_U32 mask = 0;
        if (t->luma_mode != SVTB200_INTRA_SKIP) mask |= PICTURE_BUFFER_DESC_LUMA_MASK;
        if (t->chroma_mode != SVTB200_INTRA_SKIP) mask |= PICTURE_BUFFER_DESC_CHROMA_MASK;
        EncodePassIntraPrediction(r, (EB_U32)x + pd.originX, (EB_U32)y + pd.originY, (EB_U32)N, (EB_U32)N >> 1, &pd, EB_YUV420, EB_FALSE,
                                  t->luma_mode == SVTB200_INTRA_SKIP ? 0 : t->luma_mode, t->chroma_mode == SVTB200_INTRA_SKIP ? EB_INTRA_CHROMA_DM : t->chroma_mode, mask);
    }
    if (r->dctor) r->dctor(r);
    free(r);
    for (int k = 0; k < 4; ++k) { if (na[k]->dctor) na[k]->dctor(na[k]); free(na[k]); }
    return 0;
}

/* 16-bit samples (10-bit video): GenerateIntraReference16bitSamplesEncodePass :1688 + EncodePassIntraPrediction16bit :4676; planes hold EB_U16 */
int svtref_intra_batch16(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    NeighborArrayUnit_t *na[4];   /* mode type, luma, cb, cr */
    for (int k = 0; k < 4; ++k) {
        na[k] = (NeighborArrayUnit_t *)calloc(1, sizeof(NeighborArrayUnit_t));
        const EB_U32 w = k >= 2 ? MAX_PICTURE_WIDTH_SIZE >> 1 : MAX_PICTURE_WIDTH_SIZE, h = k >= 2 ? MAX_PICTURE_HEIGHT_SIZE >> 1 : MAX_PICTURE_HEIGHT_SIZE;
        const EB_U32 gran = k == 0 ? PU_NEIGHBOR_ARRAY_GRANULARITY : SAMPLE_NEIGHBOR_ARRAY_GRANULARITY;
        if (NeighborArrayUnitCtor(na[k], w, h, k == 0 ? sizeof(EB_U8) : sizeof(EB_U16), gran, gran, NEIGHBOR_ARRAY_UNIT_FULL_MASK) != EB_ErrorNone) return -1;
    }
    IntraReference16bitSamples_t *r = (IntraReference16bitSamples_t *)calloc(1, sizeof(*r));
    if (IntraReference16bitSamplesCtor(r, EB_YUV420) != EB_ErrorNone) return -1;
    EbPictureBufferDesc_t pd;
    yuv_desc(&pd, pred);
    pd.bitDepth = EB_16BIT;
    const uint16_t *ry = (const uint16_t *)recon->y.buf + (size_t)recon->y.origin_y * recon->y.stride + recon->y.origin_x;
    const uint16_t *rcb = (const uint16_t *)recon->cb.buf + (size_t)recon->cb.origin_y * recon->cb.stride + recon->cb.origin_x;
    const uint16_t *rcr = (const uint16_t *)recon->cr.buf + (size_t)recon->cr.origin_y * recon->cr.stride + recon->cr.origin_x;
    /* scratch copies: prediction may be in place, and each job must see the neighbours as they are before it */
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const svtb200_intra_job *t = &jobs[j];
        const int x = t->x, y = t->y, N = t->size;
        NeighborArrayUnitReset(na[0]);
        for (int i = 0; i < 2 * N; ++i) {
            ((EB_U16 *)na[1]->topArray)[x + i] = ry[(ptrdiff_t)(y - 1) * (ptrdiff_t)recon->y.stride + x + i];
            ((EB_U16 *)na[1]->leftArray)[y + i] = ry[(ptrdiff_t)(y + i) * (ptrdiff_t)recon->y.stride + x - 1];
        }
        ((EB_U16 *)na[1]->topLeftArray)[MAX_PICTURE_HEIGHT_SIZE + x - y] = ry[(ptrdiff_t)(y - 1) * (ptrdiff_t)recon->y.stride + x - 1];
        const int xc = x >> 1, yc = y >> 1;
        for (int i = 0; i < N; ++i) {
            ((EB_U16 *)na[2]->topArray)[xc + i] = rcb[(ptrdiff_t)(yc - 1) * (ptrdiff_t)recon->cb.stride + xc + i];
            ((EB_U16 *)na[2]->leftArray)[yc + i] = rcb[(ptrdiff_t)(yc + i) * (ptrdiff_t)recon->cb.stride + xc - 1];
            ((EB_U16 *)na[3]->topArray)[xc + i] = rcr[(ptrdiff_t)(yc - 1) * (ptrdiff_t)recon->cr.stride + xc + i];
            ((EB_U16 *)na[3]->leftArray)[yc + i] = rcr[(ptrdiff_t)(yc + i) * (ptrdiff_t)recon->cr.stride + xc - 1];
        }
        ((EB_U16 *)na[2]->topLeftArray)[((MAX_PICTURE_HEIGHT_SIZE - y) >> 1) + xc] = rcb[(ptrdiff_t)(yc - 1) * (ptrdiff_t)recon->cb.stride + xc - 1];
        ((EB_U16 *)na[3]->topLeftArray)[((MAX_PICTURE_HEIGHT_SIZE - y) >> 1) + xc] = rcr[(ptrdiff_t)(yc - 1) * (ptrdiff_t)recon->cr.stride + xc - 1];
        const int nu = 2 * N / 4;
        for (int u = 0; u < nu; ++u) {
            if ((t->left_avail >> u) & 1) na[0]->leftArray[(y + 2 * N - 4 - 4 * u) >> 2] = INTRA_MODE;
            if ((t->top_avail >> u) & 1) na[0]->topArray[(x + 4 * u) >> 2] = INTRA_MODE;
        }
        if (t->flags & SVTB200_INTRA_TOPLEFT_AVAIL) na[0]->topLeftArray[GetNeighborArrayUnitTopLeftIndex(na[0], x, y)] = INTRA_MODE;
        EB_U32 depth = 0;
        while ((64u >> depth) > (EB_U32)N) ++depth;
        GenerateIntraReference16bitSamplesEncodePass(EB_FALSE, (t->flags & SVTB200_INTRA_STRONG_SMOOTHING) ? EB_TRUE : EB_FALSE, (EB_U32)x, (EB_U32)y, (EB_U32)N, 64, depth,
                                                na[0], na[1], na[2], na[3], r, EB_YUV420, EB_FALSE, EB_FALSE, EB_FALSE);
        EB_U32 mask = 0;
        if (t->luma_mode != SVTB200_INTRA_SKIP) mask |= PICTURE_BUFFER_DESC_LUMA_MASK;
        if (t->chroma_mode != SVTB200_INTRA_SKIP) mask |= PICTURE_BUFFER_DESC_CHROMA_MASK;
        EncodePassIntraPrediction16bit(r, (EB_U32)x + pd.originX, (EB_U32)y + pd.originY, (EB_U32)N, (EB_U32)N >> 1, &pd, EB_YUV420, EB_FALSE,
                                  t->luma_mode == SVTB200_INTRA_SKIP ? 0 : t->luma_mode, t->chroma_mode == SVTB200_INTRA_SKIP ? EB_INTRA_CHROMA_DM : t->chroma_mode, mask);
    }
    if (r->dctor) r->dctor(r);
    free(r);
    for (int k = 0; k < 4; ++k) { if (na[k]->dctor) na[k]->dctor(na[k]); free(na[k]); }
    return 0;
}


int svtref_ois_picture(const svtb200_ois_config *cfg, const svtb200_plane *src, const svtb200_me_pu_result *me_results, uint8_t *total, uint32_t *cand,
                       uint32_t cap, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 W = cfg->pic_width, H = cfg->pic_height, cols = (W + 63) / 64, rows = (H + 63) / 64, nLcu = cols * rows;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    PictureParentControlSet_t *pcs = (PictureParentControlSet_t *)calloc(1, sizeof(*pcs));
    EbObjectWrapper_t scsWrap;
    memset(&scsWrap, 0, sizeof(scsWrap));
    scs->lumaWidth = (EB_U16)W; scs->lumaHeight = (EB_U16)H; scs->lcuSize = 64;
    scs->pictureWidthInLcu = (EB_U8)cols; scs->pictureHeightInLcu = (EB_U8)rows;
    scs->inputResolution = (W * H >= 3840u * 2160u / 2) ? INPUT_SIZE_4K_RANGE : INPUT_SIZE_1080p_RANGE;
    if (LcuParamsInit(scs) != EB_ErrorNone) return -1;
    scsWrap.objectPtr = scs;
    pcs->sequenceControlSetWrapperPtr = &scsWrap;
    pcs->sliceType = cfg->slice_type == 2 ? EB_I_PICTURE : cfg->slice_type == 1 ? EB_P_PICTURE : EB_B_PICTURE;
    pcs->temporalLayerIndex = cfg->temporal_layer;
    pcs->limitOisToDcModeFlag = cfg->limit_ois_to_dc_mode;
    pcs->skipOis8x8 = cfg->skip_ois_8x8;
    pcs->cu8x8Mode = cfg->cu8x8_mode;
    pcs->meResults = (MeCuResults_t **)calloc(nLcu, sizeof(MeCuResults_t *));
    pcs->oisCu32Cu16Results = (OisCu32Cu16Results_t **)calloc(nLcu, sizeof(void *));
    pcs->oisCu8Results = (OisCu8Results_t **)calloc(nLcu, sizeof(void *));
    OisCandidate_t *pool = (OisCandidate_t *)calloc((size_t)nLcu * 85 * MAX_OIS_2, sizeof(OisCandidate_t));
    for (EB_U32 i = 0; i < nLcu; ++i) {
        pcs->meResults[i] = (MeCuResults_t *)calloc(85, sizeof(MeCuResults_t));
        if (me_results)
            for (int p = 0; p < 85; ++p) pcs->meResults[i][p].distortionDirection[0].distortion = me_results[(size_t)i * 85 + p].distortion[0];
        pcs->oisCu32Cu16Results[i] = (OisCu32Cu16Results_t *)calloc(1, sizeof(OisCu32Cu16Results_t));
        pcs->oisCu8Results[i] = (OisCu8Results_t *)calloc(1, sizeof(OisCu8Results_t));
        for (int c = 0; c < 21; ++c) pcs->oisCu32Cu16Results[i]->sortedOisCandidate[c] = pool + ((size_t)i * 85 + c) * MAX_OIS_2;
        for (int c = 0; c < 64; ++c) pcs->oisCu8Results[i]->sortedOisCandidate[c] = pool + ((size_t)i * 85 + 21 + c) * MAX_OIS_2;
    }
    MotionEstimationContext_t ctx;
    memset(&ctx, 0, sizeof(ctx));
    ctx.meContextPtr = (MeContext_t *)calloc(1, sizeof(MeContext_t));
    if (MeContextCtor(ctx.meContextPtr) != EB_ErrorNone) return -1;
    ctx.intraRefPtr = (IntraReferenceSamplesOpenLoop_t *)calloc(1, sizeof(IntraReferenceSamplesOpenLoop_t));
    if (IntraOpenLoopReferenceSamplesCtor(ctx.intraRefPtr) != EB_ErrorNone) return -1;
    ctx.oisKernelLevel = cfg->ois_kernel_level;
    ctx.oisThSet = cfg->ois_th_set;
    ctx.setBestOisDistortionToValid = cfg->set_best_ois_distortion_to_valid;

    EbPictureBufferDesc_t in;
    memset(&in, 0, sizeof(in));
    in.bufferY = src->buf; in.strideY = (EB_U16)src->stride; in.originX = src->origin_x; in.originY = src->origin_y;
    in.width = src->width; in.height = src->height; in.maxWidth = src->width; in.maxHeight = src->height; in.bitDepth = EB_8BIT;

    memset(total, 0, (size_t)nLcu * 85);
    memset(cand, 0, (size_t)nLcu * 85 * cap * sizeof(uint32_t));
    for (EB_U32 lcu = 0; lcu < nLcu; ++lcu) {
        OpenLoopIntraSearchLcu(pcs, lcu, &ctx, &in);
        for (int c = 1; c < 85; ++c) {
            total[(size_t)lcu * 85 + c] = c < 21 ? pcs->oisCu32Cu16Results[lcu]->totalIntraLumaMode[c] : pcs->oisCu8Results[lcu]->totalIntraLumaMode[c - 21];
            const OisCandidate_t *o = pool + ((size_t)lcu * 85 + c) * MAX_OIS_2;
            for (uint32_t k = 0; k < cap && k < MAX_OIS_2; ++k)
                cand[((size_t)lcu * 85 + c) * cap + k] = (uint32_t)o[k].distortion | ((uint32_t)o[k].validDistortion << 20) | ((uint32_t)o[k].intraMode << 24);
        }
    }
    for (EB_U32 i = 0; i < nLcu; ++i) { free(pcs->meResults[i]); free(pcs->oisCu32Cu16Results[i]); free(pcs->oisCu8Results[i]); }
    free(pcs->meResults); free(pcs->oisCu32Cu16Results); free(pcs->oisCu8Results); free(pool);
    if (ctx.intraRefPtr->dctor) ctx.intraRefPtr->dctor(ctx.intraRefPtr);
    free(ctx.intraRefPtr);
    if (ctx.meContextPtr->dctor) ctx.meContextPtr->dctor(ctx.meContextPtr);
    free(ctx.meContextPtr);
    free(scs->lcuParamsArray);
    free(pcs); free(scs);
    return 0;
}
