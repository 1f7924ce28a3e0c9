/*
 * oracle/ref_me_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Path-level oracle for motion estimation: drives the reference's OWN, unmodified
 * MotionEstimateLcu (Source/Lib/Codec/EbMotionEstimation.c:3671) over every LCU of a picture,
 * reproducing what MotionEstimationKernel does around it
 * (Source/Lib/Codec/EbMotionEstimationProcess.c:710-781: LCU / quarter-LCU / sixteenth-LCU
 * buffer fills, lcuSrcPtr, hmeSearchType).  It is compiled against the reference headers where
 * they lie and linked into oracle/_ref/libsvtref_*.so by oracle/build_ref.sh.  Arguments use
 * the product's own ABI structs (include/svt_hevc_b200.h) so tests hand both sides the same data.
 */
#include <stdlib.h>
#include <string.h>

#include "EbDefinitions.h"
#include "EbPictureControlSet.h"
#include "EbSequenceControlSet.h"
#include "EbMotionEstimation.h"
#include "EbMotionEstimationContext.h"
#include "EbMotionEstimationLcuResults.h"
#include "EbReferenceObject.h"
#include "EbSystemResourceManager.h"
#include "EbPictureBufferDesc.h"

#include "svt_hevc_b200.h"

extern EB_U32 EbHevcGetCpuAsmType(void);

static void fill_desc(EbPictureBufferDesc_t *d, const svtb200_plane *p)
{
    memset(d, 0, sizeof(*d));
    d->bufferY = p->buf;
    d->strideY = (EB_U16)p->stride;
    d->originX = p->origin_x;
    d->originY = p->origin_y;
    d->width = p->width;
    d->height = p->height;
    d->maxWidth = p->width;
    d->maxHeight = p->height;
    d->bitDepth = EB_8BIT;
}

/* asm_mode: 0 = C rows (ASM_TYPES = 0), 1 = best SIMD the CPU supports (EbHevcGetCpuAsmType). */
/* LCU rows [row_begin, row_end): lets bench.py spread one picture over all host threads the way the reference's
 * own ME segments do (EbMotionEstimationProcess.c:679-683).  Re-entrant: every call builds its own contexts. */
int svtref_me_picture_rows(const svtb200_me_config *cfg, const svtb200_pa_picture *src,
                           const svtb200_pa_picture *ref_l0, const svtb200_pa_picture *ref_l1,
                           svtb200_me_pu_result *results, uint32_t *lcu_best_sad, uint32_t *lcu_best_mv,
                           uint32_t *rc_me_distortion, int asm_mode, int row_begin, int row_end)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;

    const EB_U32 W = cfg->pic_width, H = cfg->pic_height;
    const EB_U32 lcuCols = (W + 63) / 64, lcuRows = (H + 63) / 64, nLcu = lcuCols * lcuRows;

    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    PictureParentControlSet_t *pcs = (PictureParentControlSet_t *)calloc(1, sizeof(*pcs));
    EbObjectWrapper_t scsWrap, refWrap[2];
    EbPaReferenceObject_t *refObj[2];
    EbPictureBufferDesc_t srcDesc[3], refDesc[2][3];
    memset(&scsWrap, 0, sizeof(scsWrap));
    memset(refWrap, 0, sizeof(refWrap));

    scs->lumaWidth = (EB_U16)W;
    scs->lumaHeight = (EB_U16)H;
    scs->lcuSize = 64;
    scs->staticConfig.unrestrictedMotionVector = 1;
    scs->staticConfig.tileRowCount = 1;
    scs->staticConfig.tileColumnCount = 1;
    scs->staticConfig.rateControlMode = cfg->rate_control_mode;
    scsWrap.objectPtr = scs;

    fill_desc(&srcDesc[0], &src->full);
    fill_desc(&srcDesc[1], &src->quarter);
    fill_desc(&srcDesc[2], &src->sixteenth);
    const svtb200_pa_picture *refs[2] = { ref_l0, ref_l1 ? ref_l1 : ref_l0 };
    for (int l = 0; l < 2; ++l) {
        fill_desc(&refDesc[l][0], &refs[l]->full);
        fill_desc(&refDesc[l][1], &refs[l]->quarter);
        fill_desc(&refDesc[l][2], &refs[l]->sixteenth);
        refObj[l] = (EbPaReferenceObject_t *)calloc(1, sizeof(EbPaReferenceObject_t));
        refObj[l]->inputPaddedPicturePtr = &refDesc[l][0];
        refObj[l]->quarterDecimatedPicturePtr = &refDesc[l][1];
        refObj[l]->sixteenthDecimatedPicturePtr = &refDesc[l][2];
        refWrap[l].objectPtr = refObj[l];
        pcs->refPaPicPtrArray[l] = &refWrap[l];
        pcs->refPicPocArray[l] = cfg->ref_poc[l];
    }

    pcs->sequenceControlSetWrapperPtr = &scsWrap;
    pcs->sliceType = (cfg->num_lists == 2) ? EB_B_PICTURE : EB_P_PICTURE;
    pcs->temporalLayerIndex = cfg->temporal_layer;
    pcs->hierarchicalLevels = cfg->hierarchical_levels;
    pcs->enableHmeFlag = cfg->enable_hme;
    pcs->enableHmeLevel0Flag = cfg->enable_hme_l0;
    pcs->enableHmeLevel1Flag = cfg->enable_hme_l1;
    pcs->enableHmeLevel2Flag = cfg->enable_hme_l2;
    pcs->cu8x8Mode = cfg->cu8x8_mode;
    pcs->cu16x16Mode = cfg->cu16x16_mode;
    pcs->maxNumberOfPusPerLcu = SVTB200_ME_PU_COUNT;
    pcs->meResults = (MeCuResults_t **)calloc(nLcu, sizeof(MeCuResults_t *));
    for (EB_U32 i = 0; i < nLcu; ++i)
        pcs->meResults[i] = (MeCuResults_t *)calloc(SVTB200_ME_PU_COUNT, sizeof(MeCuResults_t));
    pcs->rcMEdistortion = (EB_U32 *)calloc(nLcu, sizeof(EB_U32));

    MeContext_t *me = (MeContext_t *)calloc(1, sizeof(*me));
    if (MeContextCtor(me) != EB_ErrorNone) return -1;
    me->updateHmeSearchCenter = cfg->update_hme_search_center;
    me->oneQuadrantHME = cfg->one_quadrant_hme;
    me->numberHmeSearchRegionInWidth = cfg->n_hme_regions_w;
    me->numberHmeSearchRegionInHeight = cfg->n_hme_regions_h;
    me->hmeLevel0TotalSearchAreaWidth = cfg->hme_l0_total_w;
    me->hmeLevel0TotalSearchAreaHeight = cfg->hme_l0_total_h;
    for (int i = 0; i < 2; ++i) {
        me->hmeLevel0SearchAreaInWidthArray[i] = cfg->hme_l0_w[i];
        me->hmeLevel0SearchAreaInHeightArray[i] = cfg->hme_l0_h[i];
        me->hmeLevel1SearchAreaInWidthArray[i] = cfg->hme_l1_w[i];
        me->hmeLevel1SearchAreaInHeightArray[i] = cfg->hme_l1_h[i];
        me->hmeLevel2SearchAreaInWidthArray[i] = cfg->hme_l2_w[i];
        me->hmeLevel2SearchAreaInHeightArray[i] = cfg->hme_l2_h[i];
    }
    me->searchAreaWidth = cfg->search_area_w;
    me->searchAreaHeight = cfg->search_area_h;
    me->fractionalSearchModel = cfg->fractional_search_model;
    me->fractionalSearchMethod = cfg->fractional_search_method;     /* same numbering: SUB_SAD_SEARCH 0, FULL_SAD_SEARCH 1, SSD_SEARCH 2 */
    me->fractionalSearch64x64 = EB_FALSE;
    me->lambda = cfg->lambda;
    for (int i = 0; i < 12; ++i) me->mvdBitsArray[i] = cfg->mvd_bits[i];

    const EbPictureBufferDesc_t *in = &srcDesc[0], *q = &srcDesc[1], *s = &srcDesc[2];
    for (EB_U32 ly = (EB_U32)row_begin; ly < lcuRows && ly < (EB_U32)row_end; ++ly) {
        for (EB_U32 lx = 0; lx < lcuCols; ++lx) {
            const EB_U32 lcuIndex = lx + ly * lcuCols;
            const EB_U32 ox = lx * 64, oy = ly * 64;
            const EB_U32 lw = (W - ox) < 64 ? W - ox : 64, lh = (H - oy) < 64 ? H - oy : 64;
            EB_U32 idx = (in->originY + oy) * in->strideY + in->originX + ox;
            me->hmeSearchType = HME_RECTANGULAR;
            for (EB_U32 r = 0; r < 64; ++r)
                memcpy(&me->lcuBuffer[r * 64], &in->bufferY[idx + r * in->strideY], 64);
            me->lcuSrcPtr = &in->bufferY[idx];
            me->lcuSrcStride = in->strideY;
            if (pcs->enableHmeLevel1Flag) {
                idx = (q->originY + (oy >> 1)) * q->strideY + q->originX + (ox >> 1);
                for (EB_U32 r = 0; r < (lh >> 1); ++r)
                    memcpy(&me->quarterLcuBuffer[r * me->quarterLcuBufferStride], &q->bufferY[idx + r * q->strideY], lw >> 1);
            }
            if (pcs->enableHmeLevel0Flag) {
                idx = (s->originY + (oy >> 2)) * s->strideY + s->originX + (ox >> 2);
                EB_U8 *fp = &s->bufferY[idx], *lp = me->sixteenthLcuBuffer;
                for (EB_U32 r = 0; r < (lh >> 2); r += 2) {
                    memcpy(lp, fp, lw >> 2);
                    lp += 16;
                    fp += s->strideY << 1;
                }
            }
            MotionEstimateLcu(pcs, lcuIndex, ox, oy, me, (EbPictureBufferDesc_t *)in);

            for (int p = 0; p < SVTB200_ME_PU_COUNT; ++p) {
                const MeCuResults_t *r = &pcs->meResults[lcuIndex][p];
                svtb200_me_pu_result *o = &results[(size_t)lcuIndex * SVTB200_ME_PU_COUNT + p];
                o->x_mv_l0 = r->xMvL0; o->y_mv_l0 = r->yMvL0;
                o->x_mv_l1 = r->xMvL1; o->y_mv_l1 = r->yMvL1;
                for (int c = 0; c < 3; ++c) {
                    o->distortion[c] = r->distortionDirection[c].distortion;
                    o->direction[c] = (uint8_t)r->distortionDirection[c].direction;
                }
                o->total_me_candidates = r->totalMeCandidateIndex;
            }
            for (int l = 0; l < 2; ++l) {
                if (lcu_best_sad)
                    memcpy(&lcu_best_sad[((size_t)lcuIndex * 2 + l) * 85], me->pLcuBestSad[l][0], 85 * sizeof(uint32_t));
                if (lcu_best_mv)
                    memcpy(&lcu_best_mv[((size_t)lcuIndex * 2 + l) * 85], me->pLcuBestMV[l][0], 85 * sizeof(uint32_t));
            }
            if (rc_me_distortion) rc_me_distortion[lcuIndex] = pcs->rcMEdistortion[lcuIndex];
        }
    }

    for (EB_U32 i = 0; i < nLcu; ++i) free(pcs->meResults[i]);
    free(pcs->meResults);
    free(pcs->rcMEdistortion);
    free(refObj[0]); free(refObj[1]);
    if (me->dctor) me->dctor(me);
    free(me); free(pcs); free(scs);
    return 0;
}

int svtref_me_picture(const svtb200_me_config *cfg, const svtb200_pa_picture *src,
                      const svtb200_pa_picture *ref_l0, const svtb200_pa_picture *ref_l1,
                      svtb200_me_pu_result *results, uint32_t *lcu_best_sad, uint32_t *lcu_best_mv,
                      uint32_t *rc_me_distortion, int asm_mode)
{
    return svtref_me_picture_rows(cfg, src, ref_l0, ref_l1, results, lcu_best_sad, lcu_best_mv, rc_me_distortion, asm_mode, 0, 1 << 30);
}
