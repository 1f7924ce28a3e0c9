/*
 * oracle/refinc_meproc_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Pins the preset -> parameter derivation the product's host side restates (svt-hevc_b200/svtb200.py: me_config(), ois_cfg()).  The reference derives
 * those parameters per picture in four places; this translation unit calls them exactly as the encoder's kernels do:
 *   SignalDerivationPreAnalysisOq     static,  Source/Lib/Codec/EbResourceCoordinationProcess.c:239   (enableHme / level flags)  -> #include'd in place
 *   SignalDerivationMultiProcessesOq  extern,  Codec/EbPictureDecisionProcess.c:376  (useSubpelFlag, limitOisToDcModeFlag, cu8x8Mode, cu16x16Mode, skipOis8x8)
 *   SignalDerivationMeKernelOq        extern,  Codec/EbMotionEstimationProcess.c:308 (-> static SetMeHmeParamsOq :53: HME / ME search areas, oneQuadrantHME,
 *                                              fractionalSearchMethod / Model, fractionalSearch64x64, oisKernelLevel, oisThSet, updateHmeSearchCenter)
 *   lambda / mvdBits                  Codec/EbMotionEstimationProcess.c:684-704: lambdaModeDecisionRaSad{Base,RefNonBase,NonRef}[qp] and the initial
 *                                     MdRateEstimationContext_t::mvdBits of (sliceType, qp) built by EncodeCabacContextModelInit + MdRateEstimationContextInit
 *                                     (Codec/EbEncodeContext.c:138-143)
 * DeriveInputResolution (Codec/EbSequenceControlSet.c:288) classifies the picture size.  Built into its own shared object
 * (oracle/_ref/libsvtref_meproc_*.so); every other symbol resolves from libsvtref_*.so.
 */
#include <stdlib.h>
#include <string.h>

#include "EbResourceCoordinationProcess.c"        /* the reference source, compiled in place (for its static SignalDerivationPreAnalysisOq) */
#include "EbMotionEstimationProcess.h"
#include "EbMotionEstimationContext.h"
#include "EbLambdaRateTables.h"
#include "EbMdRateEstimation.h"
#include "EbCabacContextModel.h"

extern EB_ERRORTYPE SignalDerivationMultiProcessesOq(SequenceControlSet_t *sequenceControlSetPtr, PictureParentControlSet_t *pictureControlSetPtr);
extern EB_ERRORTYPE SignalDerivationMeKernelOq(SequenceControlSet_t *sequenceControlSetPtr, PictureParentControlSet_t *pictureControlSetPtr,
                                               MotionEstimationContext_t *contextPtr);
extern EB_ERRORTYPE DeriveInputResolution(SequenceControlSet_t *sequenceControlSetPtr, EB_U32 inputSize);

typedef struct svtref_me_derived {
    int32_t input_resolution;
    int32_t enable_hme, enable_hme_l0, enable_hme_l1, enable_hme_l2;
    int32_t use_subpel, limit_ois_to_dc, cu8x8_mode, cu16x16_mode, skip_ois_8x8;
    int32_t n_hme_regions_w, n_hme_regions_h;
    int32_t hme_l0_total_w, hme_l0_total_h, hme_l0_w[2], hme_l0_h[2], hme_l1_w[2], hme_l1_h[2], hme_l2_w[2], hme_l2_h[2];
    int32_t search_area_w, search_area_h;
    int32_t one_quadrant_hme, update_hme_search_center, fractional_search_method, fractional_search_model, fractional_search_64x64;
    int32_t ois_kernel_level, ois_th_set;
    uint32_t lambda;
    uint32_t mvd_bits[12];
} svtref_me_derived;

int svtref_derive_me(int width, int height, int enc_mode, int temporal_layer, int hierarchical_levels, int is_reference, int slice_type,
                     int frame_rate, int qp, svtref_me_derived *out)
{
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    PictureParentControlSet_t *pcs = (PictureParentControlSet_t *)calloc(1, sizeof(*pcs));
    MotionEstimationContext_t *ctx = (MotionEstimationContext_t *)calloc(1, sizeof(*ctx));
    MeContext_t *me = (MeContext_t *)calloc(1, sizeof(*me));
    ContextModelEncContext_t *cabac = (ContextModelEncContext_t *)malloc(sizeof(ContextModelEncContext_t) * TOTAL_NUMBER_OF_CABAC_CONTEXT_MODEL_BUFFERS);
    MdRateEstimationContext_t *md = (MdRateEstimationContext_t *)calloc(TOTAL_NUMBER_OF_MD_RATE_ESTIMATION_CASE_BUFFERS, sizeof(MdRateEstimationContext_t));
    if (!scs || !pcs || !ctx || !me || !cabac || !md || !out) return -1;
    memset(out, 0, sizeof *out);
    /* the encoder's defaults (Codec/EbEncHandle.c:2744-2748: enableHmeFlag, useDefaultMeHme, tune 1) */
    scs->lumaWidth = (EB_U16)width; scs->lumaHeight = (EB_U16)height; scs->lcuSize = 64;
    scs->staticConfig.frameRate = (EB_U32)frame_rate << 16;
    scs->staticConfig.tune = 1; scs->staticConfig.useDefaultMeHme = EB_TRUE; scs->staticConfig.enableHmeFlag = EB_TRUE;
    DeriveInputResolution(scs, (EB_U32)width * (EB_U32)height);
    pcs->encMode = (EB_U8)enc_mode; pcs->temporalLayerIndex = (EB_U8)temporal_layer; pcs->hierarchicalLevels = (EB_U8)hierarchical_levels;
    pcs->isUsedAsReferenceFlag = (EB_BOOL)(is_reference != 0); pcs->sliceType = (EB_PICTURE)slice_type; pcs->pictureQp = (EB_U8)qp;
    SignalDerivationPreAnalysisOq(scs, pcs);
    SignalDerivationMultiProcessesOq(scs, pcs);
    ctx->meContextPtr = me;
    SignalDerivationMeKernelOq(scs, pcs, ctx);
    EncodeCabacContextModelInit(cabac);
    MdRateEstimationContextInit(md, cabac);
    const MdRateEstimationContext_t *mdp = md + slice_type * TOTAL_NUMBER_OF_QP_VALUES + qp;

    out->input_resolution = scs->inputResolution;
    out->enable_hme = pcs->enableHmeFlag; out->enable_hme_l0 = pcs->enableHmeLevel0Flag; out->enable_hme_l1 = pcs->enableHmeLevel1Flag; out->enable_hme_l2 = pcs->enableHmeLevel2Flag;
    out->use_subpel = pcs->useSubpelFlag; out->limit_ois_to_dc = pcs->limitOisToDcModeFlag; out->cu8x8_mode = pcs->cu8x8Mode; out->cu16x16_mode = pcs->cu16x16Mode;
    out->skip_ois_8x8 = pcs->skipOis8x8;
    out->n_hme_regions_w = me->numberHmeSearchRegionInWidth; out->n_hme_regions_h = me->numberHmeSearchRegionInHeight;
    out->hme_l0_total_w = me->hmeLevel0TotalSearchAreaWidth; out->hme_l0_total_h = me->hmeLevel0TotalSearchAreaHeight;
    for (int i = 0; i < 2; ++i) {
        out->hme_l0_w[i] = me->hmeLevel0SearchAreaInWidthArray[i]; out->hme_l0_h[i] = me->hmeLevel0SearchAreaInHeightArray[i];
        out->hme_l1_w[i] = me->hmeLevel1SearchAreaInWidthArray[i]; out->hme_l1_h[i] = me->hmeLevel1SearchAreaInHeightArray[i];
        out->hme_l2_w[i] = me->hmeLevel2SearchAreaInWidthArray[i]; out->hme_l2_h[i] = me->hmeLevel2SearchAreaInHeightArray[i];
    }
    out->search_area_w = me->searchAreaWidth; out->search_area_h = me->searchAreaHeight;
    out->one_quadrant_hme = me->oneQuadrantHME; out->update_hme_search_center = me->updateHmeSearchCenter;
    out->fractional_search_method = me->fractionalSearchMethod; out->fractional_search_model = me->fractionalSearchModel;
    out->fractional_search_64x64 = me->fractionalSearch64x64;
    out->ois_kernel_level = ctx->oisKernelLevel; out->ois_th_set = ctx->oisThSet;
    out->lambda = temporal_layer == 0 ? lambdaModeDecisionRaSadBase[qp] : (is_reference ? lambdaModeDecisionRaSadRefNonBase[qp] : lambdaModeDecisionRaSadNonRef[qp]);
    for (int i = 0; i < NUMBER_OF_MVD_CASES && i < 12; ++i) out->mvd_bits[i] = mdp->mvdBits[i];
    free(md); free(cabac); free(me); free(ctx); free(pcs); free(scs);
    return 0;
}
