/*
 * oracle/refinc_saodec_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Path-level oracle for the SAO rate-distortion decision.  DetermineSaoLumaModeOffsets,
 * DetermineSaoChromaModeOffsets and TestSaoCopyModes are `static` in the reference's
 * Source/Lib/Codec/EbSampleAdaptiveOffsetGenerationDecision.c (:44, :182, :437), so this translation unit #includes that file WHERE IT
 * LIES (no copy) and calls them in the order SaoGenerationDecision does for saoMode = 1 (:736-767), LCU after LCU in raster order with
 * the previous decisions as left / top neighbours (what epSaoNeighborArray hands EncodePass, Codec/EbCodingLoop.c:4645-4670, one tile).
 * Built into its own shared object (oracle/_ref/libsvtref_saodec_*.so); every other symbol resolves from libsvtref_*.so.
 */
#include <stdlib.h>
#include <string.h>

#include "EbSampleAdaptiveOffsetGenerationDecision.c"        /* the reference source, compiled in place */

#include "../oracle/sao_decision.h"

int svtref_sao_decision_picture(const svtb200_sao_stats *stats, int lcu_cols, int lcu_rows, const sao_rate_tables *rates,
                                uint64_t full_lambda, uint64_t full_chroma_lambda_sao, int is10bit, sao_decision *out)
{
    const int n = lcu_cols * lcu_rows;
    MdRateEstimationContext_t *md = (MdRateEstimationContext_t *)calloc(1, sizeof(*md));
    SaoParameters_t *prm = (SaoParameters_t *)calloc((size_t)n, sizeof(*prm));
    EB_S32 bo_diff[3][SAO_BO_INTERVALS];
    EB_U16 bo_count[3][SAO_BO_INTERVALS];
    EB_S32 *bo_diff_p[3] = {bo_diff[0], bo_diff[1], bo_diff[2]};
    EB_U16 *bo_count_p[3] = {bo_count[0], bo_count[1], bo_count[2]};
    SaoStats_t st;
    if (!md || !prm) return -1;
    memcpy(md->saoMergeFlagBits, rates->merge_flag_bits, sizeof md->saoMergeFlagBits);
    memcpy(md->saoTypeIndexBits, rates->type_index_bits, sizeof md->saoTypeIndexBits);
    memcpy(md->saoOffsetTrunUnaryBits, rates->offset_trun_unary_bits, sizeof md->saoOffsetTrunUnaryBits);
    memset(&st, 0, sizeof st);
    st.boDiff = bo_diff_p; st.boCount = bo_count_p;
    for (int ly = 0; ly < lcu_rows; ++ly)
        for (int lx = 0; lx < lcu_cols; ++lx) {
            const int lcu = ly * lcu_cols + lx;
            for (int c = 0; c < 3; ++c) {
                const svtb200_sao_stats *s = stats + (size_t)lcu * 3 + c;
                memcpy(bo_diff[c], s->bo_diff, sizeof bo_diff[c]);
                memcpy(bo_count[c], s->bo_count, sizeof bo_count[c]);
                memcpy(st.eoDiff[c], s->eo_diff, sizeof st.eoDiff[c]);
                memcpy(st.eoCount[c], s->eo_count, sizeof st.eoCount[c]);
            }
            SaoParameters_t *p = prm + lcu;                      /* calloc'ed: the state SaoGenerationDecision :683-694 establishes */
            EB_S64 luma = 0, chroma = 0;
            DetermineSaoLumaModeOffsets(&st, p, full_lambda, md, &luma, (EB_BOOL)(is10bit != 0));
            DetermineSaoChromaModeOffsets(&st, p, full_chroma_lambda_sao, md, &chroma, (EB_BOOL)(is10bit != 0));
            TestSaoCopyModes(st.boDiff, st.boCount, st.eoDiff, st.eoCount, p, lx ? p - 1 : (SaoParameters_t *)EB_NULL,
                             ly ? p - lcu_cols : (SaoParameters_t *)EB_NULL, full_lambda, md, &luma, &chroma, (EB_BOOL)(is10bit != 0));
            sao_decision *o = out + lcu;
            memset(o, 0, sizeof *o);
            o->merge_left = p->saoMergeLeftFlag; o->merge_up = p->saoMergeUpFlag;
            o->type[0] = (uint8_t)p->saoTypeIndex[0]; o->type[1] = (uint8_t)p->saoTypeIndex[1];
            for (int c = 0; c < 3; ++c) {
                o->band_position[c] = (uint8_t)p->saoBandPosition[c];
                for (int k = 0; k < 4; ++k) o->offset[c][k] = p->saoOffset[c][k];
            }
            o->luma_cost = luma; o->chroma_cost = chroma;
        }
    free(prm); free(md);
    return 0;
}
