/*
 * oracle/sao_decision.h -- TEST INFRASTRUCTURE ONLY.  Interface shared by the plain-C restatement (oracle/port/sao_decision_port.c)
 * and the reference harness (oracle/refinc_saodec_harness.c) of the SAO rate-distortion decision, SURVEY §8 row f2 / next-row N3:
 * SaoGenerationDecision with saoMode = 1 (every preset <= M10), Codec/EbSampleAdaptiveOffsetGenerationDecision.c:647-767.
 * Not part of the product ABI yet: the decision still runs on the host.
 */
#ifndef ORACLE_SAO_DECISION_H
#define ORACLE_SAO_DECISION_H
#include <stdint.h>
#include "svt_hevc_b200.h"

/* the three CABAC bit-estimate tables the decision reads (MdRateEstimationContext_t, Codec/EbMdRateEstimation.h:148-150), 1 bit = 32768 */
typedef struct sao_rate_tables {
    uint32_t merge_flag_bits[2];
    uint32_t type_index_bits[6];
    uint32_t offset_trun_unary_bits[8];
} sao_rate_tables;

/* per LCU: SaoParameters_t (Codec/EbCodingUnit.h:137-146) + the two costs SaoGenerationDecision returns */
typedef struct sao_decision {
    uint8_t  merge_left, merge_up;
    uint8_t  type[2];              /* saoTypeIndex[luma, chroma] */
    uint8_t  band_position[3];
    uint8_t  reserved;
    int32_t  offset[3][4];
    int64_t  luma_cost, chroma_cost;
} sao_decision;

/* stats[(lcu * 3 + component)]; LCUs in raster order, left / top neighbours = the previous decisions (single tile).  Returns 0. */
int port_sao_decision_picture(const svtb200_sao_stats *stats, int lcu_cols, int lcu_rows, const sao_rate_tables *rates,
                              uint64_t full_lambda, uint64_t full_chroma_lambda_sao, int is10bit, sao_decision *out);
int svtref_sao_decision_picture(const svtb200_sao_stats *stats, int lcu_cols, int lcu_rows, const sao_rate_tables *rates,
                                uint64_t full_lambda, uint64_t full_chroma_lambda_sao, int is10bit, sao_decision *out);
#endif
