// saodec.cu -- SAO rate-distortion decision on the device: removes the host round trip between the SAO statistics and
// the SAO application of a picture.
//
// Replaces: DetermineSaoLumaModeOffsets / DetermineSaoChromaModeOffsets / TestSaoCopyModes (static, Codec/EbSampleAdaptiveOffsetGenerationDecision.c:44,182,437)
// in the order SaoGenerationDecision :647-767 calls them for saoMode = 1 (presets <= M10), GetSaoOffsetsFractionBits Codec/EbMdRateEstimation.c:250.
//
// Design: the offsets / type of an LCU depend on that LCU's statistics only -> one thread per LCU (sao_offsets_kernel).  The merge test reads the FINAL
// parameters of the left and top neighbours -> anti-diagonal wavefront inside ONE CTA (sao_merge_kernel: cols + rows - 1 steps, a barrier between steps).
// Scalar 64-bit integer arithmetic with the reference's own conversions (see oracle/port/sao_decision_port.c for the quirks that are kept).
#include "common.cuh"

namespace svtb200 {

#define MD_SHIFT 23                      /* BIT_ESTIMATE_PRECISION 15 + LAMBDA_PRECISION 16 - COST_PRECISION 8, Codec/EbLambdaRateTables.h:27-29 */
#define MD_OFFSET (1 << (MD_SHIFT - 1))
#define COST_PRECISION 8
#define SQR(x) ((x) * (x))

__constant__ int32_t kMinEO[2][4] = {{0, 0, -7, -7}, {0, 0, -31, -31}};
__constant__ int32_t kMaxEO[2][4] = {{7, 7, 0, 0}, {31, 31, 0, 0}};
__constant__ int32_t kMaxBO[2] = {7, 31};
__constant__ int32_t kMinBO[2] = {-7, -31};

__device__ __forceinline__ int32_t clip3(int32_t lo, int32_t hi, int32_t v) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ uint32_t absu(int32_t v) { return (uint32_t)(v < 0 ? -v : v); }

__device__ uint64_t offsets_fraction_bits(uint32_t sao_type, const int32_t *offset, const svtb200_sao_rate_tables *r)
{
    uint64_t bits = 0;
    if (sao_type != 0) {
        uint32_t o[4];
        for (int k = 0; k < 4; ++k) { o[k] = absu(offset[k]); if (o[k] > 7) o[k] = 7; }
        if (sao_type == 5) {
            bits += 163840;                                     /* band position: 5 bits */
            for (int k = 0; k < 4; ++k) bits += r->offset_trun_unary_bits[o[k]];
            for (int k = 0; k < 4; ++k) bits += o[k] != 0 ? 32768 : 0;
        } else {
            for (int k = 0; k < 4; ++k) bits += r->offset_trun_unary_bits[o[k]];
        }
    }
    return bits;
}

__device__ void luma_mode_offsets(const svtb200_sao_stats *st /* component 0 */, svtb200_sao_decision *p, uint64_t lambda, const svtb200_sao_rate_tables *r, int is10bit)
{
    const uint64_t maxCost = ~(uint64_t)0;
    const uint32_t dshift = is10bit ? 4 : 0;
    int32_t boOffset[32];
    int64_t boIntervalDistortion[32], boBestCost = (int64_t)(maxCost >> 1), eoBestCost = (int64_t)(maxCost >> 1);
    uint32_t bestBand = 0, bestEo = 0;
    int32_t eoOffset[4][4];
    const int64_t saoOffRate = r->type_index_bits[0];
    const int64_t saoOffCost = (int64_t)(((saoOffRate * lambda) + MD_OFFSET) >> MD_SHIFT);

    for (int b = 0; b < 32; ++b) {
        boOffset[b] = st->bo_count[b] == 0 ? 0 : st->bo_diff[b] / (int32_t)st->bo_count[b];
        boOffset[b] = clip3(kMinBO[is10bit], kMaxBO[is10bit], boOffset[b]);
        boIntervalDistortion[b] = (-(2 * boOffset[b] * st->bo_diff[b]) + ((int32_t)st->bo_count[b] * SQR(boOffset[b]))) >> dshift;
    }
    for (uint32_t iv = 0; iv < 32 - 4 + 1; ++iv) {
        int64_t d = 0;
        for (uint32_t b = iv; b < iv + 4; ++b) d += boIntervalDistortion[b];
        const int64_t rate = (int64_t)offsets_fraction_bits(5, &boOffset[iv], r);
        const int64_t cost = (d << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        if (cost < boBestCost) { boBestCost = cost; bestBand = iv; }
    }
    for (uint32_t t = 0; t < 4; ++t) {
        int64_t dist = 0;
        for (uint32_t k = 0; k < 4; ++k) {
            eoOffset[t][k] = st->eo_count[t][k] == 0 ? 0 : st->eo_diff[t][k] / (int32_t)st->eo_count[t][k];
            eoOffset[t][k] = clip3(kMinEO[is10bit][k], kMaxEO[is10bit][k], eoOffset[t][k]);
            dist += (int64_t)((-(2 * eoOffset[t][k] * st->eo_diff[t][k]) + ((int32_t)st->eo_count[t][k] * SQR(eoOffset[t][k]))) >> dshift);
        }
        const int64_t rate = (int64_t)offsets_fraction_bits(t + 1, eoOffset[t], r);
        const int64_t cost = (dist << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        if (cost < eoBestCost) { eoBestCost = cost; bestEo = t; }
    }
    boBestCost += (((r->type_index_bits[5] * lambda) + MD_OFFSET) >> MD_SHIFT);
    eoBestCost += (((r->type_index_bits[bestEo + 1] * lambda) + MD_OFFSET) >> MD_SHIFT);
    if (boBestCost < saoOffCost || eoBestCost < saoOffCost) {
        p->luma_cost = boBestCost < eoBestCost ? boBestCost : eoBestCost;
        if (p->luma_cost == boBestCost) {
            p->type[0] = 5; p->band_position[0] = (uint8_t)bestBand;
            for (int k = 0; k < 4; ++k) p->offset[0][k] = boOffset[bestBand + k];
        } else if (p->luma_cost == eoBestCost) {
            p->type[0] = (uint8_t)(bestEo + 1);
            for (int k = 0; k < 4; ++k) p->offset[0][k] = eoOffset[bestEo][k];
        }
    } else {
        p->type[0] = 0;
        p->luma_cost = saoOffCost;
    }
}

__device__ void chroma_mode_offsets(const svtb200_sao_stats *st /* components 0..2 of the LCU */, svtb200_sao_decision *p, uint64_t lambda, const svtb200_sao_rate_tables *r, int is10bit)
{
    const uint64_t maxCost = ~(uint64_t)0;
    const uint32_t dshift = is10bit ? 4 : 0;
    const int64_t boChromaBestCost = (int64_t)(maxCost >> 1);          /* BO for chroma is compiled out in the reference */
    int64_t eoChromaBestCost = (int64_t)(maxCost >> 1);
    uint32_t bestEo = 1;
    int32_t eoOffset[2][4][4];
    int64_t eoTypeDistortion[4] = {0, 0, 0, 0}, eoTypeCost[4] = {0, 0, 0, 0};
    const int64_t saoOffRate = r->type_index_bits[0];
    const int64_t saoOffCost = (int64_t)(((saoOffRate * lambda) + MD_OFFSET) >> MD_SHIFT);

    for (uint32_t t = 0; t < 4; ++t) {
        for (uint32_t c = 1; c < 3; ++c) {
            const svtb200_sao_stats *s = st + c;
            for (uint32_t k = 0; k < 4; ++k) {
                int32_t o = s->eo_count[t][k] == 0 ? 0 : s->eo_diff[t][k] / (int32_t)s->eo_count[t][k];
                o = clip3(kMinEO[is10bit][k], kMaxEO[is10bit][k], o);
                eoOffset[c - 1][t][k] = o;
                eoTypeDistortion[t] += (int64_t)((-(2 * o * s->eo_diff[t][k]) + ((int32_t)s->eo_count[t][k] * SQR(o))) >> dshift);   /* never reset between Cb and Cr */
            }
            const int64_t rate = (int64_t)offsets_fraction_bits(t + 1, eoOffset[c - 1][t], r);
            eoTypeCost[t] += (eoTypeDistortion[t] << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        }
        if (eoTypeCost[t] < eoChromaBestCost) { eoChromaBestCost = eoTypeCost[t]; bestEo = t; }
    }
    eoChromaBestCost += (((r->type_index_bits[bestEo + 1] * lambda) + MD_OFFSET) >> MD_SHIFT);
    if (boChromaBestCost < saoOffCost || eoChromaBestCost < saoOffCost) {
        p->chroma_cost = boChromaBestCost < eoChromaBestCost ? boChromaBestCost : eoChromaBestCost;
        if (p->chroma_cost == boChromaBestCost) {
            p->type[1] = 5;                                             /* unreachable: band positions / offsets would be 0 */
            for (int c = 1; c < 3; ++c) { p->band_position[c] = 0; for (int k = 0; k < 4; ++k) p->offset[c][k] = 0; }
        } else {
            p->type[1] = (uint8_t)(bestEo + 1);
            for (int c = 1; c < 3; ++c) for (int k = 0; k < 4; ++k) p->offset[c][k] = eoOffset[c - 1][bestEo][k];
        }
    } else {
        p->type[1] = 0;
        p->chroma_cost = saoOffCost;
    }
}

/* distortion of re-using a neighbour's parameters on this LCU's statistics */
__device__ void merge_distortion(const svtb200_sao_stats *st, const svtb200_sao_decision *n, int64_t *luma, int64_t *chroma)
{
    *luma = 0; *chroma = 0;
    if (n->type[0] > 0) {
        for (int k = 0; k < 4; ++k) {
            if (n->type[0] == 5)
                *luma += -(2 * n->offset[0][k] * st[0].bo_diff[n->band_position[0] + k]) + ((int32_t)st[0].bo_count[n->band_position[0] + k] * SQR(n->offset[0][k]));
            else
                *luma += -(2 * n->offset[0][k] * st[0].eo_diff[n->type[0] - 1][k]) + ((int32_t)st[0].eo_count[n->type[0] - 1][k] * SQR(n->offset[0][k]));
        }
    }
    if (n->type[1] > 0) {
        for (int c = 1; c < 3; ++c)
            for (int k = 0; k < 4; ++k) {
                if (n->type[1] == 5)
                    *chroma += -(2 * n->offset[c][k] * st[c].bo_diff[n->band_position[c] + k]) + ((int32_t)st[c].bo_count[n->band_position[c] + k] * SQR(n->offset[c][k]));
                else
                    *chroma += -(2 * n->offset[c][k] * st[c].eo_diff[n->type[1] - 1][k]) + ((int32_t)st[c].eo_count[n->type[1] - 1][k] * SQR(n->offset[c][k]));
            }
    }
}

__device__ void copy_modes(const svtb200_sao_stats *st, svtb200_sao_decision *p, const svtb200_sao_decision *left, const svtb200_sao_decision *up, uint64_t lambda,
                       const svtb200_sao_rate_tables *r, int is10bit)
{
    const uint64_t maxCost = ~(uint64_t)0;
    const uint32_t dshift = is10bit ? 4 : 0;
    int64_t lumaLeftCost = 0, chromaLeftCost = 0, leftCost = (int64_t)(maxCost >> 1);
    int64_t lumaUpCost = 0, chromaUpCost = 0, upCost = (int64_t)(maxCost >> 1);
    const uint64_t leftFlagBits = left ? r->merge_flag_bits[0] : 0;
    uint64_t upFlagBits = up ? r->merge_flag_bits[0] : 0;
    int64_t best = p->luma_cost + p->chroma_cost + ((((leftFlagBits + upFlagBits) * lambda) + MD_OFFSET) >> MD_SHIFT);
    p->luma_cost = p->luma_cost + ((((leftFlagBits + upFlagBits) * lambda) + MD_OFFSET) >> MD_SHIFT);
    p->chroma_cost = p->chroma_cost + ((((leftFlagBits + upFlagBits) * lambda) + MD_OFFSET) >> MD_SHIFT);
    if (left) {
        int64_t dl, dc;
        merge_distortion(st, left, &dl, &dc);
        const int64_t rate = r->merge_flag_bits[1];
        dl >>= dshift; dc >>= dshift;
        lumaLeftCost = (dl << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        chromaLeftCost = (dc << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        leftCost = (dl << COST_PRECISION) + (dc << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
    }
    if (up) {
        int64_t dl, dc;
        merge_distortion(st, up, &dl, &dc);
        dl >>= dshift; dc >>= dshift;
        upFlagBits = r->merge_flag_bits[1];
        const int64_t rate = (int64_t)(leftFlagBits + upFlagBits);
        lumaUpCost = (dl << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        chromaUpCost = (dc << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        upCost = (dl << COST_PRECISION) + (dc << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
    }
    if (leftCost < best || upCost < best) {
        best = leftCost < upCost ? leftCost : upCost;
        const svtb200_sao_decision *n = NULL;
        if (best == leftCost && left) { p->merge_left = 1; n = left; p->luma_cost = lumaLeftCost; p->chroma_cost = chromaLeftCost; }
        else if (best == upCost && up) { p->merge_up = 1; n = up; p->luma_cost = lumaUpCost; p->chroma_cost = chromaUpCost; }
        if (n) {
            p->type[0] = n->type[0]; p->type[1] = n->type[1];
            memcpy(p->band_position, n->band_position, sizeof p->band_position);
            memcpy(p->offset, n->offset, sizeof p->offset);
        }
    }
}


__global__ void __launch_bounds__(128) sao_offsets_kernel(const svtb200_sao_stats *stats, int n_lcu, svtb200_sao_rate_tables rates, uint64_t lambda, uint64_t lambda_c,
                                                          int is10bit, svtb200_sao_decision *out)
{
    const int lcu = blockIdx.x * 128 + threadIdx.x;
    if (lcu >= n_lcu) return;
    svtb200_sao_decision p;
    memset(&p, 0, sizeof p);                                                    // SaoGenerationDecision :683-694
    luma_mode_offsets(stats + (size_t)lcu * 3, &p, lambda, &rates, is10bit);
    chroma_mode_offsets(stats + (size_t)lcu * 3, &p, lambda_c, &rates, is10bit);
    out[lcu] = p;
}

__global__ void __launch_bounds__(256) sao_merge_kernel(const svtb200_sao_stats *stats, int cols, int rows, svtb200_sao_rate_tables rates, uint64_t lambda, int is10bit,
                                                        svtb200_sao_decision *out, svtb200_sao_params *params)
{
    for (int d = 0; d < cols + rows - 1; ++d) {
        const int y0 = max(0, d - (cols - 1)), y1 = min(rows - 1, d);
        for (int y = y0 + threadIdx.x; y <= y1; y += 256) {
            const int x = d - y, lcu = y * cols + x;
            svtb200_sao_decision p = out[lcu];
            svtb200_sao_decision l, u;
            if (x) l = out[lcu - 1];
            if (y) u = out[lcu - cols];
            copy_modes(stats + (size_t)lcu * 3, &p, x ? &l : nullptr, y ? &u : nullptr, lambda, &rates, is10bit);
            out[lcu] = p;
            if (params) {
                svtb200_sao_params q;
                memset(&q, 0, sizeof q);
                q.type[0] = p.type[0]; q.type[1] = p.type[1];
                for (int c = 0; c < 3; ++c) { q.band_position[c] = p.band_position[c]; for (int k = 0; k < 4; ++k) q.offset[c][k] = (int8_t)p.offset[c][k]; }
                params[lcu] = q;
            }
        }
        __syncthreads();      // the next diagonal reads what this one wrote (same CTA: the barrier orders the global accesses)
    }
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

int svtb200_sao_decision_picture_dev(const svtb200_sao_stats *stats_dev, int lcu_cols, int lcu_rows, const svtb200_sao_rate_tables *rates, uint64_t full_lambda,
                                     uint64_t full_chroma_lambda_sao, int is10bit, svtb200_sao_decision *out_dev, svtb200_sao_params *params_out_dev, void *cuda_stream)
{
    if (!stats_dev || !rates || !out_dev || lcu_cols <= 0 || lcu_rows <= 0) { set_error("sao_decision_picture_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    const int n = lcu_cols * lcu_rows;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    sao_offsets_kernel<<<(n + 127) / 128, 128, 0, st>>>(stats_dev, n, *rates, full_lambda, full_chroma_lambda_sao, is10bit != 0, out_dev);
    SVTB200_COUNT_LAUNCH();
    sao_merge_kernel<<<1, 256, 0, st>>>(stats_dev, lcu_cols, lcu_rows, *rates, full_lambda, is10bit != 0, out_dev, params_out_dev);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

int svtb200_sao_decision_picture(const svtb200_sao_stats *stats, int lcu_cols, int lcu_rows, const svtb200_sao_rate_tables *rates, uint64_t full_lambda,
                                 uint64_t full_chroma_lambda_sao, int is10bit, svtb200_sao_decision *out)
{
    if (!stats || !rates || !out || lcu_cols <= 0 || lcu_rows <= 0) { set_error("sao_decision_picture: bad argument"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t n = (size_t)lcu_cols * lcu_rows, sb = n * 3 * sizeof(svtb200_sao_stats), ob = n * sizeof(svtb200_sao_decision);
    void *d_st, *d_out;
    if ((rc = c->dev_reserve(sb + ob + 1024)) || (rc = c->dev_alloc(sb, &d_st)) || (rc = c->dev_alloc(ob, &d_out))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_st, stats, sb, cudaMemcpyHostToDevice, c->stream));
    if ((rc = svtb200_sao_decision_picture_dev((const svtb200_sao_stats *)d_st, lcu_cols, lcu_rows, rates, full_lambda, full_chroma_lambda_sao, is10bit,
                                               (svtb200_sao_decision *)d_out, nullptr, c->stream))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

}  // extern "C"
