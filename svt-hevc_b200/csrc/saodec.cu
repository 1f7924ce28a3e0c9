// saodec.cu -- SAO rate-distortion decision on the device: removes the host round trip between the SAO statistics and
// the SAO application of a picture.
//
// Replaces: DetermineSaoLumaModeOffsets / DetermineSaoChromaModeOffsets / TestSaoCopyModes (static, Codec/EbSampleAdaptiveOffsetGenerationDecision.c:44,182,437)
// in the order SaoGenerationDecision :647-767 calls them for saoMode = 1 (presets <= M10), GetSaoOffsetsFractionBits Codec/EbMdRateEstimation.c:250.
//
// Design: the offsets / type of an LCU depend on that LCU's statistics only -> one warp per LCU, lanes over bands / (type, category) pairs
// (sao_offsets_kernel).  The merge test reads the FINAL parameters of the left and top neighbours -> anti-diagonal wavefront inside ONE CTA working out of
// shared memory (sao_merge_kernel: cols + rows - 1 steps, a barrier between steps).
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

// TestSaoCopyModes with the four merge distortions already summed (the group of lanes that owns the LCU computed them, see sao_merge_kernel)
__device__ void copy_modes(svtb200_sao_decision *p, const svtb200_sao_decision *left, const svtb200_sao_decision *up, int64_t dlL, int64_t dcL, int64_t dlU, int64_t dcU,
                           uint64_t lambda, const svtb200_sao_rate_tables *r, int is10bit)
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
        int64_t dl = dlL, dc = dcL;
        const int64_t rate = r->merge_flag_bits[1];
        dl >>= dshift; dc >>= dshift;
        lumaLeftCost = (dl << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        chromaLeftCost = (dc << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
        leftCost = (dl << COST_PRECISION) + (dc << COST_PRECISION) + (((rate * lambda) + MD_OFFSET) >> MD_SHIFT);
    }
    if (up) {
        int64_t dl = dlU, dc = dcU;
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
// one of the 12 terms of the distortion of re-using neighbour n's parameters on this LCU's statistics: term 0-3 luma category k, 4-7 Cb, 8-11 Cr
// (the reference sums -(2 * offset * diff) + count * offset^2 in 32 bits per term)
__device__ __forceinline__ int32_t merge_term(const svtb200_sao_stats *st, const svtb200_sao_decision *n, int term)
{
    const int c = term < 4 ? 0 : 1 + ((term - 4) >> 2), k = term & 3, ty = n->type[c ? 1 : 0];
    if (ty == 0) return 0;
    const int32_t o = n->offset[c][k];
    if (ty == 5) { const int b = n->band_position[c] + k; return -(2 * o * st[c].bo_diff[b]) + ((int32_t)st[c].bo_count[b] * SQR(o)); }
    return -(2 * o * st[c].eo_diff[ty - 1][k]) + ((int32_t)st[c].eo_count[ty - 1][k] * SQR(o));
}

// ---- offsets / type of an LCU: one WARP per LCU --------------------------------------------------------------------------------------------
// DetermineSaoLumaModeOffsets / DetermineSaoChromaModeOffsets / GetSaoOffsetsFractionBits with their arithmetic spread over the lanes: a lane per band for the band offsets (interval sums, rates and the
// first-minimum search by shuffles), a lane per (type, category) for the luma edge offsets, a lane per (component, type, category) for the chroma ones.
// Every expression keeps the reference's operand widths (32-bit products, 64-bit costs), so the results are the scalar path's bit for bit.
__device__ __forceinline__ int64_t shfl_down64(int64_t v, int d) { return (int64_t)__shfl_down_sync(0xffffffffu, (unsigned long long)v, d); }
__device__ __forceinline__ int64_t shfl_xor64(int64_t v, int d) { return (int64_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)v, d); }
__device__ __forceinline__ int64_t shfl64(int64_t v, int l) { return (int64_t)__shfl_sync(0xffffffffu, (unsigned long long)v, l); }
// (cost, index) of the first minimum among the lanes with valid set: smaller cost wins, equal costs keep the smaller index (the reference scans upwards with <)
__device__ __forceinline__ void first_min(int64_t &cost, int &idx, bool valid)
{
    if (!valid) { cost = INT64_MAX; idx = 1 << 20; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int64_t c2 = shfl_xor64(cost, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (c2 < cost || (c2 == cost && i2 < idx)) { cost = c2; idx = i2; }
    }
}
__device__ __forceinline__ int64_t rate_cost(int64_t rate, uint64_t lambda) { return (int64_t)(((rate * lambda) + MD_OFFSET) >> MD_SHIFT); }

__global__ void __launch_bounds__(128) sao_offsets_kernel(const svtb200_sao_stats *stats, int n_lcu, const __grid_constant__ svtb200_sao_rate_tables rates, uint64_t lambda,
                                                          uint64_t lambda_c, int is10bit, svtb200_sao_decision *out)
{
    const int lcu = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (lcu >= n_lcu) return;
    const svtb200_sao_stats *st = stats + (size_t)lcu * 3;
    const uint32_t dshift = is10bit ? 4 : 0;
    const int64_t huge = (int64_t)((~(uint64_t)0) >> 1);
    const int64_t offCostL = rate_cost(rates.type_index_bits[0], lambda), offCostC = rate_cost(rates.type_index_bits[0], lambda_c);
    // ---- luma band offsets: lane = band ----
    int32_t boOff;
    int64_t boDist;
    {
        const int32_t cnt = (int32_t)st[0].bo_count[lane], diff = st[0].bo_diff[lane];
        boOff = clip3(kMinBO[is10bit], kMaxBO[is10bit], cnt == 0 ? 0 : diff / cnt);
        boDist = (int64_t)((-(2 * boOff * diff) + (cnt * SQR(boOff))) >> dshift);
    }
    int64_t boCost;
    int bestBand;
    {
        int64_t d = boDist;
        uint32_t o = min(absu(boOff), 7u);
        int64_t rate = 163840 + (int64_t)rates.offset_trun_unary_bits[o] + (o != 0 ? 32768 : 0);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            d += shfl_down64(boDist, k);
            const uint32_t ok = min(absu(__shfl_down_sync(0xffffffffu, boOff, k)), 7u);
            rate += (int64_t)rates.offset_trun_unary_bits[ok] + (ok != 0 ? 32768 : 0);
        }
        boCost = (d << COST_PRECISION) + rate_cost(rate, lambda);
        bestBand = lane;
        first_min(boCost, bestBand, lane < 32 - 4 + 1);
    }
    // ---- luma edge offsets: lane = type * 4 + category (lanes 16-31 idle) ----
    int32_t eoOff = 0;
    int64_t eoCost;
    int bestEo;
    {
        const int t = (lane >> 2) & 3, k = lane & 3;
        const int32_t cnt = (int32_t)st[0].eo_count[t][k], diff = st[0].eo_diff[t][k];
        eoOff = clip3(kMinEO[is10bit][k], kMaxEO[is10bit][k], cnt == 0 ? 0 : diff / cnt);
        int64_t dist = (int64_t)((-(2 * eoOff * diff) + (cnt * SQR(eoOff))) >> dshift);
        int64_t rate = (int64_t)rates.offset_trun_unary_bits[min(absu(eoOff), 7u)];
        dist += shfl_xor64(dist, 1); dist += shfl_xor64(dist, 2);
        rate += shfl_xor64(rate, 1); rate += shfl_xor64(rate, 2);
        eoCost = (dist << COST_PRECISION) + rate_cost(rate, lambda);
        bestEo = t;
        first_min(eoCost, bestEo, lane < 16 && k == 0);
    }
    // ---- chroma edge offsets: lane = component * 16 + type * 4 + category ----
    int32_t ceOff;
    int64_t ceCost;
    int bestCe;
    {
        const int c = lane >> 4, t = (lane >> 2) & 3, k = lane & 3;
        const int32_t cnt = (int32_t)st[1 + c].eo_count[t][k], diff = st[1 + c].eo_diff[t][k];
        ceOff = clip3(kMinEO[is10bit][k], kMaxEO[is10bit][k], cnt == 0 ? 0 : diff / cnt);
        int64_t dist = (int64_t)((-(2 * ceOff * diff) + (cnt * SQR(ceOff))) >> dshift);
        int64_t rate = (int64_t)rates.offset_trun_unary_bits[min(absu(ceOff), 7u)];
        dist += shfl_xor64(dist, 1); dist += shfl_xor64(dist, 2);
        rate += shfl_xor64(rate, 1); rate += shfl_xor64(rate, 2);
        // the reference never resets the type's distortion between Cb and Cr: cost(t) = cost(D_cb, R_cb) + cost(D_cb + D_cr, R_cr)
        const int64_t dOther = shfl_xor64(dist, 16), rOther = shfl_xor64(rate, 16);
        const int64_t dCb = c ? dOther : dist, dCr = c ? dist : dOther, rCb = c ? rOther : rate, rCr = c ? rate : rOther;
        ceCost = (dCb << COST_PRECISION) + rate_cost(rCb, lambda_c) + ((dCb + dCr) << COST_PRECISION) + rate_cost(rCr, lambda_c);
        bestCe = t;
        first_min(ceCost, bestCe, lane < 16 && k == 0);
    }
    // ---- selection (DetermineSaoLumaModeOffsets / ...ChromaModeOffsets tails); the chosen offsets come from the lanes that hold them ----
    boCost += rate_cost(rates.type_index_bits[5], lambda);
    eoCost += rate_cost(rates.type_index_bits[bestEo + 1], lambda);
    ceCost += rate_cost(rates.type_index_bits[bestCe + 1], lambda_c);
    int32_t lumaBo[4], lumaEo[4], cb[4], cr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lumaBo[k] = __shfl_sync(0xffffffffu, boOff, (bestBand + k) & 31);
        lumaEo[k] = __shfl_sync(0xffffffffu, eoOff, (bestEo * 4 + k) & 31);
        cb[k] = __shfl_sync(0xffffffffu, ceOff, (bestCe * 4 + k) & 31);
        cr[k] = __shfl_sync(0xffffffffu, ceOff, (16 + bestCe * 4 + k) & 31);
    }
    if (lane == 0) {
        svtb200_sao_decision p;
        memset(&p, 0, sizeof p);                                                    // SaoGenerationDecision :683-694
        if (boCost < offCostL || eoCost < offCostL) {
            p.luma_cost = boCost < eoCost ? boCost : eoCost;
            if (p.luma_cost == boCost) {
                p.type[0] = 5; p.band_position[0] = (uint8_t)bestBand;
#pragma unroll
                for (int k = 0; k < 4; ++k) p.offset[0][k] = lumaBo[k];
            } else {
                p.type[0] = (uint8_t)(bestEo + 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) p.offset[0][k] = lumaEo[k];
            }
        } else { p.type[0] = 0; p.luma_cost = offCostL; }
        // chroma: band offsets are compiled out in the reference (their cost stays at the maximum)
        if (huge < offCostC || ceCost < offCostC) {
            p.chroma_cost = huge < ceCost ? huge : ceCost;
            if (p.chroma_cost == huge) p.type[1] = 5;                                // unreachable: band positions / offsets would be 0
            else {
                p.type[1] = (uint8_t)(bestCe + 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) { p.offset[1][k] = cb[k]; p.offset[2][k] = cr[k]; }
            }
        } else { p.type[1] = 0; p.chroma_cost = offCostC; }
        out[lcu] = p;
    }
}

// The merge step of LCU (x, y) needs the FINAL decisions of (x - 1, y) and (x, y - 1): both lie on the previous anti-diagonal.  One CTA walks the
// diagonals; everything a step touches sits in shared memory -- the decisions of the previous diagonal (indexed by row) and the statistics of the current
// one, fetched cooperatively one diagonal ahead -- so a step costs a barrier and a few hundred dependent instructions instead of dependent L2 round trips.
constexpr int SAOM_G = 8;                  // lanes per LCU in the merge step
constexpr int SAO_STAT_CHUNKS = 3 * (int)sizeof(svtb200_sao_stats) / 8;      // 117 eight-byte chunks per LCU (936 bytes)
constexpr int SAO_DEC_CHUNKS = (int)sizeof(svtb200_sao_decision) / 8;         // 5
static_assert(3 * sizeof(svtb200_sao_stats) % 8 == 0 && sizeof(svtb200_sao_decision) % 8 == 0, "8-byte asynchronous copies");
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// blockDim.x = SAOM_G * rows rounded up to whole warps: the LCU of row y0 + t / SAOM_G of the diagonal belongs to the SAOM_G lanes t, t + 1, ...: each
// computes 3 of the 24 distortion terms (12 per neighbour), the four sums are formed by shuffles, the group's first lane finishes TestSaoCopyModes
__global__ void __launch_bounds__(1024) sao_merge_kernel(const svtb200_sao_stats *stats, int cols, int rows, const __grid_constant__ svtb200_sao_rate_tables rates,
                                                         uint64_t lambda, int is10bit, svtb200_sao_decision *out, svtb200_sao_params *params)
{
    extern __shared__ __align__(16) uint8_t saom_smem[];
    // [3][rows] statistics of the current and the next two diagonals (fetched two steps ahead: a global -> shared copy takes longer than a step), [3][rows]
    // first-phase decisions of the same, [2][rows] final decisions of the previous / current diagonal
    uint8_t *s_stats = saom_smem;
    svtb200_sao_decision *s_in = (svtb200_sao_decision *)(s_stats + (size_t)3 * rows * SAO_STAT_CHUNKS * 8);
    svtb200_sao_decision *s_dec = s_in + 3 * rows;
    const int NT = blockDim.x;
    auto fetch = [&](int d) {      // statistics and first-phase decisions of diagonal d -> buffer d % 3, asynchronously (LDGSTS: no register round trip, all in flight)
        if (d < cols + rows - 1) {
        const int y0 = max(0, d - (cols - 1)), y1 = min(rows - 1, d), per = SAO_STAT_CHUNKS + SAO_DEC_CHUNKS;
        for (int i = threadIdx.x; i < (y1 - y0 + 1) * per; i += NT) {
            const int y = y0 + i / per, c = i - (y - y0) * per, lcu = y * cols + d - y;
            if (c < SAO_STAT_CHUNKS) cp_async8(s_stats + ((size_t)((d % 3) * rows + y) * SAO_STAT_CHUNKS + c) * 8, (const uint8_t *)(stats + (size_t)lcu * 3) + 8 * c);
            else cp_async8((uint8_t *)(s_in + (d % 3) * rows + y) + 8 * (c - SAO_STAT_CHUNKS), (const uint8_t *)(out + lcu) + 8 * (c - SAO_STAT_CHUNKS));
        }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");      // (an empty group past the last diagonal keeps the wait arithmetic uniform)
    };
    fetch(0);
    fetch(1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const int g = threadIdx.x & (SAOM_G - 1);
    for (int d = 0; d < cols + rows - 1; ++d) {
        const int y0 = max(0, d - (cols - 1)), y1 = min(rows - 1, d);
        fetch(d + 2);
        const int y = y0 + (int)(threadIdx.x / SAOM_G);
        const bool on = y <= y1;
        const int x = d - y;
        const svtb200_sao_decision *prev = s_dec + ((d + 1) & 1) * rows;
        const svtb200_sao_decision *left = on && x ? &prev[y] : nullptr, *up = on && y ? &prev[y - 1] : nullptr;
        const svtb200_sao_stats *st = (const svtb200_sao_stats *)(s_stats + (size_t)((d % 3) * rows + (on ? y : 0)) * SAO_STAT_CHUNKS * 8);
        int64_t sLL = 0, sLC = 0, sUL = 0, sUC = 0;      // left luma, left chroma, up luma, up chroma
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int t = g + SAOM_G * q, nb = t >= 12, term = t - 12 * nb;
            const svtb200_sao_decision *n = nb ? up : left;
            const int64_t v = n ? (int64_t)merge_term(st, n, term) : 0;
            const bool chroma = term >= 4;
            sLL += (!nb && !chroma) ? v : 0; sLC += (!nb && chroma) ? v : 0; sUL += (nb && !chroma) ? v : 0; sUC += (nb && chroma) ? v : 0;
        }
#pragma unroll
        for (int o = SAOM_G / 2; o > 0; o >>= 1) { sLL += shfl_xor64(sLL, o); sLC += shfl_xor64(sLC, o); sUL += shfl_xor64(sUL, o); sUC += shfl_xor64(sUC, o); }
        if (on && g == 0) {
            const int lcu = y * cols + x;
            svtb200_sao_decision p = s_in[(d % 3) * rows + y];
            copy_modes(&p, left, up, sLL, sLC, sUL, sUC, lambda, &rates, is10bit);
            s_dec[(d & 1) * rows + y] = p;
            out[lcu] = p;
            if (params) {
                svtb200_sao_params q;
                memset(&q, 0, sizeof q);
                q.type[0] = p.type[0]; q.type[1] = p.type[1];
                for (int c = 0; c < 3; ++c) { q.band_position[c] = p.band_position[c]; for (int k = 0; k < 4; ++k) q.offset[c][k] = (int8_t)p.offset[c][k]; }
                params[lcu] = q;
            }
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // everything but the newest group: diagonal d + 1 has landed
        __syncthreads();      // the next diagonal reads what this one wrote
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
    if (lcu_rows * SAOM_G > 1024) { set_error("sao_decision_picture_dev: more than %d LCU rows", 1024 / SAOM_G); return SVTB200_ERR_UNSUPPORTED; }
    sao_offsets_kernel<<<(n + 3) / 4, 128, 0, st>>>(stats_dev, n, *rates, full_lambda, full_chroma_lambda_sao, is10bit != 0, out_dev);
    SVTB200_COUNT_LAUNCH();
    const size_t smem = (size_t)lcu_rows * (3 * SAO_STAT_CHUNKS * 8 + 5 * sizeof(svtb200_sao_decision));
    // opt in to more than 48 KB of dynamic shared memory once per device and size (the attribute call is a context-level operation: not per launch)
    static std::atomic<size_t> opted[64];
    int dev = 0;
    SVTB200_CUDA_TRY(cudaGetDevice(&dev));
    if (smem > 48 * 1024 && opted[dev & 63].load(std::memory_order_relaxed) < smem) {
        SVTB200_CUDA_TRY(cudaFuncSetAttribute(sao_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        opted[dev & 63].store(smem, std::memory_order_relaxed);
    }
    sao_merge_kernel<<<1, (lcu_rows * SAOM_G + 31) / 32 * 32, smem, st>>>(stats_dev, lcu_cols, lcu_rows, *rates, full_lambda, is10bit != 0, out_dev, params_out_dev);
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
