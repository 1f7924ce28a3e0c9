// intra.cu -- HEVC intra prediction: mode kernels, encode-pass batches, open-loop intra search (OIS) of a whole picture.  8-bit 4:2:0.
//
// Replaces:
//   IntraModePlanar / DCLuma / DCChroma / Vertical{Luma,Chroma} / Horizontal{Luma,Chroma} / Angular_{2,18,34} /
//     Angular_{Vertical,Horizontal}_Kernel                       C_DEFAULT/EbIntraPrediction_C.c:15-1015 (+ SSE2/SSSE3/SSE4.1/AVX2 rows)
//   IntraModeAngular_all and its 3To9 / 11To17 / 19To25 / 27To33 projections            Codec/EbIntraPrediction.c:3080-3600
//   GenerateIntraReferenceSamplesEncodePass :212, EncodePassIntraPrediction :4395
//   UpdateNeighborSamplesArrayOpenLoop :5222, IntraPredictionOpenLoop :5339
//   OpenLoopIntraSearchLcu, IntraOpenLoopSearchTheseModesOutputBest, InjectIntraCandidatesBasedOnBestMode[Islice], GetOisPoint,
//     SortIntraModesOpenLoop, SortOisCandidateOpenLoop, OpenLoopIntraDC                  Codec/EbMotionEstimation.c:4458-5330
//
// Design.  A block's reference line (4N+1 samples, bottom-left -> top-left -> top-right, the order of
// IntraReferenceSamples_t::yIntraReferenceArray) lives in shared memory; every predicted sample is evaluated directly from it
// (the negative-angle projection is computed per access instead of materialising refAbove / refLeft), 4 horizontally adjacent
// samples per thread, so no scratch prediction buffers and no mode-specific kernels exist.
//   encode pass: one CTA per job; substitution of unavailable neighbours is a bit scan over the per-unit availability word
//                (no serial walk); HBM sees the 4N+1 neighbours in and the block out.
//   OIS:         one CTA per LCU; the source LCU and its 32-sample apron are staged once in shared memory, warps take CUs from a
//                shared counter, build the CU's line once and run all its modes with VABSDIFF4 SADs against the staged source; the
//                reference's candidate injection / sorting then runs on the SAD table.  Neighbours come from the source picture,
//                so all LCUs of a picture are independent.
#include <type_traits>

#include "common.cuh"

namespace svtb200 {

__constant__ int8_t c_ang[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};                   // intraModeAngularTable, EbIntraPrediction.c:25
__constant__ int16_t c_inv_ang[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};   // invIntraModeAngularTable :48
__constant__ int8_t c_filt_thr[4] = {35, 7, 1, 0};                                   // intraLumaFilterTable :60

struct ModeSetup {
    int mode, N, lg;
    bool edge, vertical;
    int ang, inv, dc, maxv;
};

__device__ __forceinline__ ModeSetup mode_setup(int mode, int N, bool luma, int dc, int maxv = 255)
{
    ModeSetup m;
    m.maxv = maxv;
    m.mode = mode; m.N = N; m.lg = 31 - __clz(N);
    m.edge = luma && N < 32;
    m.vertical = mode >= 18;
    const int d = m.vertical ? mode - 26 : 10 - mode, ad = d < 0 ? -d : d;
    m.ang = (mode >= 2) ? (d < 0 ? -(int)c_ang[ad] : (int)c_ang[ad]) : 0;
    m.inv = (mode >= 2) ? (int)c_inv_ang[ad] : 0;
    m.dc = dc;
    return m;
}

// sample k of the main reference: 0 = top-left, > 0 the main side, < 0 projected from the other side
template <typename LT>
__device__ __forceinline__ int main_ref(const LT *tl, bool vertical, int k, int inv)
{
    if (k >= 0) return vertical ? tl[k] : tl[-k];
    const int j = (128 - k * inv) >> 8;
    return vertical ? tl[-j] : tl[j];
}

template <typename LT>
__device__ __forceinline__ int intra_pel(const LT *line, const ModeSetup &m, int x, int y)
{
    const int N = m.N;
    const LT *tl = line + 2 * N;
    if (m.mode == 0)
        return ((N - 1 - x) * (int)tl[-1 - y] + (x + 1) * (int)tl[1 + N] + (N - 1 - y) * (int)tl[1 + x] + (y + 1) * (int)tl[-1 - N] + N) >> (m.lg + 1);
    if (m.mode == 1) {
        if (m.edge) {
            if ((x | y) == 0) return ((int)tl[-1] + (int)tl[1] + 2 * m.dc + 2) >> 2;
            if (y == 0) return ((int)tl[1 + x] + 3 * m.dc + 2) >> 2;
            if (x == 0) return ((int)tl[-1 - y] + 3 * m.dc + 2) >> 2;
        }
        return m.dc;
    }
    if (m.mode == 26) { const int t = tl[1 + x]; return (m.edge && x == 0) ? min(max(t + (((int)tl[-1 - y] - (int)tl[0]) >> 1), 0), m.maxv) : t; }
    if (m.mode == 10) { const int l = tl[-1 - y]; return (m.edge && y == 0) ? min(max(l + (((int)tl[1 + x] - (int)tl[0]) >> 1), 0), m.maxv) : l; }
    const int along = m.vertical ? x : y, across = m.vertical ? y : x;
    const int delta = (across + 1) * m.ang, di = delta >> 5, df = delta & 31;
    const int a = main_ref(tl, m.vertical, along + di + 1, m.inv);
    if (df == 0) return a;
    const int b = main_ref(tl, m.vertical, along + di + 2, m.inv);
    return ((32 - df) * a + df * b + 16) >> 5;
}

__device__ __forceinline__ uint32_t intra_pel4(const uint8_t *line, const ModeSetup &m, int x, int y)
{
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) w |= (uint32_t)intra_pel(line, m, x + k, y) << (8 * k);
    return w;
}

// DC value of a line, computed by one full warp
template <typename LT>
__device__ __forceinline__ int warp_dc(const LT *line, int N)
{
    const int lane = threadIdx.x & 31;
    uint32_t s = 0;
    for (int i = lane; i < 2 * N; i += 32) s += (i < N) ? line[2 * N + 1 + i] : line[2 * N - 1 - (i - N)];
    s = warp_sum(s);
    return (int)((s + N) >> (32 - __clz(N)));     // log2(2N) = lg + 1
}

// ---- leaf -------------------------------------------------------------------------------------------------------------------
constexpr int IN_NT = 128;

__global__ void __launch_bounds__(IN_NT) intra_leaf_kernel(int mode, int N, int chroma, const uint8_t *line_g, uint8_t *pred, int stride)
{
    __shared__ uint8_t s_line[4 * 32 + 4];
    __shared__ int s_dc;
    for (int i = threadIdx.x; i < 4 * N + 1; i += IN_NT) s_line[i] = line_g[i];
    __syncthreads();
    if (threadIdx.x < 32) { const int dc = warp_dc(s_line, N); if (threadIdx.x == 0) s_dc = dc; }
    __syncthreads();
    const ModeSetup m = mode_setup(mode, N, !chroma, s_dc);
    for (int i = threadIdx.x; i < N * N; i += IN_NT) pred[(i / N) * stride + (i % N)] = (uint8_t)intra_pel(s_line, m, i % N, i / N);
}

// ---- encode pass ------------------------------------------------------------------------------------------------------------
struct PlaneW { uint8_t *p; int stride; };   // p points at picture sample (0,0)
struct IntraArgs {
    PlaneW recon[3], pred[3];
    const svtb200_intra_job *jobs;
    int n_jobs;
};

// Reference line of one block from a reconstructed plane.  unit = samples per availability bit (4 luma, 2 chroma).
// Scan order of the units: left column bottom-up (nu units), top-left (1), top row left to right (nu units).
template <bool HBD>
__device__ void build_line(typename std::conditional<HBD, uint16_t, uint8_t>::type *line, const PlaneW &pl, int x, int y, int N, int unit, uint32_t left_avail, uint32_t top_avail, bool tl_avail)
{
    using PIX = typename std::conditional<HBD, uint16_t, uint8_t>::type;
    const int nu = 2 * N / unit, lgu = unit == 4 ? 2 : 1;
    const uint64_t M = (uint64_t)(left_avail & ((1u << nu) - 1)) | ((uint64_t)(tl_avail ? 1 : 0) << nu) | ((uint64_t)(top_avail & ((1u << nu) - 1)) << (nu + 1));
    const PIX *o = (const PIX *)pl.p + (ptrdiff_t)y * pl.stride + x;
    for (int i = threadIdx.x; i < 4 * N + 1; i += IN_NT) {
        // this sample's unit and its position inside the unit (both in scan order)
        int u, k;
        if (i < 2 * N) { u = i >> lgu; k = i & (unit - 1); }
        else if (i == 2 * N) { u = nu; k = 0; }
        else { u = nu + 1 + ((i - 2 * N - 1) >> lgu); k = (i - 2 * N - 1) & (unit - 1); }
        int v = HBD ? 512 : 128;
        if (M) {
            const uint64_t below = M & ((2ull << u) - 1);
            int su, sk;                                   // source unit / sample within it
            if (below == 0) { su = __ffsll((long long)M) - 1; sk = 0; }                  // nothing before: first available sample overall
            else { su = 63 - __clzll((long long)below); sk = su == u ? k : (su == nu ? 0 : unit - 1); }   // own sample, or the last one of the nearest unit before
            ptrdiff_t off;
            if (su < nu) off = (ptrdiff_t)(2 * N - 1 - (su * unit + sk)) * pl.stride - 1;
            else if (su == nu) off = -(ptrdiff_t)pl.stride - 1;
            else off = -(ptrdiff_t)pl.stride + (su - nu - 1) * unit + sk;
            v = o[off];
        }
        line[i] = (PIX)v;
    }
}

// [1 2 1] smoothing, or the bilinear "strong" smoothing of 32x32 blocks (GenerateIntraReferenceSamplesEncodePass part 3)
template <typename LT>
__device__ void filter_line(const LT *in, LT *out, int N, bool strong_enabled, int thr)
{
    const int bl = in[0], tl = in[2 * N], tr = in[4 * N];
    const bool strong = strong_enabled && N >= 32 && abs(bl + tl - 2 * (int)in[N]) < thr && abs(tl + tr - 2 * (int)in[3 * N]) < thr;
    const int sh = 32 - __clz(N);
    for (int i = threadIdx.x; i < 4 * N + 1; i += IN_NT) {
        int v;
        if (i == 0 || i == 4 * N || (strong && i == 2 * N)) v = in[i];
        else if (strong) v = i < 2 * N ? ((2 * N - i) * bl + i * tl + N) >> sh : ((4 * N - i) * tl + (i - 2 * N) * tr + N) >> sh;
        else v = ((int)in[i - 1] + 2 * (int)in[i] + (int)in[i + 1] + 2) >> 2;
        out[i] = (LT)v;
    }
}

__device__ void predict_block16(const uint16_t *line, const ModeSetup &m, const PlaneW &dst, int x, int y)
{
    const int N = m.N;
    uint16_t *d = (uint16_t *)dst.p + (ptrdiff_t)y * dst.stride + x;
    for (int i = threadIdx.x; i < N * N; i += IN_NT) { const int r = i / N, c = i - r * N; d[(ptrdiff_t)r * dst.stride + c] = (uint16_t)intra_pel(line, m, c, r); }
}

__device__ void predict_block(const uint8_t *line, const ModeSetup &m, const PlaneW &dst, int x, int y)
{
    const int N = m.N;
    uint8_t *d = dst.p + (ptrdiff_t)y * dst.stride + x;
    if ((((uintptr_t)d | (uintptr_t)dst.stride) & 3) == 0) {
        const int q = N >> 2;
        for (int i = threadIdx.x; i < N * q; i += IN_NT) {
            const int r = i / q, c = (i - r * q) * 4;
            *(uint32_t *)(d + (ptrdiff_t)r * dst.stride + c) = intra_pel4(line, m, c, r);
        }
    } else {
        for (int i = threadIdx.x; i < N * N; i += IN_NT) { const int r = i / N, c = i - r * N; d[(ptrdiff_t)r * dst.stride + c] = (uint8_t)intra_pel(line, m, c, r); }
    }
}

template <bool HBD>
__global__ void __launch_bounds__(IN_NT) intra_batch_kernel(const __grid_constant__ IntraArgs A)
{
    using LT = typename std::conditional<HBD, uint16_t, uint8_t>::type;
    constexpr int MAXV = HBD ? 1023 : 255, THR = HBD ? 32 : 8;      // SMOOTHING_THRESHOLD[_10BIT], Codec/EbIntraPrediction.h:119-120
    __shared__ __align__(4) LT s_line[3][4 * 32 + 4], s_filt[4 * 32 + 4];
    __shared__ int s_dc[3];
    const svtb200_intra_job t = A.jobs[blockIdx.x];
    const bool do_luma = t.luma_mode != SVTB200_INTRA_SKIP, do_chroma = t.chroma_mode != SVTB200_INTRA_SKIP;
    const bool tl = t.flags & SVTB200_INTRA_TOPLEFT_AVAIL;
    const int N = t.size, Nc = N >> 1;
    // all neighbour reads first (the prediction may be written in place)
    if (do_luma) build_line<HBD>(s_line[0], A.recon[0], t.x, t.y, N, 4, t.left_avail, t.top_avail, tl);
    if (do_chroma) {
        build_line<HBD>(s_line[1], A.recon[1], t.x >> 1, t.y >> 1, Nc, 2, t.left_avail, t.top_avail, tl);
        build_line<HBD>(s_line[2], A.recon[2], t.x >> 1, t.y >> 1, Nc, 2, t.left_avail, t.top_avail, tl);
    }
    __syncthreads();
    const int lmode = t.luma_mode;
    int cmode = lmode;
    switch (t.chroma_mode) { case 0: cmode = 0; break; case 1: cmode = 26; break; case 2: cmode = 10; break; case 3: cmode = 1; break; default: break; }
    const LT *luma_line = s_line[0];
    if (do_luma) {
        const int dmin = min(abs(lmode - 10), abs(lmode - 26));
        if (dmin > c_filt_thr[31 - __clz(N) - 2] && lmode != 1) { filter_line<LT>(s_line[0], s_filt, N, t.flags & SVTB200_INTRA_STRONG_SMOOTHING, THR); luma_line = s_filt; }
    }
    const int w = threadIdx.x >> 5;
    if (w == 0 && do_luma && lmode == 1) { const int dc = warp_dc(s_line[0], N); if ((threadIdx.x & 31) == 0) s_dc[0] = dc; }
    if ((w == 1 || w == 2) && do_chroma && cmode == 1) { const int dc = warp_dc(s_line[w], Nc); if ((threadIdx.x & 31) == 0) s_dc[w] = dc; }
    __syncthreads();
    if (HBD) {
        if (do_luma) predict_block16((const uint16_t *)luma_line, mode_setup(lmode, N, true, s_dc[0], MAXV), A.pred[0], t.x, t.y);
        if (do_chroma) {
            predict_block16((const uint16_t *)s_line[1], mode_setup(cmode, Nc, false, s_dc[1], MAXV), A.pred[1], t.x >> 1, t.y >> 1);
            predict_block16((const uint16_t *)s_line[2], mode_setup(cmode, Nc, false, s_dc[2], MAXV), A.pred[2], t.x >> 1, t.y >> 1);
        }
    } else {
        if (do_luma) predict_block((const uint8_t *)luma_line, mode_setup(lmode, N, true, s_dc[0]), A.pred[0], t.x, t.y);
        if (do_chroma) {
            predict_block((const uint8_t *)s_line[1], mode_setup(cmode, Nc, false, s_dc[1]), A.pred[1], t.x >> 1, t.y >> 1);
            predict_block((const uint8_t *)s_line[2], mode_setup(cmode, Nc, false, s_dc[2]), A.pred[2], t.x >> 1, t.y >> 1);
        }
    }
}

// ---- open-loop intra search -------------------------------------------------------------------------------------------------
#ifndef OIS_NT_V
#define OIS_NT_V 256
#endif
constexpr int OIS_NT = OIS_NT_V, OIS_NW = OIS_NT / 32;
constexpr int SRC_P = 128, SRC_R = 97;      // staged source: rows / cols -1 .. 95 relative to the LCU origin; pitch = win_pitch(97)
constexpr int SAD_SLOTS = 36;

struct OisArgs {
    svtb200_ois_config cfg;
    const uint8_t *src;        // picture sample (0,0)
    int stride;
    const svtb200_me_pu_result *me;
    uint8_t *total;
    uint32_t *cand;
    int cap, lcu_cols;
    int padded;               // the plane has >= 16 columns of padding left and right and a row above: stage with 128-bit loads
};

__constant__ uint8_t c_islice_modes[7] = {0, 1, 10, 26, 2, 18, 34};           // iSliceModesArray, EbMotionEstimation.c:148
__constant__ uint8_t c_stage1_modes[9] = {10, 26, 2, 18, 34, 6, 14, 22, 30};  // stage1ModesArray :149
__constant__ int16_t c_ois_th[3][6][4] = {                                    // EbHevcOisPointTh :26
    {{-20, 50, 150, 200}, {-20, 50, 150, 200}, {-20, 50, 100, 150}, {-20, 50, 200, 300}, {-20, 50, 200, 300}, {-20, 50, 200, 300}},
    {{-150, 0, 150, 200}, {-150, 0, 150, 200}, {-125, 0, 100, 150}, {-50, 50, 200, 300}, {-50, 50, 200, 300}, {-50, 50, 200, 300}},
    {{-400, -300, -200, 0}, {-400, -300, -200, 0}, {-400, -300, -200, 0}, {-400, -300, -200, 0}, {-400, -300, -200, 0}, {-400, -300, -200, 0}}};
// InjectIntraCandidatesBasedOnBestMode (:4562): [0] = index into stage1SadArray, [1..9] = injected modes; rows: best = 2, 10, 18, 26, 34, 6, 14, 22, other
__constant__ uint8_t c_inject[9][10] = {{2, 2, 1, 0, 3, 4, 5, 7, 8, 9},        {0, 10, 1, 0, 9, 11, 8, 12, 7, 13},    {3, 18, 1, 0, 17, 19, 16, 20, 15, 21},
                                        {1, 26, 1, 0, 25, 27, 24, 28, 23, 29}, {4, 34, 1, 0, 33, 32, 29, 31, 27, 28}, {5, 6, 1, 0, 7, 5, 4, 8, 3, 9},
                                        {6, 14, 1, 0, 13, 15, 12, 16, 11, 17}, {7, 22, 1, 0, 21, 23, 20, 24, 19, 25}, {8, 30, 1, 0, 29, 31, 28, 32, 27, 33}};
__constant__ uint8_t c_islice_inject[5][3] = {{2, 4, 6}, {10, 6, 14}, {18, 14, 22}, {26, 22, 30}, {34, 32, 30}};

__device__ __forceinline__ uint32_t ois_pack(uint32_t dist, bool valid, uint32_t mode) { return (dist & 0xFFFFFu) | ((uint32_t)valid << 20) | (mode << 24); }

__device__ __forceinline__ void cu_geometry(int cu, int &N, int &cx, int &cy)
{
    int k, lp;      // lp = log2(CUs per LCU row)
    if (cu < 5) { N = 32; k = cu - 1; lp = 1; } else if (cu < 21) { N = 16; k = cu - 5; lp = 2; } else { N = 8; k = cu - 21; lp = 3; }
    cx = (k & ((1 << lp) - 1)) * N; cy = (k >> lp) * N;
}

// SAD of one predicted mode against the staged source, one warp; lanes take 4 horizontally adjacent samples each
__device__ __forceinline__ uint32_t warp_mode_sad(const uint8_t *line, const ModeSetup &m, const uint8_t *s_src, int cx, int cy)   // s_src: sample (-1, -1) of the LCU
{
    const int N = m.N, q = N >> 2, lq = 31 - __clz(q), lane = threadIdx.x & 31;      // N is 8, 16 or 32
    uint32_t acc = 0;
    for (int i = lane; i < N * q; i += 32) {
        const int r = i >> lq, c = (i & (q - 1)) * 4;
        const uintptr_t sp = (uintptr_t)(s_src + (cy + r + 1) * SRC_P + cx + c + 1);   // not 4-byte aligned (the apron column): two words + funnel shift
        const uint32_t *sw2 = (const uint32_t *)(sp & ~(uintptr_t)3);
        const uint32_t sw = __funnelshift_r(sw2[0], sw2[1], (uint32_t)(sp & 3) * 8);
        acc = vsad4_acc(intra_pel4(line, m, c, r), sw, acc);
    }
    return warp_sum(acc);
}

// how many of the nine stage-1 modes the inter-picture search evaluates for a CU: 0 when the DC distortion is close enough to the ME distortion (or
// either is 0), else 1 + 2 * point with point from the thresholds of the picture's set and layer (OpenLoopIntraSearchLcu, EbMotionEstimation.c:5117-5160)
__device__ __forceinline__ int ois_stage1_count(const svtb200_ois_config &cfg, uint32_t dc_sad, uint32_t me_sad)
{
    const int32_t diff = (int32_t)(me_sad - dc_sad) * 100;
    const int32_t dist = dc_sad ? diff / (int32_t)dc_sad : 0;
    const int16_t *th = c_ois_th[cfg.ois_th_set][cfg.temporal_layer];
    int point = 4;
    if (dc_sad == 0 || me_sad == 0 || dist <= th[0]) point = 0;
    else if (dist <= th[1]) point = 1;
    else if (dist <= th[2]) point = 2;
    else if (dist <= th[3]) point = 3;
    return point == 0 ? 0 : 1 + 2 * point;
}

__global__ void __launch_bounds__(OIS_NT) ois_picture_kernel(const __grid_constant__ OisArgs A)
{
    __shared__ __align__(16) uint8_t s_raw[SRC_R * SRC_P];
    __shared__ uint8_t s_line[OIS_NW][4 * 32 + 4];
    __shared__ uint32_t s_sad[85][SAD_SLOTS];
    __shared__ uint32_t s_sort[84][37];            // [CU][18 distortions, 18 modes]; odd pitch: a warp's CUs fall into different banks
    __shared__ int s_next;
    const svtb200_ois_config &cfg = A.cfg;
    const int W = cfg.pic_width, H = cfg.pic_height;
    const int lcu = blockIdx.x, ox = (lcu % A.lcu_cols) * 64, oy = (lcu / A.lcu_cols) * 64;
    const bool is_i = cfg.slice_type == 2;
    const int max_cu = is_i ? 84 : ((cfg.skip_ois_8x8 || cfg.cu8x8_mode == 1) ? 20 : 84);
    // which mode list this picture runs per CU: 0 I slice, 1 all 35 (kernel level), 2 DC only, 3 DC + the nine stage-1 modes
    const int path = is_i ? 0 : cfg.ois_kernel_level ? 1 : cfg.limit_ois_to_dc_mode ? 2 : 3;

    int a0 = 0;
    if (A.padded) {
        // only samples inside the picture are ever used (the line builder tests the picture bounds), so stop at its edge
        a0 = load_window_t<OIS_NT>(s_raw, SRC_P, A.src + (ptrdiff_t)(oy - 1) * A.stride + ox - 1, A.stride, min(SRC_R, W - ox + 1), min(SRC_R, H - oy + 1));
    } else {
        for (int i = threadIdx.x; i < SRC_R * SRC_R; i += OIS_NT) {
            const int r = i / SRC_R, c = i - r * SRC_R, gx = ox + c - 1, gy = oy + r - 1;
            s_raw[r * SRC_P + c] = (gx >= 0 && gy >= 0 && gx < W && gy < H) ? __ldg(A.src + (ptrdiff_t)gy * A.stride + gx) : 128;
        }
    }
    const uint8_t *s_src = s_raw + a0;
    if (threadIdx.x == 0) s_next = 1;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *line = s_line[warp];
    for (;;) {
        int cu = 0;
        if (lane == 0) cu = atomicAdd(&s_next, 1);
        cu = __shfl_sync(0xffffffffu, cu, 0);
        if (cu > max_cu) break;
        int N, cx, cy;
        cu_geometry(cu, N, cx, cy);
        const int x = ox + cx, y = oy + cy;
        if (x + N > W || y + N > H) continue;                       // rasterScanCuValidity
        // UpdateNeighborSamplesArrayOpenLoop: source neighbours, 128 where the picture ends (no substitution)
        for (int i = lane; i < 4 * N + 1; i += 32) {
            int v = 128;
            if (i < 2 * N) { const int r = 2 * N - 1 - i; if (x != 0 && y + r < H) v = s_src[(cy + r + 1) * SRC_P + cx]; }
            else if (i == 2 * N) { if (x != 0 && y != 0) v = s_src[cy * SRC_P + cx]; }
            else { const int c = i - 2 * N - 1; if (y != 0 && x + c < W) v = s_src[cy * SRC_P + cx + c + 1]; }
            line[i] = (uint8_t)v;
        }
        __syncwarp();
        const int dc = warp_dc(line, N);
        int n_modes;
        if (path == 0) n_modes = N == 32 ? 1 : 7; else if (path == 1) n_modes = 35; else if (path == 2) n_modes = 1; else n_modes = 10;
        for (int k = 0; k < n_modes; ++k) {
            int mode;
            if (path == 0) mode = c_islice_modes[k]; else if (path == 1) mode = k; else if (path == 2) mode = 1; else mode = k == 0 ? 1 : c_stage1_modes[k - 1];
            const uint32_t sad = warp_mode_sad(line, mode_setup(mode, N, true, dc), s_src, cx, cy);
            if (lane == 0) s_sad[cu][k] = sad;
            // the "OIS point" path evaluates only the 1 + 2 * point stage-1 modes its DC-vs-ME test asks for (OpenLoopIntraSearchLcu :5117-5160): most CUs
            // of an inter picture stop at the DC
            if (path == 3 && k == 0) n_modes = 1 + ois_stage1_count(cfg, sad, A.me[(size_t)lcu * 85 + cu].distortion[0]);
        }
        __syncwarp();
    }
    __syncthreads();

    uint8_t *total = A.total + (size_t)lcu * 85;
    uint32_t *cand = A.cand + (size_t)lcu * 85 * A.cap;
    // every slot of this LCU is defined: zero first, then the reference's writes
    for (int i = threadIdx.x; i < 85; i += OIS_NT) total[i] = 0;
    for (int i = threadIdx.x; i < 85 * A.cap; i += OIS_NT) cand[i] = 0;
    __syncthreads();

    if (path == 2 || path == 1) {
        for (int cu = 1 + threadIdx.x; cu <= max_cu; cu += OIS_NT) {
            int N, cx, cy;
            cu_geometry(cu, N, cx, cy);
            if (ox + cx + N > W || oy + cy + N > H) continue;
            uint32_t *c = cand + (size_t)cu * A.cap;
            if (path == 2) { c[0] = ois_pack(s_sad[cu][0], true, 1); total[cu] = 1; continue; }    // OpenLoopIntraDC :4964
            // SortIntraModesOpenLoop :4871 (keep the best 18 of 35) then SortOisCandidateOpenLoop :4843
            uint32_t *d = s_sort[cu - 1], *md = d + 18;      // per-CU scratch in shared memory (dynamically indexed: registers would spill to local memory)
            for (int m = 0; m < 18; ++m) { d[m] = s_sad[cu][m]; md[m] = m; }
            for (int m = 18; m < 35; ++m) {
                const uint32_t sad = s_sad[cu][m];
                uint32_t worst = d[0]; int wi = 0;
                for (int q = 1; q < 18; ++q) if (d[q] > worst) { worst = d[q]; wi = q; }
                if (sad < worst) { d[wi] = sad; md[wi] = m; }
            }
            for (int a = 0; a < 18; ++a)
                for (int b = a; b < 18; ++b)
                    if (d[a] > d[b]) { const uint32_t td = d[a], tm = md[a]; d[a] = d[b]; md[a] = md[b]; d[b] = td; md[b] = tm; }
            for (int m = 0; m < 18; ++m) c[m] = ois_pack(d[m], false, md[m]);
            total[cu] = 18;
        }
        return;
    }
    if (threadIdx.x != 0) return;
    // I slices and the inter/intra "OIS point" path: stage1SadArray and bestMode live across the CUs of the LCU (:5069,5073), so the
    // reference's loop is replayed in order on the SAD table
    uint32_t *stage1 = s_sort[0], *modes = s_sort[1];      // (the sort scratch is free on these paths)
    for (int k = 0; k < 11; ++k) stage1[k] = 0;
    int best_mode = 0;
    for (int cu = 1; cu <= max_cu; ++cu) {
        int N, cx, cy;
        cu_geometry(cu, N, cx, cy);
        if (ox + cx + N > W || oy + cy + N > H) continue;
        uint32_t *c = cand + (size_t)cu * A.cap;
        if (path == 0) {
            if (N == 32) { c[0] = ois_pack(s_sad[cu][0], true, 0); continue; }
            uint32_t best_sad = 32 * 32 * 255;
            for (int m = 0; m < 7; ++m) {
                stage1[m] = s_sad[cu][m];
                if (stage1[m] < best_sad) { best_mode = c_islice_modes[m]; best_sad = stage1[m]; }
            }
            int cnt = 2;
            modes[0] = 0; modes[1] = 1; modes[2] = modes[3] = modes[4] = 0;
            if (best_mode != 0 && best_mode != 1) {
                const int r = best_mode == 2 ? 0 : best_mode == 10 ? 1 : best_mode == 18 ? 2 : best_mode == 26 ? 3 : 4;
                for (int q = 0; q < 3; ++q) modes[cnt++] = c_islice_inject[r][q];
            }
            c[0] = ois_pack(stage1[0], true, 0);
            for (int q = 1; q < cnt; ++q) c[q] = ois_pack(0, false, modes[q]);
            total[cu] = (uint8_t)cnt;
            continue;
        }
        const uint32_t me_sad = A.me[(size_t)lcu * 85 + cu].distortion[0];
        stage1[0] = s_sad[cu][0];
        const int n = ois_stage1_count(cfg, stage1[0], me_sad);
        if (n == 0) { c[0] = ois_pack(stage1[0], false, 1); total[cu] = 1; continue; }
        uint32_t best_sad = 32 * 32 * 255;
        for (int m = 0; m < n; ++m) {
            stage1[m] = s_sad[cu][1 + m];
            if (stage1[m] < best_sad) { best_mode = c_stage1_modes[m]; best_sad = stage1[m]; }
        }
        int row;
        switch (best_mode) { case 2: row = 0; break; case 10: row = 1; break; case 18: row = 2; break; case 26: row = 3; break; case 34: row = 4; break;
                             case 6: row = 5; break; case 14: row = 6; break; case 22: row = 7; break; default: row = 8; break; }
        c[0] = ois_pack(stage1[c_inject[row][0]], cfg.set_best_ois_distortion_to_valid != 0, c_inject[row][1]);
        for (int q = 1; q < 9; ++q) c[q] = ois_pack(0, false, c_inject[row][1 + q]);
        total[cu] = (uint8_t)n;
    }
}

static size_t pbytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }
static PlaneW plane_w(const svtb200_plane &p, uint8_t *base, size_t bps = 1) { PlaneW w; w.p = base + ((size_t)p.origin_y * p.stride + p.origin_x) * bps; w.stride = (int)p.stride; return w; }

static int intra_check(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n, bool host_jobs)
{
    if (!recon || !pred || (!jobs && n)) { set_error("intra_batch: null argument"); return SVTB200_ERR_BAD_ARG; }
    const svtb200_plane *a[3] = {&recon->y, &recon->cb, &recon->cr}, *b[3] = {&pred->y, &pred->cb, &pred->cr};
    for (int k = 0; k < 3; ++k)
        if (!a[k]->buf || !b[k]->buf || a[k]->width != b[k]->width || a[k]->height != b[k]->height) { set_error("intra_batch: recon / pred planes missing or of different size"); return SVTB200_ERR_BAD_ARG; }
    if (recon->y.origin_x < 1 || recon->y.origin_y < 1 || recon->cb.origin_x < 1 || recon->cb.origin_y < 1) { set_error("intra_batch: recon needs at least one sample of padding (neighbour addresses are formed before the availability test)"); return SVTB200_ERR_BAD_ARG; }
    if (!host_jobs) return SVTB200_OK;
    for (uint32_t j = 0; j < n; ++j) {
        const svtb200_intra_job &t = jobs[j];
        const bool lu = t.luma_mode != SVTB200_INTRA_SKIP, ch = t.chroma_mode != SVTB200_INTRA_SKIP;
        const bool ok = (t.size == 4 || t.size == 8 || t.size == 16 || t.size == 32) && (!lu || t.luma_mode <= 34) && (!ch || t.chroma_mode <= 4) && !(ch && t.size == 4) &&
                        !(ch && !lu && t.chroma_mode == 4) && (t.x % t.size) == 0 && (t.y % t.size) == 0 && t.x + t.size <= recon->y.width && t.y + t.size <= recon->y.height;
        if (!ok) { set_error("intra_batch: job %u malformed (size 4/8/16/32, modes 0..34 / 0..4, aligned and inside the picture, no chroma at size 4, DM needs a luma mode)", j); return SVTB200_ERR_BAD_ARG; }
    }
    return SVTB200_OK;
}

static int intra_launch(const IntraArgs &A, cudaStream_t st, bool hbd = false)
{
    if (A.n_jobs == 0) return SVTB200_OK;
    if (hbd) intra_batch_kernel<true><<<A.n_jobs, IN_NT, 0, st>>>(A); else intra_batch_kernel<false><<<A.n_jobs, IN_NT, 0, st>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

// ---- DC-only open-loop search (OpenLoopIntraDC, EbMotionEstimation.c:4964): the path every non-I picture of M10 takes ---------------------------------
// (limitOisToDcModeFlag, EbPictureDecisionProcess.c:417-422; 8x8 CUs skipped).  One CTA of four warps per LCU, a warp per 32x32 quadrant, a lane per
// row: the row's eight source words feed the 32x32 CU (flat DC) and the two 16x16 CUs it crosses (DC with the [1 3] edge filter on their first row /
// column, IntraModeDCLuma), neighbour sums and SADs reduced with shuffles.  Same availability rule as UpdateNeighborSamplesArrayOpenLoop (:5222):
// neighbours outside the picture read 128, CUs crossing the picture edge are skipped.  Needs a padded plane (interior 4-byte aligned, >= 1 sample around).
constexpr int DCK_P = 80;      // tile pitch: byte 15 = column -1, bytes 16..79 = columns 0..63; row 0 = picture row -1
__global__ void __launch_bounds__(128) ois_dc_kernel(const __grid_constant__ OisArgs A)
{
    __shared__ __align__(16) uint8_t s_t[65 * DCK_P];
    const int W = A.cfg.pic_width, H = A.cfg.pic_height;
    const int lcu = blockIdx.x, ox = (lcu % A.lcu_cols) * 64, oy = (lcu / A.lcu_cols) * 64;
    // stage rows -1..63: 16 interior words + the left apron byte per row (the plane is padded: every address exists)
    for (int i = threadIdx.x; i < 65 * 17; i += 128) {
        const int r = i / 17, c = i - r * 17;
        const uint8_t *g = A.src + (ptrdiff_t)(oy + r - 1) * A.stride + ox;
        if (c < 16) ((uint32_t *)(s_t + r * DCK_P + 16))[c] = __ldg((const uint32_t *)g + c);
        else s_t[r * DCK_P + 15] = __ldg(g - 1);
    }
    uint8_t *total = A.total + (size_t)lcu * 85;
    uint32_t *cand = A.cand + (size_t)lcu * 85 * A.cap;
    for (int i = threadIdx.x; i < 85; i += 128) total[i] = 0;           // every slot of this LCU is defined: zero first, then the reference's writes
    for (int i = threadIdx.x; i < 85 * A.cap; i += 128) cand[i] = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cx = (w & 1) * 32, cy = (w >> 1) * 32, x = ox + cx, y = oy + cy;
    const uint8_t *rowp = s_t + (cy + lane + 1) * DCK_P + 16 + cx;
    const uint4 wa = *(const uint4 *)rowp, wb = *(const uint4 *)(rowp + 16);
    const uint32_t sw[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    // ---- 32x32 CU ----
    {
        const int top = (y != 0 && x + lane < W) ? s_t[cy * DCK_P + 16 + cx + lane] : 128;
        const int left = (x != 0 && y + lane < H) ? s_t[(cy + lane + 1) * DCK_P + 15 + cx] : 128;
        const int dc = (int)((warp_sum((uint32_t)(top + left)) + 32) >> 6);
        const uint32_t dc4 = (uint32_t)dc * 0x01010101u;
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = vsad4_acc(sw[k], dc4, acc);
        acc = warp_sum(acc);
        if (lane == 0 && x + 32 <= W && y + 32 <= H) { const int cu = 1 + w; cand[(size_t)cu * A.cap] = ois_pack(acc, true, 1); total[cu] = 1; }
    }
    // ---- the four 16x16 CUs of the quadrant: lanes 0-15 rows of the upper pair, 16-31 of the lower pair; words 0-3 -> left CU, 4-7 -> right CU ----
    {
        const int qy = lane >> 4, cr = lane & 15, y16 = y + 16 * qy;
        uint32_t sad[2];
#pragma unroll
        for (int qx = 0; qx < 2; ++qx) {
            const int x16 = x + 16 * qx;
            const int tcol = 16 + cx + 16 * qx, trow = cy + 16 * qy;                       // tile coordinates of the CU's sample (0, 0): row trow + 1
            const int top = (y16 != 0 && x16 + cr < W) ? s_t[trow * DCK_P + tcol + cr] : 128;
            const int left = (x16 != 0 && y16 + cr < H) ? s_t[(trow + cr + 1) * DCK_P + tcol - 1] : 128;
            uint32_t ssum = (uint32_t)(top + left);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);      // over the 16 lanes of the CU
            const int dc = (int)((ssum + 16) >> 5);
            const uint32_t dc4 = (uint32_t)dc * 0x01010101u;
            uint32_t p[4] = {dc4, dc4, dc4, dc4};
            if (cr == 0) {      // first row: (top[x] + 3 dc + 2) >> 2, corner (left[0] + top[0] + 2 dc + 2) >> 2 (IntraModeDCLuma edge filter, N < 32)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t wv = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int xx = 4 * k + b;
                        const int t = (y16 != 0 && x16 + xx < W) ? s_t[trow * DCK_P + tcol + xx] : 128;
                        const int v = xx == 0 ? (left + t + 2 * dc + 2) >> 2 : (t + 3 * dc + 2) >> 2;
                        wv |= (uint32_t)v << (8 * b);
                    }
                    p[k] = wv;
                }
            } else {
                p[0] = (dc4 & 0xffffff00u) | (uint32_t)((left + 3 * dc + 2) >> 2);       // first column
            }
            uint32_t acc = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = vsad4_acc(sw[4 * qx + k], p[k], acc);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            sad[qx] = acc;
        }
        if (cr == 0) {
#pragma unroll
            for (int qx = 0; qx < 2; ++qx) {
                const int x16 = x + 16 * qx;
                if (x16 + 16 <= W && y16 + 16 <= H) {
                    const int cu = 5 + ((cy >> 4) + qy) * 4 + (cx >> 4) + qx;
                    cand[(size_t)cu * A.cap] = ois_pack(sad[qx], true, 1); total[cu] = 1;
                }
            }
        }
    }
}

static int ois_check(const svtb200_ois_config *cfg, const svtb200_plane *src, const void *me, const void *total, const void *cand, uint32_t cap, const char *who)
{
    if (!cfg || !src || !src->buf || !total || !cand) { set_error("%s: null argument", who); return SVTB200_ERR_BAD_ARG; }
    if (cfg->pic_width == 0 || cfg->pic_height == 0 || src->width != cfg->pic_width || src->height != cfg->pic_height) { set_error("%s: picture size / plane size mismatch", who); return SVTB200_ERR_BAD_ARG; }
    if (cfg->slice_type > 2 || cfg->temporal_layer > 5 || cfg->ois_th_set > 2) { set_error("%s: slice_type 0..2, temporal_layer 0..5, ois_th_set 0..2", who); return SVTB200_ERR_BAD_ARG; }
    const bool is_i = cfg->slice_type == 2;
    const uint32_t need = (!is_i && cfg->ois_kernel_level) ? 18 : (!is_i && cfg->limit_ois_to_dc_mode) ? 1 : 9;
    if (cap < need) { set_error("%s: cand_capacity %u < %u required by this configuration", who, cap, need); return SVTB200_ERR_BAD_ARG; }
    if (!is_i && !cfg->ois_kernel_level && !cfg->limit_ois_to_dc_mode && !me) { set_error("%s: me_results required on this path", who); return SVTB200_ERR_BAD_ARG; }
    return SVTB200_OK;
}

// the staging window starts 1 sample left of / above the LCU and the 128-bit loader rounds rows out to 16 bytes on both sides
static int ois_padded(const svtb200_plane &p) { return p.origin_x >= 16 && p.origin_y >= 1 && p.stride >= (uint32_t)p.origin_x + p.width + 16; }

static int ois_launch(const svtb200_ois_config &cfg, const uint8_t *src00, int stride, int padded, const svtb200_me_pu_result *me, uint8_t *total, uint32_t *cand, uint32_t cap, cudaStream_t st)
{
    OisArgs A;
    A.cfg = cfg; A.src = src00; A.stride = stride; A.me = me; A.total = total; A.cand = cand; A.cap = (int)cap;
    A.lcu_cols = (cfg.pic_width + 63) / 64;
    A.padded = padded;
    const int n_lcu = A.lcu_cols * ((cfg.pic_height + 63) / 64);
    // non-I pictures limited to DC without 8x8 CUs (every inter picture of M10) on a padded, word-aligned plane: the dedicated kernel
    const bool dc_only = cfg.slice_type != 2 && !cfg.ois_kernel_level && cfg.limit_ois_to_dc_mode && (cfg.skip_ois_8x8 || cfg.cu8x8_mode == 1);
    if (dc_only && padded && (((uintptr_t)src00 | (uintptr_t)stride) & 3) == 0) ois_dc_kernel<<<n_lcu, 128, 0, st>>>(A);
    else ois_picture_kernel<<<n_lcu, OIS_NT, 0, st>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

int svtb200_IntraPrediction(uint32_t mode, uint32_t size, int is_chroma, const uint8_t *refSamples, uint8_t *predictionPtr, uint32_t predictionBufferStride)
{
    if (!refSamples || !predictionPtr || mode > 34 || !(size == 4 || size == 8 || size == 16 || size == 32) || predictionBufferStride < size) {
        set_error("IntraPrediction: mode 0..34, size 4/8/16/32, stride >= size, non-null pointers"); return SVTB200_ERR_BAD_ARG;
    }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    void *d_line, *d_pred, *h;
    const size_t pb = (size_t)size * size;
    if ((rc = c->dev_alloc(256, &d_line)) || (rc = c->dev_alloc(pb, &d_pred)) || (rc = c->pin_alloc(pb, &h))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_line, refSamples, 4 * size + 1, cudaMemcpyHostToDevice, c->stream));
    intra_leaf_kernel<<<1, IN_NT, 0, c->stream>>>((int)mode, (int)size, is_chroma ? 1 : 0, (const uint8_t *)d_line, (uint8_t *)d_pred, (int)size);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    SVTB200_CUDA_TRY(cudaMemcpyAsync(h, d_pred, pb, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (uint32_t y = 0; y < size; ++y) memcpy(predictionPtr + (size_t)y * predictionBufferStride, (const uint8_t *)h + (size_t)y * size, size);
    return SVTB200_OK;
}

static int intra_batch_dev_impl(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs_dev, uint32_t n_jobs, void *cuda_stream, bool hbd)
{
    const size_t bps = hbd ? 2 : 1;
    int rc = intra_check(recon, pred, jobs_dev, n_jobs, false);
    if (rc) return rc;
    if ((rc = ensure_device())) return rc;
    IntraArgs A;
    const svtb200_plane *r[3] = {&recon->y, &recon->cb, &recon->cr}, *p[3] = {&pred->y, &pred->cb, &pred->cr};
    for (int k = 0; k < 3; ++k) { A.recon[k] = plane_w(*r[k], r[k]->buf, bps); A.pred[k] = plane_w(*p[k], p[k]->buf, bps); }
    A.jobs = jobs_dev; A.n_jobs = (int)n_jobs;
    return intra_launch(A, (cudaStream_t)cuda_stream, hbd);
}

static int intra_batch_impl(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs, bool hbd)
{
    const size_t bps = hbd ? 2 : 1;
    int rc = intra_check(recon, pred, jobs, n_jobs, true);
    if (rc) return rc;
    if (n_jobs == 0) return SVTB200_OK;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const svtb200_plane *r[3] = {&recon->y, &recon->cb, &recon->cr}, *p[3] = {&pred->y, &pred->cb, &pred->cr};
    const bool in_place = recon->y.buf == pred->y.buf;
    size_t need = 8192 + align_up((size_t)n_jobs * sizeof(svtb200_intra_job), 256);
    for (int k = 0; k < 3; ++k) need += align_up(bps * pbytes(*r[k]), 256) + (in_place ? 0 : align_up(bps * pbytes(*p[k]), 256));
    if ((rc = c->dev_reserve(need))) return rc;
    void *dr[3], *dp[3], *dj;
    for (int k = 0; k < 3; ++k) {
        if ((rc = c->dev_alloc(bps * pbytes(*r[k]), &dr[k]))) return rc;
        SVTB200_CUDA_TRY(cudaMemcpyAsync(dr[k], r[k]->buf, bps * pbytes(*r[k]), cudaMemcpyHostToDevice, c->stream));
        dp[k] = dr[k];
        if (!in_place) {
            if ((rc = c->dev_alloc(bps * pbytes(*p[k]), &dp[k]))) return rc;
            SVTB200_CUDA_TRY(cudaMemcpyAsync(dp[k], p[k]->buf, bps * pbytes(*p[k]), cudaMemcpyHostToDevice, c->stream));
        }
    }
    if ((rc = c->dev_alloc((size_t)n_jobs * sizeof(svtb200_intra_job), &dj))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(dj, jobs, (size_t)n_jobs * sizeof(svtb200_intra_job), cudaMemcpyHostToDevice, c->stream));
    IntraArgs A;
    for (int k = 0; k < 3; ++k) { A.recon[k] = plane_w(*r[k], (uint8_t *)dr[k], bps); A.pred[k] = plane_w(*p[k], (uint8_t *)dp[k], bps); }
    A.jobs = (const svtb200_intra_job *)dj; A.n_jobs = (int)n_jobs;
    if ((rc = intra_launch(A, c->stream, hbd))) return rc;
    for (int k = 0; k < 3; ++k) SVTB200_CUDA_TRY(cudaMemcpyAsync(p[k]->buf, dp[k], bps * pbytes(*p[k]), cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_intra_batch_dev(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs_dev, uint32_t n_jobs, void *cuda_stream)
{ return intra_batch_dev_impl(recon, pred, jobs_dev, n_jobs, cuda_stream, false); }
int svtb200_intra_batch(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs)
{ return intra_batch_impl(recon, pred, jobs, n_jobs, false); }
int svtb200_intra_batch16_dev(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs_dev, uint32_t n_jobs, void *cuda_stream)
{ return intra_batch_dev_impl(recon, pred, jobs_dev, n_jobs, cuda_stream, true); }
int svtb200_intra_batch16(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs)
{ return intra_batch_impl(recon, pred, jobs, n_jobs, true); }

int svtb200_ois_picture_dev(const svtb200_ois_config *cfg, const svtb200_plane *src_luma_dev, const svtb200_me_pu_result *me_results_dev, uint8_t *total_modes_dev,
                            uint32_t *cand_dev, uint32_t cand_capacity, void *cuda_stream)
{
    int rc = ois_check(cfg, src_luma_dev, me_results_dev, total_modes_dev, cand_dev, cand_capacity, "ois_picture_dev");
    if (rc) return rc;
    if ((rc = ensure_device())) return rc;
    const uint8_t *s00 = src_luma_dev->buf + (size_t)src_luma_dev->origin_y * src_luma_dev->stride + src_luma_dev->origin_x;
    return ois_launch(*cfg, s00, (int)src_luma_dev->stride, ois_padded(*src_luma_dev), me_results_dev, total_modes_dev, cand_dev, cand_capacity, (cudaStream_t)cuda_stream);
}

int svtb200_ois_picture(const svtb200_ois_config *cfg, const svtb200_plane *src_luma, const svtb200_me_pu_result *me_results, uint8_t *total_modes, uint32_t *cand,
                        uint32_t cand_capacity)
{
    int rc = ois_check(cfg, src_luma, me_results, total_modes, cand, cand_capacity, "ois_picture");
    if (rc) return rc;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t n_lcu = (size_t)((cfg->pic_width + 63) / 64) * ((cfg->pic_height + 63) / 64);
    const size_t sb = pbytes(*src_luma), mb = me_results ? n_lcu * 85 * sizeof(svtb200_me_pu_result) : 0, tb = n_lcu * 85, cb = n_lcu * 85 * cand_capacity * sizeof(uint32_t);
    if ((rc = c->dev_reserve(8192 + align_up(sb, 256) + align_up(mb, 256) + align_up(tb, 256) + align_up(cb, 256)))) return rc;
    void *ds, *dm = nullptr, *dt, *dc;
    if ((rc = c->dev_alloc(sb, &ds)) || (rc = c->dev_alloc(tb, &dt)) || (rc = c->dev_alloc(cb, &dc))) return rc;
    if (mb && (rc = c->dev_alloc(mb, &dm))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(ds, src_luma->buf, sb, cudaMemcpyHostToDevice, c->stream));
    if (mb) SVTB200_CUDA_TRY(cudaMemcpyAsync(dm, me_results, mb, cudaMemcpyHostToDevice, c->stream));
    const uint8_t *s00 = (const uint8_t *)ds + (size_t)src_luma->origin_y * src_luma->stride + src_luma->origin_x;
    if ((rc = ois_launch(*cfg, s00, (int)src_luma->stride, ois_padded(*src_luma), (const svtb200_me_pu_result *)dm, (uint8_t *)dt, (uint32_t *)dc, cand_capacity, c->stream))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(total_modes, dt, tb, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(cand, dc, cb, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

}
