// me.cu -- motion estimation on B200: hierarchical ME block SAD search (HME level 0/1/2),
// 85-PU full-pel search, integer-pel bi-prediction SAD and MeCuResults assembly.
//
// Replaces:
//   SadLoopKernel*                       C_DEFAULT/EbComputeSAD_C.c:165, ASM_AVX2/EbComputeSAD_Intrinsic_AVX2.c:28,
//                                        ASM_AVX2/EbComputeSAD_SadLoopKernel_AVX512.c:59,1121
//   GetEightHorizontalSearchPointResults_8x8_16x16_PU / _32x32_64x64   C_DEFAULT/EbComputeSAD_C.c:244,355
//   SadCalculation_8x8_16x16 / _32x32_64x64                            C_DEFAULT/EbMeSadCalculation_C.c:16,73
//   FastLoop_NxMSadKernel / CombinedAveragingSAD                       C_DEFAULT/EbComputeSAD_C.c:141,14
//   MotionEstimateLcu + helpers (driver logic)                         Codec/EbMotionEstimation.c:3671 (and :586,
//                                        :1847,:2012,:2194,:2315,:2946,:3363,:2742,:2870,:4321-4438)
//
// Design: ONE CTA PER LCU for the whole picture (grid = #LCUs; 128 threads on integer-pel pictures, 256 with sub-pel refinement: me_kernels.inc).  The CTA stages the 64x64 source
// block and each reference search window in shared memory exactly once (every reference byte is read from
// HBM/L2 once per LCU), evaluates search positions with VABSDIFF4 (__vsadu4: 4 byte-SADs per lane-op) on
// funnel-shifted words, and reduces candidates with warp shuffles.  The reference's sequential tie-break rules
// are kept by encoding the raster index in the reduction key (first minimum wins) and by replaying the 85-PU
// best-update scan position by position in scan order.  No tensor-core work exists on this path.
#include "common.cuh"

namespace svtb200 {

constexpr int CHUNK = 64;               // full-pel search positions evaluated per pass
constexpr uint32_t MAX_SAD_INIT = 64 * 64 * 255;   // MAX_SAD_VALUE, EbMotionEstimation.c:99

struct PlaneD {
    const uint8_t *buf;
    int stride, ox, oy, w, h;
};
struct MeArgs {
    svtb200_me_config cfg;
    int sp_fp, sp_sp, sp_f_bytes, sp_p_bytes;   // sub-pel planes: pitches and bytes of the F plane / of each of B, H, J
    int sp_alias;         // the second list's planes overlay the search scratch
    PlaneD src[3];        // full, quarter, sixteenth
    PlaneD ref[2][3];
    svtb200_me_pu_result *results;
    uint32_t *best_sad, *best_mv, *rc;
    int cols, rows;
    int scratch_bytes;    // bytes of the search-scratch union in dynamic shared memory
    int sq_bytes;         // bytes of the quarter-resolution source block behind it (HME level 1 only)
    // fast paths (me_kernels.inc): HME level 0 on byte-phase copies, 85-PU full-pel search of 8 x 7 / 16 x 9 areas
    int hme_fast, hme_p, hme_band, hme_cs;      // window pitch (= TMA box width), search rows per band, bytes between phase copies
    int fp_yt, fp_nxh, fp_cs;                   // full-pel: search rows (7 / 9, 0 = generic path only), search width / 8, bytes between phase copies
    CUtensorMap map_hme[2];                     // per list: (hme_p + 16) x (hme_band + 14) box of the sixteenth-resolution reference plane
    CUtensorMap map_fp[2];                      // per list: 96 x (fp_yt + 63) box of the full-resolution reference plane
};

constexpr int ME_SCRATCH_OFF = 6272;            // fixed part (6184) + mbarrier, rounded up to the 128-byte TMA destination alignment

}  // namespace svtb200

#define ME_NS k256
#define ME_NT 256
#define ME_MIN_CTAS 4
namespace svtb200 {
#include "me_kernels.inc"
}
#undef ME_NS
#undef ME_NT
#undef ME_MIN_CTAS
#define ME_NS k128
#define ME_NT 128
#define ME_MIN_CTAS 7
namespace svtb200 {
#include "me_kernels.inc"
}
#undef ME_NS
#undef ME_NT
#undef ME_MIN_CTAS

namespace svtb200 {
using namespace k256;      // leaf kernels and the sub-pel picture kernel

// worst-case window bytes over every search the GENERIC path of this configuration can issue
static int me_window_bytes(const svtb200_me_config &c)
{
    static const int mult_tab[6][6] = {{100}, {100, 100}, {100, 100, 100}, {200, 140, 100, 70}, {350, 200, 100, 100, 100}, {525, 350, 200, 100, 100, 100}};
    const int mult = mult_tab[c.hierarchical_levels][c.temporal_layer];
    int best = 0;
    auto upd = [&](int saw, int sah, int bw, int bh_rows_span) {
        const int wb = saw + bw - 1, rows = sah + bh_rows_span;
        const int wp = win_pitch(wb);
        if (wp * rows > best) best = wp * rows;
    };
    if (c.enable_hme && c.enable_hme_l0) {
        int w = (c.hme_l0_total_w * mult) / 100, h = (c.hme_l0_total_h * mult) / 100;
        for (int q = 0; q < 2; ++q) { w = std::max(w, (c.hme_l0_w[q] * mult) / 100); h = std::max(h, (c.hme_l0_h[q] * mult) / 100); }
        upd(w, h, 16, 14);
    }
    auto rnd8 = [](int w) { return w < 8 ? 8 : (w & 7) ? w + (w - ((w >> 3) << 3)) : w; };
    if (c.enable_hme && c.enable_hme_l1) upd(rnd8(std::max(c.hme_l1_w[0], c.hme_l1_w[1])), std::max(c.hme_l1_h[0], c.hme_l1_h[1]), 32, 30);
    if (c.enable_hme && c.enable_hme_l2) upd(rnd8(std::max(c.hme_l2_w[0], c.hme_l2_w[1])), std::max(c.hme_l2_h[0], c.hme_l2_h[1]), 64, 62);
    upd(std::min<int>(c.search_area_w, 127), std::min<int>(c.search_area_h, 127), 64, 63);
    return (best + 15) & ~15;
}

// geometry of the fast paths for this configuration (MeArgs::hme_* / fp_*) and the scratch union they share with the generic path
static void me_plan(const svtb200_me_config &c, MeArgs &A)
{
    static const int mult_tab[6][6] = {{100}, {100, 100}, {100, 100, 100}, {200, 140, 100, 70}, {350, 200, 100, 100, 100}, {525, 350, 200, 100, 100, 100}};
    const int mult = mult_tab[c.hierarchical_levels][c.temporal_layer];
    size_t need = (size_t)CHUNK * 64 * 2 + (size_t)CHUNK * 21 * 4 + (size_t)me_window_bytes(c);
    A.hme_fast = 0;
    if (c.enable_hme && c.enable_hme_l0) {
        const bool one = c.one_quadrant_hme && !c.enable_hme_l1 && !c.enable_hme_l2;
        int w = 0, h = 0;
        if (one) { w = (c.hme_l0_total_w * mult) / 100; h = (c.hme_l0_total_h * mult) / 100; }
        else for (int q = 0; q < 2; ++q) { w = std::max(w, (c.hme_l0_w[q] * mult) / 100); h = std::max(h, (c.hme_l0_h[q] * mult) / 100); }
        // a row of 8-position items reads 32 (G - 1) + 48 bytes of each phase copy; copies 1..3 are shifted from copy 0, which costs its last word:
        // pitch >= that + 4, a multiple of 16 with an ODD number of 16-byte chunks (8 consecutive rows -> 8 different bank groups)
        const int G = (std::max(w, 1) + 31) / 32;
        int P = (32 * (G - 1) + 52 + 15) & ~15;
        if (((P >> 4) & 1) == 0) P += 16;
        const int nbands = (std::max(h, 1) + 23) / 24, band = (std::max(h, 1) + nbands - 1) / nbands;
        if (w > 0 && h > 0 && P + 16 <= 256 && band + 14 <= 256 && (size_t)w * h <= 65536) {
            A.hme_fast = 1; A.hme_p = P; A.hme_band = band;
            A.hme_cs = (P * (band + 14) + 127) & ~127;
            need = std::max(need, (size_t)4 * A.hme_cs + 64);      // four phase copies (the raw TMA box is parked on copies 2 / 3)
        }
    }
    A.fp_yt = 0;
    const int saw = std::min<int>(c.search_area_w, 127), sah = std::min<int>(c.search_area_h, 127);
    if ((saw == 8 && sah == 7) || (saw == 16 && sah == 9)) {
        A.fp_yt = sah; A.fp_nxh = saw / 8;
        int cs = (saw == 8 ? 72 : 80) * (sah + 63);
        cs = (cs + 15) & ~15;
        while ((cs & 127) != 64) cs += 16;       // phases 0 / 1 (and 2 / 3) of a half warp land in different halves of the 32 banks
        A.fp_cs = cs;
        need = std::max(need, (size_t)4 * cs + (size_t)(88 + 12 * saw * sah) * 4);
    }
    A.scratch_bytes = (int)((need + 127) & ~(size_t)127);
}

static int validate_cfg(const svtb200_me_config *c)
{
    if (!c || c->pic_width < 8 || c->pic_height < 8) { set_error("me: bad picture size"); return SVTB200_ERR_BAD_ARG; }
    if (c->num_lists < 1 || c->num_lists > 2 || c->hierarchical_levels > 5 || c->temporal_layer > c->hierarchical_levels) { set_error("me: bad list/layer configuration"); return SVTB200_ERR_BAD_ARG; }
    if (c->fractional_search_model > 2 || c->fractional_search_method > 1) { set_error("me: fractional_search_model 0..2; fractional_search_method 0 (SUB_SAD_SEARCH) or 1 (FULL_SAD_SEARCH) -- SSD_SEARCH (enc modes <= 4) is not built"); return SVTB200_ERR_UNSUPPORTED; }
    if (c->n_hme_regions_w > 2 || c->n_hme_regions_h > 2 || c->search_area_w == 0 || c->search_area_h == 0) { set_error("me: bad search configuration"); return SVTB200_ERR_BAD_ARG; }
    return SVTB200_OK;
}

static PlaneD to_dev_plane(const svtb200_plane &p, const uint8_t *dbuf)
{
    PlaneD d = {dbuf, (int)p.stride, p.origin_x, p.origin_y, p.width, p.height};
    return d;
}

static int launch_me(const svtb200_me_config *cfg, const PlaneD src[3], const PlaneD ref[2][3], svtb200_me_pu_result *results,
                     uint32_t *bs, uint32_t *bm, uint32_t *rc, cudaStream_t st)
{
    MeArgs A;
    memset(&A, 0, sizeof A);
    A.cfg = *cfg;
    for (int i = 0; i < 3; ++i) { A.src[i] = src[i]; A.ref[0][i] = ref[0][i]; A.ref[1][i] = ref[1][i]; }
    A.results = results; A.best_sad = bs; A.best_mv = bm; A.rc = rc;
    A.cols = (cfg->pic_width + 63) / 64; A.rows = (cfg->pic_height + 63) / 64;
    me_plan(*cfg, A);
    // tensor maps over the padded plane buffers (pitch and base multiples of 16: svtb200_pa_picture_create / the host entry point's re-pitched copies)
    auto rows_of = [](const PlaneD &p) { return (uint32_t)(p.h + 2 * p.oy); };
    int rc2 = 0;
    for (int li = 0; li < 2; ++li) {
        const PlaneD &rF = ref[li][0], &rS = ref[li][2];
        if (A.hme_fast && rS.buf) { if ((rc2 = make_plane_tensor_map(&A.map_hme[li], rS.buf, (uint32_t)rS.stride, rows_of(rS), (uint32_t)A.hme_p + 16, (uint32_t)A.hme_band + 14))) return rc2; }
        else if (li == 0) A.hme_fast = 0;
        if (A.fp_yt) { if ((rc2 = make_plane_tensor_map(&A.map_fp[li], rF.buf, (uint32_t)rF.stride, rows_of(rF), 96, (uint32_t)A.fp_yt + 63))) return rc2; }
    }
    const bool subpel = cfg->fractional_search_model != 2;
    A.sq_bytes = (cfg->enable_hme && cfg->enable_hme_l1) ? 2048 : 0;
    size_t smem = (size_t)ME_SCRATCH_OFF + (size_t)A.scratch_bytes + (size_t)A.sq_bytes + 128;
    if (subpel) {   // [84][8] distortions + per list the integer plane with apron and the three half-sample planes
        const int saw = std::min<int>(cfg->search_area_w, 127), sah = std::min<int>(cfg->search_area_h, 127);
        A.sp_fp = win_pitch(saw + 68);
        A.sp_sp = (saw + 68 + 4 + 3) & ~3;
        A.sp_f_bytes = (A.sp_fp * (sah + 68) + 15) & ~15;
        A.sp_p_bytes = (A.sp_sp * (sah + 68) + 15) & ~15;
        const size_t per_list = (size_t)A.sp_f_bytes + 3 * (size_t)A.sp_p_bytes;
        A.sp_alias = cfg->num_lists == 2 && per_list <= (size_t)A.scratch_bytes;
        smem += 84 * 8 * 4 + (A.sp_alias || cfg->num_lists < 2 ? 1 : 2) * per_list;
    }
    if (smem > 227 * 1024) { set_error("me: this search configuration needs %zu B of shared memory (> 227 KB)", smem); return SVTB200_ERR_UNSUPPORTED; }
    static thread_local size_t smem_set[2] = {0, 0};
    if (smem > smem_set[subpel]) {
        if (subpel) SVTB200_CUDA_TRY(cudaFuncSetAttribute(k256::me_picture_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        else SVTB200_CUDA_TRY(cudaFuncSetAttribute(k128::me_picture_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        smem_set[subpel] = smem;
    }
    if (subpel) k256::me_picture_kernel<true><<<A.cols * A.rows, k256::NT, smem, st>>>(A);
    else k128::me_picture_kernel<false><<<A.cols * A.rows, k128::NT, smem, st>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

// tensor maps need 16-byte pitches and bases; the picture kernel has no other loader for the source block
static int check_dev_planes(const svtb200_me_config *cfg, const svtb200_pa_picture *p, const char *what)
{
    const svtb200_plane *pl[3] = {&p->full, &p->quarter, &p->sixteenth};
    for (int k = 0; k < 3; ++k) {
        const bool used = k == 0 || (k == 1 && cfg->enable_hme && cfg->enable_hme_l1) || (k == 2 && cfg->enable_hme && cfg->enable_hme_l0);
        if (!used) continue;
        if (!pl[k]->buf) { set_error("me: plane %d of the %s picture is null", k, what); return SVTB200_ERR_BAD_ARG; }
        if (((uintptr_t)pl[k]->buf & 15) || (pl[k]->stride & 15)) {
            set_error("me: plane %d of the %s picture: device planes need a 16-byte-aligned base and pitch (svtb200_pa_picture_create provides both)", k, what);
            return SVTB200_ERR_UNSUPPORTED;
        }
    }
    return SVTB200_OK;
}

static size_t plane_bytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }
static size_t plane_bytes16(const svtb200_plane &p) { return align_up(p.stride, 16) * (size_t)(p.height + 2 * p.origin_y) + 64; }

// ---- leaf kernels --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) sad_loop_leaf_kernel(const uint8_t *src, int sstride, const uint8_t *ref, int rstride,
                                                           int h, int w, int raw, int saw, int sah, int row_step, uint32_t *out /*[3]*/)
{
    extern __shared__ __align__(16) uint8_t smem[];
    unsigned long long *s_red64 = (unsigned long long *)smem;
    uint8_t *s_blk = smem + 64;
    const int bp = (w + 3 + 3) & ~3;
    uint8_t *s_win = s_blk + ((bp * h + 15) & ~15);   // smem base and the 64-byte header keep this 16-byte aligned
    for (int i = threadIdx.x; i < bp * h; i += NT) { const int r = i / bp, c = i - r * bp; s_blk[i] = c < w ? src[r * sstride + c] : 0; }
    const int wbytes = saw + w - 1, wrows = sah + (h - 1) * row_step, wp = win_pitch(wbytes);
    const int a = load_window_t<NT>(s_win, wp, ref, raw, wbytes, wrows);
    __syncthreads();
    uint32_t bs;
    int bx = (int)out[1], by = (int)out[2];
    sad_search(s_blk, bp, w, h, s_win, wp, a, row_step, saw, sah, s_red64, bs, bx, by);
    if (threadIdx.x == 0) { out[0] = bs; out[1] = (uint32_t)bx; out[2] = (uint32_t)by; }
}

__global__ void __launch_bounds__(NT) sad_nxm_leaf_kernel(const uint8_t *src, int ss, const uint8_t *r1, int r1s, const uint8_t *r2, int r2s,
                                                          int h, int w, uint32_t *out)
{
    __shared__ uint32_t s_red[8];
    uint32_t acc[1] = {0};
    for (int i = threadIdx.x; i < h * w; i += NT) {
        const int r = i / w, c = i - r * w;
        int p = r1[r * r1s + c];
        if (r2) p = (p + r2[r * r2s + c] + 1) >> 1;
        acc[0] += (uint32_t)abs((int)src[r * ss + c] - p);
    }
    block_sum<1>(acc, s_red);
    if (threadIdx.x == 0) *out = acc[0];
}

__global__ void __launch_bounds__(NT) fullpel_leaf_kernel(const uint8_t *src, int ss, const uint8_t *ref, int rs, int sox, int soy,
                                                          int saw, int sah, uint32_t *bsad, uint32_t *bmv)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_src = smem;
    uint16_t *s_sad8 = (uint16_t *)(s_src + 4096);
    uint32_t *s_sum = (uint32_t *)(s_sad8 + CHUNK * 64);
    uint32_t *s_bsad = s_sum + CHUNK * 21;
    uint32_t *s_bmv = s_bsad + 85;
    uint8_t *s_win = (uint8_t *)(((uintptr_t)(s_bmv + 85 + 2) + 15) & ~(uintptr_t)15);
    for (int i = threadIdx.x; i < 4096; i += NT) s_src[i] = src[(i >> 6) * ss + (i & 63)];
    if (threadIdx.x < 85) { s_bsad[threadIdx.x] = bsad[threadIdx.x]; s_bmv[threadIdx.x] = bmv[threadIdx.x]; }
    const int wbytes = saw + 63, wrows = sah + 63, wp = win_pitch(wbytes);
    const int a = load_window_t<NT>(s_win, wp, ref, rs, wbytes, wrows);
    __syncthreads();
    fullpel_search(s_src, s_win, wp, a, sox, soy, saw, sah, s_sad8, s_sum, s_bsad, s_bmv);
    if (threadIdx.x < 85) { bsad[threadIdx.x] = s_bsad[threadIdx.x]; bmv[threadIdx.x] = s_bmv[threadIdx.x]; }
}

}  // namespace svtb200

using namespace svtb200;

extern "C" {

int svtb200_me_picture_dev(const svtb200_me_config *cfg, const svtb200_pa_picture *src, const svtb200_pa_picture *ref_l0,
                           const svtb200_pa_picture *ref_l1, svtb200_me_pu_result *results, uint32_t *lcu_best_sad,
                           uint32_t *lcu_best_mv, uint32_t *rc_me_distortion, void *cuda_stream)
{
    int rc = validate_cfg(cfg);
    if (rc) return rc;
    if (!src || !ref_l0 || !results || (cfg->num_lists == 2 && !ref_l1)) { set_error("me: null picture/result pointer"); return SVTB200_ERR_BAD_ARG; }
    if ((rc = ensure_device())) return rc;
    const svtb200_pa_picture *r1 = ref_l1 ? ref_l1 : ref_l0;
    if ((rc = check_dev_planes(cfg, src, "source")) || (rc = check_dev_planes(cfg, ref_l0, "list-0 reference")) || (rc = check_dev_planes(cfg, r1, "list-1 reference"))) return rc;
    PlaneD s[3] = {to_dev_plane(src->full, src->full.buf), to_dev_plane(src->quarter, src->quarter.buf), to_dev_plane(src->sixteenth, src->sixteenth.buf)};
    PlaneD r[2][3] = {{to_dev_plane(ref_l0->full, ref_l0->full.buf), to_dev_plane(ref_l0->quarter, ref_l0->quarter.buf), to_dev_plane(ref_l0->sixteenth, ref_l0->sixteenth.buf)},
                      {to_dev_plane(r1->full, r1->full.buf), to_dev_plane(r1->quarter, r1->quarter.buf), to_dev_plane(r1->sixteenth, r1->sixteenth.buf)}};
    return launch_me(cfg, s, r, results, lcu_best_sad, lcu_best_mv, rc_me_distortion, (cudaStream_t)cuda_stream);
}

int svtb200_me_picture(const svtb200_me_config *cfg, const svtb200_pa_picture *src, const svtb200_pa_picture *ref_l0,
                       const svtb200_pa_picture *ref_l1, svtb200_me_pu_result *results, uint32_t *lcu_best_sad,
                       uint32_t *lcu_best_mv, uint32_t *rc_me_distortion)
{
    int rc = validate_cfg(cfg);
    if (rc) return rc;
    if (!src || !ref_l0 || !results || (cfg->num_lists == 2 && !ref_l1)) { set_error("me: null picture/result pointer"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init())) return rc;
    c->reset();
    const int nl = cfg->num_lists;
    const svtb200_pa_picture *pics[3] = {src, ref_l0, nl == 2 ? ref_l1 : nullptr};
    const bool same_ref = nl == 2 && ref_l1->full.buf == ref_l0->full.buf && ref_l1->quarter.buf == ref_l0->quarter.buf && ref_l1->sixteenth.buf == ref_l0->sixteenth.buf;
    const size_t n_lcu = (size_t)((cfg->pic_width + 63) / 64) * ((cfg->pic_height + 63) / 64);
    size_t need = 0;
    for (int p = 0; p < 3; ++p) {
        if (!pics[p] || (p == 2 && same_ref)) continue;
        need += align_up(plane_bytes16(pics[p]->full), 256) + align_up(plane_bytes16(pics[p]->quarter), 256) + align_up(plane_bytes16(pics[p]->sixteenth), 256);
    }
    const size_t res_bytes = n_lcu * 85 * sizeof(svtb200_me_pu_result), arr_bytes = n_lcu * 170 * 4, rc_bytes = n_lcu * 4;
    need += align_up(res_bytes, 256) + 2 * align_up(arr_bytes, 256) + align_up(rc_bytes, 256) + 1024;
    if ((rc = c->dev_reserve(need))) return rc;
    PlaneD dp[3][3];
    for (int p = 0; p < 3; ++p) {
        if (!pics[p]) continue;
        if (p == 2 && same_ref) { for (int k = 0; k < 3; ++k) dp[2][k] = dp[1][k]; continue; }
        const svtb200_plane *pl[3] = {&pics[p]->full, &pics[p]->quarter, &pics[p]->sixteenth};
        for (int k = 0; k < 3; ++k) {
            const bool used = k == 0 || (k == 1 && cfg->enable_hme && cfg->enable_hme_l1) || (k == 2 && cfg->enable_hme && cfg->enable_hme_l0);
            void *d = nullptr;
            if (used) {
                if (!pl[k]->buf) { set_error("me: plane %d of picture %d is null", k, p); return SVTB200_ERR_BAD_ARG; }
                // the device copy gets a 16-byte pitch (TMA tensor maps); the geometry (origin, size) stays the caller's
                const size_t rows = (size_t)pl[k]->height + 2 * pl[k]->origin_y, p16 = align_up(pl[k]->stride, 16);
                if ((rc = c->dev_alloc(plane_bytes16(*pl[k]), &d))) return rc;
                SVTB200_CUDA_TRY(cudaMemcpy2DAsync(d, p16, pl[k]->buf, pl[k]->stride, pl[k]->stride, rows, cudaMemcpyHostToDevice, c->stream));
            }
            dp[p][k] = to_dev_plane(*pl[k], (const uint8_t *)d);
            dp[p][k].stride = (int)align_up(pl[k]->stride, 16);
        }
    }
    if (nl == 1) for (int k = 0; k < 3; ++k) dp[2][k] = dp[1][k];
    void *d_res, *d_bs, *d_bm, *d_rc;
    if ((rc = c->dev_alloc(res_bytes, &d_res)) || (rc = c->dev_alloc(arr_bytes, &d_bs)) || (rc = c->dev_alloc(arr_bytes, &d_bm)) || (rc = c->dev_alloc(rc_bytes, &d_rc))) return rc;
    PlaneD refs[2][3];
    for (int k = 0; k < 3; ++k) { refs[0][k] = dp[1][k]; refs[1][k] = dp[2][k]; }
    if ((rc = launch_me(cfg, dp[0], refs, (svtb200_me_pu_result *)d_res, lcu_best_sad ? (uint32_t *)d_bs : nullptr,
                        lcu_best_mv ? (uint32_t *)d_bm : nullptr, rc_me_distortion ? (uint32_t *)d_rc : nullptr, c->stream))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(results, d_res, res_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (lcu_best_sad) SVTB200_CUDA_TRY(cudaMemcpyAsync(lcu_best_sad, d_bs, arr_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (lcu_best_mv) SVTB200_CUDA_TRY(cudaMemcpyAsync(lcu_best_mv, d_bm, arr_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (rc_me_distortion) SVTB200_CUDA_TRY(cudaMemcpyAsync(rc_me_distortion, d_rc, rc_bytes, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_SadLoopKernel(const uint8_t *src, uint32_t srcStride, const uint8_t *ref, uint32_t refStride, uint32_t height, uint32_t width,
                          uint64_t *bestSad, int16_t *xSearchCenter, int16_t *ySearchCenter, uint32_t srcStrideRaw,
                          int16_t searchAreaWidth, int16_t searchAreaHeight)
{
    if (!src || !ref || !bestSad || !xSearchCenter || !ySearchCenter || width == 0 || height == 0 || width > 64 || height > 64) { set_error("SadLoopKernel: bad argument"); return SVTB200_ERR_BAD_ARG; }
    if (srcStrideRaw == 0 || refStride % srcStrideRaw) { set_error("SadLoopKernel: refStride must be a multiple of srcStrideRaw"); return SVTB200_ERR_UNSUPPORTED; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    const int saw = searchAreaWidth > 0 ? searchAreaWidth : 0, sah = searchAreaHeight > 0 ? searchAreaHeight : 0;
    if (saw == 0 || sah == 0) { *bestSad = 0xffffff; return SVTB200_OK; }
    const int row_step = (int)(refStride / srcStrideRaw);
    const size_t src_bytes = (size_t)(height - 1) * srcStride + width;
    const int wrows = sah + (int)(height - 1) * row_step;
    const size_t ref_bytes = (size_t)(wrows - 1) * srcStrideRaw + saw + width - 1;
    const int wp = win_pitch(saw + (int)width - 1), bp = ((int)width + 6) & ~3;
    const size_t smem = 64 + (((size_t)bp * height + 15) & ~15) + (size_t)wp * wrows + 96;
    if (smem > 227 * 1024) { set_error("SadLoopKernel: window of %zu B exceeds shared memory", smem); return SVTB200_ERR_UNSUPPORTED; }
    void *d_src, *d_ref, *d_out, *h_out;
    // keep the reference region's byte alignment so that the window loader's word path is exercised as in the picture kernel
    const size_t mis = (uintptr_t)ref & 3;
    if ((rc = c->dev_reserve(src_bytes + ref_bytes + 4096)) || (rc = c->dev_alloc(src_bytes, &d_src)) || (rc = c->dev_alloc(ref_bytes + 8, &d_ref)) ||
        (rc = c->dev_alloc(16, &d_out)) || (rc = c->pin_alloc(16, &h_out))) return rc;
    uint32_t *ho = (uint32_t *)h_out;
    ho[0] = 0; ho[1] = (uint32_t)(int)*xSearchCenter; ho[2] = (uint32_t)(int)*ySearchCenter;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync((uint8_t *)d_ref + mis, ref, ref_bytes, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_out, ho, 12, cudaMemcpyHostToDevice, c->stream));
    static thread_local size_t smem_set = 48 * 1024;
    if (smem > smem_set) { SVTB200_CUDA_TRY(cudaFuncSetAttribute(sad_loop_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set = smem; }
    sad_loop_leaf_kernel<<<1, NT, smem, c->stream>>>((const uint8_t *)d_src, (int)srcStride, (const uint8_t *)d_ref + mis, (int)refStride, (int)height,
                                                      (int)width, (int)srcStrideRaw, saw, sah, row_step, (uint32_t *)d_out);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    SVTB200_CUDA_TRY(cudaMemcpyAsync(ho, d_out, 12, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *bestSad = ho[0];
    *xSearchCenter = (int16_t)(int)ho[1];
    *ySearchCenter = (int16_t)(int)ho[2];
    return SVTB200_OK;
}

static int sad_leaf(const uint8_t *src, uint32_t ss, const uint8_t *r1, uint32_t r1s, const uint8_t *r2, uint32_t r2s, uint32_t h, uint32_t w, uint32_t *sad)
{
    if (!src || !r1 || !sad || h == 0 || w == 0) { set_error("sad: bad argument"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    const size_t sb = (size_t)(h - 1) * ss + w, b1 = (size_t)(h - 1) * r1s + w, b2 = r2 ? (size_t)(h - 1) * r2s + w : 0;
    void *d_s, *d_1, *d_2 = nullptr, *d_o, *h_o;
    if ((rc = c->dev_reserve(sb + b1 + b2 + 4096)) || (rc = c->dev_alloc(sb, &d_s)) || (rc = c->dev_alloc(b1, &d_1)) || (rc = c->dev_alloc(b2 + 4, &d_2)) ||
        (rc = c->dev_alloc(4, &d_o)) || (rc = c->pin_alloc(4, &h_o))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_s, src, sb, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_1, r1, b1, cudaMemcpyHostToDevice, c->stream));
    if (r2) SVTB200_CUDA_TRY(cudaMemcpyAsync(d_2, r2, b2, cudaMemcpyHostToDevice, c->stream));
    sad_nxm_leaf_kernel<<<1, NT, 0, c->stream>>>((const uint8_t *)d_s, (int)ss, (const uint8_t *)d_1, (int)r1s, r2 ? (const uint8_t *)d_2 : nullptr, (int)r2s, (int)h, (int)w, (uint32_t *)d_o);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    SVTB200_CUDA_TRY(cudaMemcpyAsync(h_o, d_o, 4, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *sad = *(uint32_t *)h_o;
    return SVTB200_OK;
}

int svtb200_NxMSadKernel(const uint8_t *src, uint32_t srcStride, const uint8_t *ref, uint32_t refStride, uint32_t height, uint32_t width, uint32_t *sad)
{
    return sad_leaf(src, srcStride, ref, refStride, nullptr, 0, height, width, sad);
}

int svtb200_CombinedAveragingSAD(const uint8_t *src, uint32_t srcStride, const uint8_t *ref1, uint32_t ref1Stride, const uint8_t *ref2,
                                 uint32_t ref2Stride, uint32_t height, uint32_t width, uint32_t *sad)
{
    if (!ref2) { set_error("CombinedAveragingSAD: ref2 is null"); return SVTB200_ERR_BAD_ARG; }
    return sad_leaf(src, srcStride, ref1, ref1Stride, ref2, ref2Stride, height, width, sad);
}

int svtb200_FullPelSearch_LCU(const uint8_t *src, uint32_t srcStride, const uint8_t *ref, uint32_t refStride, int16_t xSearchAreaOrigin,
                              int16_t ySearchAreaOrigin, uint32_t searchAreaWidth, uint32_t searchAreaHeight, uint32_t best_sad[SVTB200_ME_PU_COUNT],
                              uint32_t best_mv[SVTB200_ME_PU_COUNT])
{
    if (!src || !ref || !best_sad || !best_mv || searchAreaWidth == 0 || searchAreaHeight == 0 || searchAreaWidth > 127 || searchAreaHeight > 127) { set_error("FullPelSearch_LCU: bad argument"); return SVTB200_ERR_BAD_ARG; }
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init())) return rc;
    c->reset();
    const int saw = (int)searchAreaWidth, sah = (int)searchAreaHeight;
    const size_t sb = (size_t)63 * srcStride + 64, rb = (size_t)(sah + 62) * refStride + saw + 63;
    const size_t mis = (uintptr_t)ref & 3;
    void *d_s, *d_r, *d_bs, *d_bm, *h_b;
    if ((rc = c->dev_reserve(sb + rb + 8192)) || (rc = c->dev_alloc(sb, &d_s)) || (rc = c->dev_alloc(rb + 8, &d_r)) || (rc = c->dev_alloc(85 * 4, &d_bs)) ||
        (rc = c->dev_alloc(85 * 4, &d_bm)) || (rc = c->pin_alloc(170 * 4, &h_b))) return rc;
    uint32_t *hb = (uint32_t *)h_b;
    memcpy(hb, best_sad, 85 * 4); memcpy(hb + 85, best_mv, 85 * 4);
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_s, src, sb, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync((uint8_t *)d_r + mis, ref, rb, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_bs, hb, 85 * 4, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_bm, hb + 85, 85 * 4, cudaMemcpyHostToDevice, c->stream));
    const size_t smem = 4096 + CHUNK * 64 * 2 + CHUNK * 21 * 4 + 172 * 4 + (size_t)win_pitch(saw + 63) * (sah + 63) + 96;
    static thread_local size_t smem_set = 48 * 1024;
    if (smem > smem_set) { SVTB200_CUDA_TRY(cudaFuncSetAttribute(fullpel_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set = smem; }
    fullpel_leaf_kernel<<<1, NT, smem, c->stream>>>((const uint8_t *)d_s, (int)srcStride, (const uint8_t *)d_r + mis, (int)refStride, (int)xSearchAreaOrigin,
                                                     (int)ySearchAreaOrigin, saw, sah, (uint32_t *)d_bs, (uint32_t *)d_bm);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    SVTB200_CUDA_TRY(cudaMemcpyAsync(hb, d_bs, 85 * 4, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(hb + 85, d_bm, 85 * 4, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(best_sad, hb, 85 * 4); memcpy(best_mv, hb + 85, 85 * 4);
    return SVTB200_OK;
}

}  // extern "C"
