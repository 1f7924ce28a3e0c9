/*
 * svt_hevc_b200.h -- C ABI of libsvthevc_b200.so, the B200 (sm_100a) implementation of the
 * SVT-HEVC per-CTU data-parallel hot path.
 *
 * This library takes the place of the reference's ASM_TYPES function-pointer dispatch
 * (Source/Lib/Codec/EbDefinitions.h:350-364, tables in EbComputeSAD.h:95-175,
 * EbMeSadCalculation.h:106-121, EbMcp.h/EbMcpTables.c, EbTransforms.h:271-510,
 * EbIntraPrediction.h:470-642, EbDeblockingFilter.h:245-268, EbSampleAdaptiveOffset.h:171-335,
 * EbPackUnPack.h:27-164).  Two kinds of entry points:
 *
 *   (1) leaf entry points `svtb200_<ReferenceName>`: the EXACT prototype of the reference
 *       typedef they replace (host pointers, by-value scalars).  One call = one H2D copy,
 *       one kernel, one D2H copy.  They exist so that every kernel can be unit-compared 1:1
 *       with the reference's C/SIMD rows and dropped into a table row for debugging.
 *   (2) batched / picture-level entry points: one call processes every CTU (or every job)
 *       of a picture.  These are what MotionEstimationKernel (EbMotionEstimationProcess.c:603)
 *       and EncDecKernel (EbEncDecProcess.c:2630) call through the shim.
 *       Each has a `_dev` twin that takes device pointers + a CUDA stream so a caller that
 *       keeps pictures resident in HBM pays no copies.
 *
 * Conventions: plain C types only; all sizes in samples unless noted; every function returns
 * SVTB200_OK (0) or a negative error.  There is NO CPU fallback: without a CUDA device every
 * compute entry point returns SVTB200_ERR_NO_DEVICE.
 */
#ifndef SVT_HEVC_B200_H
#define SVT_HEVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVTB200_OK               0
#define SVTB200_ERR_NO_DEVICE   -1   /* maps to EB_ErrorInsufficientResources at EbInitEncoder */
#define SVTB200_ERR_CUDA        -2
#define SVTB200_ERR_BAD_ARG     -3   /* maps to EB_ErrorBadParameter */
#define SVTB200_ERR_UNSUPPORTED -4

#define SVTB200_ME_PU_COUNT 85      /* MAX_ME_PU_COUNT, EbDefinitions.h */
#define SVTB200_LCU_SIZE    64      /* MAX_LCU_SIZE,   EbDefinitions.h:827 */

/* ---------------------------------------------------------------------------------------
 * Library lifetime (hook: EbInitEncoder right after ASM_TYPES is set, EbEncHandle.c:702-707)
 * ------------------------------------------------------------------------------------- */
int         svtb200_init(int device_ordinal);      /* creates context, streams, staging pools */
void        svtb200_shutdown(void);
int         svtb200_device_count(void);            /* 0 when no CUDA device is visible        */
const char *svtb200_version(void);
const char *svtb200_last_error(void);
/* number of kernel launches issued by this library since load (bench.py "gpu_launches") */
uint64_t    svtb200_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Plane descriptor: mirror of the luma part of EbPictureBufferDesc_t
 * (EbPictureBufferDesc.h:26-70): buf points at the top-left of the PADDED allocation,
 * the visible picture starts at (origin_x, origin_y), row pitch is `stride` samples.
 * Indices computed by the reference (origin + x + (origin + y) * stride) are valid as-is.
 * ------------------------------------------------------------------------------------- */
typedef struct svtb200_plane {
    uint8_t *buf;        /* host pointer (plain API) or device pointer (_dev API) */
    uint32_t stride;
    uint16_t origin_x;
    uint16_t origin_y;
    uint16_t width;
    uint16_t height;
} svtb200_plane;

/* =======================================================================================
 *  A. Motion estimation
 * ===================================================================================== */

/* a1: leaf, prototype of EB_SADLOOPKERNELNxM_TYPE (EbComputeSAD.h:36-48);
 *     semantics of SadLoopKernel (C_DEFAULT/EbComputeSAD_C.c:165). The caller passes the size
 *     of the readable reference region so the host->device copy is bounded:
 *     the kernel reads ref[x + y*refStride + sy*srcStrideRaw] for x < width+searchAreaWidth-1. */
int svtb200_SadLoopKernel(
    const uint8_t *src, uint32_t srcStride,
    const uint8_t *ref, uint32_t refStride,
    uint32_t height, uint32_t width,
    uint64_t *bestSad, int16_t *xSearchCenter, int16_t *ySearchCenter,
    uint32_t srcStrideRaw, int16_t searchAreaWidth, int16_t searchAreaHeight);

/* a4: leaf, EB_SADKERNELNxM_TYPE (EbComputeSAD.h:26-32), FastLoop_NxMSadKernel
 *     (C_DEFAULT/EbComputeSAD_C.c:141).  Returns the SAD through *sad. */
int svtb200_NxMSadKernel(
    const uint8_t *src, uint32_t srcStride, const uint8_t *ref, uint32_t refStride,
    uint32_t height, uint32_t width, uint32_t *sad);

/* a5: leaf, CombinedAveragingSAD (C_DEFAULT/EbComputeSAD_C.c:14). */
int svtb200_CombinedAveragingSAD(
    const uint8_t *src, uint32_t srcStride, const uint8_t *ref1, uint32_t ref1Stride,
    const uint8_t *ref2, uint32_t ref2Stride, uint32_t height, uint32_t width, uint32_t *sad);

/* a2+a3: leaf, one call of FullPelSearch_LCU (EbMotionEstimation.c:586) over a
 *     searchAreaWidth x searchAreaHeight window: the 8-position path
 *     (GetEightHorizontalSearchPointResults_8x8_16x16_PU / _32x32_64x64,
 *     C_DEFAULT/EbComputeSAD_C.c:244,355) for the multiple-of-8 part and the single-position
 *     path (SadCalculation_8x8_16x16 / _32x32_64x64, C_DEFAULT/EbMeSadCalculation_C.c:16,73)
 *     for the remainder.  best_sad/best_mv are the 85-entry pLcuBestSad/pLcuBestMV arrays
 *     (n-order: 64x64, 4x32x32, 16x16x16, 64x8x8 in the reference's tab order,
 *     EbMotionEstimationContext.h:352-360); they are READ (running best) and UPDATED.
 *     src = 64x64 LCU (stride srcStride); ref points at the top-left search position. */
int svtb200_FullPelSearch_LCU(
    const uint8_t *src, uint32_t srcStride, const uint8_t *ref, uint32_t refStride,
    int16_t xSearchAreaOrigin, int16_t ySearchAreaOrigin,
    uint32_t searchAreaWidth, uint32_t searchAreaHeight,
    uint32_t best_sad[SVTB200_ME_PU_COUNT], uint32_t best_mv[SVTB200_ME_PU_COUNT]);

/* Per-picture ME configuration: the fields MotionEstimateLcu (EbMotionEstimation.c:3671)
 * reads from MeContext_t (EbMotionEstimationContext.h:309-464), PictureParentControlSet_t and
 * SequenceControlSet_t after SignalDerivationMeKernelOq (EbMotionEstimationProcess.c:308) has run. */
typedef struct svtb200_me_config {
    uint16_t pic_width, pic_height;        /* lumaWidth / lumaHeight                           */
    uint8_t  num_lists;                    /* 1 = P picture (list 0), 2 = B picture            */
    uint8_t  temporal_layer;               /* temporalLayerIndex                               */
    uint8_t  hierarchical_levels;          /* hierarchicalLevels (row of the multiplier table) */
    uint8_t  enable_hme;                   /* enableHmeFlag                                    */
    uint8_t  enable_hme_l0, enable_hme_l1, enable_hme_l2;
    uint8_t  one_quadrant_hme;             /* oneQuadrantHME                                   */
    uint8_t  update_hme_search_center;     /* updateHmeSearchCenter (TRUE for >576p, tune 1)   */
    uint8_t  n_hme_regions_w, n_hme_regions_h;   /* numberHmeSearchRegionInWidth/Height (<=2)  */
    uint16_t hme_l0_total_w, hme_l0_total_h;
    uint16_t hme_l0_w[2], hme_l0_h[2];
    uint16_t hme_l1_w[2], hme_l1_h[2];
    uint16_t hme_l2_w[2], hme_l2_h[2];
    uint8_t  search_area_w, search_area_h; /* searchAreaWidth / searchAreaHeight               */
    uint8_t  fractional_search_model;      /* 0 all, 1 selective, 2 off (integer-pel only)     */
    uint8_t  cu8x8_mode, cu16x16_mode;     /* CU_8x8_MODE_x / CU_16x16_MODE_x                  */
    uint8_t  rate_control_mode;            /* staticConfig.rateControlMode                     */
    uint8_t  fractional_search_method;     /* MeContext_t::fractionalSearchMethod: 0 SUB_SAD_SEARCH, 1 FULL_SAD_SEARCH (2 SSD_SEARCH: unsupported) */
    uint32_t lambda;                       /* MeContext_t::lambda                              */
    uint32_t mvd_bits[12];                 /* MeContext_t::mvdBitsArray[NUMBER_OF_MVD_CASES]   */
    uint64_t ref_poc[2];                   /* refPicPocArray                                   */
} svtb200_me_config;

/* Per-PU output record: the information MotionEstimateLcu writes into MeCuResults_t
 * (EbMotionEstimationLcuResults.h:50-74, written at EbMotionEstimation.c:4348-4438),
 * without C bitfields so the layout is ABI-stable. 24 bytes. */
typedef struct svtb200_me_pu_result {
    int16_t  x_mv_l0, y_mv_l0, x_mv_l1, y_mv_l1;  /* quarter-pel units */
    uint32_t distortion[3];
    uint8_t  direction[3];                          /* 0 = L0, 1 = L1, 2 = BI */
    uint8_t  total_me_candidates;
} svtb200_me_pu_result;

/* Source ("picture analysis") pictures used by ME: full, 1/4 and 1/16 resolution luma
 * (EbPaReferenceObject_t: inputPaddedPicturePtr, quarterDecimatedPicturePtr,
 * sixteenthDecimatedPicturePtr, EbReferenceObject.c:182-195). */
typedef struct svtb200_pa_picture {
    svtb200_plane full, quarter, sixteenth;
} svtb200_pa_picture;

/* a6: picture-level ME: every LCU of the picture, every list.
 *   results        [n_lcu][85] in PU order (0 = 64x64, 1-4 = 32x32, 5-20 = 16x16, 21-84 = 8x8 raster)
 *   lcu_best_sad/mv optional [n_lcu][2][85] raw per-list pLcuBestSad/pLcuBestMV (n-order); may be NULL
 *   rc_me_distortion optional [n_lcu] (rcMEdistortion), may be NULL
 * Host variant copies the planes in and the results out (e2e); _dev variant takes device
 * pointers in every svtb200_plane::buf and output pointer, and launches on `cuda_stream`. */
int svtb200_me_picture(const svtb200_me_config *cfg, const svtb200_pa_picture *src,
                       const svtb200_pa_picture *ref_l0, const svtb200_pa_picture *ref_l1,
                       svtb200_me_pu_result *results, uint32_t *lcu_best_sad, uint32_t *lcu_best_mv,
                       uint32_t *rc_me_distortion);
int svtb200_me_picture_dev(const svtb200_me_config *cfg, const svtb200_pa_picture *src,
                           const svtb200_pa_picture *ref_l0, const svtb200_pa_picture *ref_l1,
                           svtb200_me_pu_result *results, uint32_t *lcu_best_sad, uint32_t *lcu_best_mv,
                           uint32_t *rc_me_distortion, void *cuda_stream);


/* =======================================================================================
 *  B. Motion-compensated prediction
 * ===================================================================================== */

/* Three-plane 8-bit 4:2:0 picture: mirror of EbPictureBufferDesc_t (bufferY/Cb/Cr, strideY/Cb/Cr,
 * originX/originY; the chroma planes carry half the luma padding, EbPictureBufferDesc.c). */
typedef struct svtb200_yuv_picture {
    svtb200_plane y, cb, cr;
} svtb200_yuv_picture;

/* One 2Nx2N prediction unit: the arguments of EncodePassInterPrediction
 * (Codec/EbInterPrediction.c:761): MvUnit_t {mv[2], predDirection}, puOrigin, puWidth/Height,
 * plus which reference picture each list points at (index into the refs[] array of the batch). */
typedef struct svtb200_mc_job {
    uint16_t x, y;          /* PU origin, luma samples, picture coordinates          */
    uint8_t  w, h;          /* PU size in luma samples: 8, 16, 32 or 64              */
    uint8_t  dir;           /* 0 = UNI_PRED_LIST_0, 1 = UNI_PRED_LIST_1, 2 = BI_PRED */
    uint8_t  ref_idx[2];    /* refs[] index used by list 0 / list 1                  */
    uint8_t  reserved;
    int16_t  mv[2][2];      /* [list][x,y], quarter-pel                              */
} svtb200_mc_job;

/* Batched HEVC inter prediction of luma + both chroma planes for n_jobs PUs into `dst` at each PU's own
 * position (what EncodePassInterPrediction writes into the prediction picture): position clamp
 * (EbInterPrediction.c:794-808), 8-tap luma / 4-tap chroma separable interpolation, uni-pred rounding or
 * bi-pred averaging (EncodeUniPredInterpolation Codec/EbMcp.c:175, EncodeBiPredInterpolation :562,
 * kernels C_DEFAULT/EbMcp_C.c).  Host variant: refs/dst are host pictures (copied in/out);
 * _dev variant: device pictures + device job array. */
int svtb200_mc_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs,
                     const svtb200_yuv_picture *dst);
int svtb200_mc_batch_dev(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs_dev, int n_jobs,
                         const svtb200_yuv_picture *dst, void *cuda_stream);
/* b5: the same batch on 16-bit sample planes (10-bit video): EncodePassInterPrediction16bit (Codec/EbInterPrediction.c:928),
 * UniPredInterpolation16bit / BiPredInterpolation16bit (Codec/EbMcp.c:249,804), kernels C_DEFAULT/EbMcp_C.c:3121-4861.
 * svtb200_plane::buf points at uint16_t samples; stride and origins stay in samples (EbReferenceObject_t::referencePicture16bit geometry). */
int svtb200_mc_batch16(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs,
                       const svtb200_yuv_picture *dst);
int svtb200_mc_batch16_dev(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs_dev, int n_jobs,
                           const svtb200_yuv_picture *dst, void *cuda_stream);

/* b6: the mode-decision fast loop's interpolation-free prediction (INTERPOLATION_FREE_PATH, Codec/EbEncDecProcess.c:2728, pictures with
 * useSubpelFlag == 0): UniPredIFreeRef8Bit / BiPredIFreeRef8Bit (Codec/EbAvcStyleMcp.c:172,321) behind Inter2Nx2NPuPredictionInterpolationFree
 * (Codec/EbInterPrediction.c:60-400) -- the copy at the integer position the vector's fraction maps to, bi-prediction = rounded average of the
 * two copies.  Same job record, same position clamp, same destination convention as svtb200_mc_batch (the fast loop writes the candidate's
 * prediction buffer; here every PU lands at its own position of `dst`). */
int svtb200_mc_ifree_batch(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs, int n_jobs, const svtb200_yuv_picture *dst);
int svtb200_mc_ifree_batch_dev(const svtb200_yuv_picture *refs, int n_refs, const svtb200_mc_job *jobs_dev, int n_jobs,
                               const svtb200_yuv_picture *dst, void *cuda_stream);

/* Leaf entry points with the reference's kernel prototypes.  The reference selects one of 16 (luma) / 64
 * (chroma) functions by fractional position (uniPredLumaIFFunctionPtrArrayNew[2][16], EbMcpTables.c:28;
 * uniPredChromaIFFunctionPtrArrayNew[2][64], :195); here the position is an argument.
 * frac = fracPosx + (fracPosy << 2) for luma; chroma takes fracPosx, fracPosy in 1/8 units.
 * refPic points at the integer sample of the block's top-left (the kernels read 3 left/above, 4 right/below). */
int svtb200_LumaInterpolationFilterNew(uint32_t frac, const uint8_t *refPic, uint32_t srcStride, uint8_t *dst,
                                       uint32_t dstStride, uint32_t puWidth, uint32_t puHeight);
int svtb200_ChromaInterpolationFilterNew(uint32_t fracPosx, uint32_t fracPosy, const uint8_t *refPic, uint32_t srcStride,
                                         uint8_t *dst, uint32_t dstStride, uint32_t puWidth, uint32_t puHeight);
/* bi-pred: raw 14-bit intermediates of both lists (biPred*IFFunctionPtrArrayNew) + BiPredClipping (C_DEFAULT/EbMcp_C.c:21) */
int svtb200_LumaBiPredInterpolation(uint32_t frac0, const uint8_t *ref0, uint32_t ref0Stride, uint32_t frac1, const uint8_t *ref1,
                                    uint32_t ref1Stride, uint8_t *dst, uint32_t dstStride, uint32_t puWidth, uint32_t puHeight);
int svtb200_ChromaBiPredInterpolation(uint32_t frac0x, uint32_t frac0y, const uint8_t *ref0, uint32_t ref0Stride, uint32_t frac1x,
                                      uint32_t frac1y, const uint8_t *ref1, uint32_t ref1Stride, uint8_t *dst, uint32_t dstStride,
                                      uint32_t puWidth, uint32_t puHeight);

/* =======================================================================================
 *  C. Residual -> transform -> quant / dequant -> inverse transform -> reconstruction
 *     
 * ===================================================================================== */

#define SVTB200_SHAPE_DEFAULT 0   /* DEFAULT_SHAPE, EbDefinitions.h (EB_TRANS_COEFF_SHAPE) */
#define SVTB200_SHAPE_N2      1   /* N2_SHAPE: low-frequency N/2 x N/2 quadrant only       */
#define SVTB200_SHAPE_N4      2   /* N4_SHAPE: low-frequency N/4 x N/4                     */
#define SVTB200_SHAPE_ONLY_DC 3   /* ONLY_DC_SHAPE                                          */

#define SVTB200_TQ_CLEAN_SPARSE   1u  /* cleanSparseCeoffPfEncDec (EbTransforms.c:3199-3243)              */
#define SVTB200_TQ_ENABLE_CBF     2u  /* enableCbflag of UpdateQiQCoef (C_DEFAULT/EbTransforms_C.c:209)   */
#define SVTB200_TQ_CONTOURING     4u  /* enableContouringQCUpdateFlag                                      */
#define SVTB200_TQ_SKIP_RECON     8u  /* quantise only (MD-style call): do not run the inverse / addition  */

/* One transform unit of one colour plane: what EncodeLoop + EncodeGenerateRecon
 * (Codec/EbCodingLoop.c:651,1084) do for it on the 8-bit path. */
typedef struct svtb200_tq_job {
    uint16_t x, y;          /* TU origin in samples of ITS OWN plane (chroma: already halved)         */
    uint8_t  plane;         /* 0 = Y, 1 = Cb, 2 = Cr                                                  */
    uint8_t  size;          /* 4, 8, 16 or 32                                                          */
    uint8_t  dst4x4;        /* 1: 4x4 DST (intra luma 4x4), EstimateTransform dstTansformFlag          */
    uint8_t  shape;         /* SVTB200_SHAPE_*                                                         */
    uint8_t  qp;            /* cu qp (luma) or mapped chroma qp, 0..51                                 */
    uint8_t  slice_type;    /* 0 = B, 1 = P, 2 = I (EB_B/P/I_PICTURE, EbApi.h:36-38), 3 = IDR          */
    uint8_t  temporal_layer;
    uint8_t  flags;         /* SVTB200_TQ_*                                                            */
    uint32_t dz_offset;     /* dZoffset of UnifiedQuantizeInvQuantize (0 = use the I/P rounding offset) */
} svtb200_tq_job;

/* Quantised coefficients, picture geometry (no padding): plane p, sample (x,y) at buf[p][y*stride[p] + x].
 * (The reference keeps them per LCU with pitch 64, EbCodingLoop.c:740; the shim copies rows, INTEGRATION.md.) */
typedef struct svtb200_coeff_picture {
    int16_t *buf[3];
    uint32_t stride[3];
} svtb200_coeff_picture;

/* Batched encode-pass chain.  For every job: residual = src - pred (PictureResidual), EstimateTransform,
 * UnifiedQuantizeInvQuantize default path (QuantizeInvQuantize + optional sparse cleaning + UpdateQiQCoef),
 * PfZeroOutUselessQuadrants, EncodeInvTransform, AdditionKernel.  recon may alias pred (the reference
 * reconstructs in place).  nz_counts[job] receives countNonZeroCoeffs.  Host and device variants. */
int svtb200_tq_batch(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                     const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts);
int svtb200_tq_batch_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                         const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs_dev, int n_jobs, uint32_t *nz_counts_dev,
                         void *cuda_stream);
/* the same chain on 16-bit sample planes (10-bit video): EncodeLoop16bit / EncodeGenerateRecon16bit (Codec/EbCodingLoop.c:1240,1640) --
 * PictureResidual16bit, EncodeTransform with BIT_INCREMENT_10BIT (EbTransforms.c:3343), UnifiedQuantizeInvQuantize at EB_10BIT with
 * qp + QP_BD_OFFSET, EncodeInvTransform, PictureAdditionKernel16bit (clip to 1023).  job.qp stays the 8-bit-scale qp (0..51).
 * svtb200_plane::buf points at uint16_t samples; strides and origins in samples. */
int svtb200_tq_batch16(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                       const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts);
int svtb200_tq_batch16_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                           const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs_dev, int n_jobs, uint32_t *nz_counts_dev,
                           void *cuda_stream);

/* Coefficient hand-off to the entropy coder.  EntropyCodingKernel reads coeffSamplesTB only for TUs whose cbf is set
 * (EbCodingLoop.c:745,767-803; EbEntropyCoding.c), so the coefficient planes of svtb200_tq_batch*_dev are packed on the device into one dense
 * stream -- TU after TU in job order, size x size int16 row-major, TUs with nz_counts == 0 skipped -- and only that stream crosses PCIe:
 *   offsets_dev[j]       first element of job j in the stream (meaningful when nz_counts[j] != 0)
 *   offsets_dev[n_jobs]  total elements; when it exceeds stream_capacity the stream is incomplete (TUs that did not fit are skipped)
 * offsets_dev holds n_jobs + 1 entries. */
int svtb200_tq_pack_coeffs_dev(const svtb200_coeff_picture *coeff_dev, const svtb200_tq_job *jobs_dev, const uint32_t *nz_counts_dev, int n_jobs,
                               int16_t *stream_dev, uint32_t stream_capacity, uint32_t *offsets_dev, void *cuda_stream);

/* ---- distortion leaves of mode decision (the first bricks of row N2) ----------------------------------------------
 * The full loop evaluates them per candidate block; here they are batched, one job per block, one result per job.
 *   spatial_distortion   SpatialFullDistortionKernel (C_DEFAULT/EbPictureOperators_C.c:697; table Codec/EbPictureOperators.h:633; callers
 *                        Codec/EbFullLoop.c:384-392,950-975 and the SSD refinement Codec/EbMotionEstimation.c:799-979): sum (input - recon)^2
 *   full_distortion      FullDistortionKernel_32bit / ..Intra_32bit / ..CbfZero_32bit (C_DEFAULT/EbPictureOperators_C.c:385-475; table
 *                        Codec/EbPictureOperators.h:520-606): out[2j] / out[2j+1] = distortionResult[0] / [1]; 32-bit wrapping sums like both rows
 *   satd                 Compute8x8Satd / Compute4x4Satd summed over the block's 8x8 / 4x4 units (C_DEFAULT/EbPictureOperators_C.c:482,
 *                        Codec/EbHmCode.c:41; drivers Codec/EbProductCodingLoop.c:2642) on the residual src - pred
 * Blocks lie inside the picture; origin / dimensions are multiples of 4 (8 for SATD_8X8 jobs); coefficient jobs need even x. */
#define SVTB200_FULLDIST_DEFAULT  0   /* [0] = sum (coeff - recon)^2, [1] = sum coeff^2 */
#define SVTB200_FULLDIST_INTRA    1   /* [0] = [1] = sum (coeff - recon)^2              */
#define SVTB200_FULLDIST_CBF_ZERO 2   /* [0] = [1] = sum coeff^2                        */
#define SVTB200_SATD_8X8 0
#define SVTB200_SATD_4X4 1
typedef struct svtb200_dist_job {
    uint16_t x, y;          /* block origin in samples of its own plane */
    uint8_t  plane;         /* 0 = Y, 1 = Cb, 2 = Cr                     */
    uint8_t  width, height; /* 4 .. 64                                   */
    uint8_t  kind;          /* SVTB200_FULLDIST_* / SVTB200_SATD_* (ignored by spatial_distortion) */
} svtb200_dist_job;         /* 8 bytes */
int svtb200_spatial_distortion_batch(const svtb200_yuv_picture *input, const svtb200_yuv_picture *recon, const svtb200_dist_job *jobs, int n_jobs,
                                     uint64_t *out);
int svtb200_spatial_distortion_batch_dev(const svtb200_yuv_picture *input, const svtb200_yuv_picture *recon, const svtb200_dist_job *jobs_dev, int n_jobs,
                                         uint64_t *out_dev, void *cuda_stream);
int svtb200_full_distortion_batch(const svtb200_coeff_picture *coeff, const svtb200_coeff_picture *recon_coeff, uint32_t width, uint32_t height,
                                  const svtb200_dist_job *jobs, int n_jobs, uint64_t *out);
int svtb200_full_distortion_batch_dev(const svtb200_coeff_picture *coeff, const svtb200_coeff_picture *recon_coeff, const svtb200_dist_job *jobs_dev,
                                      int n_jobs, uint64_t *out_dev, void *cuda_stream);
int svtb200_satd_batch(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_dist_job *jobs, int n_jobs, uint64_t *out);
int svtb200_satd_batch_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_dist_job *jobs_dev, int n_jobs,
                           uint64_t *out_dev, void *cuda_stream);

/* Leaf entry points (reference prototypes; host pointers). */
/* EstimateTransform, Codec/EbTransforms.c:3268 (EB_TRANSFORM_FUNC tables EbTransforms.h:339-472) */
int svtb200_EstimateTransform(const int16_t *residualBuffer, uint32_t residualStride, int16_t *coeffBuffer, uint32_t coeffStride,
                              uint32_t transformSize, uint32_t bitIncrement, uint8_t dstTransformFlag, uint8_t transCoeffShape);
/* EB_QIQ_TYPE (EbTransforms.h:213-225), QuantizeInvQuantize C_DEFAULT/EbTransforms_C.c:89 */
int svtb200_QuantizeInvQuantize(const int16_t *coeff, uint32_t coeffStride, int16_t *quantCoeff, int16_t *reconCoeff, uint32_t qFunc,
                                uint32_t q_offset, int32_t shiftedQBits, int32_t shiftedFFunc, int32_t iq_offset, int32_t shiftNum,
                                uint32_t areaSize, uint32_t *nonzerocoeff);
/* EncodeInvTransform, Codec/EbTransforms.c:3502 (isOnlyDc gives identical samples, so it is not an argument) */
int svtb200_EncodeInvTransform(const int16_t *coeffBuffer, uint32_t coeffStride, int16_t *reconBuffer, uint32_t reconStride,
                               uint32_t transformSize, uint32_t bitIncrement, uint8_t dstTransformFlag);

/* =======================================================================================
 *  E. In-loop deblocking filter
 * ===================================================================================== */

/* Picture-level deblocking parameters: the PictureControlSet_t fields the reference's DLF drivers read
 * (Codec/EbDeblockingFilter.c:2222,2828,3518: qpArrayStride, tcOffset, betaOffset, cbQpOffset, crQpOffset). */
typedef struct svtb200_dlf_params {
    uint16_t pic_width, pic_height;   /* luma, multiples of 8                                             */
    uint16_t qp_array_stride;         /* entries per row of qp_array (one entry per 8x8 luma block)       */
    int8_t   tc_offset, beta_offset;  /* slice-level offsets (already multiplied by 2, as the PCS stores) */
    int8_t   cb_qp_offset, cr_qp_offset;
} svtb200_dlf_params;

/* Deblocks a whole picture in place, HEVC order: every vertical edge of the picture, then every horizontal edge on
 * the result.  This equals what the reference obtains LCU by LCU with LCUInternalAreaDLFCore + LCUBoundaryDLFCore +
 * LCUPictureEdgeDLFCore (Codec/EbCodingLoop.c:4604-4635) once all LCUs are done.
 *   vertical_bs / horizontal_bs : [n_lcu][256] boundary strengths, LCU raster order, entry (x>>2) + (y>>2)*16 inside the
 *       64x64 LCU (verticalEdgeBSArray / horizontalEdgeBSArray, BLK4X4_ADDR_TO_*_EDGE_BS_ARRAY_IDX EbDeblockingFilter.h:19-20);
 *       a vertical entry is the strength of the edge on the LEFT side of that 4x4 block, a horizontal one of its TOP side
 *   qp_array : one qp per 8x8 luma block (SetQpArrayBasedOnCU :50)
 * Edge filters: Luma4SampleEdgeDLFCore / Chroma2SampleEdgeDLFCore (C_DEFAULT/EbDeblockingFilter_C.c:39,469). */
int svtb200_dlf_picture(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs,
                        const uint8_t *qp_array, const svtb200_dlf_params *prm);
int svtb200_dlf_picture_dev(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs,
                            const uint8_t *qp_array, const svtb200_dlf_params *prm, void *cuda_stream);
/* the same on 16-bit sample planes (10-bit video): LCUInternalAreaDLFCore16bit / LCUBoundaryDLFCore16bit / LCUPictureEdgeDLFCore16bit
 * (Codec/EbDeblockingFilter.c:2514,3174,3922), edge cores C_DEFAULT/EbDeblockingFilter_C.c:240,524 -- tc and beta scaled by 4, clip to 1023 */
int svtb200_dlf_picture16(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs,
                          const uint8_t *qp_array, const svtb200_dlf_params *prm);
int svtb200_dlf_picture16_dev(const svtb200_yuv_picture *recon, const uint8_t *vertical_bs, const uint8_t *horizontal_bs,
                              const uint8_t *qp_array, const svtb200_dlf_params *prm, void *cuda_stream);

/* N3: the boundary strengths on the device, so that deblocking needs no host-built arrays.  CalculateBSForPUBoundary (Codec/EbDeblockingFilter.c:109),
 * SetBSArrayBasedOnPUBoundary (:339), SetBSArrayBasedOnTUBoundary (:472), call sites Codec/EbCodingLoop.c:3554,4329,4364.  The encode pass sets a CU's
 * strengths when that CU is coded, from its neighbour arrays; its own cbf entries and everything left of / above it are final by then, so the result
 * is a function of the picture's final maps:
 *   cu_map  : one svtb200_cu_info per 8x8 luma block, raster, cu_stride entries per row -- CodingUnit_t::predictionModeFlag, PredictionUnit_t::
 *             interPredDirectionIndex and mv[2] (quarter-pel) of the CU covering the block, and that CU's size
 *   cbf_map : one byte per 4x4 luma block (PictureControlSet_t::cbfMapArray: lumaCbf of the TU covering it), cbf_stride entries per row
 *   slice_type 0 = B, 1 = P, 2 = I; ref_poc_l0 / l1 = refPOC of refPicPtrArray[REF_LIST_0 / 1]
 * Output: the [n_lcu][256] arrays svtb200_dlf_picture consumes.  One tile; transform depth 0 except 64x64 CUs (four 32x32 TUs), as the encoder codes. */
typedef struct svtb200_cu_info {
    int16_t mv[2][2];       /* [list][x, y], quarter-pel */
    uint8_t pred_dir;       /* 0 = list 0, 1 = list 1, 2 = bi (inter CUs) */
    uint8_t intra;          /* predictionModeFlag == INTRA_MODE */
    uint8_t log2_cu_size;   /* 3 .. 6 */
    uint8_t reserved;
} svtb200_cu_info;          /* 12 bytes */
int svtb200_dlf_bs_picture(const svtb200_cu_info *cu_map, uint32_t cu_stride, const uint8_t *cbf_map, uint32_t cbf_stride, uint16_t pic_width,
                           uint16_t pic_height, uint8_t slice_type, uint64_t ref_poc_l0, uint64_t ref_poc_l1, uint8_t *vertical_bs,
                           uint8_t *horizontal_bs);
int svtb200_dlf_bs_picture_dev(const svtb200_cu_info *cu_map_dev, uint32_t cu_stride, const uint8_t *cbf_map_dev, uint32_t cbf_stride,
                               uint16_t pic_width, uint16_t pic_height, uint8_t slice_type, uint64_t ref_poc_l0, uint64_t ref_poc_l1,
                               uint8_t *vertical_bs_dev, uint8_t *horizontal_bs_dev, void *cuda_stream);

/* Leaf entry points, reference prototypes (EbDeblockingFilter.h:245-273). In-place on host memory; the caller states how
 * far the plane extends around the edge start via the fixed footprint: 4 samples either side, 4 (luma) / 2 (chroma) along. */
int svtb200_Luma4SampleEdgeDLFCore(uint8_t *edgeStartFilteredSamplePtr, uint32_t reconLumaPicStride, uint8_t isVerticalEdge,
                                   int32_t tc, int32_t beta);
int svtb200_Chroma2SampleEdgeDLFCore(uint8_t *edgeStartSampleCb, uint8_t *edgeStartSampleCr, uint32_t reconChromaPicStride,
                                     uint8_t isVerticalEdge, uint8_t cbTc, uint8_t crTc);

/* =======================================================================================
 *  F. Sample adaptive offset (the RD decision f2 stays on the host)
 * ===================================================================================== */

/* Per-LCU, per-component statistics: the outputs of GatherSaoStatisticsLcuLossy_62x62
 * (C_DEFAULT/EbSampleAdaptiveOffset_C.c:125; SaoStats_t in EbSampleAdaptiveOffsetGenerationDecision.h). */
typedef struct svtb200_sao_stats {
    int32_t  bo_diff[32];
    uint16_t bo_count[32];
    int32_t  eo_diff[4][5];     /* [EO type 0/90/135/45][category], categories already compacted as the reference does */
    uint16_t eo_count[4][5];
} svtb200_sao_stats;

/* Statistics of every LCU and component of a picture: out[(lcu * 3 + component)].  `recon` is read as it is -- the
 * caller decides which deblocking state it holds (the reference gathers them per LCU right after that LCU's DLF,
 * SURVEY §7 H3 / Codec/EbCodingLoop.c:4692-4748; on pictures whose DLF is off that state is simply the un-deblocked recon). */
int svtb200_sao_stats_picture(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out);
int svtb200_sao_stats_picture_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out_dev, void *cuda_stream);
/* H3: the statistics the encoder gathers on a picture WITH deblocking -- LCU n is gathered right after its own deblocking step
 * (Codec/EbCodingLoop.c:4600-4760), when its last 4 columns / rows still wait for the right / bottom neighbour: `deblocked` everywhere except those
 * strips, which are read from `unfiltered` (same geometry; the reconstruction before svtb200_dlf_picture). */
int svtb200_sao_stats_picture_h3(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered,
                                 svtb200_sao_stats *out);
int svtb200_sao_stats_picture_h3_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered,
                                     svtb200_sao_stats *out_dev, void *cuda_stream);

/* Per-LCU SAO parameters: mirror of SaoParameters_t (Codec/EbCodingUnit.h:137-146). */
typedef struct svtb200_sao_params {
    uint8_t type[2];            /* saoTypeIndex[luma, chroma]: 0 off, 1 EO-0, 2 EO-90, 3 EO-135, 4 EO-45, 5 BO */
    uint8_t band_position[3];   /* saoBandPosition[Y, Cb, Cr]                                                   */
    int8_t  offset[3][4];       /* saoOffset[Y, Cb, Cr][4]                                                      */
    uint8_t reserved[3];
} svtb200_sao_params;

/* f2 / N3: the per-LCU rate-distortion decision between the statistics and the application (saoMode = 1, presets <= M10):
 * DetermineSaoLumaModeOffsets / DetermineSaoChromaModeOffsets / TestSaoCopyModes in the order SaoGenerationDecision calls them
 * (Codec/EbSampleAdaptiveOffsetGenerationDecision.c:44,182,437,647-767).  rates = saoMergeFlagBits / saoTypeIndexBits / saoOffsetTrunUnaryBits of
 * the picture's MdRateEstimationContext_t (Codec/EbMdRateEstimation.h:148-150); lambdas = contextPtr->fullLambda / fullChromaLambdaSao
 * (Codec/EbCodingLoop.c:4732-4748).  LCUs in raster order, one tile: left / top neighbour = the previous decisions.  The shutSaoCondition* and
 * noise-class overrides around the call (EbCodingLoop.c:4671-4760) stay with the caller. */
typedef struct svtb200_sao_rate_tables {
    uint32_t merge_flag_bits[2];
    uint32_t type_index_bits[6];
    uint32_t offset_trun_unary_bits[8];
} svtb200_sao_rate_tables;
typedef struct svtb200_sao_decision {
    uint8_t  merge_left, merge_up;   /* saoMergeLeftFlag / saoMergeUpFlag */
    uint8_t  type[2];                /* saoTypeIndex[luma, chroma]        */
    uint8_t  band_position[3];
    uint8_t  reserved;
    int32_t  offset[3][4];
    int64_t  luma_cost, chroma_cost; /* saoLumaBestCost / saoChromaBestCost */
} svtb200_sao_decision;
int svtb200_sao_decision_picture(const svtb200_sao_stats *stats, int lcu_cols, int lcu_rows, const svtb200_sao_rate_tables *rates, uint64_t full_lambda,
                                 uint64_t full_chroma_lambda_sao, int is10bit, svtb200_sao_decision *out);
/* params_out_dev (optional): the same decisions in the form svtb200_sao_apply_picture_dev consumes */
int svtb200_sao_decision_picture_dev(const svtb200_sao_stats *stats_dev, int lcu_cols, int lcu_rows, const svtb200_sao_rate_tables *rates, uint64_t full_lambda,
                                     uint64_t full_chroma_lambda_sao, int is10bit, svtb200_sao_decision *out_dev, svtb200_sao_params *params_out_dev, void *cuda_stream);

/* Applies SAO to a whole picture: out = SAO(in), every sample classified on the UNMODIFIED input (what the reference
 * achieves in place with its saoUpBuffer / saoLeftBuffer line buffers, ApplySaoOffsetsPicture Codec/EbEncDecProcess.c:522,
 * ApplySaoOffsetsLcu :215, kernels SAOApplyBO / SAOApplyEO_{0,90,135,45} C_DEFAULT/EbSampleAdaptiveOffset_C.c:394-861).
 * in and out must be different pictures of identical geometry; picture-border samples of edge classes stay unmodified.
 * sao_flag[0] / [1]: pictureControlSetPtr->saoFlag (luma / chroma enabled for this picture). */
int svtb200_sao_apply_picture(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params,
                              const uint8_t sao_flag[2]);
int svtb200_sao_apply_picture_dev(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params_dev,
                                  const uint8_t sao_flag[2], void *cuda_stream);

/* the same two on 16-bit sample planes (10-bit video): GatherSaoStatisticsLcu_62x62_16bit (C_DEFAULT/EbSampleAdaptiveOffset_C.c:23: differences are
 * not clipped, band = sample >> 5), SAOApplyBO16bit / SAOApplyEO_*16bit (:431 ff, clip to 1023), ApplySaoOffsetsPicture16bit (Codec/EbEncDecProcess.c:1063) */
int svtb200_sao_stats_picture16(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out);
int svtb200_sao_stats_picture16_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *recon, svtb200_sao_stats *out_dev, void *cuda_stream);
/* H3 statistics on 16-bit sample planes (same rule, in each plane's own samples) */
int svtb200_sao_stats_picture16_h3(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered,
                                   svtb200_sao_stats *out);
int svtb200_sao_stats_picture16_h3_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *deblocked, const svtb200_yuv_picture *unfiltered,
                                       svtb200_sao_stats *out_dev, void *cuda_stream);
int svtb200_sao_apply_picture16(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params,
                                const uint8_t sao_flag[2]);
int svtb200_sao_apply_picture16_dev(const svtb200_yuv_picture *in, const svtb200_yuv_picture *out, const svtb200_sao_params *lcu_params_dev,
                                    const uint8_t sao_flag[2], void *cuda_stream);

/* =======================================================================================
 *  D. Intra prediction
 * ===================================================================================== */

/* Leaf: one HEVC intra mode on an explicit reference line -- IntraModePlanar / DCLuma / DCChroma / VerticalLuma|Chroma /
 * HorizontalLuma|Chroma / Angular_{2,18,34} / Angular_{Vertical,Horizontal}_Kernel behind IntraModeAngular_all
 * (C_DEFAULT/EbIntraPrediction_C.c:15-1015, Codec/EbIntraPrediction.c:3516).  refSamples holds 4*size+1 samples in the order of
 * IntraReferenceSamples_t::yIntraReferenceArray: [0] = bottom-most sample of the bottom-left column ... [2*size-1] = left sample of
 * row 0, [2*size] = top-left, [2*size+1 ...4*size] = top row then top-right.  is_chroma selects the variants without the
 * DC / horizontal / vertical edge filters.  mode 0 planar, 1 DC, 2..34 angular; size 4/8/16/32. */
int svtb200_IntraPrediction(uint32_t mode, uint32_t size, int is_chroma, const uint8_t *refSamples, uint8_t *predictionPtr,
                            uint32_t predictionBufferStride);

/* Encode-pass intra prediction of a batch of mutually independent blocks (one wavefront diagonal):
 * GenerateIntraReferenceSamplesEncodePass + EncodePassIntraPrediction (Codec/EbIntraPrediction.c:212,4395), 8-bit 4:2:0.
 * The neighbour samples are read from `recon` (what the reference keeps in its recon neighbour arrays); which of them may be used
 * is given per 4-luma-sample unit, in the blockIndex order of the reference: left_avail bit i = i-th unit of the left column counted
 * from the BOTTOM of the bottom-left block (2*size/4 units), top_avail bit i = i-th unit of the top row from the left (top then
 * top-right).  The caller folds picture/tile/slice boundaries, z-scan order (isBottomLeftAvailable / isUpperRightAvailable,
 * Codec/EbAvailability.c) and constrained-intra into these bits.  Luma is predicted at `size`, chroma (Cb and Cr) at size/2.
 * recon and pred may be the same picture (the reference predicts into its recon buffer, EbCodingLoop.c:3336). */
#define SVTB200_INTRA_SKIP 0xFF               /* luma_mode / chroma_mode: do not predict this component              */
#define SVTB200_INTRA_TOPLEFT_AVAIL 1         /* flags                                                                */
#define SVTB200_INTRA_STRONG_SMOOTHING 2      /* flags: sps strong_intra_smoothing_enable_flag                        */
#define SVTB200_INTRA_CHROMA_PLANAR 0         /* chroma_mode: EB_INTRA_CHROMA_* (Codec/EbSyntaxElements.h:58-62)      */
#define SVTB200_INTRA_CHROMA_VERTICAL 1
#define SVTB200_INTRA_CHROMA_HORIZONTAL 2
#define SVTB200_INTRA_CHROMA_DC 3
#define SVTB200_INTRA_CHROMA_DM 4
typedef struct svtb200_intra_job {
    uint16_t x, y;           /* luma origin of the block, picture coordinates                                         */
    uint8_t  size;           /* luma size 4/8/16/32 (4 only with chroma_mode = SKIP: the chroma of an NxN CU is one size-8 job) */
    uint8_t  luma_mode;      /* 0..34 or SVTB200_INTRA_SKIP                                                           */
    uint8_t  chroma_mode;    /* SVTB200_INTRA_CHROMA_* or SVTB200_INTRA_SKIP                                          */
    uint8_t  flags;
    uint16_t left_avail, top_avail;
} svtb200_intra_job;         /* 12 bytes */
int svtb200_intra_batch(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs);
int svtb200_intra_batch_dev(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs_dev,
                            uint32_t n_jobs, void *cuda_stream);
/* the same on 16-bit sample planes (10-bit video): GenerateIntraReference16bitSamplesEncodePass :1688 + EncodePassIntraPrediction16bit :4676
 * (midrange 512, SMOOTHING_THRESHOLD_10BIT, edge filters clipped to 1023); svtb200_plane::buf points at uint16_t samples */
int svtb200_intra_batch16(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs, uint32_t n_jobs);
int svtb200_intra_batch16_dev(const svtb200_yuv_picture *recon, const svtb200_yuv_picture *pred, const svtb200_intra_job *jobs_dev,
                              uint32_t n_jobs, void *cuda_stream);

/* Open-loop intra search of a whole picture: OpenLoopIntraSearchLcu (Codec/EbMotionEstimation.c:5053) for every LCU, with
 * UpdateNeighborSamplesArrayOpenLoop / IntraPredictionOpenLoop (Codec/EbIntraPrediction.c:5222,5339) and the candidate injection
 * and sorting helpers (:4458-4960).  Neighbours come from the SOURCE picture, so all LCUs are independent.
 * Outputs, indexed [lcu * 85 + rasterScanCuIndex] (index 0, the 64x64, is never written):
 *   total_modes[.]              OisCu32Cu16Results_t / OisCu8Results_t::totalIntraLumaMode
 *   cand[. * cand_capacity + k] OisCandidate_t packed as the reference's oisResults word: bits 0-19 distortion,
 *                               bit 20 validDistortion, bits 24-31 intraMode.  Fields the reference leaves unwritten read 0.
 * cand_capacity >= 18 with ois_kernel_level, >= 9 otherwise (>= 1 suffices with limit_ois_to_dc_mode on non-I pictures).
 * me_results (the svtb200_me_picture output of the same picture) is read only on non-I pictures without ois_kernel_level and
 * without limit_ois_to_dc_mode; it may be NULL otherwise. */
typedef struct svtb200_ois_config {
    uint16_t pic_width, pic_height;
    uint8_t  slice_type;                        /* 0 B, 1 P, 2 I (EB_SLICE)                                            */
    uint8_t  temporal_layer;
    uint8_t  ois_kernel_level;                  /* MotionEstimationContext_t::oisKernelLevel                           */
    uint8_t  ois_th_set;                        /* ::oisThSet 0..2                                                     */
    uint8_t  set_best_ois_distortion_to_valid;  /* ::setBestOisDistortionToValid                                       */
    uint8_t  limit_ois_to_dc_mode;              /* PictureParentControlSet_t::limitOisToDcModeFlag                     */
    uint8_t  skip_ois_8x8;                      /* ::skipOis8x8                                                        */
    uint8_t  cu8x8_mode;                        /* ::cu8x8Mode                                                         */
} svtb200_ois_config;
#define SVTB200_OIS_CU_COUNT 85
int svtb200_ois_picture(const svtb200_ois_config *cfg, const svtb200_plane *src_luma, const svtb200_me_pu_result *me_results,
                        uint8_t *total_modes, uint32_t *cand, uint32_t cand_capacity);
int svtb200_ois_picture_dev(const svtb200_ois_config *cfg, const svtb200_plane *src_luma_dev, const svtb200_me_pu_result *me_results_dev,
                            uint8_t *total_modes_dev, uint32_t *cand_dev, uint32_t cand_capacity, void *cuda_stream);

/* ---------------------------------------------------------------------------------------
 * 10-bit sample packing.  Same argument order as the reference typedefs in
 * Codec/EbPackUnPack.h (EB_ENC_Pack2D_TYPE :17, COMPPack_TYPE :52, EB_ENC_UnPack2D_TYPE :70,
 * EB_ENC_UnPack8BitData_TYPE :136, EB_ENC_UnpackAvg_TYPE :101); strides in elements; HOST pointers.
 *   msbPack2D        out16 = (in8 << 2) | ((inn >> 6) & 3)                  C_DEFAULT/EbPackUnPack_C.c:12
 *   CompressedPackmsb same, inn holds four 2-bit pels per byte, first pel in bits 7:6 (:44); width % 4 == 0
 *   CPack            unpacked 2-bit plane -> compressed 2-bit plane (:93); width % 4 == 0; localCache unused
 *   msbUnPack2D      out8 = in16 >> 2, outn = (u8)(in16 << 6) (:130);  UnPack8BitData[SafeSub]: out8 only (:153,203)
 *   UnpackAvg[SafeSub] dst = ((u8)(l0 >> 2) + (u8)(l1 >> 2) + 1) >> 1 (:174,226)
 * svtb200_pack_plane_dev runs one of them on DEVICE planes (op = SVTB200_PACK_*; unused planes NULL / 0):
 *   in0 / in1: the two inputs in the order of the host signature, out0 / out1: the outputs.
 * ------------------------------------------------------------------------------------- */
int svtb200_EB_ENC_msbPack2D(const uint8_t *in8BitBuffer, uint32_t in8Stride, const uint8_t *innBitBuffer, uint16_t *out16BitBuffer,
                             uint32_t innStride, uint32_t outStride, uint32_t width, uint32_t height);
int svtb200_CompressedPackmsb(const uint8_t *in8BitBuffer, uint32_t in8Stride, const uint8_t *innBitBuffer, uint16_t *out16BitBuffer,
                              uint32_t innStride, uint32_t outStride, uint32_t width, uint32_t height);
int svtb200_CPack(const uint8_t *innBitBuffer, uint32_t innStride, uint8_t *inCompnBitBuffer, uint32_t outStride, uint8_t *localCache,
                  uint32_t width, uint32_t height);
int svtb200_EB_ENC_msbUnPack2D(const uint16_t *in16BitBuffer, uint32_t inStride, uint8_t *out8BitBuffer, uint8_t *outnBitBuffer,
                               uint32_t out8Stride, uint32_t outnStride, uint32_t width, uint32_t height);
int svtb200_UnPack8BitData(const uint16_t *in16BitBuffer, uint32_t inStride, uint8_t *out8BitBuffer, uint32_t out8Stride, uint32_t width,
                           uint32_t height);
int svtb200_UnpackAvg(const uint16_t *ref16L0, uint32_t refL0Stride, const uint16_t *ref16L1, uint32_t refL1Stride, uint8_t *dstPtr,
                      uint32_t dstStride, uint32_t width, uint32_t height);
#define SVTB200_PACK_MSB 0
#define SVTB200_PACK_COMPRESSED_MSB 1
#define SVTB200_PACK_CPACK 2
#define SVTB200_PACK_UNPACK 3
#define SVTB200_PACK_UNPACK8 4
#define SVTB200_PACK_UNPACK_AVG 5
int svtb200_pack_plane_dev(int op, const void *in0, uint32_t in0_stride, const void *in1, uint32_t in1_stride, void *out0,
                           uint32_t out0_stride, void *out1, uint32_t out1_stride, uint32_t width, uint32_t height, void *cuda_stream);

/* ---------------------------------------------------------------------------------------
 * Device-resident source pictures (the PictureAnalysis feeders).
 * svtb200_pa_picture_create allocates padded full / quarter / sixteenth luma planes in HBM with the
 * reference's geometry (pads 68 / 32 / 16: EbEncHandle.c:901-925); svtb200_pa_picture_upload copies an
 * UNPADDED host luma plane into the full plane and runs, on the device, what PictureAnalysisKernel does
 * on the host: edge padding (GeneratePadding, EbMcp.c:1017) and point decimation by 2 and 4
 * (Decimation2D, EbPictureAnalysisProcess.c:173; DecimateInputPicture :4139).  The resulting
 * svtb200_pa_picture holds DEVICE pointers and is what the *_dev entry points take, so a source frame
 * crosses PCIe once and is reused as a reference by later pictures without further copies.
 * ------------------------------------------------------------------------------------- */
int svtb200_pa_picture_create(uint16_t width, uint16_t height, svtb200_pa_picture *out_dev);
int svtb200_pa_picture_upload(const svtb200_pa_picture *dev, const uint8_t *host_luma, uint32_t host_stride, void *cuda_stream);
/* same, source already in device memory (unpadded, pitch dev_stride) */
int svtb200_pa_picture_build_dev(const svtb200_pa_picture *dev, const uint8_t *dev_luma, uint32_t dev_stride, void *cuda_stream);
int svtb200_pa_picture_download(const svtb200_pa_picture *dev, const svtb200_pa_picture *host_out);
/* Transfers of the picture-level binding: the interiors of a device Y/Cb/Cr picture from UNPADDED host planes (strides in
 * samples, 1 or 2 bytes per sample), and a plain asynchronous copy for job lists / results (pinned host memory on the host side). */
int svtb200_yuv_picture_upload(const svtb200_yuv_picture *dev, const void *host_y, uint32_t host_y_stride, const void *host_cb,
                               uint32_t host_cb_stride, const void *host_cr, uint32_t host_cr_stride, uint32_t bytes_per_sample,
                               void *cuda_stream);
/* b6: GeneratePadding (Codec/EbMcp.c:1017) on a device-resident reconstructed picture, in place, before it serves as a reference
 * (PadRefAndSetFlags, Codec/EbEncDecProcess.c): every padding sample takes the nearest interior sample.  8-bit planes. */
int svtb200_yuv_picture_pad_dev(const svtb200_yuv_picture *dev, void *cuda_stream);
/* Device picture -> device picture of the same geometry (whole padded planes, bytes_per_sample 1 or 2): keeps the reconstruction a later stage needs
 * unmodified, e.g. the picture before svtb200_dlf_picture_dev for svtb200_sao_stats_picture_h3_dev (the reference keeps it implicitly: its LCU n is
 * gathered before LCU n + 1 is deblocked, Codec/EbCodingLoop.c:4604-4748). */
int svtb200_yuv_picture_copy_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *dst, uint32_t bytes_per_sample, void *cuda_stream);
/* N1: the per-LCU mean / variance pyramid of PictureAnalysis (ComputeBlockMeanComputeVariance, Codec/EbPictureAnalysisProcess.c:1646; 8x8 figures of
 * ComputeSubMean8x8 / ComputeSubdMeanOfSquaredValues8x8, ASM_SSE2/EbComputeMean_Intrinsic_SSE2.c:53,10) of the padded input luma:
 * y_mean[lcu][85] = PictureParentControlSet_t::yMean, variance[lcu][85] = ::variance, PU order of ME_TIER_ZERO_PU_* (64x64, 32x32 x 4, 16x16 x 16, 8x8 x 64). */
int svtb200_variance_picture(const svtb200_plane *luma_padded, uint8_t *y_mean, uint16_t *variance);
int svtb200_variance_picture_dev(const svtb200_plane *luma_padded_dev, uint8_t *y_mean_dev, uint16_t *variance_dev, void *cuda_stream);
int svtb200_copy_async(void *dst, const void *src, size_t bytes, int device_to_host, void *cuda_stream);
/* Peer memory for the GOP-boundary hand-off to the entropy-coding rank (one process per GPU): the receiver allocates and exports its buffers,
 * every producer opens them (CUDA IPC, peer access enabled lazily) and moves a finished picture's outputs device to device over NVLink with the copy
 * engines.  The 64-byte handle travels between the processes by whatever the host uses (torch.distributed in bench.py). */
int svtb200_peer_alloc(size_t bytes, void **dev_ptr);
int svtb200_peer_free(void *dev_ptr);
int svtb200_peer_export(const void *dev_ptr, uint8_t handle[64]);
int svtb200_peer_open(const uint8_t handle[64], void **dev_ptr);
int svtb200_peer_close(void *dev_ptr);
int svtb200_peer_copy_async(void *dst_dev, const void *src_dev, size_t bytes, void *cuda_stream);
void svtb200_pa_picture_destroy(svtb200_pa_picture *dev);

#ifdef __cplusplus
}
#endif
#endif /* SVT_HEVC_B200_H */
