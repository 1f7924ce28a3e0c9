/*
 * oracle/ref_tq_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Path-level oracle for the residual/transform/quant/inverse/recon chain: for every TU job it performs the calls
 * EncodeLoop and EncodeGenerateRecon make on the 8-bit path (Source/Lib/Codec/EbCodingLoop.c:651-1083, :1084-1240),
 * using the reference's OWN functions and dispatch tables: PictureResidual, EstimateTransform,
 * UnifiedQuantizeInvQuantize (default branch: no RDOQ, no perceptual masking), PfZeroOutUselessQuadrants,
 * EncodeInvTransform, AdditionKernel_funcPtrArray.
 */
#include <stdlib.h>
#include <string.h>

#include "EbDefinitions.h"
#include "EbPictureControlSet.h"
#include "EbSequenceControlSet.h"
#include "EbEncDecProcess.h"
#include "EbModeDecisionProcess.h"
#include "EbTransforms.h"
#include "EbPictureOperators.h"

#include "svt_hevc_b200.h"

extern EB_U32 EbHevcGetCpuAsmType(void);

static const svtb200_plane *pl(const svtb200_yuv_picture *p, int k) { return k == 0 ? &p->y : k == 1 ? &p->cb : &p->cr; }

int svtref_tq_batch(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                    const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts, int asm_mode,
                    int job_begin, int job_end)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureParentControlSet_t *ppcs = (PictureParentControlSet_t *)calloc(1, sizeof(*ppcs));
    ppcs->sequenceControlSetWrapperPtr = &scsWrap;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->ParentPcsPtr = ppcs;
    ModeDecisionContext_t *md = (ModeDecisionContext_t *)calloc(1, sizeof(*md));
    md->rdoqPmCoreMethod = 0;
    EncDecContext_t *ctx = (EncDecContext_t *)calloc(1, sizeof(*ctx));
    ctx->mdContext = md;
    ctx->pmMethod = 0;

    EB_S16 *res = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16)), *tr = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16));
    EB_S16 *q = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16)), *scratch = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16));
    if (job_end > n_jobs) job_end = n_jobs;
    for (int j = job_begin; j < job_end; ++j) {
        const svtb200_tq_job *t = &jobs[j];
        const svtb200_plane *sp = pl(src, t->plane), *pp = pl(pred, t->plane), *rp = pl(recon, t->plane);
        const EB_U32 N = t->size;
        EB_U8 *s = sp->buf + (size_t)(sp->origin_y + t->y) * sp->stride + sp->origin_x + t->x;
        EB_U8 *p = pp->buf + (size_t)(pp->origin_y + t->y) * pp->stride + pp->origin_x + t->x;
        EB_U8 *r = rp->buf + (size_t)(rp->origin_y + t->y) * rp->stride + rp->origin_x + t->x;
        memset(q, 0, 64 * 64 * sizeof(EB_S16));
        PictureResidual(s, sp->stride, p, pp->stride, res, 64, N, N);
        EstimateTransform(res, 64, tr, 64, N, scratch, BIT_INCREMENT_8BIT, (EB_BOOL)t->dst4x4, (EB_TRANS_COEFF_SHAPE)t->shape);
        EB_U32 nz = 0;
        const EB_PICTURE st = t->slice_type == 3 ? EB_IDR_PICTURE : (EB_PICTURE)t->slice_type;
        UnifiedQuantizeInvQuantize(ctx, pcs, tr, 64, q, res, t->qp, 8, N, st, &nz, t->shape,
                                   (t->flags & SVTB200_TQ_CLEAN_SPARSE) ? 1 : 0, 0, 0, (t->flags & SVTB200_TQ_ENABLE_CBF) ? 1 : 0,
                                   (t->flags & SVTB200_TQ_CONTOURING) ? 1 : 0, t->plane ? 1 : 0, t->temporal_layer, t->dz_offset,
                                   NULL, 0, 0, 4, NULL);
        nz_counts[j] = nz;
        /* quantised coefficients out (active area; zeros elsewhere) */
        const EB_U32 act = t->shape == SVTB200_SHAPE_ONLY_DC ? 1 : (N >> t->shape);
        int16_t *co = coeff->buf[t->plane] + (size_t)t->y * coeff->stride[t->plane] + t->x;
        for (EB_U32 y = 0; y < N; ++y)
            for (EB_U32 x = 0; x < N; ++x) co[(size_t)y * coeff->stride[t->plane] + x] = (y < act && x < act) ? q[y * 64 + x] : 0;
        const EB_BOOL cbf = nz ? EB_TRUE : EB_FALSE;
        EB_BOOL onlyDc = EB_FALSE;
        const EB_BOOL isMin = (EB_BOOL)(t->plane == 0 ? (N == MIN_PU_SIZE) : (N == MIN_PU_SIZE && 0));
        /* isOnlyDc / quadrant zeroing exactly as EncodeLoop does it (luma: :767-803, chroma :869-900) */
        if (!(t->plane == 0 && N == MIN_PU_SIZE)) {
            onlyDc = (nz == 1 && res[0] != 0 && !(t->plane == 0 && N == 32)) ? EB_TRUE : EB_FALSE;
            if (t->shape && cbf && !onlyDc) {
                if (t->shape == N2_SHAPE || t->shape == N4_SHAPE) PfZeroOutUselessQuadrants(res, 64, N >> 1);
                if (t->shape == N4_SHAPE) PfZeroOutUselessQuadrants(res, 64, N >> 2);
            }
        } else if (t->shape && cbf) {
            PfZeroOutUselessQuadrants(res, 64, N >> 1);
            if (t->shape == N4_SHAPE) PfZeroOutUselessQuadrants(res, 64, N >> 2);
        }
        (void)isMin;
        if (!(t->flags & SVTB200_TQ_SKIP_RECON)) {
            if (cbf) {
                EncodeInvTransform((t->plane == 0 && N == MIN_PU_SIZE) ? EB_FALSE : (EB_BOOL)(t->shape == ONLY_DC_SHAPE || onlyDc), res, 64, res, 64, N, scratch,
                                   BIT_INCREMENT_8BIT, (EB_BOOL)t->dst4x4);
                AdditionKernel_funcPtrArray[!!(ASM_TYPES & PREAVX2_MASK)][N >> 3](p, pp->stride, res, 64, r, rp->stride, N, N);
            } else if (r != p) {
                for (EB_U32 y = 0; y < N; ++y) memcpy(r + (size_t)y * rp->stride, p + (size_t)y * pp->stride, N);
            }
        }
    }
    free(res); free(tr); free(q); free(scratch); free(ctx); free(md); free(pcs); free(ppcs); free(scs);
    return 0;
}

/* 16-bit samples (10-bit video): EncodeLoop16bit / EncodeGenerateRecon16bit (Codec/EbCodingLoop.c:1240-1500, :1640-1800): PictureResidual16bit,
 * EncodeTransform with BIT_INCREMENT_10BIT, UnifiedQuantizeInvQuantize at EB_10BIT with qp + QP_BD_OFFSET, EncodeInvTransform, AdditionKernel16bit.
 * Planes hold EB_U16 samples; strides / origins in samples. */
int svtref_tq_batch16(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                    const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts, int asm_mode,
                    int job_begin, int job_end)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    SequenceControlSet_t *scs = (SequenceControlSet_t *)calloc(1, sizeof(*scs));
    EbObjectWrapper_t scsWrap; memset(&scsWrap, 0, sizeof scsWrap); scsWrap.objectPtr = scs;
    PictureParentControlSet_t *ppcs = (PictureParentControlSet_t *)calloc(1, sizeof(*ppcs));
    ppcs->sequenceControlSetWrapperPtr = &scsWrap;
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->ParentPcsPtr = ppcs;
    ModeDecisionContext_t *md = (ModeDecisionContext_t *)calloc(1, sizeof(*md));
    md->rdoqPmCoreMethod = 0;
    EncDecContext_t *ctx = (EncDecContext_t *)calloc(1, sizeof(*ctx));
    ctx->mdContext = md;
    ctx->pmMethod = 0;

    EB_S16 *res = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16)), *tr = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16));
    EB_S16 *q = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16)), *scratch = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16));
    if (job_end > n_jobs) job_end = n_jobs;
    for (int j = job_begin; j < job_end; ++j) {
        const svtb200_tq_job *t = &jobs[j];
        const svtb200_plane *sp = pl(src, t->plane), *pp = pl(pred, t->plane), *rp = pl(recon, t->plane);
        const EB_U32 N = t->size;
        EB_U16 *s = (EB_U16 *)sp->buf + (size_t)(sp->origin_y + t->y) * sp->stride + sp->origin_x + t->x;
        EB_U16 *p = (EB_U16 *)pp->buf + (size_t)(pp->origin_y + t->y) * pp->stride + pp->origin_x + t->x;
        EB_U16 *r = (EB_U16 *)rp->buf + (size_t)(rp->origin_y + t->y) * rp->stride + rp->origin_x + t->x;
        memset(q, 0, 64 * 64 * sizeof(EB_S16));
        PictureResidual16bit(s, sp->stride, p, pp->stride, res, 64, N, N);
        EncodeTransform(res, 64, tr, 64, N, scratch, BIT_INCREMENT_10BIT, (EB_BOOL)t->dst4x4, (EB_TRANS_COEFF_SHAPE)t->shape);
        EB_U32 nz = 0;
        const EB_PICTURE st = t->slice_type == 3 ? EB_IDR_PICTURE : (EB_PICTURE)t->slice_type;
        UnifiedQuantizeInvQuantize(ctx, pcs, tr, 64, q, res, t->qp + QP_BD_OFFSET, EB_10BIT, N, st, &nz, t->shape,
                                   (t->flags & SVTB200_TQ_CLEAN_SPARSE) ? 1 : 0, 0, 0, (t->flags & SVTB200_TQ_ENABLE_CBF) ? 1 : 0,
                                   (t->flags & SVTB200_TQ_CONTOURING) ? 1 : 0, t->plane ? 1 : 0, t->temporal_layer, t->dz_offset,
                                   NULL, 0, 0, 4, NULL);
        nz_counts[j] = nz;
        /* quantised coefficients out (active area; zeros elsewhere) */
        const EB_U32 act = t->shape == SVTB200_SHAPE_ONLY_DC ? 1 : (N >> t->shape);
        int16_t *co = coeff->buf[t->plane] + (size_t)t->y * coeff->stride[t->plane] + t->x;
        for (EB_U32 y = 0; y < N; ++y)
            for (EB_U32 x = 0; x < N; ++x) co[(size_t)y * coeff->stride[t->plane] + x] = (y < act && x < act) ? q[y * 64 + x] : 0;
        const EB_BOOL cbf = nz ? EB_TRUE : EB_FALSE;
        EB_BOOL onlyDc = EB_FALSE;
        const EB_BOOL isMin = (EB_BOOL)(t->plane == 0 ? (N == MIN_PU_SIZE) : (N == MIN_PU_SIZE && 0));
        /* isOnlyDc / quadrant zeroing exactly as EncodeLoop does it (luma: :767-803, chroma :869-900) */
        if (!(t->plane == 0 && N == MIN_PU_SIZE)) {
            onlyDc = (nz == 1 && res[0] != 0 && !(t->plane == 0 && N == 32)) ? EB_TRUE : EB_FALSE;
            if (t->shape && cbf && !onlyDc) {
                if (t->shape == N2_SHAPE || t->shape == N4_SHAPE) PfZeroOutUselessQuadrants(res, 64, N >> 1);
                if (t->shape == N4_SHAPE) PfZeroOutUselessQuadrants(res, 64, N >> 2);
            }
        } else if (t->shape && cbf) {
            PfZeroOutUselessQuadrants(res, 64, N >> 1);
            if (t->shape == N4_SHAPE) PfZeroOutUselessQuadrants(res, 64, N >> 2);
        }
        (void)isMin;
        if (!(t->flags & SVTB200_TQ_SKIP_RECON)) {
            if (cbf) {
                EncodeInvTransform((t->plane == 0 && N == MIN_PU_SIZE) ? EB_FALSE : (EB_BOOL)(t->shape == ONLY_DC_SHAPE || onlyDc), res, 64, res, 64, N, scratch,
                                   BIT_INCREMENT_10BIT, (EB_BOOL)t->dst4x4);
                AdditionKernel_funcPtrArray16bit[!!(ASM_TYPES & PREAVX2_MASK)](p, pp->stride, res, 64, r, rp->stride, N, N);
            } else if (r != p) {
                for (EB_U32 y = 0; y < N; ++y) memcpy(r + (size_t)y * rp->stride, p + (size_t)y * pp->stride, N * sizeof(EB_U16));
            }
        }
    }
    free(res); free(tr); free(q); free(scratch); free(ctx); free(md); free(pcs); free(ppcs); free(scs);
    return 0;
}

/* The mode-decision full loop's quantiser: the reference's OWN ProductUnifiedQuantizeInvQuantizeMd (Codec/EbFullLoop.c:77) after EstimateTransform, per TU of
 * the job list -- at enc modes >= 5 (rdoqPmCoreMethod == EB_NO_RDOQ, Codec/EbEncDecProcess.c:2201) it is QuantizeInvQuantize + UpdateQiQCoef with the I / P
 * rounding offset, no dead-zone offset and enableCbflag = 0, i.e. what svtb200_tq_batch computes for a job with SVTB200_TQ_SKIP_RECON, dz_offset = 0 and no
 * ENABLE_CBF / CLEAN_SPARSE flag.  Outputs: quantised coefficients (active area) and nz counts, like svtref_tq_batch. */
#include "EbFullLoop.h"
int svtref_tq_md_batch(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs,
                       uint32_t *nz_counts, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    PictureParentControlSet_t *ppcs = (PictureParentControlSet_t *)calloc(1, sizeof(*ppcs));
    PictureControlSet_t *pcs = (PictureControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->ParentPcsPtr = ppcs;
    EB_S16 *res = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16)), *tr = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16));
    EB_S16 *q = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16)), *rc = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16)), *scratch = (EB_S16 *)calloc(64 * 64, sizeof(EB_S16));
    for (int j = 0; j < n_jobs; ++j) {
        const svtb200_tq_job *t = &jobs[j];
        if (t->shape == SVTB200_SHAPE_ONLY_DC) { nz_counts[j] = 0; continue; }      /* (the full loop has no ONLY_DC shape) */
        const svtb200_plane *sp = pl(src, t->plane), *pp = pl(pred, t->plane);
        const EB_U32 N = t->size;
        EB_U8 *s = sp->buf + (size_t)(sp->origin_y + t->y) * sp->stride + sp->origin_x + t->x;
        EB_U8 *p = pp->buf + (size_t)(pp->origin_y + t->y) * pp->stride + pp->origin_x + t->x;
        memset(q, 0, 64 * 64 * sizeof(EB_S16));
        PictureResidual(s, sp->stride, p, pp->stride, res, 64, N, N);
        EstimateTransform(res, 64, tr, 64, N, scratch, BIT_INCREMENT_8BIT, (EB_BOOL)t->dst4x4, (EB_TRANS_COEFF_SHAPE)t->shape);
        pcs->sliceType = t->slice_type == 3 ? EB_IDR_PICTURE : (EB_PICTURE)t->slice_type;
        pcs->temporalLayerIndex = t->temporal_layer;
        EB_U32 nz = 0;
        ProductUnifiedQuantizeInvQuantizeMd(pcs, tr, 64, q, rc, t->qp, N, &nz, t->shape == SVTB200_SHAPE_N2 ? PF_N2 : t->shape == SVTB200_SHAPE_N4 ? PF_N4 : PF_OFF,
                                            (t->flags & SVTB200_TQ_CONTOURING) ? 1 : 0, t->plane ? 1 : 0, EB_NO_RDOQ, NULL, 0, INTER_MODE, 0, 0, NULL);
        nz_counts[j] = nz;
        const EB_U32 act = N >> t->shape;
        int16_t *co = coeff->buf[t->plane] + (size_t)t->y * coeff->stride[t->plane] + t->x;
        for (EB_U32 y = 0; y < N; ++y)
            for (EB_U32 x = 0; x < N; ++x) co[(size_t)y * coeff->stride[t->plane] + x] = (y < act && x < act) ? q[y * 64 + x] : 0;
    }
    free(res); free(tr); free(q); free(rc); free(scratch); free(pcs); free(ppcs);
    return 0;
}
