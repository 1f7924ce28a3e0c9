/*
 * oracle/refinc_pa_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * PictureAnalysis feeder of the path: the per-LCU mean / variance pyramid.  ComputeBlockMeanComputeVariance is `static` in the
 * reference's Source/Lib/Codec/EbPictureAnalysisProcess.c (:1646), so this translation unit #includes that file WHERE IT LIES and calls it for every
 * LCU exactly like ComputePictureSpatialStatistics does (:3897-3912); built into oracle/_ref/libsvtref_pa_*.so (see build_ref.sh).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "EbPictureAnalysisProcess.c"        /* the reference source, compiled in place */

#include "svt_hevc_b200.h"

extern EB_U32 EbHevcGetCpuAsmType(void);

/* luma: the padded input picture (svtb200_plane: buf, stride, origin, size).  y_mean[n_lcu][85] (EB_U8), variance[n_lcu][85] (EB_U16), PU order of
 * ME_TIER_ZERO_PU_* (64x64, 32x32 x 4, 16x16 x 16, 8x8 x 64, each raster). */
int svtref_variance_picture(const svtb200_plane *luma, uint8_t *y_mean, uint16_t *variance, int asm_mode)
{
    ASM_TYPES = asm_mode ? EbHevcGetCpuAsmType() : 0;
    const EB_U32 cols = (luma->width + 63) / 64, rows = (luma->height + 63) / 64, n = cols * rows;
    EbPictureBufferDesc_t pic; memset(&pic, 0, sizeof pic);
    pic.bufferY = luma->buf; pic.strideY = (EB_U16)luma->stride; pic.originX = luma->origin_x; pic.originY = luma->origin_y;
    pic.width = luma->width; pic.height = luma->height; pic.bitDepth = EB_8BIT;
    PictureParentControlSet_t *pcs = (PictureParentControlSet_t *)calloc(1, sizeof(*pcs));
    pcs->variance = (EB_U16 **)calloc(n, sizeof(EB_U16 *));
    pcs->yMean = (EB_U8 **)calloc(n, sizeof(EB_U8 *));
    for (EB_U32 i = 0; i < n; ++i) { pcs->variance[i] = variance + (size_t)i * 85; pcs->yMean[i] = y_mean + (size_t)i * 85; }
    pcs->disableVarianceFlag = EB_FALSE;
    for (EB_U32 lcu = 0; lcu < n; ++lcu) {
        const EB_U32 ox = (lcu % cols) * 64, oy = (lcu / cols) * 64;
        ComputeBlockMeanComputeVariance(pcs, &pic, lcu, (pic.originY + oy) * pic.strideY + pic.originX + ox);
    }
    free(pcs->variance); free(pcs->yMean); free(pcs);
    return 0;
}
