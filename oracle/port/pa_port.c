/*
 * oracle/port/pa_port.c -- TEST INFRASTRUCTURE ONLY (see me_port.c header for the rules).
 *
 * Plain-C restatement of the per-LCU mean / variance pyramid of PictureAnalysis:
 *   ComputeBlockMeanComputeVariance          Codec/EbPictureAnalysisProcess.c:1646-2400
 *   ComputeSubMean8x8_SSE2_INTRIN            ASM_SSE2/EbComputeMean_Intrinsic_SSE2.c:53   (rows 0, 2, 4, 6 of the 8x8 block: sum << 3  = mean * 2^8)
 *   ComputeSubdMeanOfSquaredValues8x8_...    ASM_SSE2/EbComputeMean_Intrinsic_SSE2.c:10   (same rows: sum of squares << 11       = mean of squares * 2^16)
 * 16x16 / 32x32 / 64x64 figures are (sum of the four children) >> 2 at every level; yMean = mean >> 8, variance = (meanSq - mean^2) >> 16, stored as
 * EB_U8 / EB_U16.  PU order ME_TIER_ZERO_PU_*: 64x64, 4 x 32x32, 16 x 16x16, 64 x 8x8, each in raster order.  The LCUs of a partial last row / column read
 * the picture's padding exactly like the reference.
 */
#include <stdint.h>

#include "svt_hevc_b200.h"

int port_variance_picture(const svtb200_plane *luma, uint8_t *y_mean, uint16_t *variance)
{
    const int cols = (luma->width + 63) / 64, rows = (luma->height + 63) / 64;
    for (int ly = 0; ly < rows; ++ly)
        for (int lx = 0; lx < cols; ++lx) {
            const uint8_t *p = luma->buf + (size_t)(luma->origin_y + ly * 64) * luma->stride + luma->origin_x + lx * 64;
            uint64_t m8[64], s8[64], m16[16], s16[16], m32[4], s32[4];
            for (int b = 0; b < 64; ++b) {
                const uint8_t *q = p + (size_t)(b >> 3) * 8 * luma->stride + (b & 7) * 8;
                uint32_t sum = 0, sq = 0;
                for (int r = 0; r < 8; r += 2)
                    for (int c = 0; c < 8; ++c) { const uint32_t v = q[(size_t)r * luma->stride + c]; sum += v; sq += v * v; }
                m8[b] = (uint64_t)sum << 3; s8[b] = (uint64_t)sq << 11;
            }
            for (int b = 0; b < 16; ++b) {
                const int i = (b >> 2) * 16 + (b & 3) * 2;
                m16[b] = (m8[i] + m8[i + 1] + m8[i + 8] + m8[i + 9]) >> 2; s16[b] = (s8[i] + s8[i + 1] + s8[i + 8] + s8[i + 9]) >> 2;
            }
            for (int b = 0; b < 4; ++b) {
                const int i = (b >> 1) * 8 + (b & 1) * 2;
                m32[b] = (m16[i] + m16[i + 1] + m16[i + 4] + m16[i + 5]) >> 2; s32[b] = (s16[i] + s16[i + 1] + s16[i + 4] + s16[i + 5]) >> 2;
            }
            const uint64_t m64 = (m32[0] + m32[1] + m32[2] + m32[3]) >> 2, s64 = (s32[0] + s32[1] + s32[2] + s32[3]) >> 2;
            uint8_t *ym = y_mean + (size_t)(ly * cols + lx) * 85;
            uint16_t *vr = variance + (size_t)(ly * cols + lx) * 85;
            ym[0] = (uint8_t)(m64 >> 8); vr[0] = (uint16_t)((s64 - m64 * m64) >> 16);
            for (int b = 0; b < 4; ++b) { ym[1 + b] = (uint8_t)(m32[b] >> 8); vr[1 + b] = (uint16_t)((s32[b] - m32[b] * m32[b]) >> 16); }
            for (int b = 0; b < 16; ++b) { ym[5 + b] = (uint8_t)(m16[b] >> 8); vr[5 + b] = (uint16_t)((s16[b] - m16[b] * m16[b]) >> 16); }
            for (int b = 0; b < 64; ++b) { ym[21 + b] = (uint8_t)(m8[b] >> 8); vr[21 + b] = (uint16_t)((s8[b] - m8[b] * m8[b]) >> 16); }
        }
    return 0;
}
