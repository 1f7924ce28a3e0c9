/*
 * oracle/yasm_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference ships two yasm files (Source/Lib/ASM_SSE2/EbPictureOperators_SSE2.asm,
 * EbGatherSaoStatistics16bit_SSE2.asm).  This image has no yasm/nasm, so the 7 global
 * symbols they define are restated here in plain C so that the otherwise UNMODIFIED
 * reference C sources link into oracle/_ref/libsvtref_*.so.
 *
 *   EbHevcLog2f_SSE2            EbPictureOperators_SSE2.asm:623  (bsr)
 *   PictureCopyKernel_SSE2      EbPictureOperators_SSE2.asm:11   (row copy, width in bytes)
 *   ZeroOutCoeff{4x4,8x8,16x16,32x32}  EbPictureOperators_SSE2.asm:212,231,253,302
 *                               (zero an NxN s16 tile at coeffbuffer + coeffOriginIndex)
 *   GatherSaoStatisticsLcu16bit_SSE2  EbGatherSaoStatistics16bit_SSE2.asm:200
 *                               (forwards to the C 62x62 16-bit gatherer,
 *                                C_DEFAULT/EbSampleAdaptiveOffset_C.c:23)
 */
#include <stdint.h>
#include <string.h>

uint32_t EbHevcLog2f_SSE2(uint32_t x)
{
    return x ? (uint32_t)(31 - __builtin_clz(x)) : 0;
}

void PictureCopyKernel_SSE2(uint8_t *src, uint32_t srcStride, uint8_t *dst, uint32_t dstStride,
                            uint32_t areaWidthBytes, uint32_t areaHeight)
{
    for (uint32_t y = 0; y < areaHeight; ++y)
        memcpy(dst + (size_t)y * dstStride, src + (size_t)y * srcStride, areaWidthBytes);
}

static void zero_tile(int16_t *buf, uint32_t stride, uint32_t origin, uint32_t n)
{
    for (uint32_t y = 0; y < n; ++y)
        memset(buf + origin + (size_t)y * stride, 0, n * sizeof(int16_t));
}
#define ZERO_FN(name, n)                                                                    \
    void name(int16_t *coeffbuffer, uint32_t coeffStride, uint32_t coeffOriginIndex,        \
              uint32_t areaWidth, uint32_t areaHeight)                                      \
    { (void)areaWidth; (void)areaHeight; zero_tile(coeffbuffer, coeffStride, coeffOriginIndex, n); }
ZERO_FN(ZeroOutCoeff4x4_SSE, 4)
ZERO_FN(ZeroOutCoeff8x8_SSE2, 8)
ZERO_FN(ZeroOutCoeff16x16_SSE2, 16)
ZERO_FN(ZeroOutCoeff32x32_SSE2, 32)

/* C_DEFAULT/EbSampleAdaptiveOffset_C.h:42 */
extern uint32_t GatherSaoStatisticsLcu_62x62_16bit(
    uint16_t *inputSamplePtr, uint32_t inputStride, uint16_t *reconSamplePtr, uint32_t reconStride,
    uint32_t lcuWidth, uint32_t lcuHeight, int32_t *boDiff, uint16_t *boCount,
    int32_t eoDiff[][5], uint16_t eoCount[][5]);

uint32_t GatherSaoStatisticsLcu16bit_SSE2(
    uint16_t *inputSamplePtr, uint32_t inputStride, uint16_t *reconSamplePtr, uint32_t reconStride,
    uint32_t lcuWidth, uint32_t lcuHeight, int32_t *boDiff, uint16_t *boCount,
    int32_t eoDiff[][5], uint16_t eoCount[][5])
{
    return GatherSaoStatisticsLcu_62x62_16bit(inputSamplePtr, inputStride, reconSamplePtr, reconStride,
                                              lcuWidth, lcuHeight, boDiff, boCount, eoDiff, eoCount);
}
