/* oracle/port/pack_port.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Plain-C restatement of the reference's 10-bit sample packing, one sample at a time.
 *   port_msb_pack2d        <- EB_ENC_msbPack2D        C_DEFAULT/EbPackUnPack_C.c:12
 *   port_compressed_packmsb<- CompressedPackmsb        :44
 *   port_cpack             <- CPack_C                  :93
 *   port_msb_unpack2d      <- EB_ENC_msbUnPack2D       :130   (outn may be NULL: UnPack8BitData[SafeSub] :153,203)
 *   port_unpack_avg        <- UnpackAvg[SafeSub]       :174,226
 * Pinned by tests/test_oracle_pack.py against the C rows and the AVX2 / AVX512 / SSE2 rows of Codec/EbPackUnPack.h. */
#include <stddef.h>
#include <stdint.h>

void port_msb_pack2d(const uint8_t *in8, uint32_t s8, const uint8_t *inn, uint16_t *out, uint32_t sn, uint32_t so, uint32_t w, uint32_t h)
{
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x)
            out[(size_t)y * so + x] = (uint16_t)((in8[(size_t)y * s8 + x] << 2) | (inn[(size_t)y * sn + x] >> 6));
}

void port_compressed_packmsb(const uint8_t *in8, uint32_t s8, const uint8_t *inc, uint16_t *out, uint32_t sc, uint32_t so, uint32_t w, uint32_t h)
{
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < (w & ~3u); ++x) {
            const unsigned two = (inc[(size_t)y * sc + (x >> 2)] >> (6 - 2 * (x & 3))) & 3u;   /* first pel of a group sits in the top bits */
            out[(size_t)y * so + x] = (uint16_t)((in8[(size_t)y * s8 + x] << 2) | two);
        }
}

void port_cpack(const uint8_t *inn, uint32_t sn, uint8_t *outc, uint32_t sc, uint32_t w, uint32_t h)
{
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t g = 0; g * 4 < w; ++g) {
            unsigned v = 0;
            for (unsigned k = 0; k < 4; ++k) v |= (unsigned)(inn[(size_t)y * sn + g * 4 + k] >> 6) << (6 - 2 * k);
            outc[(size_t)y * sc + g] = (uint8_t)v;
        }
}

void port_msb_unpack2d(const uint16_t *in, uint32_t si, uint8_t *o8, uint8_t *on, uint32_t s8, uint32_t sn, uint32_t w, uint32_t h)
{
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const uint16_t v = in[(size_t)y * si + x];
            o8[(size_t)y * s8 + x] = (uint8_t)(v >> 2);
            if (on) on[(size_t)y * sn + x] = (uint8_t)((v & 3u) << 6);
        }
}

void port_unpack_avg(const uint16_t *l0, uint32_t s0, const uint16_t *l1, uint32_t s1, uint8_t *dst, uint32_t sd, uint32_t w, uint32_t h)
{
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const unsigned a = (uint8_t)(l0[(size_t)y * s0 + x] >> 2), b = (uint8_t)(l1[(size_t)y * s1 + x] >> 2);
            dst[(size_t)y * sd + x] = (uint8_t)((a + b + 1) >> 1);
        }
}
