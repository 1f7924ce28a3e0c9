// tq.cu -- residual -> forward DCT/DST -> quantise + dequantise -> inverse transform -> reconstruction, fused per TU.
//
// Replaces:
//   Transform4x4 / DstTransform4x4 / Transform8x8 / Transform16x16Estimate / Transform32x32Estimate   C_DEFAULT/EbTransforms_C.c:1634-1889
//     (+ SIMD rows lowPrecisionTransform{16x16,32x32}_AVX2_INTRIN, Transform8x8_SSE4_1_INTRIN, ..., Pfreq[N4]Transform*), EstimateTransform
//     Codec/EbTransforms.c:3268
//   QuantizeInvQuantize C_DEFAULT/EbTransforms_C.c:89 (+ QuantizeInvQuantize{4x4_SSE3,8x8_AVX2,NxN_AVX2}), UnifiedQuantizeInvQuantize default
//     branch Codec/EbTransforms.c:3094-3261, UpdateQiQCoef C_DEFAULT/EbTransforms_C.c:209
//   InvTransform{4,8,16,32} / InvDstTransform4x4 C_DEFAULT/EbTransforms_C.c:1910-2119 (+ PFinvTransform*_SSSE3, ...), EncodeInvTransform
//     Codec/EbTransforms.c:3502
//   PictureResidual / AdditionKernel / PfZeroOutUselessQuadrants (Codec/EbPictureOperators.h:242,349; EbTransforms.c:3627)
//
// Design: one CTA (64 threads: the per-TU parameter derivation every warp repeats and the barriers cost half of what 128 threads paid, measured -15 %) per TU.  The TU never leaves shared memory between the stages, so HBM sees exactly
// src + pred in, quantised coefficients + reconstruction out (30 720 B per CTU, the algorithmic figure of SURVEY §8d).
// The 1-D passes are plain integer matrix products against the HEVC core matrix held in __constant__ memory (lanes of a
// warp share the matrix row -> broadcast); the 16/32-point forward passes fold the first butterfly stage in 16 bits exactly
// like PartialButterfly{16,32}Estimate.  Only the `active` low-frequency corner is computed for the N2 / N4 shapes, and
// the inverse skips the zero rows/columns.  No tensor cores: operands are s16 x s8 with a rounding shift per pass.
#include <mutex>

#include "common.cuh"

namespace svtb200 {

#ifndef TQ_NT_V
#define TQ_NT_V 64
#endif
constexpr int TQ_NT = TQ_NT_V;     // threads per TU
#ifndef TQ_MIN_CTAS
#define TQ_MIN_CTAS 8
#endif
constexpr int TP = 34;   // s16 pitch of the 32x32 tiles: 17 words -> conflict-free column walks

struct YuvW {
    uint8_t *p[3];
    int stride[3];
    int ox[3], oy[3];
};
struct TqArgs {
    YuvW src, pred, recon;
    int16_t *coeff[3];
    int cstride[3];
    const svtb200_tq_job *jobs;
    int n_jobs;
    uint32_t *nz;
};

__constant__ int8_t c_dct32[32][32];    // M32[k][n]; M_N[k][n] = M32[k * 32 / N][n]
// the same matrix and its per-size transposes T_N[n][k] = M_N[k][n] (pitch 32) in global memory: every CTA stages both with 128-bit loads
__device__ __align__(16) int8_t g_dct32[32 * 32];
__device__ __align__(16) int8_t g_dct_t[4][32 * 32];   // [log2 N - 2]
__constant__ int8_t c_dst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
__constant__ uint32_t c_qfunc[6] = {26214, 23302, 20560, 18396, 16384, 14564};   // QFunc, Codec/EbTransforms.h:26
__constant__ uint32_t c_ffunc[6] = {40, 45, 51, 57, 64, 72};                      // FFunc, :27
#include "tq_tables.inc"

static void build_dct32(int8_t m[32][32])
{
    static const int T[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0};
    for (int k = 0; k < 32; ++k)
        for (int n = 0; n < 32; ++n) {
            const int a = (k * (2 * n + 1)) & 127;
            m[k][n] = (int8_t)(a <= 32 ? T[a] : a <= 64 ? -T[64 - a] : a <= 96 ? -T[a - 64] : T[128 - a]);
        }
}

__device__ __forceinline__ int clip16i(int v) { return min(max(v, -32768), 32767); }

// forward 1-D pass: out[k][r] = (s16)((sum_n M[k][n] * in[r][n] + off) >> shift) for k < kmax, r < rows
// s_m: the 32x32 core matrix staged in shared memory.  Lanes of a warp either share the matrix row (forward pass: broadcast)
// or walk along it (inverse pass: consecutive bytes), so neither pass serialises on the matrix the way divergent
// __constant__ reads would.
__device__ void fwd_pass(const int8_t *s_m, const int16_t *in, int16_t *out, int N, int rows, int kmax, int shift, bool dst)
{
    const int off = 1 << (shift - 1), kstep = 32 / N;
    for (int i = threadIdx.x; i < kmax * rows; i += TQ_NT) {
        const int k = i / rows, r = i - k * rows;
        const int16_t *x = in + r * TP;
        int acc = 0;
        if (dst) {
#pragma unroll
            for (int n = 0; n < 4; ++n) acc += c_dst4[k][n] * x[n];
        } else if (N >= 16) {
            const int8_t *m = s_m + k * kstep * 32;
            if (k & 1) for (int n = 0; n < N / 2; ++n) acc += m[n] * (int)(short)(x[n] - x[N - 1 - n]);
            else for (int n = 0; n < N / 2; ++n) acc += m[n] * (int)(short)(x[n] + x[N - 1 - n]);
        } else {
            const int8_t *m = s_m + k * kstep * 32;
            for (int n = 0; n < N; ++n) acc += m[n] * x[n];
        }
        out[k * TP + r] = (int16_t)((acc + off) >> shift);
    }
}

// inverse 1-D pass: out[r][n] = clip16((sum_{k<kact} M[k][n] * in[k][r] + off) >> shift) for r < rmax, n < N
__device__ void inv_pass(const int8_t *s_m, const int16_t *in, int16_t *out, int N, int rmax, int kact, int shift, bool dst)
{
    const int off = 1 << (shift - 1), kstep = 32 / N;
    for (int i = threadIdx.x; i < rmax * N; i += TQ_NT) {
        const int r = i / N, n = i - r * N;
        int acc = 0;
        if (dst) {
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += c_dst4[k][n] * in[k * TP + r];
        } else {
            for (int k = 0; k < kact; ++k) acc += s_m[k * kstep * 32 + n] * in[k * TP + r];
        }
        out[r * TP + n] = (int16_t)clip16i((acc + off) >> shift);
    }
}


__device__ __forceinline__ int dp2a_lo(uint32_t a_s16x2, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t a_s16x2, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_s8x4), "r"(c));
    return d;
}

__device__ __forceinline__ int ilog2d(int v) { return 31 - __clz(v); }
__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16); }

// First butterfly stage of the 16/32-point forward passes, kept in 16 bits like PartialButterfly{16,32}Estimate:
// out[r][0..N/2) = x[n] + x[N-1-n], out[r][N/2..N) = x[n] - x[N-1-n].
__device__ void fold_rows(const int16_t *in, int16_t *out, int N, int rows)
{
    const int h = N >> 1, lgh = ilog2d(h);
    for (int i = threadIdx.x; i < rows * h; i += TQ_NT) {
        const int r = i >> lgh, n = i & (h - 1);
        const int a = in[r * TP + n], b = in[r * TP + N - 1 - n];
        out[r * TP + n] = (int16_t)(a + b);
        out[r * TP + h + n] = (int16_t)(a - b);
    }
}

// Row-by-row dot products on DP2A:  C(i, j) = sum_{n < Nn} A[j][n] * B[i][n]   (A: s8 rows, pitch lda bytes; B: s16 rows, pitch TP)
//   out[j * osj + i * osi] = CLIP ? clip16((C + off) >> shift) : (s16)((C + off) >> shift)
// `odd_off`: when != 0, odd j read their B row at +odd_off samples (the odd half of a folded row).
// Lanes run along i (B rows 17 words apart: conflict-free; the A row is a broadcast); each thread produces outputs j and j + J/2,
// which share the B words: 4 shared-memory loads feed 4 DP2A = 8 MACs.
template <bool CLIP>
__device__ void rowdot(const int8_t *A, int lda, const int16_t *B, int odd_off, int I, int J, int Nn, int off, int shift, int16_t *out, int osi, int osj)
{
    const int Jh = J >> 1, n4 = Nn >> 2, lgI = ilog2d(I);
    for (int it = threadIdx.x; it < I * Jh; it += TQ_NT) {
        const int j = it >> lgI, i = it & (I - 1);
        const uint32_t *b = (const uint32_t *)(B + i * TP + ((j & 1) ? odd_off : 0));
        const uint32_t *a0 = (const uint32_t *)(A + j * lda), *a1 = (const uint32_t *)(A + (j + Jh) * lda);
        int c0 = 0, c1 = 0;
        for (int q = 0; q < n4; ++q) {
            const uint32_t b0 = b[2 * q], b1 = b[2 * q + 1], w0 = a0[q], w1 = a1[q];
            c0 = dp2a_lo(b0, w0, c0); c0 = dp2a_hi(b1, w0, c0);
            c1 = dp2a_lo(b0, w1, c1); c1 = dp2a_hi(b1, w1, c1);
        }
        c0 = (c0 + off) >> shift; c1 = (c1 + off) >> shift;
        if (CLIP) { c0 = clip16i(c0); c1 = clip16i(c1); }
        out[j * osj + i * osi] = (int16_t)c0;
        out[(j + Jh) * osj + i * osi] = (int16_t)c1;
    }
}



// Register-blocked variant for the large passes (I * J >= 1024): every thread owns a 2 (i) x 4 (j) block of outputs, so that per 4 taps
// 4 + 4 shared-memory words feed 16 DP2A (32 MACs) and a 32x32 pass is exactly one block per thread.  j_m = jq + m * J/4 keeps the parity
// of jq (J/4 is even whenever the folded passes use this), which is what selects the sum / difference half of a folded row.
template <bool CLIP>
__device__ void rowdot24(const int8_t *A, int lda, const int16_t *B, int odd_off, int I, int J, int Nn, int off, int shift, int16_t *out, int osi, int osj)
{
    const int Ih = I >> 1, Jq = J >> 2, n4 = Nn >> 2, lgIh = ilog2d(Ih);
    for (int it = threadIdx.x; it < Ih * Jq; it += TQ_NT) {
        const int jq = it >> lgIh, i = it & (Ih - 1);
        const uint32_t *b0p = (const uint32_t *)(B + i * TP + ((jq & 1) ? odd_off : 0)), *b1p = b0p + (Ih * TP) / 2;
        const uint32_t *ap = (const uint32_t *)(A + jq * lda);
        const int astep = (Jq * lda) >> 2;
        int c[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 2
        for (int q = 0; q < n4; ++q) {
            const uint32_t x0 = b0p[2 * q], x1 = b0p[2 * q + 1], y0 = b1p[2 * q], y1 = b1p[2 * q + 1];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const uint32_t w = ap[m * astep + q];
                c[0][m] = dp2a_hi(x1, w, dp2a_lo(x0, w, c[0][m]));
                c[1][m] = dp2a_hi(y1, w, dp2a_lo(y0, w, c[1][m]));
            }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            int v0 = (c[0][m] + off) >> shift, v1 = (c[1][m] + off) >> shift;
            if (CLIP) { v0 = clip16i(v0); v1 = clip16i(v1); }
            const int j = jq + m * Jq;
            out[j * osj + i * osi] = (int16_t)v0;
            out[j * osj + (i + Ih) * osi] = (int16_t)v1;
        }
    }
}

// picks the blocking by pass size
template <bool CLIP>
__device__ __forceinline__ void rowdot_any(const int8_t *A, int lda, const int16_t *B, int odd_off, int I, int J, int Nn, int off, int shift, int16_t *out, int osi, int osj)
{
    if (I * J >= 1024) rowdot24<CLIP>(A, lda, B, odd_off, I, J, Nn, off, shift, out, osi, osj);
    else rowdot<CLIP>(A, lda, B, odd_off, I, J, Nn, off, shift, out, osi, osj);
}

// The whole chain for one TU held in shared memory.  s_a: residual in (pitch TP); returns nz.
// On exit s_a holds the reconstructed residual (if nz && want_recon), s_q the quantised coefficients (act x act corner).
__device__ uint32_t tq_chain(const int8_t *s_m, const int8_t *s_mt, int16_t *s_a, int16_t *s_b, int16_t *s_q, uint32_t *s_cnt, const svtb200_tq_job &t, bool want_recon, bool prefolded, int bit_inc)
{
    const int N = t.size, lg = ilog2d(N);
    const bool dst = t.dst4x4 != 0;
    const bool only_dc = t.shape == SVTB200_SHAPE_ONLY_DC;
    const int act = only_dc ? 1 : (N == 4 ? 4 : (N >> t.shape));
    if (threadIdx.x == 0) *s_cnt = 0;
    // ---- forward ----
    if (only_dc) {
        // EstimateTransform ONLY_DC branch (Codec/EbTransforms.c:3316-3332): one rounding for both passes
        int part = 0;
        for (int i = threadIdx.x; i < N * N; i += TQ_NT) part += s_a[(i / N) * TP + (i % N)];
        part = (int)warp_sum((uint32_t)part);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) atomicAdd(s_cnt, (uint32_t)part);
        __syncthreads();
        const int sum = (int)*s_cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            if (bit_inc) {   // EncodeTransform ONLY_DC branch (Codec/EbTransforms.c:3425-3447): the 16-bit path rounds after each pass
                const int sh1 = lg - 1 + bit_inc, sh2 = lg + 6;
                const int d1 = (64 * sum + (1 << (sh1 - 1))) >> sh1;
                s_a[0] = (int16_t)((64 * d1 + (1 << (sh2 - 1))) >> sh2);
            } else {
                const int shift = (lg - 1) + (lg + 6);
                s_a[0] = (int16_t)((64 * 64 * sum + (1 << (shift - 1))) >> shift);
            }
            *s_cnt = 0;
        }
        __syncthreads();
    } else {
        int s1, s2;
        if (N == 32) { s1 = 6 + bit_inc; s2 = 9; } else if (N == 16) { s1 = 4 + bit_inc; s2 = 9; } else { s1 = lg - 1 + bit_inc; s2 = lg + 6; }
        if (dst) {
            fwd_pass(s_m, s_a, s_b, N, N, act, s1, dst);     // s_b[k][r]
            __syncthreads();
            fwd_pass(s_m, s_b, s_a, N, act, act, s2, dst);   // s_a[v][u]
            __syncthreads();
        } else {
            const bool fold = N >= 16;
            const int Nn = fold ? N >> 1 : N, lda = (32 / N) * 32;
            if (fold && !prefolded) { fold_rows(s_a, s_q, N, N); __syncthreads(); }
            rowdot_any<false>(s_m, lda, fold ? s_q : s_a, fold ? Nn : 0, N, act, Nn, 1 << (s1 - 1), s1, s_b, 1, TP);       // s_b[k][r]
            __syncthreads();
            if (fold) { fold_rows(s_b, s_q, N, act); __syncthreads(); }
            rowdot_any<false>(s_m, lda, fold ? s_q : s_b, fold ? Nn : 0, act, act, Nn, 1 << (s2 - 1), s2, s_a, 1, TP);     // s_a[v][u]
            __syncthreads();
        }
    }
    // ---- quantise / dequantise (UnifiedQuantizeInvQuantize parameter derivation, Codec/EbTransforms.c:3094-3110) ----
    const int qp = t.qp + (bit_inc ? 12 : 0);      // QP_BD_OFFSET on the 10-bit path (EbCodingLoop.c:1307)
    const int qpRem = qp % 6, qpPer = qp / 6;
    const int tshift = 15 - (8 + bit_inc) - lg;
    const int shiftedQBits = 14 + qpPer + tshift;
    const uint32_t qoff_ip = (uint32_t)(t.slice_type >= 2 ? 171 : 85) << (shiftedQBits - 9);
    const uint32_t qoff = (!only_dc && t.dz_offset) ? (uint32_t)(t.dz_offset * (1u << shiftedQBits) / 20) : qoff_ip;
    const int shiftedFFunc = qpPer > 8 ? (int)c_ffunc[qpRem] << (qpPer - 2) : (int)c_ffunc[qpRem] << qpPer;
    const int shiftNum = qpPer > 8 ? 20 - 14 - tshift - 2 : 20 - 14 - tshift;
    const int iqoff = 1 << (shiftNum - 1);
    const uint32_t qf = c_qfunc[qpRem];
    uint32_t mynz = 0;
    const bool fast = !dst && !only_dc;          // DP2A inverse: dequantised coefficients are kept TRANSPOSED (s_b[u][v])
    const int rsy = fast ? 1 : TP, rsx = fast ? TP : 1;
    for (int i = threadIdx.x; i < N * N; i += TQ_NT) {
        const int y = i >> lg, x = i & (N - 1);
        int16_t qv = 0, rv = 0;
        if (y < act && x < act) {
            const int c = s_a[y * TP + x];
            int a = (int)((uint32_t)abs(c) * qf + qoff);
            a >>= shiftedQBits;
            qv = (int16_t)clip16i(c < 0 ? -a : a);
            rv = (int16_t)clip16i((qv * shiftedFFunc + iqoff) >> shiftNum);
            mynz += qv != 0;
        }
        s_q[y * TP + x] = qv;
        s_b[y * rsy + x * rsx] = rv;     // dequantised; zero outside the active corner (PfZeroOutUselessQuadrants)
    }
    mynz = warp_sum(mynz);
    if ((threadIdx.x & 31) == 0 && mynz) atomicAdd(s_cnt, mynz);
    __syncthreads();
    if (!only_dc) {
        if ((t.flags & SVTB200_TQ_CLEAN_SPARSE) && N >= 8 && *s_cnt && t.slice_type != 2 && act >= 4) {
            // isolated-coefficient cleaning, Codec/EbTransforms.c:3199-3243: one thread per 4x4 group
            __syncthreads();
            const int g = act / 4;
            for (int b = threadIdx.x; b < g * g; b += TQ_NT) {
                const int by = b / g, bx = b - by * g;
                int cnt = 0, loc = 0, locb = 0;
                for (int y = 0; y < 4; ++y)
                    for (int x = 0; x < 4; ++x)
                        if (s_q[(by * 4 + y) * TP + bx * 4 + x]) { ++cnt; loc = (by * 4 + y) * TP + bx * 4 + x; locb = (by * 4 + y) * rsy + (bx * 4 + x) * rsx; }
                if (cnt == 1 && loc != 0) { s_q[loc] = 0; s_b[locb] = 0; atomicSub(s_cnt, 1u); }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            // UpdateQiQCoef, C_DEFAULT/EbTransforms_C.c:209 -- the (EB_S16) cast binds before the shift there
            uint32_t nz = *s_cnt;
            const int16_t forced = (int16_t)((int16_t)(shiftedFFunc + iqoff) >> shiftNum);
            if (nz < 10 && (t.flags & SVTB200_TQ_CONTOURING) && t.slice_type == 2 && t.temporal_layer == 0 && t.plane == 0) {
                const int loc = (act - 1) * TP + (act - 1);
                if (s_q[loc] == 0) { ++nz; s_q[loc] = 1; s_b[(act - 1) * rsy + (act - 1) * rsx] = forced; }
            }
            if (nz == 0 && (t.flags & SVTB200_TQ_ENABLE_CBF)) {
                const int loc = (act - 2) * TP + (act - 1);
                nz = 1; s_q[loc] = 1; s_b[(act - 2) * rsy + (act - 1) * rsx] = forced;
            }
            *s_cnt = nz;
        }
        __syncthreads();
    }
    const uint32_t nz = *s_cnt;
    if (nz && want_recon && fast) {
        // pass 1: tmp(u, y) = sum_v M[v][y] * deq[v][u], written as s_a[y][u]; pass 2: res(y, x) = sum_u M[u][x] * tmp(u, y) -> s_b[y][x]
        rowdot_any<true>(s_mt, 32, s_b, 0, act, N, act, 64, 7, s_a, 1, TP);
        __syncthreads();
        rowdot_any<true>(s_mt, 32, s_a, 0, N, N, act, 2048 >> bit_inc, 12 - bit_inc, s_b, TP, 1);
        __syncthreads();
    } else if (nz && want_recon) {
        inv_pass(s_m, s_b, s_a, N, dst ? 4 : act, act, 7, dst);     // s_a[r][n], rows r >= act would be all zero
        __syncthreads();
        if (!dst && act < N) {   // zero rows act..N-1 of the intermediate so the second pass may stop at act as well
            for (int i = threadIdx.x; i < (N - act) * N; i += TQ_NT) s_a[(act + i / N) * TP + (i % N)] = 0;
            __syncthreads();
        }
        inv_pass(s_m, s_a, s_b, N, N, dst ? 4 : act, 12 - bit_inc, dst);      // s_b[y][x] residual
        __syncthreads();
    }
    return nz;
}

// 4 horizontally adjacent samples at sample index i of a plane (8-bit: one 32-bit access, 16-bit: one 64-bit access)
template <bool HBD>
__device__ __forceinline__ void ld4(const uint8_t *base, size_t i, int (&v)[4])
{
    if (HBD) { const uint2 w = *(const uint2 *)(base + 2 * i); v[0] = w.x & 0xffff; v[1] = w.x >> 16; v[2] = w.y & 0xffff; v[3] = w.y >> 16; }
    else { const uint32_t w = *(const uint32_t *)(base + i); v[0] = w & 0xff; v[1] = (w >> 8) & 0xff; v[2] = (w >> 16) & 0xff; v[3] = w >> 24; }
}
template <bool HBD>
__device__ __forceinline__ void st4(uint8_t *base, size_t i, const int (&v)[4])
{
    if (HBD) *(uint2 *)(base + 2 * i) = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
    else *(uint32_t *)(base + i) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
}
template <bool HBD>
__device__ __forceinline__ int ld1(const uint8_t *base, size_t i) { return HBD ? (int)((const uint16_t *)base)[i] : (int)base[i]; }

// HBD: 16-bit sample planes (10-bit video): EncodeLoop16bit / EncodeGenerateRecon16bit (Codec/EbCodingLoop.c:1240,1640); strides stay in samples
// The general path: the whole CTA (TQ_NT threads) on one TU, every size / shape / flag.
template <bool HBD>
__device__ void tq_cta_path(const TqArgs &A, const int j, int16_t *s_a, int16_t *s_b, int16_t *s_q, uint32_t &s_cnt, int8_t *s_m, int8_t *s_mt)
{
    constexpr int BPS = HBD ? 2 : 1, BIT_INC = HBD ? 2 : 0, MAXV = HBD ? 1023 : 255;
    const svtb200_tq_job t = A.jobs[j];
    const int N = t.size, pl = t.plane, lg = ilog2d(N);
    // core matrix and the transpose of this TU size: 64 + 64 128-bit loads
    for (int i = threadIdx.x; i < 128; i += TQ_NT) {
        if (i < 64) ((uint4 *)s_m)[i] = ((const uint4 *)g_dct32)[i];
        else ((uint4 *)s_mt)[i - 64] = ((const uint4 *)g_dct_t[lg - 2])[i - 64];
    }
    const int ss = A.src.stride[pl], ps = A.pred.stride[pl], rs = A.recon.stride[pl], cs = A.cstride[pl];
    const uint8_t *s = A.src.p[pl] + ((size_t)(A.src.oy[pl] + t.y) * ss + A.src.ox[pl] + t.x) * BPS;
    const uint8_t *p = A.pred.p[pl] + ((size_t)(A.pred.oy[pl] + t.y) * ps + A.pred.ox[pl] + t.x) * BPS;
    uint8_t *r = A.recon.p[pl] + ((size_t)(A.recon.oy[pl] + t.y) * rs + A.recon.ox[pl] + t.x) * BPS;
    int16_t *co = A.coeff[pl] + (size_t)t.y * cs + t.x;
    // 4 samples per access: 4 bytes (8-bit) or 8 bytes (16-bit) must be aligned
    const bool vec_in = ((((uintptr_t)s | (uintptr_t)p) | (uintptr_t)((ss | ps) * BPS)) & (4 * BPS - 1)) == 0;
    // residual; for the 16- and 32-point DCT the first butterfly stage (x[n] +- x[N-1-n], kept in 16 bits like
    // PartialButterfly{16,32}Estimate) is applied on the way in: s_q[r][0..N/2) sums, [N/2..N) differences
    const bool prefold = vec_in && N >= 16 && !t.dst4x4 && t.shape != SVTB200_SHAPE_ONLY_DC;
    if (prefold) {
        const int h = N >> 1, w8 = N >> 3, lw8 = lg - 3;
        for (int i = threadIdx.x; i < N * w8; i += TQ_NT) {
            const int y = i >> lw8, q = i & (w8 - 1), xl = 4 * q, xr = N - 4 - 4 * q;
            int sl[4], sr[4], pl_[4], pr[4];
            ld4<HBD>(s, (size_t)y * ss + xl, sl); ld4<HBD>(s, (size_t)y * ss + xr, sr);
            ld4<HBD>(p, (size_t)y * ps + xl, pl_); ld4<HBD>(p, (size_t)y * ps + xr, pr);
            int a[4], b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[e] = sl[e] - pl_[e];
                b[e] = sr[3 - e] - pr[3 - e];     // mirror sample N-1-(xl+e)
            }
            uint32_t *o = (uint32_t *)(s_q + y * TP);
            o[(xl >> 1)] = pack2(a[0] + b[0], a[1] + b[1]); o[(xl >> 1) + 1] = pack2(a[2] + b[2], a[3] + b[3]);
            o[((h + xl) >> 1)] = pack2(a[0] - b[0], a[1] - b[1]); o[((h + xl) >> 1) + 1] = pack2(a[2] - b[2], a[3] - b[3]);
        }
    } else if (vec_in) {
        const int w4 = N >> 2, lw4 = lg - 2;
        for (int i = threadIdx.x; i < N * w4; i += TQ_NT) {
            const int y = i >> lw4, x = (i & (w4 - 1)) * 4;
            int sw[4], pw[4];
            ld4<HBD>(s, (size_t)y * ss + x, sw); ld4<HBD>(p, (size_t)y * ps + x, pw);
            uint32_t *o = (uint32_t *)(s_a + y * TP + x);
            o[0] = pack2(sw[0] - pw[0], sw[1] - pw[1]);
            o[1] = pack2(sw[2] - pw[2], sw[3] - pw[3]);
        }
    } else {
        for (int i = threadIdx.x; i < N * N; i += TQ_NT) {
            const int y = i >> lg, x = i & (N - 1);
            s_a[y * TP + x] = (int16_t)(ld1<HBD>(s, (size_t)y * ss + x) - ld1<HBD>(p, (size_t)y * ps + x));
        }
    }
    __syncthreads();
    const bool want_recon = !(t.flags & SVTB200_TQ_SKIP_RECON);
    const uint32_t nz = tq_chain(s_m, s_mt, s_a, s_b, s_q, &s_cnt, t, want_recon, prefold, BIT_INC);
    if (threadIdx.x == 0) A.nz[j] = nz;
    const bool vec_out = vec_in && ((((uintptr_t)r | (uintptr_t)(rs * BPS)) & (4 * BPS - 1)) == 0) && ((((uintptr_t)co | (uintptr_t)(cs * 2)) & 7) == 0);
    if (vec_out) {
        const int w4 = N >> 2, lw4 = lg - 2;
        for (int i = threadIdx.x; i < N * w4; i += TQ_NT) {
            const int y = i >> lw4, x = (i & (w4 - 1)) * 4;
            const uint32_t *q = (const uint32_t *)(s_q + y * TP + x);
            *(uint2 *)(co + (size_t)y * cs + x) = make_uint2(q[0], q[1]);
            if (want_recon && (nz || r != p)) {
                int pw[4];
                ld4<HBD>(p, (size_t)y * ps + x, pw);
                if (nz) {
                    const int16_t *e = s_b + y * TP + x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) pw[k] = min(max(pw[k] + e[k], 0), MAXV);
                }
                st4<HBD>(r, (size_t)y * rs + x, pw);
            }
        }
    } else {
        for (int i = threadIdx.x; i < N * N; i += TQ_NT) {
            const int y = i >> lg, x = i & (N - 1);
            co[(size_t)y * cs + x] = s_q[y * TP + x];
            if (want_recon) {
                const int pv = ld1<HBD>(p, (size_t)y * ps + x);
                const int v = nz ? min(max(pv + s_b[y * TP + x], 0), MAXV) : pv;
                if (nz || r != p) { if (HBD) ((uint16_t *)r)[(size_t)y * rs + x] = (uint16_t)v; else r[(size_t)y * rs + x] = (uint8_t)v; }
            }
        }
    }
}



// =====================================================================================================================================
// Warp-level fast path (round 2): ONE WARP PER TU for the TUs that make up a picture -- 16x16 / 32x32, full (DEFAULT) shape, DCT, no post-quantisation
// flags -- with no block barrier anywhere.  Lane r owns row r (forward pass 1: residual row -> folded -> N outputs into a transposed s16 tile), then row u
// of that tile (forward pass 2 -> the N coefficients of column u, quantised and dequantised in registers), then column u of the inverse pass 1 (even /
// odd partial butterflies: out[n] = E[n] + O[n], out[N-1-n] = E[n] - O[n], half the multiplies of the dense product and exact in integers) into the same
// tile, then row y of the inverse pass 2, the prediction row and the reconstruction.  Every tap is a compile-time-indexed word of the static
// __constant__ tables (tq_tables.inc): a constant operand, not a load.  Two shared-memory transposes (2 x N x N s16) are all the data exchange there is.
// Arithmetic identical to tq_chain (same shifts, offsets, 16-bit wraps and clips); the general path above keeps every other size / shape / flag.
// =====================================================================================================================================
__device__ __forceinline__ uint32_t swap16(uint32_t w) { return __byte_perm(w, w, 0x1032); }
__device__ __forceinline__ uint32_t cw(const int8_t *tab, int word) { return ((const uint32_t *)tab)[word]; }

// out[k] for k < N from the folded row (fs: sums, fd: differences; N/4 words each): taps M_N[k][0 .. N/2)
template <int N>
__device__ __forceinline__ int fwd_dot(const uint32_t (&fs)[N / 4], const uint32_t (&fd)[N / 4], int k)
{
    // (k is a compile-time constant at every call site: the loops over k are fully unrolled)
    const int row = (N == 32 ? k : 2 * k) * 8;
    int acc = 0;
#pragma unroll
    for (int q = 0; q < N / 8; ++q) {
        const uint32_t w = cw(c_fw32, row + q);
        acc = (k & 1) ? dp2a_hi(fd[2 * q + 1], w, dp2a_lo(fd[2 * q], w, acc)) : dp2a_hi(fs[2 * q + 1], w, dp2a_lo(fs[2 * q], w, acc));
    }
    return acc;
}
// E[n], O[n] of the inverse butterflies from the even- / odd-index input words (we[2q] = (in[8q], in[8q+2]), we[2q+1] = (in[8q+4], in[8q+6]); wo likewise + 1)
template <int N>
__device__ __forceinline__ void inv_eo(const uint32_t (&we)[N / 4], const uint32_t (&wo)[N / 4], int n, int &E, int &O)
{
    const int8_t *te = N == 32 ? c_te32 : c_te16, *to = N == 32 ? c_to32 : c_to16;
    int e = 0, o = 0;
#pragma unroll
    for (int q = 0; q < N / 8; ++q) {
        const uint32_t a = cw(te, n * (N / 8) + q), b = cw(to, n * (N / 8) + q);
        e = dp2a_hi(we[2 * q + 1], a, dp2a_lo(we[2 * q], a, e));
        o = dp2a_hi(wo[2 * q + 1], b, dp2a_lo(wo[2 * q], b, o));
    }
    E = e; O = o;
}
// a row of N s16 (natural order, N/2 words) -> even / odd index words for inv_eo
template <int N>
__device__ __forceinline__ void split_eo(const uint32_t (&x)[N / 2], uint32_t (&we)[N / 4], uint32_t (&wo)[N / 4])
{
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        we[i] = __byte_perm(x[2 * i], x[2 * i + 1], 0x5410);      // (x[4i], x[4i+2])
        wo[i] = __byte_perm(x[2 * i], x[2 * i + 1], 0x7632);      // (x[4i+1], x[4i+3])
    }
}
// first butterfly stage in 16 bits (wrapping, like PartialButterfly{16,32}Estimate): x = N/2 words of s16 pairs in natural order
template <int N>
__device__ __forceinline__ void fold16(const uint32_t (&x)[N / 2], uint32_t (&fs)[N / 4], uint32_t (&fd)[N / 4])
{
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        const uint32_t m = swap16(x[N / 2 - 1 - i]);              // (X[N-1-2i], X[N-2-2i])
        fs[i] = __vadd2(x[i], m);
        fd[i] = __vsub2(x[i], m);
    }
}

// PAIR (N = 16 only): the two half-warps run two TUs at once -- lanes 0-15 the TU `t` / `j` / `T` they were handed, lanes 16-31 theirs; every
// per-TU quantity below is per lane, the tile exchanges stay inside a half-warp's own tile, and what used to be an early return is a predicate.
// The function is a real call (three instantiations inlined into the job loop of the kernel spilled at the 128-register budget): what it needs of the
// kernel's parameter block travels by value in TuPtrs, the tiles are namespace-scope __shared__ arrays, and the picture pointers are declared global
// to the compiler -- a callee would otherwise reach all of them through generic addresses.
struct TuPtrs { const uint8_t *s, *p; uint8_t *r; int16_t *co; uint32_t *nz; int ss, ps, rs, cs; };
template <bool HBD>
__device__ __forceinline__ TuPtrs tu_ptrs(const TqArgs &A, const svtb200_tq_job &t, int j)
{
    constexpr int BPS = HBD ? 2 : 1;
    const int pl = t.plane;
    TuPtrs P;
    P.ss = A.src.stride[pl]; P.ps = A.pred.stride[pl]; P.rs = A.recon.stride[pl]; P.cs = A.cstride[pl];
    P.s = A.src.p[pl] + ((size_t)(A.src.oy[pl] + t.y) * P.ss + A.src.ox[pl] + t.x) * BPS;
    P.p = A.pred.p[pl] + ((size_t)(A.pred.oy[pl] + t.y) * P.ps + A.pred.ox[pl] + t.x) * BPS;
    P.r = A.recon.p[pl] + ((size_t)(A.recon.oy[pl] + t.y) * P.rs + A.recon.ox[pl] + t.x) * BPS;
    P.co = A.coeff[pl] + (size_t)t.y * P.cs + t.x;
    P.nz = A.nz + j;
    return P;
}
__shared__ __align__(16) int16_t s_tile[2][32 * TP];      // one 32x32 (or two 16x16) s16 tile per warp of the batch kernel

template <int N, bool HBD, bool PAIR = false>
__device__ __noinline__ void tq_warp_fast(const TuPtrs P, const svtb200_tq_job t, const int tile_off, const bool live = true)
{
    static_assert(!PAIR || N == 16, "two TUs per warp: 16x16 only");
    constexpr int BPS = HBD ? 2 : 1, BIT_INC = HBD ? 2 : 0, MAXV = HBD ? 1023 : 255, LG = N == 32 ? 5 : 4;
    const int lane = PAIR ? (int)(threadIdx.x & 15) : (int)(threadIdx.x & 31);      // the lane's row / column inside its TU
    const bool on = live && lane < N;
    int16_t *T = &s_tile[0][0] + tile_off;
    const int ss = P.ss, ps = P.ps, rs = P.rs, cs = P.cs;
    const uint8_t *s = P.s, *p = P.p;
    uint8_t *r = P.r;
    int16_t *co = P.co;
    __builtin_assume(__isGlobal(s)); __builtin_assume(__isGlobal(p)); __builtin_assume(__isGlobal(r)); __builtin_assume(__isGlobal(co)); __builtin_assume(__isGlobal(P.nz));
    const bool want_recon = !(t.flags & SVTB200_TQ_SKIP_RECON);
    const int s1 = (N == 32 ? 6 : 4) + BIT_INC, s2 = 9;
    uint32_t x[N / 2];
    // ---- residual row `lane` as s16 pairs ----
    if (on) {
        const uint8_t *sr = s + (size_t)lane * ss * BPS, *pr = p + (size_t)lane * ps * BPS;
        if (HBD) {
#pragma unroll
            for (int c = 0; c < N / 8; ++c) {
                const uint4 a = ((const uint4 *)sr)[c], b = ((const uint4 *)pr)[c];
                x[4 * c] = __vsub2(a.x, b.x); x[4 * c + 1] = __vsub2(a.y, b.y); x[4 * c + 2] = __vsub2(a.z, b.z); x[4 * c + 3] = __vsub2(a.w, b.w);
            }
        } else {
#pragma unroll
            for (int c = 0; c < N / 8; ++c) {      // 8-byte vectors: the chroma planes of a picture padded by 40 samples are 8-, not 16-byte aligned
                const uint2 a = ((const uint2 *)sr)[c], b = ((const uint2 *)pr)[c];
                const uint32_t aw[2] = {a.x, a.y}, bw[2] = {b.x, b.y};
#pragma unroll
                for (int q = 0; q < 2; ++q) {      // bytes -> u16 pairs, then the 16-bit difference (|difference| <= 255)
                    x[4 * c + 2 * q] = __vsub2(__byte_perm(aw[q], 0, 0x4140), __byte_perm(bw[q], 0, 0x4140));
                    x[4 * c + 2 * q + 1] = __vsub2(__byte_perm(aw[q], 0, 0x4342), __byte_perm(bw[q], 0, 0x4342));
                }
            }
        }
    }
    // ---- forward pass 1: T[k][lane] ----
    {
        uint32_t fs[N / 4], fd[N / 4];
        fold16<N>(x, fs, fd);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int v = (fwd_dot<N>(fs, fd, k) + (1 << (s1 - 1))) >> s1;
            if (on) T[k * TP + lane] = (int16_t)v;
        }
    }
    __syncwarp();
    // ---- forward pass 2 on row `lane` of T: coefficient column u = lane ----
    if (on) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) x[i] = ((const uint32_t *)(T + lane * TP))[i];
    }
    __syncwarp();                       // every lane holds its row: the tile is free for the inverse pass
    // quantiser parameters (UnifiedQuantizeInvQuantize, Codec/EbTransforms.c:3094-3110; same derivation as tq_chain)
    const int qp = t.qp + (BIT_INC ? 12 : 0);
    const int qpRem = qp % 6, qpPer = qp / 6;
    const int tshift = 15 - (8 + BIT_INC) - LG;
    const int shiftedQBits = 14 + qpPer + tshift;
    const uint32_t qoff_ip = (uint32_t)(t.slice_type >= 2 ? 171 : 85) << (shiftedQBits - 9);
    const uint32_t qoff = t.dz_offset ? (uint32_t)(t.dz_offset * (1u << shiftedQBits) / 20) : qoff_ip;
    const int shiftedFFunc = qpPer > 8 ? (int)c_ffunc[qpRem] << (qpPer - 2) : (int)c_ffunc[qpRem] << qpPer;
    const int shiftNum = qpPer > 8 ? 20 - 14 - tshift - 2 : 20 - 14 - tshift;
    const int iqoff = 1 << (shiftNum - 1);
    const uint32_t qf = c_qfunc[qpRem];
    uint32_t mynz = 0;
    uint32_t d[N / 2];                  // dequantised column, natural order
    {
        uint32_t fs[N / 4], fd[N / 4];
        fold16<N>(x, fs, fd);
#pragma unroll
        for (int v = 0; v < N; ++v) {
            const int c = (int)(short)((fwd_dot<N>(fs, fd, v) + (1 << (s2 - 1))) >> s2);
            int a = (int)((uint32_t)abs(c) * qf + qoff);
            a >>= shiftedQBits;
            const int qv = clip16i(c < 0 ? -a : a);
            const int rv = clip16i((qv * shiftedFFunc + iqoff) >> shiftNum);
            mynz += (on && qv != 0);
            if (on) co[(size_t)v * cs + lane] = (int16_t)qv;
            if (v & 1) d[v >> 1] = (d[v >> 1] & 0xffffu) | ((uint32_t)(uint16_t)rv << 16); else d[v >> 1] = (uint32_t)(uint16_t)rv;
        }
    }
    uint32_t nz;
    if (PAIR) {
        nz = mynz;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);      // stays inside the half-warp
    } else nz = __reduce_add_sync(0xffffffffu, mynz);
    if (lane == 0 && live) *P.nz = nz;
    if (want_recon && nz == 0 && on && r != p) {      // prediction = reconstruction (copied when the two pictures differ)
        const uint2 *pr = (const uint2 *)(p + (size_t)lane * ps * BPS);
        uint2 *rr = (uint2 *)(r + (size_t)lane * rs * BPS);
#pragma unroll
        for (int c = 0; c < N * BPS / 8; ++c) rr[c] = pr[c];
    }
    const bool inv = on && want_recon && nz != 0;
    if (!__any_sync(0xffffffffu, inv)) return;
    // ---- inverse pass 1 (along v): T[y][lane] = clip16((sum_v M[v][y] d[v] + 64) >> 7) ----
    {
        uint32_t we[N / 4], wo[N / 4];
        split_eo<N>(d, we, wo);
#pragma unroll
        for (int y = 0; y < N / 2; ++y) {
            int E, O;
            inv_eo<N>(we, wo, y, E, O);
            if (inv) {
                T[y * TP + lane] = (int16_t)clip16i((E + O + 64) >> 7);
                T[(N - 1 - y) * TP + lane] = (int16_t)clip16i((E - O + 64) >> 7);
            }
        }
    }
    __syncwarp();
    // ---- inverse pass 2 (along u) on row y = lane, prediction, reconstruction ----
    if (inv) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) x[i] = ((const uint32_t *)(T + lane * TP))[i];
        uint32_t we[N / 4], wo[N / 4];
        split_eo<N>(x, we, wo);
        uint32_t rw[N / 2];              // the residual row as s16 pairs (32 registers less than one int per sample)
        constexpr int off2 = 2048 >> BIT_INC, sh2 = 12 - BIT_INC;
#pragma unroll
        for (int n = 0; n < N / 2; ++n) {
            int E, O;
            inv_eo<N>(we, wo, n, E, O);
            const uint32_t a = (uint32_t)(uint16_t)clip16i((E + O + off2) >> sh2), b = (uint32_t)(uint16_t)clip16i((E - O + off2) >> sh2);
            const int m = N - 1 - n;
            if (n & 1) rw[n >> 1] = (rw[n >> 1] & 0xffffu) | (a << 16); else rw[n >> 1] = a;
            if (m & 1) rw[m >> 1] = b << 16; else rw[m >> 1] = (rw[m >> 1] & 0xffff0000u) | b;      // (m runs downwards: the odd half of a word comes first)
        }
        const uint8_t *pr = p + (size_t)lane * ps * BPS;
        uint8_t *rr = r + (size_t)lane * rs * BPS;
        if (HBD) {
#pragma unroll
            for (int c = 0; c < N / 8; ++c) {
                const uint4 b = ((const uint4 *)pr)[c];
                const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t e = rw[4 * c + q];
                    const int v0 = min(max((int)(bw[q] & 0xffff) + (int)(short)(e & 0xffff), 0), MAXV), v1 = min(max((int)(bw[q] >> 16) + (int)(short)(e >> 16), 0), MAXV);
                    o[q] = (uint32_t)v0 | ((uint32_t)v1 << 16);
                }
                ((uint4 *)rr)[c] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < N / 8; ++c) {
                const uint2 b = ((const uint2 *)pr)[c];
                const uint32_t bw[2] = {b.x, b.y};
                uint32_t o[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    int v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t w = rw[4 * c + 2 * q + (e >> 1)];
                        v[e] = min(max((int)((bw[q] >> (8 * e)) & 0xff) + (int)(short)((e & 1) ? (w >> 16) : (w & 0xffff)), 0), 255);
                    }
                    o[q] = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
                }
                ((uint2 *)rr)[c] = make_uint2(o[0], o[1]);
            }
        }
    }
}

// eligibility of the warp path (uniform per TU): a full-shape 16x16 / 32x32 DCT TU without post-quantisation flags whose rows are vector accessible
template <bool HBD>
__device__ __forceinline__ bool tq_fast_ok(const TqArgs &A, const svtb200_tq_job &t)
{
    constexpr int BPS = HBD ? 2 : 1;
    if (!(t.size == 16 || t.size == 32) || t.shape != SVTB200_SHAPE_DEFAULT || t.dst4x4 || (t.flags & ~SVTB200_TQ_SKIP_RECON)) return false;
    const int pl = t.plane;
    const uintptr_t a = (uintptr_t)(A.src.p[pl] + ((size_t)(A.src.oy[pl] + t.y) * A.src.stride[pl] + A.src.ox[pl] + t.x) * BPS) | (uintptr_t)(A.src.stride[pl] * BPS) |
                        (uintptr_t)(A.pred.p[pl] + ((size_t)(A.pred.oy[pl] + t.y) * A.pred.stride[pl] + A.pred.ox[pl] + t.x) * BPS) | (uintptr_t)(A.pred.stride[pl] * BPS) |
                        (uintptr_t)(A.recon.p[pl] + ((size_t)(A.recon.oy[pl] + t.y) * A.recon.stride[pl] + A.recon.ox[pl] + t.x) * BPS) | (uintptr_t)(A.recon.stride[pl] * BPS);
    return (a & (HBD ? 15 : 7)) == 0;      // 8-bit rows move as 8-byte vectors, 16-bit rows as 16-byte vectors
}

// Batch kernel: a CTA of two warps takes 2 x TQ_CHUNK consecutive TUs; each warp runs its TUs alone when the warp path applies, the rest go through the
// general path afterwards with the whole CTA.
constexpr int TQ_CHUNK = 3;
template <bool HBD>
__global__ void __launch_bounds__(TQ_NT, TQ_MIN_CTAS) tq_batch_kernel(const __grid_constant__ TqArgs A)
{
    static_assert(TQ_NT == 64, "two warps, two TUs");
    int16_t *s_a = s_tile[0], *s_b = s_tile[1];
    __shared__ __align__(16) int16_t s_q[32 * TP];
    __shared__ uint32_t s_cnt;
    __shared__ __align__(16) int8_t s_m[32 * 32], s_mt[32 * 32];
    __shared__ int s_slow[2];
    // a warp takes TQ_CHUNK consecutive jobs (a 32x32 CU's luma TU and its two chroma TUs in the usual job order) one after the other; two neighbouring
    // 16x16 TUs run side by side in its half-warps
    const int wid = threadIdx.x >> 5, j0 = TQ_CHUNK * (2 * blockIdx.x + wid), half = (threadIdx.x >> 4) & 1;
    const int tile = wid * 32 * TP;
    uint32_t slow = 0;
    for (int i = 0; i < TQ_CHUNK && j0 + i < A.n_jobs; ++i) {
        const int j = j0 + i;
        const svtb200_tq_job t = A.jobs[j];
        if (!tq_fast_ok<HBD>(A, t)) { slow |= 1u << i; continue; }
        if (t.size == 32) { tq_warp_fast<32, HBD>(tu_ptrs<HBD>(A, t, j), t, tile); __syncwarp(); continue; }
        bool pair = false;
        svtb200_tq_job t2 = t;
        if (i + 1 < TQ_CHUNK && j + 1 < A.n_jobs) { t2 = A.jobs[j + 1]; pair = t2.size == 16 && tq_fast_ok<HBD>(A, t2); }
        if (pair) { const svtb200_tq_job th = half ? t2 : t; tq_warp_fast<16, HBD, true>(tu_ptrs<HBD>(A, th, j + half), th, tile + half * 16 * TP); ++i; }
        else tq_warp_fast<16, HBD>(tu_ptrs<HBD>(A, t, j), t, tile);
        __syncwarp();
    }
    if ((threadIdx.x & 31) == 0) s_slow[wid] = (int)slow;
    __syncthreads();
    const uint32_t all_slow = (uint32_t)s_slow[0] | ((uint32_t)s_slow[1] << TQ_CHUNK);
    for (int k = 0; k < 2 * TQ_CHUNK; ++k)
        if ((all_slow >> k) & 1) { tq_cta_path<HBD>(A, TQ_CHUNK * 2 * blockIdx.x + k, s_a, s_b, s_q, s_cnt, s_m, s_mt); __syncthreads(); }
}

// ---- coefficient hand-off: only coded TUs cross PCIe -------------------------------------------------------------------
// The entropy coder reads coeffSamplesTB of TUs whose cbf is set (EbCodingLoop.c:745, EbEntropyCoding.c) and nothing else, so the coefficient
// planes are packed on the device into one dense stream: TU after TU in job order, size x size int16 row-major, TUs with nz == 0 skipped.
// offsets[j] = first element of job j in the stream, offsets[n_jobs] = total elements.
constexpr int PK_SCAN_NT = 1024;
constexpr int PK_TILE = 32 * 1024;      // jobs per pass of the scan kernel (one size code byte each in shared memory)
// One CTA: per pass the size codes of PK_TILE jobs are fetched with coalesced loads (jobs of TUs without coefficients count 0) and kept as
// bytes in shared memory; every thread then owns a run of consecutive jobs, the run totals are scanned across the CTA with shuffles, and the
// offsets of the run are written.  (The earlier version walked the job array with a per-thread stride: ~50 us of dependent, uncoalesced loads.)
__global__ void __launch_bounds__(PK_SCAN_NT) tq_pack_scan_kernel(const svtb200_tq_job *jobs, const uint32_t *nz, int n_jobs, uint32_t *offsets)
{
    __shared__ uint8_t s_code[PK_TILE];          // 0: not coded, else log2(size) (size * size = 1 << 2 * code)
    __shared__ uint32_t s_part[PK_SCAN_NT / 32];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    for (int t0 = 0; t0 < n_jobs; t0 += PK_TILE) {
        const int nt = min(PK_TILE, n_jobs - t0);
        __syncthreads();
        const uint32_t base = s_base;          // stable here: its last update is followed by a barrier
        // both loads unconditional and unrolled: the batch of independent requests is in flight together (a conditional second load serialised
        // ~50 dependent DRAM round trips)
#pragma unroll 8
        for (int i = threadIdx.x; i < nt; i += PK_SCAN_NT) {
            const uint32_t z = __ldg(nz + t0 + i);
            const int sz = __ldg(&jobs[t0 + i].size);
            s_code[i] = z ? (uint8_t)(31 - __clz(sz)) : (uint8_t)0;
        }
        __syncthreads();
        const int per = (nt + PK_SCAN_NT - 1) / PK_SCAN_NT, j0 = min(nt, (int)threadIdx.x * per), j1 = min(nt, j0 + per);
        uint32_t local = 0;
        for (int j = j0; j < j1; ++j) { const uint32_t c = s_code[j]; local += c ? 1u << (2 * c) : 0u; }
        // exclusive scan of the per-thread sums: warp scan, then scan of the 32 warp totals
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_part[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t v = s_part[lane], w = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            s_part[lane] = w - v;      // exclusive
            if (lane == 31) s_base = base + w;
        }
        __syncthreads();
        uint32_t run = base + s_part[wid] + incl - local;
        for (int j = j0; j < j1; ++j) { offsets[t0 + j] = run; const uint32_t c = s_code[j]; run += c ? 1u << (2 * c) : 0u; }
    }
    __syncthreads();
    if (threadIdx.x == 0) offsets[n_jobs] = s_base;
}

__global__ void __launch_bounds__(128) tq_pack_copy_kernel(const __grid_constant__ TqArgs A, const uint32_t *offsets, int16_t *stream, uint32_t capacity)
{
    const int j = blockIdx.x;
    if (j >= A.n_jobs || A.nz[j] == 0) return;
    const svtb200_tq_job t = A.jobs[j];
    const int N = t.size, cs = A.cstride[t.plane];
    const uint32_t off = offsets[j];
    if (off + (uint32_t)(N * N) > capacity) return;      // the caller sees offsets[n_jobs] > capacity
    const int16_t *co = A.coeff[t.plane] + (size_t)t.y * cs + t.x;
    int16_t *o = stream + off;
    if (((((uintptr_t)co | (uintptr_t)(cs * 2)) & 7) == 0) && (((uintptr_t)o & 7) == 0)) {      // 4 coefficients per access
        const int w4 = N >> 2;
        for (int i = threadIdx.x; i < N * w4; i += 128) { const int y = i / w4, x = (i - y * w4) * 4; *(uint2 *)(o + y * N + x) = *(const uint2 *)(co + (size_t)y * cs + x); }
    } else {
        for (int i = threadIdx.x; i < N * N; i += 128) { const int y = i / N, x = i - y * N; o[i] = co[(size_t)y * cs + x]; }
    }
}

// leaf kernels: mode 0 forward (EstimateTransform), 1 QiQ with explicit parameters, 2 inverse
struct LeafArgs {
    const int16_t *in; int16_t *out, *out2; uint32_t *nz;
    int istride, ostride, N, bit_inc, dst, shape, mode;
    uint32_t qFunc, q_offset; int shiftedQBits, shiftedFFunc, iq_offset, shiftNum, area;
};
__global__ void __launch_bounds__(TQ_NT) tq_leaf_kernel(const LeafArgs L)
{
    __shared__ __align__(16) int16_t s_a[32 * TP], s_b[32 * TP];
    __shared__ uint32_t s_cnt;
    __shared__ __align__(16) int8_t s_m[32 * 32];
    for (int i = threadIdx.x; i < 256; i += TQ_NT) ((uint32_t *)s_m)[i] = ((const uint32_t *)c_dct32)[i];
    const int N = L.N, lg = ilog2d(N);
    if (L.mode == 1) {
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        uint32_t mynz = 0;
        for (int i = threadIdx.x; i < L.area * L.area; i += TQ_NT) {
            const int y = i / L.area, x = i - y * L.area;
            const int c = L.in[y * L.istride + x];
            int a = (int)((uint32_t)abs(c) * L.qFunc + L.q_offset);
            a >>= L.shiftedQBits;
            const int16_t qv = (int16_t)clip16i(c < 0 ? -a : a);
            L.out[y * L.istride + x] = qv;
            L.out2[y * L.istride + x] = (int16_t)clip16i((qv * L.shiftedFFunc + L.iq_offset) >> L.shiftNum);
            mynz += qv != 0;
        }
        mynz = warp_sum(mynz);
        if ((threadIdx.x & 31) == 0 && mynz) atomicAdd(&s_cnt, mynz);
        __syncthreads();
        if (threadIdx.x == 0) *L.nz = s_cnt;
        return;
    }
    for (int i = threadIdx.x; i < N * N; i += TQ_NT) s_a[(i / N) * TP + (i % N)] = L.in[(i / N) * L.istride + (i % N)];
    __syncthreads();
    if (L.mode == 0) {
        const bool dst = L.dst != 0;
        if (L.shape == SVTB200_SHAPE_ONLY_DC) {
            if (threadIdx.x == 0) {
                int sum = 0;
                for (int i = 0; i < N * N; ++i) sum += s_a[(i / N) * TP + (i % N)];
                const int shift = (lg - 1 + L.bit_inc) + (lg + 6);
                L.out[0] = (int16_t)((64 * 64 * sum + (1 << (shift - 1))) >> shift);
            }
            return;
        }
        const int act = N == 4 ? 4 : (N >> L.shape);
        int s1, s2;
        if (N == 32) { s1 = 6 + L.bit_inc; s2 = 9; } else if (N == 16) { s1 = 4 + L.bit_inc; s2 = 9; } else { s1 = lg - 1 + L.bit_inc; s2 = lg + 6; }
        fwd_pass(s_m, s_a, s_b, N, N, act, s1, dst);
        __syncthreads();
        fwd_pass(s_m, s_b, s_a, N, act, act, s2, dst);
        __syncthreads();
        for (int i = threadIdx.x; i < act * act; i += TQ_NT) L.out[(i / act) * L.ostride + (i % act)] = s_a[(i / act) * TP + (i % act)];
    } else {
        const bool dst = L.dst != 0;
        inv_pass(s_m, s_a, s_b, N, N, N, 7, dst);
        __syncthreads();
        inv_pass(s_m, s_b, s_a, N, N, N, 12 - L.bit_inc, dst);
        __syncthreads();
        for (int i = threadIdx.x; i < N * N; i += TQ_NT) L.out[(i / N) * L.ostride + (i % N)] = s_a[(i / N) * TP + (i % N)];
    }
}

// The transform matrices live in __constant__ / __device__ symbols of the CURRENT device: they are built into local buffers and uploaded once per device
// under a lock (entry points are called concurrently from many EncDec threads), and the upload is complete -- cudaMemcpyToSymbol from pageable memory
// followed by a device synchronisation -- before any thread launches on its own non-blocking stream.
static int ensure_tables()
{
    static std::mutex mu;
    static std::atomic<int> uploaded_dev{-1};
    int dev = -1;
    SVTB200_CUDA_TRY(cudaGetDevice(&dev));
    if (uploaded_dev.load(std::memory_order_acquire) == dev) return SVTB200_OK;
    std::lock_guard<std::mutex> lk(mu);
    if (uploaded_dev.load(std::memory_order_relaxed) == dev) return SVTB200_OK;
    int8_t m[32][32];
    build_dct32(m);
    std::vector<int8_t> mt(4 * 32 * 32, 0);
    for (int l = 0; l < 4; ++l) {
        const int N = 4 << l, ks = 32 / N;
        for (int k = 0; k < N; ++k)
            for (int n = 0; n < N; ++n) mt[(size_t)l * 1024 + n * 32 + k] = m[k * ks][n];
    }
    SVTB200_CUDA_TRY(cudaMemcpyToSymbol(c_dct32, m, sizeof m));
    SVTB200_CUDA_TRY(cudaMemcpyToSymbol(g_dct32, m, sizeof m));
    SVTB200_CUDA_TRY(cudaMemcpyToSymbol(g_dct_t, mt.data(), mt.size()));
    SVTB200_CUDA_TRY(cudaDeviceSynchronize());
    uploaded_dev.store(dev, std::memory_order_release);
    return SVTB200_OK;
}

static YuvW to_yuvw(const svtb200_yuv_picture &p, uint8_t *y, uint8_t *cb, uint8_t *cr)
{
    YuvW w;
    w.p[0] = y; w.p[1] = cb; w.p[2] = cr;
    const svtb200_plane *pl[3] = {&p.y, &p.cb, &p.cr};
    for (int k = 0; k < 3; ++k) { w.stride[k] = (int)pl[k]->stride; w.ox[k] = pl[k]->origin_x; w.oy[k] = pl[k]->origin_y; }
    return w;
}
static size_t pbytes(const svtb200_plane &p) { return (size_t)p.stride * (p.height + 2 * p.origin_y); }

static int check_jobs(const svtb200_tq_job *jobs, int n, const svtb200_yuv_picture &pic)
{
    for (int j = 0; j < n; ++j) {
        const svtb200_tq_job &t = jobs[j];
        const svtb200_plane &pl = t.plane == 0 ? pic.y : t.plane == 1 ? pic.cb : pic.cr;
        const bool ok = t.plane <= 2 && (t.size == 4 || t.size == 8 || t.size == 16 || t.size == 32) && t.shape <= 3 && t.qp <= 51 && t.slice_type <= 3 &&
                        (!t.dst4x4 || t.size == 4) && (t.size > 4 || t.shape == 0) && !(t.size == 8 && t.shape == SVTB200_SHAPE_N4) &&
                        t.x + t.size <= pl.width && t.y + t.size <= pl.height;      // TUs lie inside the picture: the coefficient planes are unpadded
        if (!ok) { set_error("tq_batch: job %d malformed (size/shape/qp/plane out of range, TU outside the picture, or 8x8 with N4 which the reference leaves undefined)", j); return SVTB200_ERR_BAD_ARG; }
    }
    return SVTB200_OK;
}

}  // namespace svtb200
using namespace svtb200;

extern "C" {

static int tq_batch_dev_impl(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                             const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs_dev, int n_jobs, uint32_t *nz_counts_dev, void *cuda_stream, bool hbd)
{
    if (!src || !pred || !recon || !coeff || (!jobs_dev && n_jobs) || n_jobs < 0 || !nz_counts_dev) { set_error("tq_batch_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    for (int k = 0; k < 3; ++k) {
        const svtb200_plane *a[3] = {k == 0 ? &src->y : k == 1 ? &src->cb : &src->cr, k == 0 ? &pred->y : k == 1 ? &pred->cb : &pred->cr, k == 0 ? &recon->y : k == 1 ? &recon->cb : &recon->cr};
        if (!a[0]->buf || !a[1]->buf || !a[2]->buf || !coeff->buf[k] || a[1]->width != a[0]->width || a[1]->height != a[0]->height || a[2]->width != a[0]->width ||
            a[2]->height != a[0]->height || coeff->stride[k] < a[0]->width) { set_error("tq_batch_dev: plane %d: null buffer, or prediction / reconstruction / coefficient geometry differs from the source's", k); return SVTB200_ERR_BAD_ARG; }
    }
    int rc = ensure_device();
    if (rc || (rc = ensure_tables())) return rc;
    if (n_jobs == 0) return SVTB200_OK;
    TqArgs A;
    A.src = to_yuvw(*src, src->y.buf, src->cb.buf, src->cr.buf);
    A.pred = to_yuvw(*pred, pred->y.buf, pred->cb.buf, pred->cr.buf);
    A.recon = to_yuvw(*recon, recon->y.buf, recon->cb.buf, recon->cr.buf);
    for (int k = 0; k < 3; ++k) { A.coeff[k] = coeff->buf[k]; A.cstride[k] = (int)coeff->stride[k]; }
    A.jobs = jobs_dev; A.n_jobs = n_jobs; A.nz = nz_counts_dev;
    if (hbd) tq_batch_kernel<true><<<(n_jobs + 2 * TQ_CHUNK - 1) / (2 * TQ_CHUNK), TQ_NT, 0, (cudaStream_t)cuda_stream>>>(A); else tq_batch_kernel<false><<<(n_jobs + 2 * TQ_CHUNK - 1) / (2 * TQ_CHUNK), TQ_NT, 0, (cudaStream_t)cuda_stream>>>(A);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

static int tq_batch_impl(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon,
                         const svtb200_coeff_picture *coeff, const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts, bool hbd)
{
    const size_t bps = hbd ? 2 : 1;
    if (!src || !pred || !recon || !coeff || (!jobs && n_jobs) || n_jobs < 0 || !nz_counts) { set_error("tq_batch: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = check_jobs(jobs, n_jobs, *src);
    if (rc) return rc;
    ThreadCtx *c = tls_ctx();
    if ((rc = c->init()) || (rc = ensure_tables())) return rc;
    c->reset();
    const bool in_place = recon->y.buf == pred->y.buf && recon->cb.buf == pred->cb.buf && recon->cr.buf == pred->cr.buf;
    const svtb200_yuv_picture *pics[3] = {src, pred, recon};
    const size_t cw = src->y.width, chh = src->y.height;
    const size_t cbytes[3] = {(size_t)coeff->stride[0] * chh * 2, (size_t)coeff->stride[1] * (chh / 2) * 2, (size_t)coeff->stride[2] * (chh / 2) * 2};
    (void)cw;
    size_t need = 8192 + (size_t)n_jobs * (sizeof(svtb200_tq_job) + 4);
    for (int i = 0; i < 3; ++i) need += align_up(bps * pbytes(pics[i]->y), 256) + align_up(bps * pbytes(pics[i]->cb), 256) + align_up(bps * pbytes(pics[i]->cr), 256);
    for (int k = 0; k < 3; ++k) need += align_up(cbytes[k], 256);
    if ((rc = c->dev_reserve(need))) return rc;
    uint8_t *dp[3][3];
    for (int i = 0; i < 3; ++i) {
        const svtb200_plane *pl[3] = {&pics[i]->y, &pics[i]->cb, &pics[i]->cr};
        for (int k = 0; k < 3; ++k) {
            if (i == 2 && in_place) { dp[2][k] = dp[1][k]; continue; }
            void *d;
            if ((rc = c->dev_alloc(bps * pbytes(*pl[k]), &d))) return rc;
            dp[i][k] = (uint8_t *)d;
            // recon (when separate) is updated only inside TU areas: bring its current content over as well
            SVTB200_CUDA_TRY(cudaMemcpyAsync(d, pl[k]->buf, bps * pbytes(*pl[k]), cudaMemcpyHostToDevice, c->stream));
        }
    }
    void *dc[3], *d_jobs, *d_nz;
    for (int k = 0; k < 3; ++k) {
        if ((rc = c->dev_alloc(cbytes[k], &dc[k]))) return rc;
        SVTB200_CUDA_TRY(cudaMemcpyAsync(dc[k], coeff->buf[k], cbytes[k], cudaMemcpyHostToDevice, c->stream));
    }
    if ((rc = c->dev_alloc(sizeof(svtb200_tq_job) * (size_t)std::max(n_jobs, 1), &d_jobs)) || (rc = c->dev_alloc(4 * (size_t)std::max(n_jobs, 1), &d_nz))) return rc;
    if (n_jobs) SVTB200_CUDA_TRY(cudaMemcpyAsync(d_jobs, jobs, sizeof(svtb200_tq_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, c->stream));
    svtb200_yuv_picture ds = *src, dpred = *pred, drec = *recon;
    ds.y.buf = dp[0][0]; ds.cb.buf = dp[0][1]; ds.cr.buf = dp[0][2];
    dpred.y.buf = dp[1][0]; dpred.cb.buf = dp[1][1]; dpred.cr.buf = dp[1][2];
    drec.y.buf = dp[2][0]; drec.cb.buf = dp[2][1]; drec.cr.buf = dp[2][2];
    svtb200_coeff_picture dcoef = *coeff;
    for (int k = 0; k < 3; ++k) dcoef.buf[k] = (int16_t *)dc[k];
    if ((rc = tq_batch_dev_impl(&ds, &dpred, &drec, &dcoef, (const svtb200_tq_job *)d_jobs, n_jobs, (uint32_t *)d_nz, c->stream, hbd))) return rc;
    const svtb200_plane *rpl[3] = {&recon->y, &recon->cb, &recon->cr};
    for (int k = 0; k < 3; ++k) {
        SVTB200_CUDA_TRY(cudaMemcpyAsync(rpl[k]->buf, dp[2][k], bps * pbytes(*rpl[k]), cudaMemcpyDeviceToHost, c->stream));
        SVTB200_CUDA_TRY(cudaMemcpyAsync(coeff->buf[k], dc[k], cbytes[k], cudaMemcpyDeviceToHost, c->stream));
    }
    if (n_jobs) SVTB200_CUDA_TRY(cudaMemcpyAsync(nz_counts, d_nz, 4 * (size_t)n_jobs, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SVTB200_OK;
}

int svtb200_tq_batch_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon, const svtb200_coeff_picture *coeff,
                         const svtb200_tq_job *jobs_dev, int n_jobs, uint32_t *nz_counts_dev, void *cuda_stream)
{ return tq_batch_dev_impl(src, pred, recon, coeff, jobs_dev, n_jobs, nz_counts_dev, cuda_stream, false); }
int svtb200_tq_batch(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon, const svtb200_coeff_picture *coeff,
                     const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts)
{ return tq_batch_impl(src, pred, recon, coeff, jobs, n_jobs, nz_counts, false); }
int svtb200_tq_batch16_dev(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon, const svtb200_coeff_picture *coeff,
                           const svtb200_tq_job *jobs_dev, int n_jobs, uint32_t *nz_counts_dev, void *cuda_stream)
{ return tq_batch_dev_impl(src, pred, recon, coeff, jobs_dev, n_jobs, nz_counts_dev, cuda_stream, true); }
int svtb200_tq_batch16(const svtb200_yuv_picture *src, const svtb200_yuv_picture *pred, const svtb200_yuv_picture *recon, const svtb200_coeff_picture *coeff,
                       const svtb200_tq_job *jobs, int n_jobs, uint32_t *nz_counts)
{ return tq_batch_impl(src, pred, recon, coeff, jobs, n_jobs, nz_counts, true); }

static int tq_leaf(LeafArgs L, const int16_t *in, size_t in_elems, int16_t *out, size_t out_elems, int16_t *out2, uint32_t *nz)
{
    ThreadCtx *c = tls_ctx();
    int rc;
    if ((rc = c->init()) || (rc = ensure_tables())) return rc;
    c->reset();
    void *d_in, *d_out, *d_out2, *d_nz, *h;
    if ((rc = c->dev_reserve((in_elems + 2 * out_elems) * 2 + 8192)) || (rc = c->dev_alloc(in_elems * 2, &d_in)) || (rc = c->dev_alloc(out_elems * 2, &d_out)) ||
        (rc = c->dev_alloc(out_elems * 2, &d_out2)) || (rc = c->dev_alloc(4, &d_nz)) || (rc = c->pin_alloc(out_elems * 4 + 16, &h))) return rc;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(d_in, in, in_elems * 2, cudaMemcpyHostToDevice, c->stream));
    SVTB200_CUDA_TRY(cudaMemsetAsync(d_out, 0, out_elems * 2, c->stream));
    SVTB200_CUDA_TRY(cudaMemsetAsync(d_out2, 0, out_elems * 2, c->stream));
    L.in = (const int16_t *)d_in; L.out = (int16_t *)d_out; L.out2 = (int16_t *)d_out2; L.nz = (uint32_t *)d_nz;
    tq_leaf_kernel<<<1, TQ_NT, 0, c->stream>>>(L);
    SVTB200_COUNT_LAUNCH();
    SVTB200_CUDA_TRY(cudaGetLastError());
    int16_t *h16 = (int16_t *)h;
    SVTB200_CUDA_TRY(cudaMemcpyAsync(h16, d_out, out_elems * 2, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(h16 + out_elems, d_out2, out_elems * 2, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaMemcpyAsync(h16 + 2 * out_elems, d_nz, 4, cudaMemcpyDeviceToHost, c->stream));
    SVTB200_CUDA_TRY(cudaStreamSynchronize(c->stream));
    const int rows = L.mode == 1 ? L.area : L.N, cols = rows, os = L.mode == 1 ? L.istride : L.ostride;
    const int keep = (L.mode == 0) ? (L.shape == SVTB200_SHAPE_ONLY_DC ? 1 : (L.N == 4 ? 4 : L.N >> L.shape)) : rows;
    for (int y = 0; y < keep; ++y)
        for (int x = 0; x < (L.mode == 0 ? keep : cols); ++x) {
            out[(size_t)y * os + x] = h16[(size_t)y * os + x];
            if (out2) out2[(size_t)y * os + x] = h16[out_elems + (size_t)y * os + x];
        }
    if (nz) memcpy(nz, h16 + 2 * out_elems, 4);
    return SVTB200_OK;
}

int svtb200_EstimateTransform(const int16_t *residualBuffer, uint32_t residualStride, int16_t *coeffBuffer, uint32_t coeffStride, uint32_t transformSize,
                              uint32_t bitIncrement, uint8_t dstTransformFlag, uint8_t transCoeffShape)
{
    const uint32_t N = transformSize;
    if (!residualBuffer || !coeffBuffer || !(N == 4 || N == 8 || N == 16 || N == 32) || transCoeffShape > 3 || (dstTransformFlag && N != 4) || bitIncrement > 4) { set_error("EstimateTransform: bad argument"); return SVTB200_ERR_BAD_ARG; }
    LeafArgs L{};
    L.istride = (int)residualStride; L.ostride = (int)coeffStride; L.N = (int)N; L.bit_inc = (int)bitIncrement; L.dst = dstTransformFlag; L.shape = transCoeffShape; L.mode = 0;
    return tq_leaf(L, residualBuffer, (size_t)(N - 1) * residualStride + N, coeffBuffer, (size_t)(N - 1) * coeffStride + N, nullptr, nullptr);
}

int svtb200_QuantizeInvQuantize(const int16_t *coeff, uint32_t coeffStride, int16_t *quantCoeff, int16_t *reconCoeff, uint32_t qFunc, uint32_t q_offset,
                                int32_t shiftedQBits, int32_t shiftedFFunc, int32_t iq_offset, int32_t shiftNum, uint32_t areaSize, uint32_t *nonzerocoeff)
{
    if (!coeff || !quantCoeff || !reconCoeff || !nonzerocoeff || areaSize == 0 || areaSize > 32 || shiftedQBits < 0 || shiftedQBits > 31 || shiftNum < 0 || shiftNum > 31) { set_error("QuantizeInvQuantize: bad argument"); return SVTB200_ERR_BAD_ARG; }
    LeafArgs L{};
    L.istride = (int)coeffStride; L.ostride = (int)coeffStride; L.N = (int)areaSize; L.mode = 1; L.area = (int)areaSize;
    L.qFunc = qFunc; L.q_offset = q_offset; L.shiftedQBits = shiftedQBits; L.shiftedFFunc = shiftedFFunc; L.iq_offset = iq_offset; L.shiftNum = shiftNum;
    const size_t n = (size_t)(areaSize - 1) * coeffStride + areaSize;
    return tq_leaf(L, coeff, n, quantCoeff, n, reconCoeff, nonzerocoeff);
}

int svtb200_EncodeInvTransform(const int16_t *coeffBuffer, uint32_t coeffStride, int16_t *reconBuffer, uint32_t reconStride, uint32_t transformSize,
                               uint32_t bitIncrement, uint8_t dstTransformFlag)
{
    const uint32_t N = transformSize;
    if (!coeffBuffer || !reconBuffer || !(N == 4 || N == 8 || N == 16 || N == 32) || (dstTransformFlag && N != 4) || bitIncrement > 4) { set_error("EncodeInvTransform: bad argument"); return SVTB200_ERR_BAD_ARG; }
    LeafArgs L{};
    L.istride = (int)coeffStride; L.ostride = (int)reconStride; L.N = (int)N; L.bit_inc = (int)bitIncrement; L.dst = dstTransformFlag; L.mode = 2;
    return tq_leaf(L, coeffBuffer, (size_t)(N - 1) * coeffStride + N, reconBuffer, (size_t)(N - 1) * reconStride + N, nullptr, nullptr);
}


int svtb200_tq_pack_coeffs_dev(const svtb200_coeff_picture *coeff_dev, const svtb200_tq_job *jobs_dev, const uint32_t *nz_counts_dev, int n_jobs,
                               int16_t *stream_dev, uint32_t stream_capacity, uint32_t *offsets_dev, void *cuda_stream)
{
    if (!coeff_dev || (!jobs_dev && n_jobs) || !nz_counts_dev || n_jobs < 0 || !stream_dev || !offsets_dev) { set_error("tq_pack_coeffs_dev: bad argument"); return SVTB200_ERR_BAD_ARG; }
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    tq_pack_scan_kernel<<<1, PK_SCAN_NT, 0, st>>>(jobs_dev, nz_counts_dev, n_jobs, offsets_dev);
    SVTB200_COUNT_LAUNCH();
    if (n_jobs) {
        TqArgs A;
        memset(&A, 0, sizeof A);
        for (int k = 0; k < 3; ++k) { A.coeff[k] = coeff_dev->buf[k]; A.cstride[k] = (int)coeff_dev->stride[k]; }
        A.jobs = jobs_dev; A.n_jobs = n_jobs; A.nz = (uint32_t *)nz_counts_dev;
        tq_pack_copy_kernel<<<n_jobs, 128, 0, st>>>(A, offsets_dev, stream_dev, stream_capacity);
        SVTB200_COUNT_LAUNCH();
    }
    SVTB200_CUDA_TRY(cudaGetLastError());
    return SVTB200_OK;
}

}  // extern "C"
