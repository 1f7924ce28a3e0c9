// ubench_pipes.cu -- issue-rate micro-benchmarks that bound the ME design on sm_100a (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_pipes.cu && /tmp/ubench
// Reports warp-instructions per cycle per SM sub-partition for VABSDIFF4(.add), SHF, IMAD, LDS.128 and mixes of them.
// The "VABSDIFF4-issue roofline" bench.py reports beside the HBM fraction is derived from the first line.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t vsad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t *out, int iters, uint32_t seed)
{
    __shared__ uint4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i * seed, i + seed, i ^ seed, i);
    __syncthreads();
    uint32_t a[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = i; w[i] = threadIdx.x * 0x01010101u + i * seed; }
    uint32_t s = seed, t = threadIdx.x;
    const uint4 *p = sm + (threadIdx.x & 31) + (threadIdx.x >> 5) * 8;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {          // 16 VABSDIFF4.add
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j & 7] = vsad(w[j & 7], s, a[j & 7]);
        } else if (MODE == 1) {   // 16 VABSDIFF4.add + 16 IMAD (fma pipe)
#pragma unroll
            for (int j = 0; j < 16; ++j) { a[j & 7] = vsad(w[j & 7], s, a[j & 7]); w[(j + 3) & 7] = w[(j + 3) & 7] * 3u + t; }
        } else if (MODE == 2) {   // 16 SHF
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j & 7] = __funnelshift_r(w[j & 7], w[(j + 1) & 7], s);
        } else if (MODE == 3) {   // 16 VABSDIFF4 + 8 SHF
#pragma unroll
            for (int j = 0; j < 16; ++j) { a[j & 7] = vsad(w[j & 7], s, a[j & 7]); if (j & 1) w[j & 7] = __funnelshift_r(w[j & 7], w[(j + 1) & 7], s); }
        } else if (MODE == 4) {   // 2 LDS.128 + 16 VABSDIFF4
            const uint4 x = p[(it & 15) * 64], y = p[(it & 15) * 64 + 32];
            const uint32_t v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j & 7] = vsad(v[j & 7], w[j >> 1], a[j & 7]);
        } else if (MODE == 5) {   // 3 LDS.128 + 32 VABSDIFF4 (the 8-position HME item)
            const uint4 x = p[(it & 15) * 64], y = p[(it & 15) * 64 + 32], z = p[(it & 15) * 64 + 40];
            const uint32_t v[12] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w, z.x, z.y, z.z, z.w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) a[j] = vsad(v[j + q], w[q], a[j]);
        } else if (MODE == 6) {   // 16 IADD3 (alu)
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j & 7] = a[j & 7] + w[j & 7] + s;
        } else if (MODE == 7) {   // 16 VABSDIFF4 without accumulate + 16 IADD
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j & 7] += __vabsdiffu4(w[j & 7], s) ;
        } else if (MODE == 8) {   // 16 PRMT
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j & 7] = __byte_perm(w[j & 7], w[(j + 1) & 7], 0x4321 + (s & 1));
        } else if (MODE == 9) {   // 16 DP4A
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j & 7] = __dp4a(w[j & 7], s, a[j & 7]);
        }
        s += 0x01000001u;
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && t1 - t0 == 1) out[0] = 1;
}

template <int MODE>
static void run(const char *name, int per_iter, int warps_per_smsp)
{
    uint32_t *d;
    cudaMalloc(&d, (1 << 22) + 8 * 1024 * 1024);
    const int iters = 4096, threads = 128 * warps_per_smsp;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads>>>(d, 64, 3);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(d, iters, 3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3;
    const double winstr = (double)iters * per_iter * warps_per_smsp;     // per SMSP
    printf("%-44s warps/SMSP %d: %.3f warp-instr/clk/SMSP (%.1f us, nominal clk %d kHz) err=%s\n", name, warps_per_smsp, winstr / cyc, ms * 1e3, clk,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main()
{
    for (int w : {2, 4, 8}) {
        run<0>("16 VABSDIFF4.add", 16, w);
        run<1>("16 VABSDIFF4.add + 16 IMAD", 32, w);
        run<2>("16 SHF", 16, w);
        run<3>("16 VABSDIFF4.add + 8 SHF", 24, w);
        run<4>("2 LDS.128 + 16 VABSDIFF4.add", 18, w);
        run<5>("3 LDS.128 + 32 VABSDIFF4.add", 35, w);
        run<6>("16 IADD3", 16, w);
        run<7>("16 VABSDIFF4 + 16 IADD", 32, w);
        run<8>("16 PRMT", 16, w);
        run<9>("16 DP4A", 16, w);
    }
    return 0;
}
