// tma_probe.cu -- checks the TMA assumptions the ME kernel relies on (run on the GPU box): a uint8 2-D tensor map over a padded plane, a box fetched at an
// arbitrary byte coordinate lands densely packed; boxes larger than the tensor / partly outside it are zero-filled and complete with the full byte count.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args { CUtensorMap map; int x, y, bytes; uint8_t *out; int pad[8]; CUtensorMap map2[2]; };

__global__ void probe(const __grid_constant__ Args A, int which, const CUtensorMap *gmap)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t *bar = (uint64_t *)(sm + 65536 - 64);
    const CUtensorMap *m = which == 0 ? &A.map : which == 3 ? gmap : &A.map2[which - 1];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(A.bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sm)),
                     "l"(m), "r"(A.x), "r"(A.y), "r"(smem_u32(bar))
                     : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    for (int spin = 0; !done && spin < (1 << 20); ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) A.out[A.bytes] = (uint8_t)done;
    for (int i = threadIdx.x; i < A.bytes; i += blockDim.x) A.out[i] = sm[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1, mode = argc > 2 ? atoi(argv[2]) : 0;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int pitch = 144, rows = 96;
    std::vector<uint8_t> h(pitch * rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d, *out;
    cudaMalloc(&d, h.size() + 64); cudaMalloc(&out, 70000);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    struct Case { int bw, bh, x, y; } cases[] = {{64, 64, 16, 4}, {64, 64, 17, 5}, {80, 36, 3, 1}, {176, 36, 3, 1}, {80, 70, 100, 60}, {64, 64, 0, 0}, {112, 30, 33, 90}};
    int ci = -1;
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(CUtensorMap));
    for (auto &c : cases) { ++ci; if (only >= 0 && ci != only) continue; const int which = mode;
        Args A;
        memset(&A, 0, sizeof A);
        const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)pitch};
        const cuuint32_t box[2] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh}, es[2] = {1, 1};
        CUresult r = fn(&A.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        A.map2[1] = A.map;
        A.x = c.x; A.y = c.y; A.bytes = c.bw * c.bh; A.out = out;
        cudaMemset(out, 0xEE, 70000);
        cudaMemcpy(gmap, &A.map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        probe<<<1, 128, 65536>>>(A, which, gmap);
        cudaError_t ke = cudaDeviceSynchronize();
        std::vector<uint8_t> o(A.bytes + 1);
        cudaMemcpy(o.data(), out, A.bytes + 1, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < c.bh; ++yy)
            for (int xx = 0; xx < c.bw; ++xx) {
                const int gx = c.x + xx, gy = c.y + yy;
                const uint8_t want = (gx < pitch && gy < rows) ? h[gy * pitch + gx] : 0;
                if (o[yy * c.bw + xx] != want) ++bad;
            }
        printf("box %3dx%3d at (%3d,%3d) map%d: encode=%d kernel=%s done=%d mismatches=%d sizeofArgs=%zu\n", c.bw, c.bh, c.x, c.y, which, (int)r, cudaGetErrorString(ke), (int)o[A.bytes], bad, sizeof(Args));
        if (ke != cudaSuccess) return 1;
    }
    return 0;
}
