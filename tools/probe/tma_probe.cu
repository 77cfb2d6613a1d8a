// TMA semantics probe: one box load at given coordinates, dump smem.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, float *out, int nfloats) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) ((float *)smem)[i] = -777.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(nfloats * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     :: "r"(dst), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_a) : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = ((float *)smem)[i];
}
int main(int argc, char **argv) {
    int W = atoi(argv[1]), H = atoi(argv[2]), C = atoi(argv[3]), N = atoi(argv[4]);
    int c0 = atoi(argv[5]), c1 = atoi(argv[6]), c2 = atoi(argv[7]), c3 = atoi(argv[8]);
    int boxc = argc > 9 ? atoi(argv[9]) : 32;
    int swz = argc > 10 ? atoi(argv[10]) : 3;
    size_t n = (size_t)W * H * C * N;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;   // value = flat index
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int nf = 32 * boxc;
    cudaMalloc(&o, nf * 4);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {32, 1, (cuuint32_t)boxc, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    if (r) return 1;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe<<<1, 128, 32 * 1024 + 1024>>>(tm, c0, c1, c2, c3, o, nf);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    if (e) return 1;
    std::vector<float> out(nf);
    cudaMemcpy(out.data(), o, nf * 4, cudaMemcpyDeviceToHost);
    // expected: row k (channel c2+k), 16B chunk j, float e -> position w = c0 + 4*(j ^ (k&7)) + e
    int bad = 0;
    for (int k = 0; k < boxc; ++k)
        for (int j = 0; j < 8; ++j)
            for (int e2 = 0; e2 < 4; ++e2) {
                int jj = swz == 3 ? (j ^ (k & 7)) : (swz == 4 ? ((((j >> 1) ^ (k & 3)) << 1) | (j & 1)) : j);
                int w = c0 + 4 * jj + e2, hh = c1, c = c2 + k, nn = c3;
                float want = 0.f;
                if (w >= 0 && w < W && hh >= 0 && hh < H && c >= 0 && c < C && nn >= 0 && nn < N)
                    want = (float)(((size_t)nn * C + c) * H * W + (size_t)hh * W + w);
                float got = out[k * 32 + j * 4 + e2];
                if (got != want) { if (bad < 8) printf("row %d chunk %d e %d: got %g want %g\n", k, j, e2, got, want); ++bad; }
            }
    printf("mismatches: %d of %d\n", bad, nf);
    printf("row0: "); for (int i = 0; i < 12; ++i) printf("%g ", out[i]); printf("\nrow1: "); for (int i = 0; i < 12; ++i) printf("%g ", out[32 + i]); printf("\n");
    return 0;
}
