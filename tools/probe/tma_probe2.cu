// TMA probe 2: NHWC tensor (c, w, h, n), box (32, bw, bh, bn) with element strides (1, s, s, 1).
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
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, float *out, int nfloats, int tx) {
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
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(tx) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     :: "r"(dst), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_a) : "memory");
    }
    __syncthreads();
    uint32_t done = 0; long long spins = 0;
    while (!done && spins < 2000000) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
        ++spins;
    }
    if (threadIdx.x == 0 && !done) printf("TIMEOUT waiting for %d tx bytes\n", tx);
    __syncthreads();
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = ((float *)smem)[i];
}
int main(int argc, char **argv) {
    int C = atoi(argv[1]), W = atoi(argv[2]), H = atoi(argv[3]), N = atoi(argv[4]);
    int bw = atoi(argv[5]), bh = atoi(argv[6]), bn = atoi(argv[7]), s = atoi(argv[8]);
    int c0 = atoi(argv[9]), c1 = atoi(argv[10]), c2 = atoi(argv[11]), c3 = atoi(argv[12]);
    int swz = argc > 13 ? atoi(argv[13]) : 3;
    size_t n = (size_t)W * H * C * N;
    std::vector<float> h(n);
    for (int nn = 0; nn < N; ++nn) for (int hh = 0; hh < H; ++hh) for (int ww = 0; ww < W; ++ww) for (int cc = 0; cc < C; ++cc)
        h[(((size_t)nn * H + hh) * W + ww) * C + cc] = nn * 1000000.f + hh * 10000.f + ww * 100.f + cc;
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int rows = bw * bh * bn, nf = 32 * rows;
    cudaMalloc(&o, nf * 4);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(bw * s), (cuuint32_t)(bh * s), (cuuint32_t)bn}; cuuint32_t es[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    if (r) return 1;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    probe<<<1, 128, nf * 4 + 1024>>>(tm, c0, c1, c2, c3, o, nf, nf * 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    if (e) return 1;
    std::vector<float> out(nf);
    cudaMemcpy(out.data(), o, nf * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < rows; ++row) {
        int ww = row % bw, hh = (row / bw) % bh, nn = row / (bw * bh);
        int gw = c1 + ww * s, gh = c2 + hh * s, gn = c3 + nn;
        for (int j = 0; j < 8; ++j) for (int e2 = 0; e2 < 4; ++e2) {
            int jj = swz == 3 ? (j ^ (row & 7)) : (swz == 4 ? ((((j >> 1) ^ (row & 3)) << 1) | (j & 1)) : j);
            int cc = c0 + 4 * jj + e2;
            float want = 0.f;
            if (gw >= 0 && gw < W && gh >= 0 && gh < H && gn >= 0 && gn < N && cc < C) want = gn * 1000000.f + gh * 10000.f + gw * 100.f + cc;
            float got = out[row * 32 + j * 4 + e2];
            if (got != want) { if (bad < 6) printf("row %d (w%d h%d n%d) chunk %d: got %g want %g\n", row, gw, gh, gn, j, got, want); ++bad; }
        }
    }
    printf("mismatches: %d of %d\n", bad, nf);
    return 0;
}
