// blas_compat.cu -- the kernel-level helpers of the reference's CUDA build that OTHER layers (the
// ones outside this path: dropout, lrn, deconv, ...) and user code still call by name:
//   bcnn_cuda_fill_f32 / copy_f32 / axpy / scal / add_bias / grad_bias / gemm
// (jnbraun/bcnn src/kernels/bcnn_mat.h:258-309; implementations bcnn_mat.cu:31-91, 348-391, on
// cuBLAS and 1-thread-per-element kernels). Same names, argument order and meaning, over this
// library's own kernels: no cuBLAS. They launch on the process-current stream
// (bcnn_b200_current_stream, include/bcnn_b200_net.h) and, like the reference's, return nothing:
// a failed launch prints and exits (bcnn_cuda_check convention, src/bcnn_utils.h:174-195).
// None of this is on the hot path of a bcnn_b200 net; the GEMM is a plain 64x64x16 shared-memory
// SIMT tile kernel, correct for every transpose / size, not tuned.
#include <stdlib.h>

#include "common.cuh"

using namespace b200;

extern "C" void *bcnn_b200_current_stream(void);   // bcnn_net.c

namespace {

void check_launch(const char *what) {
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "[ERROR] [CUDA] %s: %s\n", what, cudaGetErrorString(e));
        exit((int)e);
    }
}

cudaStream_t cur() { return as_stream(bcnn_b200_current_stream()); }

__global__ void __launch_bounds__(256)
fill_strided_kernel(float *x, int n, float a, int incx) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        x[(size_t)i * incx] = a;
}
__global__ void __launch_bounds__(256)
copy_strided_kernel(const float *x, int incx, float *y, int incy, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        y[(size_t)i * incy] = x[(size_t)i * incx];
}
__global__ void __launch_bounds__(256)
axpy_strided_kernel(float a, const float *x, int incx, float *y, int incy, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        y[(size_t)i * incy] += a * x[(size_t)i * incx];
}
__global__ void __launch_bounds__(256)
scal_strided_kernel(float a, float *x, int incx, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        x[(size_t)i * incx] *= a;
}

// gb[c] += sum over (batch, position) of dy: one CTA per channel, fixed order => deterministic
// (the reference's kernel, bcnn_mat.cu:370-391, races on its shared partials).
__global__ void __launch_bounds__(256)
grad_bias_kernel(float *gb, const float *dy, int batch, int c, int hw) {
    __shared__ float red[8];
    const int ch = blockIdx.x;
    float acc[1] = {0.f};
    for (int b = 0; b < batch; ++b) {
        const float *p = dy + ((size_t)b * c + ch) * hw;
        for (int i = threadIdx.x; i < hw; i += 256) acc[0] += p[i];
    }
    block_sum<1, 256>(acc, red);
    if (threadIdx.x == 0) gb[ch] += acc[0];
}

// C[M x N] (row-major, ldc) = alpha * op(A) * op(B) + beta * C. op(A) is M x K: element (m, k) at
// A[m * lda + k] (TA == 0) or A[k * lda + m] (TA != 0); likewise op(B), K x N.
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256)
sgemm_kernel(int ta, int tb, int M, int N, int K, float alpha, const float *__restrict__ A, int lda,
             const float *__restrict__ B, int ldb, float beta, float *__restrict__ C, int ldc) {
    __shared__ float As[GK][GT + 1], Bs[GK][GT + 1];
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += GK) {
        for (int i = threadIdx.x; i < GT * GK; i += 256) {
            // walk the contiguous direction of each operand with the fast thread index
            int mm, kk;
            if (ta) { mm = i % GT; kk = i / GT; } else { kk = i % GK; mm = i / GK; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? A[ta ? (size_t)k * lda + m : (size_t)m * lda + k] : 0.f;
            int nn, kb;
            if (tb) { kb = i % GK; nn = i / GK; } else { nn = i % GT; kb = i / GT; }
            const int n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < N && k2 < K) ? B[tb ? (size_t)n * ldb + k2 : (size_t)k2 * ldb + n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float *c = C + (size_t)m * ldc + n;
            *c = alpha * acc[i][j] + (beta == 0.f ? 0.f : beta * *c);
        }
    }
}

}  // namespace

extern "C" {

BCNN_B200_API void bcnn_cuda_fill_f32(int n, float alpha, float *x, int incx) {
    if (n <= 0) return;
    fill_strided_kernel<<<stream_grid((size_t)n, 256), 256, 0, cur()>>>(x, n, alpha, incx);
    check_launch("bcnn_cuda_fill_f32");
}

BCNN_B200_API void bcnn_cuda_copy_f32(int n, float *x, int incx, float *y, int incy) {
    if (n <= 0) return;
    copy_strided_kernel<<<stream_grid((size_t)n, 256), 256, 0, cur()>>>(x, incx, y, incy, n);
    check_launch("bcnn_cuda_copy_f32");
}

BCNN_B200_API void bcnn_cuda_axpy(int n, float alpha, float *x, int incx, float *y, int incy) {
    if (n <= 0) return;
    axpy_strided_kernel<<<stream_grid((size_t)n, 256), 256, 0, cur()>>>(alpha, x, incx, y, incy, n);
    check_launch("bcnn_cuda_axpy");
}

BCNN_B200_API void bcnn_cuda_scal(int n, float alpha, float *x, int incx) {
    if (n <= 0) return;
    scal_strided_kernel<<<stream_grid((size_t)n, 256), 256, 0, cur()>>>(alpha, x, incx, n);
    check_launch("bcnn_cuda_scal");
}

BCNN_B200_API void bcnn_cuda_add_bias(float *output, float *bias, int batch_size, int num_channels,
                                      int spatial_size) {
    int err = bcnn_b200_add_bias(output, bias, batch_size, num_channels, spatial_size,
                                 bcnn_b200_current_stream());
    if (err) {
        fprintf(stderr, "[ERROR] [CUDA] bcnn_cuda_add_bias: %s\n", cudaGetErrorString((cudaError_t)err));
        exit(err);
    }
}

BCNN_B200_API void bcnn_cuda_grad_bias(float *grad_bias, float *grad_data, int batch_size,
                                       int num_channels, int spatial_size) {
    if (num_channels <= 0) return;
    grad_bias_kernel<<<num_channels, 256, 0, cur()>>>(grad_bias, grad_data, batch_size, num_channels,
                                                      spatial_size);
    check_launch("bcnn_cuda_grad_bias");
}

// Like the reference (bcnn_mat.cu:31-45) the leading dimensions follow from the shapes, whatever
// the caller passes: lda = K (or M when A is transposed), ldb = N (or K), ldc = N.
BCNN_B200_API void bcnn_cuda_gemm(int trans_a, int trans_b, int m, int n, int k, float alpha, float *a,
                                  int lda, float *b, int ldb, float beta, float *c, int ldc) {
    (void)lda; (void)ldb; (void)ldc;
    if (m <= 0 || n <= 0) return;
    const int ldaa = trans_a ? m : k, ldbb = trans_b ? k : n;
    dim3 grid(ceil_div(n, GT), ceil_div(m, GT));
    sgemm_kernel<<<grid, 256, 0, cur()>>>(trans_a, trans_b, m, n, k, alpha, a, ldaa, b, ldbb, beta, c, n);
    check_launch("bcnn_cuda_gemm");
}

}  // extern "C"
