/*
 * bcnn_b200_compat.h -- kernel-level helpers of the reference's CUDA build kept by name,
 * argument order and meaning (jnbraun/bcnn src/kernels/bcnn_mat.h:258-309), for the layers
 * outside this path and user code that still call them. They run this library's own kernels
 * (no cuBLAS) on the process-current stream (bcnn_b200_current_stream, bcnn_b200_net.h) and
 * return nothing: a failed launch prints and exits like the reference's bcnn_cuda_check.
 * Implemented in bcnn_b200/csrc/blas_compat.cu.
 *
 * The internal layer entry points with the reference's signatures (bcnn_forward_activation_gpu,
 * bcnn_forward_batchnorm_gpu, ...: src/layers/bcnn_activation_layer.h:48-51,
 * bcnn_batchnorm_layer.h:72-95) are exported too; their prototypes live in
 * the headers under bcnn_b200/src/layers because they take the internal bcnn_tensor / bcnn_mode types.
 */
#ifndef BCNN_B200_COMPAT_H
#define BCNN_B200_COMPAT_H

#include <bcnn_b200.h>

#ifdef __cplusplus
extern "C" {
#endif

/* x[i * incx] = alpha; replaces bcnn_cuda_fill_f32, bcnn_mat.h:264 */
BCNN_B200_API void bcnn_cuda_fill_f32(int n, float alpha, float *x, int incx);
/* y[i * incy] = x[i * incx]; bcnn_mat.h:265 */
BCNN_B200_API void bcnn_cuda_copy_f32(int n, float *x, int incx, float *y, int incy);
/* y += alpha * x; bcnn_mat.h:266 */
BCNN_B200_API void bcnn_cuda_axpy(int n, float alpha, float *x, int incx, float *y, int incy);
/* x *= alpha; bcnn_mat.h:267 */
BCNN_B200_API void bcnn_cuda_scal(int n, float alpha, float *x, int incx);
/* output[b, c, :] += bias[c]; bcnn_mat.h:298 */
BCNN_B200_API void bcnn_cuda_add_bias(float *output, float *bias, int batch_size, int num_channels,
                                      int spatial_size);
/* grad_bias[c] += sum over (b, position) of grad_data[b, c, :]; bcnn_mat.h:300 (deterministic
 * here; the reference's kernel races) */
BCNN_B200_API void bcnn_cuda_grad_bias(float *grad_bias, float *grad_data, int batch_size,
                                       int num_channels, int spatial_size);
/* Row-major C[m x n] = alpha * op(A)[m x k] * op(B)[k x n] + beta * C; bcnn_mat.h:258. As in the
 * reference (bcnn_mat.cu:31-45) the leading dimensions are derived from the shapes (lda = k or m,
 * ldb = n or k, ldc = n); the lda / ldb / ldc arguments are ignored. */
BCNN_B200_API void bcnn_cuda_gemm(int trans_a, int trans_b, int m, int n, int k, float alpha, float *a,
                                  int lda, float *b, int ldb, float beta, float *c, int ldc);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_B200_COMPAT_H */
