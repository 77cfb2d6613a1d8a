// conv_impl.cuh -- internal interface between the conv dispatcher (conv.cu) and the two
// implicit-GEMM implementations.
#pragma once
#include <cuda_runtime.h>

#include "bcnn_b200.h"

namespace b200 {

// FP32 SIMT path (conv_simt.cu) -- covers every shape.
size_t conv_simt_workspace_bytes(const bcnn_b200_conv_desc *d);
int conv_simt_forward(const bcnn_b200_conv_desc *d, const float *x, const float *w,
                      const float *bias, int act, float *y, void *workspace, size_t workspace_bytes,
                      cudaStream_t st);
int conv_simt_backward_data(const bcnn_b200_conv_desc *d, const float *w, const float *dy,
                            float *dx, int accumulate, void *workspace, size_t workspace_bytes,
                            cudaStream_t st);
int conv_simt_backward_weights(const bcnn_b200_conv_desc *d, const float *x, const float *dy,
                               float *gw, void *workspace, size_t workspace_bytes,
                               cudaStream_t st);

// BF16 tcgen05 path (conv_tc.cu) -- covers the shapes conv_tc_supports_* accepts.
bool conv_tc_supports_fprop(const bcnn_b200_conv_desc *d);
bool conv_tc_supports_dgrad(const bcnn_b200_conv_desc *d);
bool conv_tc_supports_wgrad(const bcnn_b200_conv_desc *d);
size_t conv_tc_workspace_bytes(const bcnn_b200_conv_desc *d);
size_t conv_tc_wgrad_workspace_bytes(const bcnn_b200_conv_desc *d);  // conv_tc_wgrad.cu
int conv_tc_forward(const bcnn_b200_conv_desc *d, const float *x, const float *w,
                    const float *bias, int act, float *y, void *workspace,
                    size_t workspace_bytes, cudaStream_t st);
int conv_tc_backward_data(const bcnn_b200_conv_desc *d, const float *w, const float *dy,
                          float *dx, int accumulate, void *workspace, size_t workspace_bytes,
                          cudaStream_t st);
int conv_tc_backward_weights(const bcnn_b200_conv_desc *d, const float *x, const float *dy,
                             float *gw, void *workspace, size_t workspace_bytes,
                             cudaStream_t st);


// TF32 tcgen05 path with TMA-staged operands (conv_tma.cu) -- stride-1, TMA-addressable shapes.
bool conv_tma_supports_fprop(const bcnn_b200_conv_desc *d);
bool conv_tma_supports_dgrad(const bcnn_b200_conv_desc *d);
bool conv_tma_supports_wgrad(const bcnn_b200_conv_desc *d);
size_t conv_tma_workspace_bytes(const bcnn_b200_conv_desc *d);
size_t conv_tma_x_shadow_bytes(const bcnn_b200_conv_desc *d);
size_t conv_tma_dy_shadow_bytes(const bcnn_b200_conv_desc *d);
// sh (may be NULL): NHWC shadows kept between the passes of one layer (bcnn_b200_conv_shadows)
int conv_tma_forward(const bcnn_b200_conv_desc *d, const float *x, const float *w,
                     const float *bias, int act, float *y, void *workspace,
                     size_t workspace_bytes, bcnn_b200_conv_shadows *sh, cudaStream_t st);
int conv_tma_forward_stats(const bcnn_b200_conv_desc *d, const float *x, const float *w, float *y,
                           void *workspace, size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                           const float **stat_partial, int *stat_rows, cudaStream_t st);
int conv_tma_backward_data(const bcnn_b200_conv_desc *d, const float *w, const float *dy,
                           float *dx, int accumulate, void *workspace, size_t workspace_bytes,
                           bcnn_b200_conv_shadows *sh, cudaStream_t st);
int conv_tma_backward_weights(const bcnn_b200_conv_desc *d, const float *x, const float *dy,
                              float *gw, void *workspace, size_t workspace_bytes,
                              bcnn_b200_conv_shadows *sh, cudaStream_t st);

// Resident BF16 NHWC activations (conv_tma.cu): source and result are BF16 NHWC tensors, the thin
// first layer reads its FP32 NCHW input through an im2col buffer. Bit mask of the passes covered:
// 1 fprop, 2 dgrad, 4 wgrad.
int conv_nhwc_supported(const bcnn_b200_conv_desc *d);
size_t conv_nhwc_workspace_bytes(const bcnn_b200_conv_desc *d);
size_t conv_nhwc_x_keep_bytes(const bcnn_b200_conv_desc *d);
int conv_nhwc_forward(const bcnn_b200_conv_desc *d, const void *x, const float *w, const float *bias,
                      int act, void *y16, void *workspace, size_t workspace_bytes,
                      bcnn_b200_conv_shadows *sh, const float **stat_partial, int *stat_rows,
                      cudaStream_t st);
int conv_nhwc_backward_data(const bcnn_b200_conv_desc *d, const float *w, const void *dy16, void *dx16,
                            int accumulate, void *workspace, size_t workspace_bytes, cudaStream_t st);
int conv_nhwc_backward_weights(const bcnn_b200_conv_desc *d, const void *x, const void *dy16, float *gw,
                               void *workspace, size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                               cudaStream_t st);

// Packed weight images kept by the caller (conv_tma.cu): jobs for one pack launch over many layers, and
// the (weights, pass) -> image registry the resident launchers consult.
size_t conv_pack_job_bytes();
int conv_nhwc_pack_jobs(const bcnn_b200_conv_desc *d, int dgrad, const float *w, void *dst, void *jobs,
                        int max_jobs, size_t *bytes);
unsigned int conv_pack_table_finish(void *jobs, int count);
int conv_pack_run(const void *jobs_dev, int count, unsigned int grid, cudaStream_t st);
void conv_prepacked_set(const float *w, int dgrad, const void *image);
void conv_prepacked_enable(int on);

// batchnorm.cu: TRAIN statistics from the per-tile partial sums a convolution epilogue left
// (partial[(row * 2 + {0: sum, 1: sum of squares}) * c + channel]), folded in a fixed order.
int bn_stats_from_partials(const float *partial, int rows, int c, double count, float *saved_mean,
                           float *saved_var, float *run_mean, float *run_var, float *scratch,
                           cudaStream_t st);

}  // namespace b200
