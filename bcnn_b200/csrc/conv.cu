// conv.cu -- C-ABI convolution entry points (include/bcnn_b200.h) and the dispatcher
// between the FP32 SIMT implicit GEMM (conv_simt.cu) and the BF16 tcgen05 implicit GEMM
// (conv_tc.cu). There is no CPU fallback and no cuDNN / cuBLAS on either path.
#include <cstdlib>

#include "common.cuh"
#include "conv_impl.cuh"

using namespace b200;

extern "C" int bcnn_b200_conv_uses_tensor_cores(const bcnn_b200_conv_desc *d, int pass) {
    switch (pass) {
        case 0: return (conv_tma_supports_fprop(d) || conv_tc_supports_fprop(d)) ? 1 : 0;
        case 1: return (conv_tma_supports_dgrad(d) || conv_tc_supports_dgrad(d)) ? 1 : 0;
        case 2: return (conv_tma_supports_wgrad(d) || conv_tc_supports_wgrad(d)) ? 1 : 0;
    }
    return 0;
}

extern "C" size_t bcnn_b200_conv_workspace_bytes(const bcnn_b200_conv_desc *d, int math) {
    size_t a = conv_simt_workspace_bytes(d);
    size_t b = (math != BCNN_B200_MATH_FP32) ? conv_tc_workspace_bytes(d) : 0;
    size_t c = (math != BCNN_B200_MATH_FP32) ? conv_tma_workspace_bytes(d) : 0;
    if (b > a) a = b;
    return a > c ? a : c;
}

extern "C" size_t bcnn_b200_conv_x_shadow_bytes(const bcnn_b200_conv_desc *d, int math) {
    return math != BCNN_B200_MATH_FP32 ? conv_tma_x_shadow_bytes(d) : 0;
}

extern "C" size_t bcnn_b200_conv_dy_shadow_bytes(const bcnn_b200_conv_desc *d, int math) {
    return math != BCNN_B200_MATH_FP32 ? conv_tma_dy_shadow_bytes(d) : 0;
}

extern "C" int bcnn_b200_conv_forward(const bcnn_b200_conv_desc *d, const float *x,
                                      const float *w, const float *bias, int act, float *y,
                                      void *workspace, size_t workspace_bytes, int math,
                                      void *stream) {
    return bcnn_b200_conv_forward_sh(d, x, w, bias, act, y, workspace, workspace_bytes, math,
                                     nullptr, stream);
}

extern "C" int bcnn_b200_conv_backward_data(const bcnn_b200_conv_desc *d, const float *w,
                                            const float *dy, float *dx, int accumulate,
                                            void *workspace, size_t workspace_bytes, int math,
                                            void *stream) {
    return bcnn_b200_conv_backward_data_sh(d, w, dy, dx, accumulate, workspace, workspace_bytes,
                                           math, nullptr, stream);
}

extern "C" int bcnn_b200_conv_backward_weights(const bcnn_b200_conv_desc *d, const float *x,
                                               const float *dy, float *gw, void *workspace,
                                               size_t workspace_bytes, int math, void *stream) {
    return bcnn_b200_conv_backward_weights_sh(d, x, dy, gw, workspace, workspace_bytes, math,
                                              nullptr, stream);
}

extern "C" int bcnn_b200_conv_forward_sh(const bcnn_b200_conv_desc *d, const float *x,
                                         const float *w, const float *bias, int act, float *y,
                                         void *workspace, size_t workspace_bytes, int math,
                                         bcnn_b200_conv_shadows *sh, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (math != BCNN_B200_MATH_FP32 && conv_tma_supports_fprop(d))
        return conv_tma_forward(d, x, w, bias, act, y, workspace, workspace_bytes, sh, st);
    if (math != BCNN_B200_MATH_FP32 && conv_tc_supports_fprop(d))
        return conv_tc_forward(d, x, w, bias, act, y, workspace, workspace_bytes, st);
    return conv_simt_forward(d, x, w, bias, act, y, workspace, workspace_bytes, st);
}

extern "C" int bcnn_b200_conv_forward_bn_stats(const bcnn_b200_conv_desc *d, const float *x,
                                               const float *w, float *y, void *workspace,
                                               size_t workspace_bytes, int math,
                                               bcnn_b200_conv_shadows *sh, float *saved_mean,
                                               float *saved_var, float *run_mean, float *run_var,
                                               float *scratch, void *stream) {
    cudaStream_t st = as_stream(stream);
    static int unfused = -1;
    if (unfused < 0) {
        const char *e = getenv("BCNN_B200_NO_FUSED_BN_STATS");
        unfused = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    if (!unfused && math != BCNN_B200_MATH_FP32 && conv_tma_supports_fprop(d)) {
        const float *partial = nullptr;
        int rows = 0;
        int err = conv_tma_forward_stats(d, x, w, y, workspace, workspace_bytes, sh, &partial, &rows, st);
        if (err) return err;
        if (partial)
            return bn_stats_from_partials(partial, rows, d->cout,
                                          (double)d->batch * d->ho * d->wo, saved_mean, saved_var,
                                          run_mean, run_var, scratch, st);
    } else {
        int err = bcnn_b200_conv_forward_sh(d, x, w, nullptr, 0, y, workspace, workspace_bytes, math, sh,
                                            stream);
        if (err) return err;
    }
    return bcnn_b200_bn_stats(y, d->batch, d->cout, d->ho * d->wo, saved_mean, saved_var, run_mean,
                              run_var, scratch, stream);
}

extern "C" int bcnn_b200_conv_backward_data_sh(const bcnn_b200_conv_desc *d, const float *w,
                                               const float *dy, float *dx, int accumulate,
                                               void *workspace, size_t workspace_bytes, int math,
                                               bcnn_b200_conv_shadows *sh, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (math != BCNN_B200_MATH_FP32 && conv_tma_supports_dgrad(d))
        return conv_tma_backward_data(d, w, dy, dx, accumulate, workspace, workspace_bytes, sh, st);
    if (math != BCNN_B200_MATH_FP32 && conv_tc_supports_dgrad(d))
        return conv_tc_backward_data(d, w, dy, dx, accumulate, workspace, workspace_bytes, st);
    return conv_simt_backward_data(d, w, dy, dx, accumulate, workspace, workspace_bytes, st);
}

extern "C" int bcnn_b200_conv_backward_weights_sh(const bcnn_b200_conv_desc *d, const float *x,
                                                  const float *dy, float *gw, void *workspace,
                                                  size_t workspace_bytes, int math,
                                                  bcnn_b200_conv_shadows *sh, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (math != BCNN_B200_MATH_FP32 && conv_tma_supports_wgrad(d))
        return conv_tma_backward_weights(d, x, dy, gw, workspace, workspace_bytes, sh, st);
    if (math != BCNN_B200_MATH_FP32 && conv_tc_supports_wgrad(d))
        return conv_tc_backward_weights(d, x, dy, gw, workspace, workspace_bytes, st);
    return conv_simt_backward_weights(d, x, dy, gw, workspace, workspace_bytes, st);
}

// ---------------------------------------------------------------- resident BF16 NHWC activations
extern "C" int bcnn_b200_conv_nhwc_supported(const bcnn_b200_conv_desc *d) { return conv_nhwc_supported(d); }
extern "C" size_t bcnn_b200_conv_nhwc_workspace_bytes(const bcnn_b200_conv_desc *d) {
    return conv_nhwc_workspace_bytes(d);
}
extern "C" size_t bcnn_b200_conv_nhwc_x_keep_bytes(const bcnn_b200_conv_desc *d) {
    return conv_nhwc_x_keep_bytes(d);
}
extern "C" int bcnn_b200_conv_forward_nhwc(const bcnn_b200_conv_desc *d, const void *x, const float *w,
                                           const float *bias, int act, void *y, void *workspace,
                                           size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                                           void *stream) {
    return conv_nhwc_forward(d, x, w, bias, act, y, workspace, workspace_bytes, sh, nullptr, nullptr,
                             as_stream(stream));
}
extern "C" int bcnn_b200_conv_forward_bn_stats_nhwc(const bcnn_b200_conv_desc *d, const void *x,
                                                    const float *w, void *y, void *workspace,
                                                    size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                                                    float *saved_mean, float *saved_var, float *run_mean,
                                                    float *run_var, float *nhwc_scratch, float *scratch,
                                                    void *stream) {
    cudaStream_t st = as_stream(stream);
    const float *partial = nullptr;
    int rows = 0;
    int err = conv_nhwc_forward(d, x, w, nullptr, 0, y, workspace, workspace_bytes, sh, &partial, &rows, st);
    if (err) return err;
    const size_t positions = (size_t)d->batch * d->ho * d->wo;
    if (partial)
        return bn_stats_from_partials(partial, rows, d->cout, (double)positions, saved_mean, saved_var,
                                      run_mean, run_var, scratch, st);
    return bcnn_b200_bn_stats_nhwc(y, positions, d->cout, saved_mean, saved_var, run_mean, run_var,
                                   nhwc_scratch, scratch, stream);
}
extern "C" size_t bcnn_b200_conv_pack_job_bytes(void) { return conv_pack_job_bytes(); }
extern "C" int bcnn_b200_conv_nhwc_pack_jobs(const bcnn_b200_conv_desc *d, int dgrad, const float *w, void *dst,
                                             void *jobs, int max_jobs, size_t *bytes) {
    return conv_nhwc_pack_jobs(d, dgrad, w, dst, jobs, max_jobs, bytes);
}
extern "C" unsigned int bcnn_b200_conv_pack_table_finish(void *jobs, int count) {
    return conv_pack_table_finish(jobs, count);
}
extern "C" int bcnn_b200_conv_pack_run(const void *jobs_dev, int count, unsigned int grid, void *stream) {
    return conv_pack_run(jobs_dev, count, grid, as_stream(stream));
}
extern "C" void bcnn_b200_conv_prepacked_set(const float *w, int dgrad, const void *image) {
    conv_prepacked_set(w, dgrad, image);
}
extern "C" void bcnn_b200_conv_prepacked_enable(int on) { conv_prepacked_enable(on); }
extern "C" int bcnn_b200_conv_backward_data_nhwc(const bcnn_b200_conv_desc *d, const float *w,
                                                 const void *dy, void *dx, int accumulate,
                                                 void *workspace, size_t workspace_bytes, void *stream) {
    return conv_nhwc_backward_data(d, w, dy, dx, accumulate, workspace, workspace_bytes, as_stream(stream));
}
extern "C" int bcnn_b200_conv_backward_weights_nhwc(const bcnn_b200_conv_desc *d, const void *x,
                                                    const void *dy, float *gw, void *workspace,
                                                    size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                                                    void *stream) {
    return conv_nhwc_backward_weights(d, x, dy, gw, workspace, workspace_bytes, sh, as_stream(stream));
}
