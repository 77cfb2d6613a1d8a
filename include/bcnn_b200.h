/*
 * bcnn_b200.h -- kernel-level C ABI of libbcnn_b200.so.
 *
 * These are the entry points the C layer files (the .c files under bcnn_b200/src/layers) call
 * where jnbraun/bcnn's layer files call its bcnn_cuda_* helpers, cuBLAS and
 * cuDNN.  Plain pointers and sizes only: device pointers are raw `float *` /
 * `int *` into cudaMalloc'd memory, `stream` is a cudaStream_t passed as
 * `void *` (NULL = legacy default stream).  Every launcher returns 0 on success
 * or the cudaError_t of the failed launch; nothing here synchronises.
 * All tensors are NCHW float32; element counts fit an int (bcnn_tensor_size,
 * reference src/bcnn_tensor.c:97).
 *
 * Each declaration cites the reference interface it replaces
 * (paths relative to jnbraun/bcnn @ 3825c4f).
 */
#ifndef BCNN_B200_H
#define BCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCNN_B200_API __attribute__((visibility("default")))

/* Convolution arithmetic selection (per net, see bcnn_b200_set_conv_math). */
enum {
    BCNN_B200_MATH_FP32 = 0, /* FP32 SIMT implicit GEMM: 1e-5 verification path */
    BCNN_B200_MATH_TC = 1,   /* BF16 / TF32 tcgen05 implicit GEMM, FP32 accumulate in TMEM:
                                2e-2 path; falls back per layer to FP32 for shapes
                                it does not cover (never to the CPU) */
    BCNN_B200_MATH_TC_BF16 = 2 /* the same tensor-core kernels with RESIDENT activations:
                                between two convolutions tensors and their gradients live
                                as BF16 NHWC (see the block at the end of this header);
                                FP32 NCHW copies are made when somebody asks for them.
                                Kernel-level entry points treat it as MATH_TC. */
};

/* Geometry of one convolution; shared by fprop / dgrad / wgrad. */
typedef struct bcnn_b200_conv_desc {
    int batch, cin, h, w;      /* input  [batch, cin, h, w]          */
    int cout, ho, wo;          /* output [batch, cout, ho, wo]       */
    int ksize, stride, pad;    /* square kernel                      */
    int groups;                /* cin % groups == cout % groups == 0 */
} bcnn_b200_conv_desc;

/* NHWC shadows of one convolution layer, kept between the passes of a training step.
 * The TMA cannot shift an NCHW row by one pixel, so k > 1 / strided / odd-plane layers read
 * an NHWC (BF16 or FP32) copy of their operand.  A layer that owns storage for them hands
 * it to the *_sh entry points: fprop leaves the x shadow (or the im2col buffer of a thin
 * first layer) in `x` for wgrad, wgrad leaves the dy shadow in `dy` for dgrad, so each
 * operand is transposed once per step instead of once per pass.  `*_fmt` says what the
 * buffer currently holds (BCNN_B200_SHADOW_*); the caller resets it to NONE whenever the
 * FP32 tensor it mirrors changes.  No reference counterpart (its im2col workspace,
 * src/layers/bcnn_conv_layer.c:141-144, plays the same role per image). */
enum {
    BCNN_B200_SHADOW_NONE = 0,
    BCNN_B200_SHADOW_NHWC_F32 = 1,
    BCNN_B200_SHADOW_NHWC_BF16 = 2,
    BCNN_B200_SHADOW_IM2COL_F32 = 3,
    BCNN_B200_SHADOW_IM2COL_BF16 = 4
};
typedef struct bcnn_b200_conv_shadows {
    void *x;  size_t x_bytes;  int x_fmt;  /* 256-byte aligned device memory or NULL */
    void *dy; size_t dy_bytes; int dy_fmt;
} bcnn_b200_conv_shadows;

/* ---- device / stream / memory helpers --------------------------------- */
/* replaces bcnn_cuda_set_device, src/bcnn_utils.c:201 */
BCNN_B200_API int bcnn_b200_set_device(int device);
BCNN_B200_API int bcnn_b200_device_count(void);
BCNN_B200_API int bcnn_b200_sm_count(void);
/* replaces bcnn_cuda_malloc_f32 / _i32 / bcnn_cuda_free, src/bcnn_utils.c:124-160
 * (memory comes back zero-filled, like the reference's calloc + H2D copy) */
BCNN_B200_API void *bcnn_b200_malloc(size_t bytes);
BCNN_B200_API void bcnn_b200_free(void *dev_ptr);
BCNN_B200_API void *bcnn_b200_malloc_host(size_t bytes); /* pinned */
BCNN_B200_API void bcnn_b200_free_host(void *host_ptr);
/* replaces bcnn_cuda_memcpy_host2dev / dev2host, src/bcnn_utils.c:189-199 */
BCNN_B200_API int bcnn_b200_memcpy_h2d(void *dst_dev, const void *src_host,
                                       size_t bytes, void *stream);
BCNN_B200_API int bcnn_b200_memcpy_d2h(void *dst_host, const void *src_dev,
                                       size_t bytes, void *stream);
BCNN_B200_API int bcnn_b200_memcpy_d2d(void *dst_dev, const void *src_dev,
                                       size_t bytes, void *stream);
BCNN_B200_API void *bcnn_b200_stream_create(void);
BCNN_B200_API void bcnn_b200_stream_destroy(void *stream);
BCNN_B200_API int bcnn_b200_stream_sync(void *stream);
BCNN_B200_API void *bcnn_b200_event_create(void);
BCNN_B200_API void bcnn_b200_event_destroy(void *event);
BCNN_B200_API int bcnn_b200_event_record(void *event, void *stream);
BCNN_B200_API int bcnn_b200_stream_wait_event(void *stream, void *event);
BCNN_B200_API float bcnn_b200_event_elapsed_ms(void *start, void *stop);
/* CUDA graphs (no counterpart in the reference, whose CUDA path launches every kernel from the
 * host on the default stream): record the work queued on `stream` between begin and end, then
 * replay it with one launch. graph_end returns NULL when the capture failed; captured work is
 * not executed until the graph is launched. */
BCNN_B200_API int bcnn_b200_graph_begin(void *stream);
BCNN_B200_API void *bcnn_b200_graph_end(void *stream);
/* kernels: how many launches the graph holds (added to bcnn_b200_launch_count per replay). */
BCNN_B200_API int bcnn_b200_graph_launch(void *graph_exec, unsigned long long kernels,
                                         void *stream);
BCNN_B200_API void bcnn_b200_graph_destroy(void *graph_exec);
BCNN_B200_API const char *bcnn_b200_error_string(int err);
/* number of kernels this library has launched in this process (bench.py's
 * gpu_launches claim) */
BCNN_B200_API uint64_t bcnn_b200_launch_count(void);

/* ---- BLAS-1 class ------------------------------------------------------ */
/* replaces bcnn_cuda_fill_f32, src/kernels/bcnn_mat.h:264 */
BCNN_B200_API int bcnn_b200_fill_f32(float *x, size_t n, float value, void *stream);
/* y += a*x ; replaces bcnn_cuda_axpy, bcnn_mat.h:266 */
BCNN_B200_API int bcnn_b200_axpy(float *y, const float *x, size_t n, float a,
                                 void *stream);

/* ---- max pooling ------------------------------------------------------- */
/* replaces bcnn_forward_maxpool_layer_kernel, src/layers/bcnn_maxpool_layer.cu:28-70
 * with the CPU rule of src/layers/bcnn_maxpool_layer.c:145-191 (first max wins,
 * int32 flat NCHW argmax, -1 when the window is empty). */
BCNN_B200_API int bcnn_b200_maxpool_forward(const float *x, float *y, int *indexes,
                                            int n, int c, int h, int w, int ksize,
                                            int stride, int ho, int wo, void *stream);
/* dx[indexes[o]] += dy[o]; replaces bcnn_backward_maxpool_layer_kernel,
 * bcnn_maxpool_layer.cu:97-138 (gather form, no atomics, CPU summation order). */
BCNN_B200_API int bcnn_b200_maxpool_backward(float *dx, const float *dy,
                                             const int *indexes, int n, int c, int h,
                                             int w, int ksize, int stride, int ho,
                                             int wo, void *stream);

/* ---- global average pooling -------------------------------------------- */
/* replaces _bcnn_forward_avgpool_layer_kernel / _backward_, src/layers/
 * bcnn_avgpool_layer.cu:29-75; planes = n*c, hw = h*w.  y = mean; dx += dy/hw. */
BCNN_B200_API int bcnn_b200_avgpool_forward(const float *x, float *y, int planes,
                                            int hw, void *stream);
BCNN_B200_API int bcnn_b200_avgpool_backward(float *dx, const float *dy, int planes,
                                             int hw, void *stream);

/* ---- activations ------------------------------------------------------- */
/* `act` is a bcnn_activation value.  slope / g_slope are the per-channel PReLU
 * parameters (NULL otherwise).  In place.  replaces bcnn_forward_activation_gpu /
 * bcnn_backward_activation_gpu, src/layers/bcnn_activation_layer.cu:64-81,115-135,
 * with all ten activations of the CPU path (bcnn_activation_layer.c:90-226). */
BCNN_B200_API int bcnn_b200_activation_forward(float *x, int sz, int act,
                                               const float *slope, int hw, int c,
                                               void *stream);
BCNN_B200_API int bcnn_b200_activation_backward(const float *y, float *dy, int sz,
                                                int act, const float *slope,
                                                float *g_slope, int hw, int c,
                                                void *stream);

/* ---- per-channel bias -------------------------------------------------- */
/* y[n,c,:] += b[c]; replaces bcnn_cuda_add_bias, src/kernels/bcnn_mat.cu:348-368 */
BCNN_B200_API int bcnn_b200_add_bias(float *y, const float *bias, int n, int c, int hw,
                                     void *stream);
/* gb[c] += sum dy[:,c,:]; replaces bcnn_cuda_grad_bias, bcnn_mat.cu:370-391 (which
 * races); optionally applies the activation derivative first (dy *= act'(y) in
 * place, y may be NULL when act == NONE). */
BCNN_B200_API int bcnn_b200_actbwd_grad_bias(float *gb, float *dy, const float *y,
                                             int act, int n, int c, int hw,
                                             float *scratch, void *stream);

/* ---- batch normalisation ----------------------------------------------- */
/* Scratch needed by the two reduction entry points, in floats. */
BCNN_B200_API size_t bcnn_b200_bn_scratch_floats(int c);
/* TRAIN statistics: mean = sum(x)/m, var = sum(x^2)/m - mean^2 (biased, as
 * _mean_variance_forward, src/layers/bcnn_batchnorm_layer.c:147-168), then
 * running = 0.9 running + 0.1 batch (:221-224).  replaces fast_mean_kernel /
 * fast_variance_kernel, src/layers/bcnn_batchnorm_layer.cu:28-89. */
BCNN_B200_API int bcnn_b200_bn_stats(const float *x, int n, int c, int hw,
                                     float *saved_mean, float *saved_var,
                                     float *run_mean, float *run_var, float *scratch,
                                     void *stream);
/* y = act(gamma * (x - mean) / sqrt(var + 1e-6) + beta)  (x may alias y).
 * replaces _norm_forward_kernel + bcnn_scales_kernel + bcnn_cuda_add_bias_kernel
 * (+ activation kernel), src/kernels/bcnn_mat.cu:179-193,393-410,348-368. */
BCNN_B200_API int bcnn_b200_bn_apply(const float *x, float *y, const float *mean,
                                     const float *var, const float *gamma,
                                     const float *beta, int n, int c, int hw, int act,
                                     void *stream);
/* PREDICT mode: y = act(gamma * x + beta) (statistics pre-folded at load time,
 * scale_and_add_bias, bcnn_batchnorm_layer.c:184-194). */
BCNN_B200_API int bcnn_b200_scale_bias(const float *x, float *y, const float *gamma,
                                       const float *beta, int n, int c, int hw,
                                       int act, void *stream);
/* Backward of (activation o batchnorm), in place on dy -> dx:
 *   dy' = dy * act'(y);  g_beta += sum dy';  g_gamma += sum dy' * xhat;
 *   d_mean, d_var as _mean_variance_backward (bcnn_batchnorm_layer.c:263-281,
 *   eps 1e-5, var*sqrt(var) form); dx as _normalize_backward (:283-299).
 * x is the pre-normalisation input kept by the forward pass; y the
 * post-activation output (may be NULL when act == NONE).  dx_out may alias dy.
 * beta (may be NULL): the shift the forward used. With it, ReLU / leaky-ReLU masks
 * are rebuilt from x with the forward's own arithmetic (y > 0 <=> gamma * xhat +
 * beta > 0) and y is not read: 20 instead of 28 bytes per element.
 * replaces fast_mean_delta_kernel / fast_variance_delta_kernel /
 * _norm_backward_kernel / bcnn_grad_scales_kernel. */
BCNN_B200_API int bcnn_b200_bn_backward(const float *x, const float *y, float *dy,
                                        float *dx_out, const float *mean,
                                        const float *var, const float *gamma,
                                        const float *beta, float *g_gamma,
                                        float *g_beta, float *d_mean, float *d_var,
                                        int n, int c, int hw, int act, float *scratch,
                                        void *stream);

/* ---- convolution ------------------------------------------------------- */
/* 1 when pass (0 fprop, 1 dgrad, 2 wgrad) of `d` runs on the tcgen05 kernel under
 * BCNN_B200_MATH_TC, 0 when it takes the FP32 SIMT kernel (thin-K first layers, groups). */
BCNN_B200_API int bcnn_b200_conv_uses_tensor_cores(const bcnn_b200_conv_desc *d, int pass);
/* Bytes of device workspace the three conv entry points may use for `d`. */
BCNN_B200_API size_t bcnn_b200_conv_workspace_bytes(const bcnn_b200_conv_desc *d,
                                                    int math);
/* Bytes of shadow storage worth keeping for `d` (0: its passes read NCHW directly). */
BCNN_B200_API size_t bcnn_b200_conv_x_shadow_bytes(const bcnn_b200_conv_desc *d, int math);
BCNN_B200_API size_t bcnn_b200_conv_dy_shadow_bytes(const bcnn_b200_conv_desc *d, int math);
/* y = act(W (*) x + bias)   (bias may be NULL, act may be NONE).
 * replaces the per-image bcnn_cuda_im2col + bcnn_cuda_gemm loop and
 * bcnn_cuda_add_bias, src/layers/bcnn_conv_layer.c:628-656 (and the cuDNN
 * branch :609-620).  One launch for the whole batch. */
BCNN_B200_API int bcnn_b200_conv_forward(const bcnn_b200_conv_desc *d, const float *x,
                                         const float *w, const float *bias, int act,
                                         float *y, void *workspace,
                                         size_t workspace_bytes, int math,
                                         void *stream);
/* dx = W^T (*) dy, overwriting dx (beta = 0 GEMM + zero-filling col2im of the
 * reference, bcnn_conv_layer.c:567-578, bcnn_mat.c:944); accumulate != 0 gives
 * dx += (used by the fully-connected layer, bcnn_fc_layer.c:217-223). */
BCNN_B200_API int bcnn_b200_conv_backward_data(const bcnn_b200_conv_desc *d,
                                               const float *w, const float *dy,
                                               float *dx, int accumulate,
                                               void *workspace, size_t workspace_bytes,
                                               int math, void *stream);
/* gw += dy (*) x over the whole batch (beta = 1, bcnn_conv_layer.c:551). */
BCNN_B200_API int bcnn_b200_conv_backward_weights(const bcnn_b200_conv_desc *d,
                                                  const float *x, const float *dy,
                                                  float *gw, void *workspace,
                                                  size_t workspace_bytes, int math,
                                                  void *stream);
/* The same three passes with shadow storage (sh may be NULL = the plain entry points). */
BCNN_B200_API int bcnn_b200_conv_forward_sh(const bcnn_b200_conv_desc *d, const float *x,
                                            const float *w, const float *bias, int act,
                                            float *y, void *workspace,
                                            size_t workspace_bytes, int math,
                                            bcnn_b200_conv_shadows *sh, void *stream);
/* Convolution feeding a TRAIN-mode batch normalisation: y = W (*) x (no bias, no
 * activation) and the statistics bcnn_b200_bn_stats would compute over y (saved mean /
 * biased variance, running update), taken from the FP32 accumulators in the convolution
 * epilogue so that y is not read back: the first of the reference's seven passes over
 * the tensor (bcnn_forward_batchnorm_cpu, src/layers/bcnn_batchnorm_layer.c:196-242,
 * called from bcnn_conv_layer.c:470) disappears. Shapes the TMA kernel does not cover
 * run convolution and bcnn_b200_bn_stats back to back. `scratch` as bcnn_b200_bn_stats. */
BCNN_B200_API int bcnn_b200_conv_forward_bn_stats(const bcnn_b200_conv_desc *d,
                                                  const float *x, const float *w, float *y,
                                                  void *workspace, size_t workspace_bytes,
                                                  int math, bcnn_b200_conv_shadows *sh,
                                                  float *saved_mean, float *saved_var,
                                                  float *run_mean, float *run_var,
                                                  float *scratch, void *stream);
BCNN_B200_API int bcnn_b200_conv_backward_data_sh(const bcnn_b200_conv_desc *d,
                                                  const float *w, const float *dy,
                                                  float *dx, int accumulate,
                                                  void *workspace, size_t workspace_bytes,
                                                  int math, bcnn_b200_conv_shadows *sh,
                                                  void *stream);
BCNN_B200_API int bcnn_b200_conv_backward_weights_sh(const bcnn_b200_conv_desc *d,
                                                     const float *x, const float *dy,
                                                     float *gw, void *workspace,
                                                     size_t workspace_bytes, int math,
                                                     bcnn_b200_conv_shadows *sh,
                                                     void *stream);

/* ---- depthwise convolution --------------------------------------------- */
/* y = act(dw(x, w) + bias); replaces _bcnn_forward_depthwise_conv_weight_kernel,
 * src/layers/bcnn_depthwise_conv_layer.cu:33-62 (+ add_bias + activation). */
BCNN_B200_API int bcnn_b200_depthwise_forward(const float *x, const float *w,
                                              const float *bias, int act, float *y,
                                              int n, int c, int h, int wd, int ksize,
                                              int stride, int pad, void *stream);
/* gw += ..., dx += ... ; replaces the racy weight kernel and the data kernel,
 * bcnn_depthwise_conv_layer.cu:88-155. */
BCNN_B200_API int bcnn_b200_depthwise_backward(const float *x, const float *w,
                                               const float *dy, float *gw, float *dx,
                                               int n, int c, int h, int wd, int ksize,
                                               int stride, int pad, float *scratch,
                                               size_t scratch_floats, void *stream);
BCNN_B200_API size_t bcnn_b200_depthwise_scratch_floats(int n, int c, int ksize);

/* ---- optimizer --------------------------------------------------------- */
/* One fused pass of bcnn_sgd_update_gpu (src/bcnn_learner.c:86-103):
 *   g += wd_scale * w;  w += step * g;  g *= g_scale
 * with wd_scale = decay*batch (0 for biases), step = -lr/batch, g_scale =
 * momentum (momentum / world_size under data parallelism). */
BCNN_B200_API int bcnn_b200_sgd_update(float *w, float *g, size_t n, float wd_scale,
                                       float step, float g_scale, void *stream);
/* The same pass over up to BCNN_B200_SGD_MULTI_MAX parameter tensors in ONE launch (a ResNet-50 step
 * has 108 of them, most of a few KB: 108 launches of ~3 us each plus the gaps between them). Every
 * tensor keeps its own wd_scale; step and g_scale are common. first_block[i] = first CTA of tensor i
 * (4096 elements per CTA), first_block[count] = grid size; w[i] / g[i] must be 16-byte aligned.
 * Element arithmetic is that of bcnn_b200_sgd_update: results are bit-identical. */
#define BCNN_B200_SGD_MULTI_MAX 96
typedef struct bcnn_b200_sgd_batch {
    float *w[BCNN_B200_SGD_MULTI_MAX];
    float *g[BCNN_B200_SGD_MULTI_MAX];
    unsigned int n[BCNN_B200_SGD_MULTI_MAX];
    float wd_scale[BCNN_B200_SGD_MULTI_MAX];
    unsigned int first_block[BCNN_B200_SGD_MULTI_MAX + 1];
    int count;
    float step, g_scale;
} bcnn_b200_sgd_batch;
BCNN_B200_API int bcnn_b200_sgd_update_multi(const bcnn_b200_sgd_batch *batch, void *stream);
/* One fused pass of bcnn_adam_update_gpu's weight branch (src/bcnn_learner.c:148-161; the CPU
 * arithmetic of :118-129 is the parity target):
 *   g += wd_scale * w;  m = (1-beta1) g + beta1 m;  v = (1-beta2) g^2 + beta2 v;
 *   w += alpha * m / (sqrt(v) + 1e-7);  g = 0
 * with wd_scale = decay*batch and alpha = -lr/batch * sqrt(1-beta2^(t+1)) / (1-beta1^(t+1))
 * computed by the caller (t = samples seen). Biases take bcnn_b200_sgd_update. */
BCNN_B200_API int bcnn_b200_adam_update(float *w, float *g, float *m, float *v, size_t n,
                                        float wd_scale, float beta1, float beta2, float alpha,
                                        void *stream);

/* ---- glue kernels ---------------------------------------- */
/* softmax over channels at each spatial position, log-sum-exp form of
 * src/layers/bcnn_softmax_layer.c:88-155 */
BCNN_B200_API int bcnn_b200_softmax_forward(const float *x, float *y, int n, int c,
                                            int hw, void *stream);
/* grad = pred - label (bcnn_euclidean_loss_forward, src/layers/bcnn_cost_layer.c
 * :111-128) and metric[0] = #misclassified (ERROR_RATE) / sum sq (SSE) /
 * SSE/input_size (MSE) / logloss, computed on device (the reference copies to
 * the host, :142-158). metric_kind is a bcnn_loss_metric. */
BCNN_B200_API int bcnn_b200_cost_forward(const float *pred, const float *label,
                                         float *grad, float *metric, int n,
                                         int input_size, int metric_kind,
                                         void *stream);
/* y = act(a + b') over sz elements, where b' = b on the first n_add elements and 0
 * beyond. n_add = sz is the batch-correct residual add; n_add = C*H*W reproduces the
 * reference, whose equal-shape path adds only sample 0 (src/layers/bcnn_eltwise_layer.c
 * :119-121). replaces bcnn_cuda_copy_f32 + bcnn_cuda_axpy + activation kernel. */
BCNN_B200_API int bcnn_b200_eltwise_forward(const float *a, const float *b, float *y,
                                            int sz, int n_add, int act, void *stream);
/* dy *= act'(y); da (+)= dy; db[:n_add] (+)= dy[:n_add]  (da / db may be NULL).
 * accumulate_flags bit 0 / bit 1: da / db already hold a partial sum of this step and are
 * added to (the reference's +=, bcnn_eltwise_layer.c:148-158); a clear bit means the buffer is
 * stale and is overwritten (db beyond n_add with zero) -- same result as += onto the zero-
 * filled gradient of bcnn_reset_gradients without the fill and the read. */
BCNN_B200_API int bcnn_b200_eltwise_backward(const float *y, float *dy, float *da,
                                             float *db, int sz, int n_add, int act,
                                             int accumulate_flags, void *stream);

/* ---- concat / upsample (the second YOLO head's glue) ---------------------- */
/* Channel concatenation of one source into the output: image j (src_sz = C_src*H*W floats)
 * goes to dst + j * dst_sz + dst_offset.  replaces the bcnn_cuda_copy_f32 loop of
 * bcnn_forward_concat_layer_gpu, src/layers/bcnn_concat_layer.c:146-160 (CPU :107-121). */
BCNN_B200_API int bcnn_b200_concat_forward(const float *src, float *dst, int n, int src_sz,
                                           int dst_sz, int dst_offset, void *stream);
/* src_grad (+)= the matching slice of dst_grad (bcnn_axpy loop, :123-142); accumulate == 0
 * stores instead (first backward writer of the step). */
BCNN_B200_API int bcnn_b200_concat_backward(const float *dst_grad, float *src_grad, int n,
                                            int src_sz, int dst_sz, int dst_offset,
                                            int accumulate, void *stream);
/* YOLOv3 head activation, the inference part of bcnn_forward_yolo_layer_cpu
 * (src/layers/bcnn_yolo.c:226-250): x, y are [n, boxes_per_cell * (coords + classes + 1), hw];
 * per anchor group the logistic function is applied to entries 0, 1 (centre offsets) and
 * coords .. coords + classes (objectness, class scores); entries 2 .. coords-1 are copied. */
BCNN_B200_API int bcnn_b200_yolo_activate(const float *x, float *y, int n, int boxes_per_cell,
                                          int classes, int coords, int hw, void *stream);
/* YOLOv3 detection loss, the TRAIN part of bcnn_forward_yolo_layer_cpu (src/layers/bcnn_yolo.c:
 * 251-416), on the device: `out` is the activated head, `label` the [n, max_boxes * (coords + 1)]
 * truth list (x, y, w, h, class per box, ended by x == 0), `anchors` 2 * total_anchors sizes,
 * `mask` the boxes_per_cell anchor indices of this head; `delta` receives d(loss)/d(out) (every
 * entry written), cost_scratch[0] the loss (sum of squared delta entries); cost_scratch holds
 * bcnn_b200_yolo_cost_scratch_floats() floats. The reference does this on the host after a
 * device -> host copy of the head (:418-431). */
BCNN_B200_API int bcnn_b200_yolo_cost_scratch_floats(void);
BCNN_B200_API int bcnn_b200_yolo_loss_forward(const float *out, const float *label,
                                              const float *anchors, const int *mask, float *delta,
                                              float *cost_scratch, int n, int boxes_per_cell,
                                              int classes, int coords, int lw, int lh, int netw,
                                              int neth, int total_anchors, int max_boxes,
                                              void *stream);
/* y[n, c, j, i] = x[n, c, j / size, i / size]   (src/layers/bcnn_upsample_layer.c:86-109;
 * replaces bcnn_cuda_upsample_kernel, bcnn_upsample_layer.cu). */
BCNN_B200_API int bcnn_b200_upsample_forward(const float *x, float *y, int n, int c, int h,
                                             int w, int size, void *stream);
/* dx[n, c, y, x] (+)= sum of the size x size block of dy, summed in the raster order of
 * bcnn_backward_upsample_layer_cpu (:119-142). */
BCNN_B200_API int bcnn_b200_upsample_backward(const float *dy, float *dx, int n, int c, int h,
                                              int w, int size, int accumulate, void *stream);


/* ---- BF16 NHWC resident tensors (BCNN_B200_MATH_TC_BF16; csrc/nhwc_bf16.cu) -------------- */
/* In this mode a convolution writes its result once as BF16 NHWC from the tensor-core epilogue and
 * the layers between two convolutions work on that format: half the bytes of the FP32 NCHW
 * kernels above and no transposition pass in front of the next convolution's TMA loads. Element
 * (n, h, w, c) lives at ((n * H + h) * W + w) * C + c, C % 8 == 0, buffers 16-byte aligned;
 * `positions` = N * H * W. Same formulas as the FP32 entry points they shadow (cited there),
 * evaluated in FP32 on BF16-rounded storage: the 2e-2 tolerance class. No reference counterpart
 * (the reference has no reduced-precision path). */
/* layout / precision converters: what materialises the FP32 NCHW tensors bcnn_get_tensor_by_index
 * hands out (reference inc/bcnn/bcnn.h:242-255) and brings FP32 producers' outputs in. c % 2 == 0. */
BCNN_B200_API int bcnn_b200_f32nchw_to_bf16nhwc(const float *in, void *out, int n, int c, int hw,
                                                void *stream);
BCNN_B200_API int bcnn_b200_bf16nhwc_to_f32nchw(const void *in, float *out, int n, int c, int hw,
                                                void *stream);
/* floats of scratch the two reductions below need for `c` channels */
BCNN_B200_API size_t bcnn_b200_nhwc_scratch_floats(int c);
/* TRAIN statistics of a BF16 NHWC tensor: what bcnn_b200_bn_stats computes (biased variance,
 * running = 0.9 running + 0.1 batch). nhwc_scratch: bcnn_b200_nhwc_scratch_floats(c) floats,
 * scratch: bcnn_b200_bn_scratch_floats(c) floats. */
BCNN_B200_API int bcnn_b200_bn_stats_nhwc(const void *x, size_t positions, int c, float *saved_mean,
                                          float *saved_var, float *run_mean, float *run_var,
                                          float *nhwc_scratch, float *scratch, void *stream);
/* y = act(gamma (x - mean) / sqrt(var + 1e-6) + beta); mean == NULL: y = act(gamma x + beta)
 * (PREDICT). act: NONE, RELU or LRELU. x may alias y. Shadows bcnn_b200_bn_apply / _scale_bias. */
BCNN_B200_API int bcnn_b200_bn_apply_nhwc(const void *x, void *y, const float *mean, const float *var,
                                          const float *gamma, const float *beta, size_t positions,
                                          int c, int act, void *stream);
/* Backward of (activation o batchnorm): shadows bcnn_b200_bn_backward with the mask rebuilt from x
 * (beta required). dx may alias dy. g_gamma / g_beta accumulate (FP32); d_mean / d_var written. */
BCNN_B200_API int bcnn_b200_bn_backward_nhwc(const void *x, void *dy, void *dx, const float *mean,
                                             const float *var, const float *gamma, const float *beta,
                                             float *g_gamma, float *g_beta, float *d_mean,
                                             float *d_var, size_t positions, int c, int act,
                                             float *scratch, void *stream);
/* dy *= act'(y) in place; g_bias[c] += sum dy. Shadows bcnn_b200_actbwd_grad_bias. */
BCNN_B200_API int bcnn_b200_actbwd_grad_bias_nhwc(float *g_bias, void *dy, const void *y, int act,
                                                  size_t positions, int c, float *scratch,
                                                  void *stream);
/* y = act(bn_a(xa) + bn_b(xb)): batch-norm apply of one or both operands fused with the residual
 * add (bcnn_forward_batchnorm_cpu's normalise + scale + shift, src/layers/bcnn_batchnorm_layer.c
 * :170-194, then bcnn_eltwise_layer.c:111-127) so that the block's last convolution and its
 * projection shortcut never store their normalised outputs. Per operand: gamma == NULL: plain
 * tensor; mean != NULL: gamma (x - mean) / sqrt(var + 1e-6) + beta; mean == NULL: gamma x + beta
 * (PREDICT, folded statistics). The branches are added in FP32 and the sum rounded once (the unfused
 * path rounds each branch to BF16 first). act: NONE, RELU, LRELU. */
BCNN_B200_API int bcnn_b200_bn_add_act_nhwc(const void *xa, const float *mean_a, const float *var_a,
                                            const float *gamma_a, const float *beta_a,
                                            const void *xb, const float *mean_b, const float *var_b,
                                            const float *gamma_b, const float *beta_b, void *y,
                                            size_t positions, int c, int act, void *stream);
/* residual add; shadows bcnn_b200_eltwise_forward / _backward (sz, n_add in elements, % 8 == 0) */
BCNN_B200_API int bcnn_b200_eltwise_forward_bf16(const void *a, const void *b, void *y, size_t sz,
                                                 size_t n_add, int act, void *stream);
BCNN_B200_API int bcnn_b200_eltwise_backward_bf16(const void *y, void *dy, void *da, void *db,
                                                  size_t sz, size_t n_add, int act,
                                                  int accumulate_flags, void *stream);
/* Weight packing of many layers in one launch. The resident fprop / dgrad entry points above turn the
 * FP32 weights of the call into the BF16 K-major image(s) their TMA loads want, one small launch per
 * call (the im2col + GEMM operand set-up of src/layers/bcnn_conv_layer.c:628-656 has no counterpart
 * that outlives a call either). A caller that knows when weights change can keep the images itself:
 *   bcnn_b200_conv_nhwc_pack_jobs  appends the jobs of one layer and pass (dgrad: 0 fprop, 1 dgrad; a
 *       strided dgrad has one image per class of input positions) to a host table of
 *       bcnn_b200_conv_pack_job_bytes()-sized entries, images laid out from `dst` on; returns the number
 *       of jobs (< 0: no such route) and the bytes of the images;
 *   bcnn_b200_conv_pack_table_finish  numbers the CTAs of the finished table and returns the grid;
 *   bcnn_b200_conv_pack_run  packs every image of the (device copy of the) table in one launch;
 *   bcnn_b200_conv_prepacked_set / _enable  registers `dst` for (w, pass): while enabled, the entry
 *       points use the registered images instead of packing (image == NULL unregisters). The images
 *       must be re-packed whenever the weights change; the net runtime does so at the start of every
 *       TRAIN-mode forward pass. */
BCNN_B200_API size_t bcnn_b200_conv_pack_job_bytes(void);
BCNN_B200_API int bcnn_b200_conv_nhwc_pack_jobs(const bcnn_b200_conv_desc *desc, int dgrad, const float *w,
                                                void *dst, void *jobs, int max_jobs, size_t *bytes);
BCNN_B200_API unsigned int bcnn_b200_conv_pack_table_finish(void *jobs, int count);
BCNN_B200_API int bcnn_b200_conv_pack_run(const void *jobs_dev, int count, unsigned int grid, void *stream);
BCNN_B200_API void bcnn_b200_conv_prepacked_set(const float *w, int dgrad, const void *image);
BCNN_B200_API void bcnn_b200_conv_prepacked_enable(int on);
/* Residual-add backward fused with the reduction pass of the batch-norm backward of the branches it
 * feeds (conv + BN without activation whose only reader is the add, reference pair
 * src/layers/bcnn_eltwise_layer.c:137-161 + src/layers/bcnn_batchnorm_layer.c:263-299): besides what
 * bcnn_b200_eltwise_backward_bf16 does (whole-tensor add only), per channel S1 = sum dy',
 * S2 = sum dy' (x - mean) of branch a and / or b (x = the branch's raw convolution result, NULL = not a
 * fused branch) go to partial_a / partial_b (bcnn_b200_nhwc_scratch_floats(c) floats each) as *rows
 * partial rows; bcnn_b200_bn_backward_nhwc_partials then finishes that branch's batch-norm backward
 * (finalize + apply) without reading x and dy' a first time. */
BCNN_B200_API int bcnn_b200_eltwise_backward_bn_reduce_bf16(const void *y, void *dy, void *da, void *db,
                                                            size_t positions, int c, int act,
                                                            int accumulate_flags, const void *xa,
                                                            const float *mean_a, float *partial_a,
                                                            const void *xb, const float *mean_b,
                                                            float *partial_b, int *rows, void *stream);
BCNN_B200_API int bcnn_b200_bn_backward_nhwc_partials(const void *x, void *dy, void *dx, const float *mean,
                                                      const float *var, const float *gamma,
                                                      const float *beta, float *g_gamma, float *g_beta,
                                                      float *d_mean, float *d_var, size_t positions, int c,
                                                      const float *partial, int rows, void *stream);
/* max pooling with the reference's rule (first max wins, bottom / right padding only); the index
 * buffer is laid out like y (NHWC), its values are the reference's flat NCHW indices (-1: empty
 * window). Shadows bcnn_b200_maxpool_forward / _backward; accumulate == 0 overwrites dx. */
BCNN_B200_API int bcnn_b200_maxpool_forward_nhwc(const void *x, void *y, int *indexes, int n, int c,
                                                 int h, int w, int ksize, int stride, int ho, int wo,
                                                 void *stream);
BCNN_B200_API int bcnn_b200_maxpool_backward_nhwc(void *dx, const void *dy, const int *indexes, int n,
                                                  int c, int h, int w, int ksize, int stride, int ho,
                                                  int wo, int accumulate, void *stream);
/* global average pooling: y is FP32 [n, c]; dx (+)= dy / hw. Shadows bcnn_b200_avgpool_*. */
BCNN_B200_API int bcnn_b200_avgpool_forward_nhwc(const void *x, float *y, int n, int c, int hw,
                                                 void *stream);
BCNN_B200_API int bcnn_b200_avgpool_backward_nhwc(void *dx, const float *dy, int n, int c, int hw,
                                                  int accumulate, void *stream);

/* Convolution on resident tensors: x, y, dy, dx are BF16 NHWC; weights, bias and the weight
 * gradient stay FP32 in the reference's [Cout, Cin, k, k] layout. Same three passes and the same
 * semantics as bcnn_b200_conv_forward / _backward_data / _backward_weights above (reference
 * src/layers/bcnn_conv_layer.c:367-587), on the tcgen05 + TMA kernels only: the TMA reads the
 * activation tensor itself (no transposed shadow), the epilogue writes BF16 NHWC through a bulk
 * tensor store. 1x1 problems see the batch as one row of N*H*W positions.
 * _supported: bit mask of the passes these kernels cover for `d` (1 fprop, 2 dgrad, 4 wgrad);
 * needs groups == 1, cout % 8 == 0 and cin % 8 == 0. A thin first layer (cin < 16) is covered too:
 * its `x` is the FP32 NCHW input, gathered into a BF16 im2col buffer that `sh->x` keeps for wgrad
 * (_x_keep_bytes = its size). */
BCNN_B200_API int bcnn_b200_conv_nhwc_supported(const bcnn_b200_conv_desc *d);
BCNN_B200_API size_t bcnn_b200_conv_nhwc_workspace_bytes(const bcnn_b200_conv_desc *d);
BCNN_B200_API size_t bcnn_b200_conv_nhwc_x_keep_bytes(const bcnn_b200_conv_desc *d);
BCNN_B200_API int bcnn_b200_conv_forward_nhwc(const bcnn_b200_conv_desc *d, const void *x,
                                              const float *w, const float *bias, int act, void *y,
                                              void *workspace, size_t workspace_bytes,
                                              bcnn_b200_conv_shadows *sh, void *stream);
/* shadows bcnn_b200_conv_forward_bn_stats: y is the raw convolution result, statistics come from
 * the FP32 accumulators in the epilogue (or, when that is not possible, from bcnn_b200_bn_stats_nhwc
 * over y). */
BCNN_B200_API int bcnn_b200_conv_forward_bn_stats_nhwc(const bcnn_b200_conv_desc *d, const void *x,
                                                       const float *w, void *y, void *workspace,
                                                       size_t workspace_bytes,
                                                       bcnn_b200_conv_shadows *sh, float *saved_mean,
                                                       float *saved_var, float *run_mean,
                                                       float *run_var, float *nhwc_scratch,
                                                       float *scratch, void *stream);
BCNN_B200_API int bcnn_b200_conv_backward_data_nhwc(const bcnn_b200_conv_desc *d, const float *w,
                                                    const void *dy, void *dx, int accumulate,
                                                    void *workspace, size_t workspace_bytes,
                                                    void *stream);
BCNN_B200_API int bcnn_b200_conv_backward_weights_nhwc(const bcnn_b200_conv_desc *d, const void *x,
                                                       const void *dy, float *gw, void *workspace,
                                                       size_t workspace_bytes,
                                                       bcnn_b200_conv_shadows *sh, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_B200_H */
