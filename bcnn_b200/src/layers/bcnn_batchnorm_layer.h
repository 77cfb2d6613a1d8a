/*
 * bcnn_batchnorm_layer.h -- standalone batchnorm node and the shared forward / backward
 * used by the fused conv path. bcnn_forward / backward_batchnorm_gpu keep the reference's
 * signatures (jnbraun/bcnn src/layers/bcnn_batchnorm_layer.h:72-95, no cuDNN descriptors); the
 * fused variants with an explicit net / activation / scratch are bcnn_b200_*_batchnorm.
 */
#ifndef BCNN_BATCHNORM_LAYER_H
#define BCNN_BATCHNORM_LAYER_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcnn_batchnorm_param {
    bcnn_tensor saved_mean;
    bcnn_tensor saved_variance;
    float *workspace; /* CPU copies of the reference: unused (NULL) */
    float *x_norm;
    float *workspace_gpu; /* not allocated: the source tensor itself is the saved input */
    float *x_norm_gpu;    /* not materialised */
    float *reduce_scratch_gpu;
} bcnn_batchnorm_param;

void bcnn_forward_batchnorm_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_batchnorm_layer(bcnn_net *net, bcnn_node *node);
void bcnn_release_param_batchnorm_layer(bcnn_node *node);
void bcnn_forward_batchnorm_layer_gpu(bcnn_net *net, bcnn_node *node);
void bcnn_backward_batchnorm_layer_gpu(bcnn_net *net, bcnn_node *node);

/* Entry points with the reference's exact signatures (src/layers/bcnn_batchnorm_layer.h:72-95,
 * CUDA flavour without cuDNN). src_tensor->data_gpu is normalised into dst_tensor->data_gpu (they
 * may alias). workspace_gpu, when given, receives a copy of the input as in the reference (its
 * backward reads the input from there; pass NULL when src stays intact until backward);
 * x_norm_gpu is accepted and left untouched (x_hat is recomputed from the saved statistics).
 * Kernels go to the process-current stream (bcnn_b200_current_stream: the stream of the net
 * whose forward / backward loop is running, else the legacy default stream); the reduction
 * scratch is a grow-only process buffer. backward: in place on dst_tensor->grad_data_gpu, then
 * copied to src_tensor->grad_data_gpu when the two differ (reference :327-330). */
void bcnn_forward_batchnorm_gpu(bcnn_tensor *src_tensor, bcnn_tensor *dst_tensor,
                                bcnn_tensor *bn_mean, bcnn_tensor *bn_var, bcnn_tensor *bn_scales,
                                bcnn_tensor *biases, bcnn_tensor *saved_mean,
                                bcnn_tensor *saved_var, float *x_norm_gpu, float *workspace_gpu,
                                bcnn_mode mode);
void bcnn_backward_batchnorm_gpu(bcnn_tensor *src_tensor, bcnn_tensor *dst_tensor,
                                 bcnn_tensor *bn_mean, bcnn_tensor *bn_var, bcnn_tensor *bn_scales,
                                 bcnn_tensor *biases, bcnn_tensor *saved_mean,
                                 bcnn_tensor *saved_variance, float *x_norm_gpu,
                                 float *workspace_gpu, bcnn_mode mode);

/* The fused variants the layer files of this library call: x_gpu = pre-normalisation input
 * (device), `act` fused into the same pass (BCNN_ACT_NONE for the standalone node), explicit
 * stream (the net's) and per-layer scratch. */
void bcnn_b200_forward_batchnorm(bcnn_net *net, const float *x_gpu, bcnn_tensor *dst_tensor,
                                 bcnn_tensor *bn_mean, bcnn_tensor *bn_var, bcnn_tensor *bn_scales,
                                 bcnn_tensor *biases, bcnn_tensor *saved_mean,
                                 bcnn_tensor *saved_var, float *scratch_gpu, bcnn_mode mode,
                                 bcnn_activation act);
/* In place on dst_tensor->grad_data_gpu; y_gpu (post-activation output) is needed only
 * when act != NONE. */
void bcnn_b200_backward_batchnorm(bcnn_net *net, const float *x_gpu, const float *y_gpu,
                                  bcnn_tensor *dst_tensor, bcnn_tensor *bn_mean,
                                  bcnn_tensor *bn_var, bcnn_tensor *bn_scales, bcnn_tensor *biases,
                                  bcnn_tensor *saved_mean, bcnn_tensor *saved_var,
                                  float *scratch_gpu, bcnn_mode mode, bcnn_activation act);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_BATCHNORM_LAYER_H */
