/*
 * bcnn_batchnorm_layer.h -- standalone batchnorm node and the shared forward / backward
 * used by the fused conv path. Signatures of the *_gpu entry points follow jnbraun/bcnn
 * src/layers/bcnn_batchnorm_layer.h:72-95 (no cuDNN descriptors), with the net prepended
 * for the stream and a trailing activation + per-layer scratch (fusion, see .c).
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

/* x_gpu: pre-normalisation input (device), y_tensor: output. `act` is fused into the
 * same pass (BCNN_ACT_NONE for the standalone node). */
void bcnn_forward_batchnorm_gpu(bcnn_net *net, const float *x_gpu, bcnn_tensor *dst_tensor,
                                bcnn_tensor *bn_mean, bcnn_tensor *bn_var, bcnn_tensor *bn_scales,
                                bcnn_tensor *biases, bcnn_tensor *saved_mean,
                                bcnn_tensor *saved_var, float *scratch_gpu, bcnn_mode mode,
                                bcnn_activation act);
/* In place on dst_tensor->grad_data_gpu; y_gpu (post-activation output) is needed only
 * when act != NONE. */
void bcnn_backward_batchnorm_gpu(bcnn_net *net, const float *x_gpu, const float *y_gpu,
                                 bcnn_tensor *dst_tensor, bcnn_tensor *bn_mean,
                                 bcnn_tensor *bn_var, bcnn_tensor *bn_scales, bcnn_tensor *biases,
                                 bcnn_tensor *saved_mean, bcnn_tensor *saved_var,
                                 float *scratch_gpu, bcnn_mode mode, bcnn_activation act);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_BATCHNORM_LAYER_H */
