/*
 * bcnn_conv_layer.h -- convolution node (optionally with fused batchnorm + activation).
 * Leading fields of bcnn_conv_param follow jnbraun/bcnn src/layers/bcnn_conv_layer.h:34-74
 * (CUDA flavour, no cuDNN) because the runtime and tools peek at num/size/stride/pad/
 * num_groups/batch_norm; the CPU-only scratch pointers stay NULL; B200 state is appended.
 */
#ifndef BCNN_CONV_LAYER_H
#define BCNN_CONV_LAYER_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcnn_conv_param {
    int num;
    int size;
    int stride;
    int pad;
    int num_groups;
    int batch_norm;
    int post_func;
    size_t workspace_size;
    bcnn_activation activation;
    bcnn_tensor saved_mean;     /* data_gpu: batch mean;  grad_data_gpu: d(mean)  */
    bcnn_tensor saved_variance; /* data_gpu: batch var;   grad_data_gpu: d(var)   */
    float *conv_workspace;      /* CPU im2col buffer of the reference: unused (NULL) */
    float *workspace;
    float *weights_workspace;
    float *biases_workspace;
    float *scales_workspace;
    float *slopes_workspace;
    float *src_workspace;
    float *dst_workspace;
    float *x_norm;
    float *adam_m;
    float *adam_v;
    float *conv_workspace_gpu; /* -> net-level shared workspace (split-K partials, packs) */
    float *bn_workspace_gpu;   /* pre-normalisation conv output, kept for backward */
    float *x_norm_gpu;         /* not materialised: x_hat is recomputed from saved stats */
    float *adam_m_gpu;
    float *adam_v_gpu;
    /* ---- B200 additions ---- */
    bcnn_b200_conv_desc desc;
    float *reduce_scratch_gpu; /* per-layer scratch of the per-channel reductions */
    /* NHWC shadows kept across the passes of a training step: x belongs to the layer
     * (written by forward, read by wgrad), dy points into the net-level buffer (written by
     * wgrad, read by dgrad of the same backward call) */
    bcnn_b200_conv_shadows shadows;
    /* resident mode (BCNN_B200_MATH_TC_BF16): raw convolution result of a conv+BN node as BF16
     * NHWC (the twin of bn_workspace_gpu), and whether this node runs on the resident kernels
     * (0 = not decided yet, 1 = yes, -1 = no: it takes the FP32-tensor path) */
    void *bn_raw16_gpu;
    int resident_state;
    /* batch-norm backward whose reduction pass the residual add behind this node has already done
     * (bcnn_b200_eltwise_backward_bn_reduce_bf16): the partial rows and how many (0 = none pending) */
    float *bn_partial_gpu;
    int bn_partial_rows;
} bcnn_conv_param;

/* For the residual add that feeds this node's backward (grad alias): the raw convolution result, the
 * saved mean and a buffer for the partial rows of the fused reduction; 0 when the node cannot take
 * them (no batch norm, not TRAIN, allocation failed). bcnn_conv_layer_bn_reduce_done records the rows. */
int bcnn_conv_layer_bn_reduce_operand(bcnn_net *net, bcnn_node *node, const void **raw, const float **mean,
                                      float **partial);
void bcnn_conv_layer_bn_reduce_done(bcnn_node *node, int rows);
void bcnn_forward_conv_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_conv_layer(bcnn_net *net, bcnn_node *node);
void bcnn_update_conv_layer(bcnn_net *net, bcnn_node *node);
void bcnn_release_param_conv_layer(bcnn_node *node);
void bcnn_forward_conv_layer_gpu(bcnn_net *net, bcnn_node *node);
int bcnn_conv_layer_is_resident(bcnn_net *net, bcnn_node *node);
/* resident mode: apply the deferred batch norm of this node (raw result -> dst's BF16 twin) */
void bcnn_conv_layer_materialize(bcnn_net *net, bcnn_node *node);
/* resident mode: tensor idx is the output of a resident conv + BN node without activation, i.e. one
 * whose backward can take its incoming gradient from a residual add's output gradient */
int bcnn_conv_layer_takes_grad_alias(bcnn_net *net, int idx);
/* resident mode: mean / variance the batch norm of this node normalises with in the net's mode */
void bcnn_conv_layer_bn_operand(bcnn_net *net, bcnn_node *node, const void **raw, const float **mean,
                                const float **var, const float **gamma, const float **beta);
void bcnn_backward_conv_layer_gpu(bcnn_net *net, bcnn_node *node);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_CONV_LAYER_H */
