/* bcnn_depthwise_conv_layer.h -- depthwise convolution node; leading param fields of
 * jnbraun/bcnn src/layers/bcnn_depthwise_conv_layer.h:33-45. */
#ifndef BCNN_DEPTHWISE_CONV_LAYER_H
#define BCNN_DEPTHWISE_CONV_LAYER_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcnn_depthwise_conv_param {
    int num;
    int size;
    int stride;
    int pad;
    bcnn_activation activation;
    float *adam_m;
    float *adam_v;
    float *adam_m_gpu;
    float *adam_v_gpu;
    /* ---- B200 additions ---- */
    float *reduce_scratch_gpu; /* bias-gradient reduction */
    float *wgrad_scratch_gpu;  /* weight-gradient reduction */
    size_t wgrad_scratch_floats;
} bcnn_depthwise_conv_param;

void bcnn_forward_depthwise_conv_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_depthwise_conv_layer(bcnn_net *net, bcnn_node *node);
void bcnn_update_depthwise_conv_layer(bcnn_net *net, bcnn_node *node);
void bcnn_release_param_depthwise_conv_layer(bcnn_node *node);
void bcnn_forward_depthwise_conv_layer_gpu(bcnn_net *net, bcnn_node *node);
void bcnn_backward_depthwise_conv_layer_gpu(bcnn_net *net, bcnn_node *node);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_DEPTHWISE_CONV_LAYER_H */
