/* bcnn_fc_layer.h -- fully-connected node (glue); entry points of
 * jnbraun/bcnn src/layers/bcnn_fc_layer.h. */
#ifndef BCNN_FC_LAYER_H
#define BCNN_FC_LAYER_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcnn_fullc_param {
    bcnn_activation activation;
    float *adam_m;
    float *adam_v;
    float *adam_m_gpu;
    float *adam_v_gpu;
    /* ---- B200 additions ---- */
    bcnn_b200_conv_desc desc;  /* the layer runs as a 1x1 convolution over a 1x1 image */
    float *reduce_scratch_gpu;
} bcnn_fullc_param;

void bcnn_forward_fullc_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_fullc_layer(bcnn_net *net, bcnn_node *node);
void bcnn_update_fullc_layer(bcnn_net *net, bcnn_node *node);
void bcnn_release_param_fullc_layer(bcnn_node *node);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_FC_LAYER_H */
