/* bcnn_glue_layers.h -- softmax, euclidean cost and residual-add nodes: the small layers
 * either side of the hot path that keep a whole training step on the device
 * (SURVEY.md 8f rank 1). */
#ifndef BCNN_GLUE_LAYERS_H
#define BCNN_GLUE_LAYERS_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcnn_cost_param {
    bcnn_loss loss;
    bcnn_loss_metric loss_metric;
    float scale;
} bcnn_cost_param;

typedef struct bcnn_eltwise_param {
    bcnn_activation activation;
    int stride[2];
    int min_dim[3];
} bcnn_eltwise_param;

typedef struct bcnn_upsample_param {
    int size;
} bcnn_upsample_param;

void bcnn_forward_concat_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_concat_layer(bcnn_net *net, bcnn_node *node);
void bcnn_forward_upsample_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_upsample_layer(bcnn_net *net, bcnn_node *node);
void bcnn_forward_softmax_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_softmax_layer(bcnn_net *net, bcnn_node *node);
void bcnn_forward_cost_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_cost_layer(bcnn_net *net, bcnn_node *node);
void bcnn_forward_eltwise_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_eltwise_layer(bcnn_net *net, bcnn_node *node);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_GLUE_LAYERS_H */
