/*
 * bcnn_activation_layer.h -- standalone (in-place) activation node; entry points of
 * jnbraun/bcnn src/layers/bcnn_activation_layer.h:37-52 with their signatures; PReLU slopes
 * are accepted on the device path of the node.
 */
#ifndef BCNN_ACTIVATION_LAYER_H
#define BCNN_ACTIVATION_LAYER_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcnn_activation_param {
    bcnn_activation activation;
} bcnn_activation_param;

void bcnn_forward_activation_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_activation_layer(bcnn_net *net, bcnn_node *node);
void bcnn_update_activation_layer(bcnn_net *net, bcnn_node *node);
void bcnn_forward_activation_layer_gpu(bcnn_net *net, bcnn_node *node);
void bcnn_backward_activation_layer_gpu(bcnn_net *net, bcnn_node *node);
/* reference signatures (src/layers/bcnn_activation_layer.h:48-51); kernels go to the
 * process-current stream (bcnn_b200_current_stream) */
void bcnn_forward_activation_gpu(float *x, int sz, bcnn_activation a);
void bcnn_backward_activation_gpu(float *x, float *dx, int sz, bcnn_activation a);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_ACTIVATION_LAYER_H */
