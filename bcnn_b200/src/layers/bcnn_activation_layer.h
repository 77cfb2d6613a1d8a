/*
 * bcnn_activation_layer.h -- standalone (in-place) activation node; entry points of
 * jnbraun/bcnn src/layers/bcnn_activation_layer.h:37-52, with the net prepended to the
 * raw-pointer helpers (for the stream) and PReLU slopes accepted on the device path.
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
void bcnn_forward_activation_gpu(bcnn_net *net, float *x, int sz, bcnn_activation a);
void bcnn_backward_activation_gpu(bcnn_net *net, float *x, float *dx, int sz, bcnn_activation a);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_ACTIVATION_LAYER_H */
