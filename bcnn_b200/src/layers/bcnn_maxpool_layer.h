/*
 * bcnn_maxpool_layer.h -- max-pooling node. Param layout of jnbraun/bcnn
 * src/layers/bcnn_maxpool_layer.h:34-47 (CUDA flavour, no cuDNN): the argmax lives in
 * param->indexes_gpu (int32 flat NCHW offsets), not in a tensor.
 */
#ifndef BCNN_MAXPOOL_LAYER_H
#define BCNN_MAXPOOL_LAYER_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcnn_maxpool_param {
    int size;
    int stride;
    bcnn_padding padding;
    int *indexes;     /* host mirror, filled on demand by bcnn_b200_maxpool_indexes */
    int *indexes_gpu;
    int indexes_nhwc; /* B200: the last forward laid indexes_gpu out like a BF16 NHWC result */
} bcnn_maxpool_param;

void bcnn_forward_maxpool_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_maxpool_layer(bcnn_net *net, bcnn_node *node);
void bcnn_release_param_maxpool_layer(bcnn_node *node);
void bcnn_forward_maxpool_layer_gpu(bcnn_net *net, bcnn_node *node);
void bcnn_backward_maxpool_layer_gpu(bcnn_net *net, bcnn_node *node);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_MAXPOOL_LAYER_H */
