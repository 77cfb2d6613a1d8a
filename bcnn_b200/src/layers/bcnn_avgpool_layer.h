/* bcnn_avgpool_layer.h -- global average pooling node (entry points of jnbraun/bcnn
 * src/layers/bcnn_avgpool_layer.h:33-37). */
#ifndef BCNN_AVGPOOL_LAYER_H
#define BCNN_AVGPOOL_LAYER_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif
void bcnn_forward_avgpool_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_avgpool_layer(bcnn_net *net, bcnn_node *node);
void bcnn_forward_avgpool_layer_gpu(bcnn_net *net, bcnn_node *node);
void bcnn_backward_avgpool_layer_gpu(bcnn_net *net, bcnn_node *node);
#ifdef __cplusplus
}
#endif
#endif /* BCNN_AVGPOOL_LAYER_H */
