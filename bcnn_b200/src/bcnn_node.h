/*
 * bcnn_node.h -- the node vtable: the plugin point of the hot path.
 * Field order is the reference's (jnbraun/bcnn src/bcnn_node.h:36-48); a layer is
 * plugged in by its bcnn_add_*_layer constructor filling the four function pointers.
 */
#ifndef BCNN_NODE_H
#define BCNN_NODE_H

#include <bcnn/bcnn.h>

#ifdef __cplusplus
extern "C" {
#endif

struct bcnn_node {
    int num_src;
    int num_dst;
    bcnn_layer_type type;
    size_t param_size;
    int *src; /* indexes into net->tensors[] (never cache bcnn_tensor*: the array reallocs) */
    int *dst;
    void *param;
    void (*forward)(struct bcnn_net *net, struct bcnn_node *node);
    void (*backward)(struct bcnn_net *net, struct bcnn_node *node);
    void (*update)(struct bcnn_net *net, struct bcnn_node *node);
    void (*release_param)(struct bcnn_node *node);
};
typedef struct bcnn_node bcnn_node;

bcnn_status bcnn_node_add_input(bcnn_net *net, bcnn_node *node, int index);
bcnn_status bcnn_node_add_output(bcnn_net *net, bcnn_node *node, int index);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_NODE_H */
