/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY. Compiled INTO oracle/_ref/libbcnn_ref.so next to
 * the unmodified reference sources (which stay under /root/reference; nothing is copied).
 *
 * The reference keeps bcnn_net opaque in its public header, so a driver needs its internal
 * headers to walk nodes and reach layer params. This shim does that
 * walking in C and exports it under the same names as the introspection helpers of
 * include/bcnn_b200_net.h, so tests drive the reference and the B200 library with one
 * Python wrapper.
 */
#include <string.h>

#include "bcnn_batchnorm_layer.h"
#include "bcnn_conv_layer.h"
#include "bcnn_maxpool_layer.h"
#include "bcnn_net.h"
#include "bcnn_tensor.h"

int bcnn_b200_num_nodes(bcnn_net *net) { return net->num_nodes; }
int bcnn_b200_num_tensors(bcnn_net *net) { return net->num_tensors; }
int bcnn_b200_node_type(bcnn_net *net, int node) {
    return (node >= 0 && node < net->num_nodes) ? (int)net->nodes[node].type : -1;
}
int bcnn_b200_node_src(bcnn_net *net, int node, int i) {
    if (node < 0 || node >= net->num_nodes || i < 0 || i >= net->nodes[node].num_src) return -1;
    return net->nodes[node].src[i];
}
int bcnn_b200_node_dst(bcnn_net *net, int node, int i) {
    if (node < 0 || node >= net->num_nodes || i < 0 || i >= net->nodes[node].num_dst) return -1;
    return net->nodes[node].dst[i];
}

int bcnn_b200_maxpool_indexes(bcnn_net *net, int node_index, int *out) {
    if (node_index < 0 || node_index >= net->num_nodes) return -1;
    bcnn_node *node = &net->nodes[node_index];
    if (node->type != BCNN_LAYER_MAXPOOL) return -1;
    bcnn_maxpool_param *param = (bcnn_maxpool_param *)node->param;
    int count = bcnn_tensor_size(&net->tensors[node->dst[0]]);
    memcpy(out, param->indexes, (size_t)count * sizeof(int));
    return count;
}

int bcnn_b200_bn_saved_stats(bcnn_net *net, int node_index, float *mean_out, float *var_out) {
    if (node_index < 0 || node_index >= net->num_nodes) return -1;
    bcnn_node *node = &net->nodes[node_index];
    bcnn_tensor *mean = NULL, *var = NULL;
    if (node->type == BCNN_LAYER_BATCHNORM) {
        bcnn_batchnorm_param *p = (bcnn_batchnorm_param *)node->param;
        mean = &p->saved_mean;
        var = &p->saved_variance;
    } else if (node->type == BCNN_LAYER_CONV2D && ((bcnn_conv_param *)node->param)->batch_norm) {
        bcnn_conv_param *p = (bcnn_conv_param *)node->param;
        mean = &p->saved_mean;
        var = &p->saved_variance;
    } else {
        return -1;
    }
    int c = bcnn_tensor_size(mean);
    memcpy(mean_out, mean->data, (size_t)c * sizeof(float));
    memcpy(var_out, var->data, (size_t)c * sizeof(float));
    return c;
}

/* bcnn_set_num_threads clamps to omp_get_max_threads(); expose the count actually used. */
int bcnn_ref_num_threads(bcnn_net *net) { return net->num_threads; }
