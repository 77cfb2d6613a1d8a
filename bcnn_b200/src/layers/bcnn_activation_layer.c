/*
 * bcnn_activation_layer.c -- standalone in-place activation node.
 * Layout of jnbraun/bcnn src/layers/bcnn_activation_layer.c:35-88: the node's dst index IS
 * its src index; PReLU adds a "<src>_w_prelu" slope tensor at src[1], stepped with SGD
 * (:262-293, batch divisor = weights->n = 1). The reference's CPU node dereferences a NULL
 * slope tensor for every non-PReLU activation; this one does not.
 */
#include "bcnn_activation_layer.h"

#include <bcnn_b200_net.h>

#include "bcnn_learner.h"
#include "bcnn_tensor.h"

bcnn_status bcnn_add_activation_layer(bcnn_net *net, bcnn_activation type, const char *src_id) {
    bcnn_node node = {0};
    BCNN_CHECK_AND_LOG(net->log_ctx, net->num_nodes >= 1, BCNN_INVALID_PARAMETER,
                       "Activation layer can't be the first layer of the network\n");
    int src = bcnn_get_tensor_index_by_name(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Activation layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    BCNN_CHECK_STATUS(bcnn_node_add_output(net, &node, src));
    node.type = BCNN_LAYER_ACTIVATION;
    node.param_size = sizeof(bcnn_activation_param);
    bcnn_activation_param *param = (bcnn_activation_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    param->activation = type;
    node.forward = bcnn_forward_activation_layer;
    node.backward = bcnn_backward_activation_layer;
    node.update = bcnn_update_activation_layer;
    if (type == BCNN_ACT_PRELU)
        BCNN_CHECK_STATUS(bcnn_net_add_param_tensor(net, &node, 1, 1, 1, net->tensors[src].c, 1,
                                                    src_id, "_w_prelu", NULL));
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx, "[%s] %-8s (%4d x%4d x%4d)\n", bcnn_act2str(type), src_id,
              net->tensors[src].w, net->tensors[src].h, net->tensors[src].c);
    return BCNN_SUCCESS;
}

void bcnn_forward_activation_gpu(float *x, int sz, bcnn_activation a) {
    bcnn_cuda_check(bcnn_b200_activation_forward(x, sz, a, NULL, 1, 1, bcnn_b200_current_stream()));
}

void bcnn_backward_activation_gpu(float *x, float *dx, int sz, bcnn_activation a) {
    bcnn_cuda_check(bcnn_b200_activation_backward(x, dx, sz, a, NULL, NULL, 1, 1,
                                                  bcnn_b200_current_stream()));
}

void bcnn_forward_activation_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_activation_param *param = (bcnn_activation_param *)node->param;
    bcnn_tensor *dst = &net->tensors[node->dst[0]];
    const float *slope =
        param->activation == BCNN_ACT_PRELU ? net->tensors[node->src[1]].data_gpu : NULL;
    bcnn_cuda_check(bcnn_b200_activation_forward(dst->data_gpu, bcnn_tensor_size(dst),
                                                 param->activation, slope, dst->w * dst->h, dst->c,
                                                 bcnn_stream(net)));
}

void bcnn_backward_activation_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_activation_param *param = (bcnn_activation_param *)node->param;
    bcnn_tensor *dst = &net->tensors[node->dst[0]];
    if (!dst->grad_data_gpu) return;
    const float *slope = NULL;
    float *g_slope = NULL;
    if (param->activation == BCNN_ACT_PRELU) {
        slope = net->tensors[node->src[1]].data_gpu;
        g_slope = net->tensors[node->src[1]].grad_data_gpu;
    }
    bcnn_cuda_check(bcnn_b200_activation_backward(dst->data_gpu, dst->grad_data_gpu,
                                                  bcnn_tensor_size(dst), param->activation, slope,
                                                  g_slope, dst->w * dst->h, dst->c,
                                                  bcnn_stream(net)));
}

void bcnn_forward_activation_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_forward_activation_layer_gpu(net, node);
}

void bcnn_backward_activation_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_backward_activation_layer_gpu(net, node);
}

void bcnn_update_activation_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_activation_param *param = (bcnn_activation_param *)node->param;
    if (param->activation != BCNN_ACT_PRELU) return;
    bcnn_tensor *w = &net->tensors[node->src[1]];
    bcnn_sgd_update_gpu(net, w->data_gpu, NULL, w->grad_data_gpu, NULL, bcnn_tensor_size(w), 0,
                        w->n, net->learner->learning_rate, net->learner->momentum,
                        net->learner->decay);
}
