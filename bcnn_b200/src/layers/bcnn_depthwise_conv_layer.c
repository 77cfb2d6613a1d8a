/*
 * bcnn_depthwise_conv_layer.c -- depthwise k x k convolution node on the B200 kernels.
 * Layout of jnbraun/bcnn src/layers/bcnn_depthwise_conv_layer.c:42-163: src[1] = weights
 * [1,1,1,C*k*k] ("<src>_w"), src[2] = bias [1,1,1,C] ("<src>_b"); the batch_norm argument
 * is accepted and ignored exactly as the reference does. Backward computes
 * the weight gradient only when the source has a gradient buffer (:318).
 */
#include "bcnn_depthwise_conv_layer.h"

#include "bcnn_learner.h"
#include "bcnn_tensor.h"

bcnn_status bcnn_add_depthwise_conv_layer(bcnn_net *net, int size, int stride, int pad,
                                          int batch_norm, bcnn_filler_type init,
                                          bcnn_activation activation, const char *src_id,
                                          const char *dst_id) {
    (void)batch_norm;
    bcnn_node node = {0};
    int src = bcnn_net_find_src(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Dephtwise convolution layer: invalid input node name %s", src_id);
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const int n = net->tensors[src].n, c = net->tensors[src].c, h = net->tensors[src].h,
              w = net->tensors[src].w;
    node.type = BCNN_LAYER_DEPTHWISE_CONV2D;
    node.param_size = sizeof(bcnn_depthwise_conv_param);
    bcnn_depthwise_conv_param *param = (bcnn_depthwise_conv_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    param->activation = activation;
    param->pad = pad;
    param->num = c;
    param->size = size;
    param->stride = stride;
    node.forward = bcnn_forward_depthwise_conv_layer;
    node.backward = bcnn_backward_depthwise_conv_layer;
    node.update = bcnn_update_depthwise_conv_layer;
    node.release_param = bcnn_release_param_depthwise_conv_layer;
    bcnn_tensor_filler wfill = {.range = size * size * c, .type = init};
    BCNN_CHECK_STATUS(
        bcnn_net_add_param_tensor(net, &node, 1, 1, 1, c * size * size, 1, src_id, "_w", &wfill));
    BCNN_CHECK_STATUS(bcnn_net_add_param_tensor(net, &node, 1, 1, 1, c, 1, src_id, "_b", NULL));
    const int ho = (h + 2 * pad - size) / stride + 1, wo = (w + 2 * pad - size) / stride + 1;
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, c, ho, wo, dst_id));
    if (net->mode != BCNN_MODE_PREDICT) {
        param->reduce_scratch_gpu =
            (float *)bcnn_b200_malloc(bcnn_b200_bn_scratch_floats(c) * sizeof(float));
        param->wgrad_scratch_floats = bcnn_b200_depthwise_scratch_floats(n, c, size);
        param->wgrad_scratch_gpu =
            (float *)bcnn_b200_malloc(param->wgrad_scratch_floats * sizeof(float));
        BCNN_CHECK(param->reduce_scratch_gpu && param->wgrad_scratch_gpu, BCNN_CUDA_FAILED_ALLOC);
    }
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx,
              "[DeptwiseConv2d][%s] %-8s (%4d x%4d x%4d) -> %-8s (%4d x%4d x%4d) %2d x %2d / %2d,%2d\n",
              bcnn_act2str(activation), src_id, w, h, c, dst_id, wo, ho, c, size, size, stride,
              pad);
    return BCNN_SUCCESS;
}

void bcnn_forward_depthwise_conv_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_depthwise_conv_param *param = (bcnn_depthwise_conv_param *)node->param;
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *src = &t[node->src[0]], *dst = &t[node->dst[0]];
    bcnn_cuda_check(bcnn_b200_depthwise_forward(
        src->data_gpu, t[node->src[1]].data_gpu, t[node->src[2]].data_gpu, param->activation,
        dst->data_gpu, src->n, src->c, src->h, src->w, param->size, param->stride, param->pad,
        bcnn_stream(net)));
}

void bcnn_backward_depthwise_conv_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_depthwise_conv_param *param = (bcnn_depthwise_conv_param *)node->param;
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *src = &t[node->src[0]], *dst = &t[node->dst[0]];
    bcnn_tensor *weights = &t[node->src[1]], *biases = &t[node->src[2]];
    void *stream = bcnn_stream(net);
    bcnn_cuda_check(bcnn_b200_actbwd_grad_bias(biases->grad_data_gpu, dst->grad_data_gpu,
                                               dst->data_gpu, param->activation, dst->n, dst->c,
                                               dst->h * dst->w, param->reduce_scratch_gpu, stream));
    if (!src->grad_data_gpu) return;
    bcnn_net_grad_prepare_accumulate(net, node->src[0]); /* the kernel does dx += */
    bcnn_cuda_check(bcnn_b200_depthwise_backward(
        src->data_gpu, weights->data_gpu, dst->grad_data_gpu, weights->grad_data_gpu,
        src->grad_data_gpu, src->n, src->c, src->h, src->w, param->size, param->stride, param->pad,
        param->wgrad_scratch_gpu, param->wgrad_scratch_floats, stream));
}

void bcnn_forward_depthwise_conv_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_forward_depthwise_conv_layer_gpu(net, node);
}

void bcnn_backward_depthwise_conv_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_backward_depthwise_conv_layer_gpu(net, node);
}

void bcnn_update_depthwise_conv_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_depthwise_conv_param *param = (bcnn_depthwise_conv_param *)node->param;
    bcnn_optimizer_step_gpu(net, &net->tensors[node->src[1]], &net->tensors[node->src[2]],
                            &param->adam_m_gpu, &param->adam_v_gpu);
}

void bcnn_release_param_depthwise_conv_layer(bcnn_node *node) {
    bcnn_depthwise_conv_param *param = (bcnn_depthwise_conv_param *)node->param;
    bcnn_b200_free(param->reduce_scratch_gpu);
    bcnn_b200_free(param->wgrad_scratch_gpu);
    bcnn_b200_free(param->adam_m_gpu);
    bcnn_b200_free(param->adam_v_gpu);
}
