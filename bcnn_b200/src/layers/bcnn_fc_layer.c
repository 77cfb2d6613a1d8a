/*
 * bcnn_fc_layer.c -- fully-connected node, run on the implicit-GEMM convolution kernels
 * as a 1x1 convolution over a 1x1 image with Cin = C*H*W of the source.
 * Layout of jnbraun/bcnn src/layers/bcnn_fc_layer.c:39-142: src[1] = W [out, C, H, W]
 * ("<src>_w"), src[2] = bias [1,1,1,out] ("<src>_b"), dst [N, out, 1, 1]. Semantics of the
 * CPU path (:144-226): y = act(W x + b); gb += sum_n dy; gW += dy^T x; dx += dy W (the data
 * gradient accumulates here, unlike the convolution's).
 */
#include "bcnn_fc_layer.h"

#include "bcnn_learner.h"
#include "bcnn_tensor.h"

bcnn_status bcnn_add_fullc_layer(bcnn_net *net, int output_size, bcnn_filler_type init,
                                 bcnn_activation activation, int quantize, const char *src_id,
                                 const char *dst_id) {
    (void)quantize;
    bcnn_node node = {0};
    int src = bcnn_net_find_src(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Full-connected layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const int n = net->tensors[src].n, c = net->tensors[src].c, h = net->tensors[src].h,
              w = net->tensors[src].w;
    const int input_size = c * h * w;
    bcnn_tensor_filler wfill = {.range = input_size, .type = init};
    BCNN_CHECK_STATUS(
        bcnn_net_add_param_tensor(net, &node, output_size, c, h, w, 1, src_id, "_w", &wfill));
    BCNN_CHECK_STATUS(
        bcnn_net_add_param_tensor(net, &node, 1, 1, 1, output_size, 1, src_id, "_b", NULL));
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, output_size, 1, 1, dst_id));
    node.type = BCNN_LAYER_FULL_CONNECTED;
    node.param_size = sizeof(bcnn_fullc_param);
    bcnn_fullc_param *param = (bcnn_fullc_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    param->activation = activation;
    bcnn_b200_conv_desc desc = {n, input_size, 1, 1, output_size, 1, 1, 1, 1, 0, 1};
    param->desc = desc;
    bcnn_net_require_workspace(net, bcnn_b200_conv_workspace_bytes(&desc, BCNN_B200_MATH_FP32));
    param->reduce_scratch_gpu =
        (float *)bcnn_b200_malloc(bcnn_b200_bn_scratch_floats(output_size) * sizeof(float));
    BCNN_CHECK(param->reduce_scratch_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);
    node.forward = bcnn_forward_fullc_layer;
    node.backward = bcnn_backward_fullc_layer;
    node.update = bcnn_update_fullc_layer;
    node.release_param = bcnn_release_param_fullc_layer;
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx, "[Dense] %-8s (%4d x%4d x%4d) -> %-8s (1 x 1 x%4d)\n", src_id, w, h,
              c, dst_id, output_size);
    return BCNN_SUCCESS;
}

void bcnn_forward_fullc_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_fullc_param *param = (bcnn_fullc_param *)node->param;
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    bcnn_tensor *t = net->tensors;
    bcnn_cuda_check(bcnn_b200_conv_forward(
        &param->desc, t[node->src[0]].data_gpu, t[node->src[1]].data_gpu, t[node->src[2]].data_gpu,
        param->activation, t[node->dst[0]].data_gpu, ctx->workspace_gpu, ctx->workspace_bytes,
        BCNN_B200_MATH_FP32, ctx->stream));
}

void bcnn_backward_fullc_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_fullc_param *param = (bcnn_fullc_param *)node->param;
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *src = &t[node->src[0]], *dst = &t[node->dst[0]];
    bcnn_tensor *weights = &t[node->src[1]], *biases = &t[node->src[2]];
    bcnn_cuda_check(bcnn_b200_actbwd_grad_bias(biases->grad_data_gpu, dst->grad_data_gpu,
                                               dst->data_gpu, param->activation, dst->n, dst->c, 1,
                                               param->reduce_scratch_gpu, ctx->stream));
    bcnn_cuda_check(bcnn_b200_conv_backward_weights(
        &param->desc, src->data_gpu, dst->grad_data_gpu, weights->grad_data_gpu,
        ctx->workspace_gpu, ctx->workspace_bytes, BCNN_B200_MATH_FP32, ctx->stream));
    if (src->grad_data_gpu)
        bcnn_cuda_check(bcnn_b200_conv_backward_data(
            &param->desc, weights->data_gpu, dst->grad_data_gpu, src->grad_data_gpu,
            /*accumulate (the reference's +=)*/ bcnn_net_grad_accumulate(net, node->src[0]),
            ctx->workspace_gpu, ctx->workspace_bytes, BCNN_B200_MATH_FP32,
            ctx->stream));
}

void bcnn_update_fullc_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_fullc_param *param = (bcnn_fullc_param *)node->param;
    bcnn_optimizer_step_gpu(net, &net->tensors[node->src[1]], &net->tensors[node->src[2]],
                            &param->adam_m_gpu, &param->adam_v_gpu);
}

void bcnn_release_param_fullc_layer(bcnn_node *node) {
    bcnn_fullc_param *param = (bcnn_fullc_param *)node->param;
    bcnn_b200_free(param->reduce_scratch_gpu);
    bcnn_b200_free(param->adam_m_gpu);
    bcnn_b200_free(param->adam_v_gpu);
}
