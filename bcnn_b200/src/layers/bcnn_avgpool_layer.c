/*
 * bcnn_avgpool_layer.c -- global average pooling: [N,C,H,W] -> [N,C,1,1].
 * Semantics of jnbraun/bcnn src/layers/bcnn_avgpool_layer.c:82-125 (mean over H*W;
 * backward adds dy/(H*W) into src.grad). The reference's CUDA forward accumulates into an
 * un-zeroed dst (bcnn_avgpool_layer.cu:29-44); this one writes the mean.
 */
#include "bcnn_avgpool_layer.h"

#include "bcnn_tensor.h"

bcnn_status bcnn_add_avgpool_layer(bcnn_net *net, const char *src_id, const char *dst_id) {
    bcnn_node node = {0};
    int src = bcnn_net_find_src(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Avgpool layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const int n = net->tensors[src].n, c = net->tensors[src].c;
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, c, 1, 1, dst_id));
    node.type = BCNN_LAYER_AVGPOOL;
    node.forward = bcnn_forward_avgpool_layer;
    node.backward = bcnn_backward_avgpool_layer;
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx, "[Avgpool] %-8s (%4d x%4d x%4d) -> %-8s (1 x 1 x%4d)\n", src_id,
              net->tensors[src].w, net->tensors[src].h, c, dst_id, c);
    return BCNN_SUCCESS;
}

void bcnn_forward_avgpool_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    if (bcnn_net_node_is_resident(net, node)) { /* BF16 NHWC in, FP32 [n, c] out */
        bcnn_cuda_check(bcnn_b200_avgpool_forward_nhwc(bcnn_net_data16_in(net, node->src[0]), dst->data_gpu,
                                                       src->n, src->c, src->h * src->w, bcnn_stream(net)));
        bcnn_net_data32_written(net, node->dst[0]);
        return;
    }
    bcnn_cuda_check(bcnn_b200_avgpool_forward(src->data_gpu, dst->data_gpu, src->n * src->c,
                                              src->h * src->w, bcnn_stream(net)));
}

void bcnn_backward_avgpool_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    if (!src->grad_data_gpu) return;
    if (bcnn_net_node_is_resident(net, node)) {
        const float *dy = bcnn_net_grad32_in(net, node->dst[0]);
        const int accumulate = bcnn_net_grad_accumulate(net, node->src[0]);
        void *dx16 = accumulate ? bcnn_net_grad16_in(net, node->src[0]) : bcnn_net_grad16_out(net, node->src[0]);
        bcnn_cuda_check(bcnn_b200_avgpool_backward_nhwc(dx16, dy, src->n, src->c, src->h * src->w,
                                                        accumulate, bcnn_stream(net)));
        bcnn_net_grad16_modified(net, node->src[0]);
        return;
    }
    bcnn_net_grad_prepare_accumulate(net, node->src[0]); /* the kernel does += */
    bcnn_cuda_check(bcnn_b200_avgpool_backward(src->grad_data_gpu, dst->grad_data_gpu,
                                               src->n * src->c, src->h * src->w,
                                               bcnn_stream(net)));
}

void bcnn_forward_avgpool_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_forward_avgpool_layer_gpu(net, node);
}

void bcnn_backward_avgpool_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_backward_avgpool_layer_gpu(net, node);
}
