/*
 * bcnn_maxpool_layer.c -- max-pooling node on the B200 kernels.
 * Output extents per padding policy as jnbraun/bcnn src/layers/bcnn_maxpool_layer.c:62-83
 * (SAME (H+s-1)/s, VALID (H-k+s)/s, CAFFE ceil((H-k)/s)+1); forward / backward semantics
 * of the CPU path (:145-191, :258-273): see csrc/pool.cu.
 */
#include "bcnn_maxpool_layer.h"

#include <math.h>

#include <bcnn_b200_net.h>

#include "bcnn_tensor.h"

static int pooled_extent(int in, int size, int stride, bcnn_padding padding) {
    switch (padding) {
        case BCNN_PADDING_SAME: return (in + stride - 1) / stride;
        case BCNN_PADDING_VALID: return (in - size + stride) / stride;
        case BCNN_PADDING_CAFFE: return (int)(ceil((float)(in - size) / stride)) + 1;
    }
    return 0;
}

bcnn_status bcnn_add_maxpool_layer(bcnn_net *net, int size, int stride, bcnn_padding padding,
                                   const char *src_id, const char *dst_id) {
    bcnn_node node = {0};
    int src = bcnn_net_find_src(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Maxpool layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, size > 0 && stride > 0, BCNN_INVALID_PARAMETER,
                       "Maxpool layer: invalid size / stride\n");
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const bcnn_tensor *s = &net->tensors[src];
    const int ho = pooled_extent(s->h, size, stride, padding);
    const int wo = pooled_extent(s->w, size, stride, padding);
    const int n = s->n, c = s->c, h = s->h, w = s->w;
    BCNN_CHECK_AND_LOG(net->log_ctx, ho > 0 && wo > 0, BCNN_INVALID_PARAMETER,
                       "Maxpool layer: empty output\n");
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, c, ho, wo, dst_id));

    node.type = BCNN_LAYER_MAXPOOL;
    node.param_size = sizeof(bcnn_maxpool_param);
    bcnn_maxpool_param *param = (bcnn_maxpool_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    param->size = size;
    param->stride = stride;
    param->padding = padding;
    param->indexes_gpu = (int *)bcnn_b200_malloc((size_t)n * c * ho * wo * sizeof(int));
    BCNN_CHECK(param->indexes_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);
    node.forward = bcnn_forward_maxpool_layer;
    node.backward = bcnn_backward_maxpool_layer;
    node.release_param = bcnn_release_param_maxpool_layer;
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx, "[Maxpool] %-8s (%4d x%4d x%4d) -> %-8s (%4d x%4d x%4d) %d x %d / %d\n",
              src_id, w, h, c, dst_id, wo, ho, c, size, size, stride);
    return BCNN_SUCCESS;
}

void bcnn_forward_maxpool_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_maxpool_param *param = (bcnn_maxpool_param *)node->param;
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    if (bcnn_net_node_is_resident(net, node)) { /* BF16 NHWC in and out; index VALUES stay NCHW-flat */
        const void *x16 = bcnn_net_data16_in(net, node->src[0]);
        void *y16 = bcnn_net_data16_out(net, node->dst[0]);
        bcnn_cuda_check(bcnn_b200_maxpool_forward_nhwc(x16, y16, param->indexes_gpu, src->n, src->c,
                                                       src->h, src->w, param->size, param->stride,
                                                       dst->h, dst->w, bcnn_stream(net)));
        param->indexes_nhwc = 1;
        return;
    }
    param->indexes_nhwc = 0;
    bcnn_cuda_check(bcnn_b200_maxpool_forward(src->data_gpu, dst->data_gpu, param->indexes_gpu,
                                              src->n, src->c, src->h, src->w, param->size,
                                              param->stride, dst->h, dst->w, bcnn_stream(net)));
}

void bcnn_backward_maxpool_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_maxpool_param *param = (bcnn_maxpool_param *)node->param;
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    if (!src->grad_data_gpu) return;
    if (bcnn_net_node_is_resident(net, node) && param->indexes_nhwc) {
        const void *dy16 = bcnn_net_grad16_in(net, node->dst[0]);
        const int accumulate = bcnn_net_grad_accumulate(net, node->src[0]);
        void *dx16 = accumulate ? bcnn_net_grad16_in(net, node->src[0]) : bcnn_net_grad16_out(net, node->src[0]);
        bcnn_cuda_check(bcnn_b200_maxpool_backward_nhwc(dx16, dy16, param->indexes_gpu, src->n, src->c,
                                                        src->h, src->w, param->size, param->stride,
                                                        dst->h, dst->w, accumulate, bcnn_stream(net)));
        bcnn_net_grad16_modified(net, node->src[0]);
        return;
    }
    bcnn_net_grad_prepare_accumulate(net, node->src[0]); /* the kernel does += */
    bcnn_cuda_check(bcnn_b200_maxpool_backward(src->grad_data_gpu, dst->grad_data_gpu,
                                               param->indexes_gpu, src->n, src->c, src->h, src->w,
                                               param->size, param->stride, dst->h, dst->w,
                                               bcnn_stream(net)));
}

void bcnn_forward_maxpool_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_forward_maxpool_layer_gpu(net, node);
}

void bcnn_backward_maxpool_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_backward_maxpool_layer_gpu(net, node);
}

void bcnn_release_param_maxpool_layer(bcnn_node *node) {
    bcnn_maxpool_param *param = (bcnn_maxpool_param *)node->param;
    free(param->indexes);
    bcnn_b200_free(param->indexes_gpu);
}

int bcnn_b200_maxpool_indexes(bcnn_net *net, int node_index, int *host_out) {
    if (node_index < 0 || node_index >= net->num_nodes) return -1;
    bcnn_node *node = &net->nodes[node_index];
    if (node->type != BCNN_LAYER_MAXPOOL) return -1;
    bcnn_maxpool_param *param = (bcnn_maxpool_param *)node->param;
    const bcnn_tensor *dst = &net->tensors[node->dst[0]];
    int count = bcnn_tensor_size(dst);
    bcnn_cuda_check(bcnn_b200_memcpy_d2h(host_out, param->indexes_gpu, (size_t)count * sizeof(int),
                                         bcnn_stream(net)));
    bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_stream(net)));
    if (param->indexes_nhwc) { /* hand the buffer out in the reference's NCHW order */
        int *tmp = (int *)malloc((size_t)count * sizeof(int));
        if (!tmp) return -1;
        memcpy(tmp, host_out, (size_t)count * sizeof(int));
        const int c = dst->c, hw = dst->h * dst->w;
        for (int n = 0; n < dst->n; ++n)
            for (int p = 0; p < hw; ++p)
                for (int ch = 0; ch < c; ++ch)
                    host_out[((size_t)n * c + ch) * hw + p] = tmp[((size_t)n * hw + p) * c + ch];
        free(tmp);
    }
    return count;
}
