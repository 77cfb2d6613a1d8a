/*
 * bcnn_glue_layers.c -- softmax, euclidean cost and residual-add nodes on the device.
 *
 * softmax : jnbraun/bcnn src/layers/bcnn_softmax_layer.c:36-166 (forward = softmax over
 *           channels; backward passes the gradient through: src.grad += dst.grad).
 * cost    : src/layers/bcnn_cost_layer.c:35-284, euclidean loss only: forward writes
 *           dst.grad = pred - label and the scalar metric into dst.data[0] (on the device;
 *           the reference copies three tensors to the host every step); backward does
 *           src.grad += scale * dst.grad. The label tensor (tensors[1]) takes the source's
 *           shape when the cost node is added (:70-74).
 * eltwise : src/layers/bcnn_eltwise_layer.c:33-161, equal-shape inputs only. The
 *           reference adds only sample 0 of the second input on this path (:119-121,
 *           148-150); that is reproduced while reference_quirks is on
 *           (default) and replaced by the batch-correct add when it is off. Source order
 *           follows the reference's reverse scan (the later-defined tensor becomes src[0]).
 */
#include "bcnn_glue_layers.h"

#include "bcnn_conv_layer.h"
#include "bcnn_tensor.h"

/* ------------------------------- softmax ------------------------------- */

bcnn_status bcnn_add_softmax_layer(bcnn_net *net, const char *src_id, const char *dst_id) {
    bcnn_node node = {0};
    BCNN_CHECK_AND_LOG(net->log_ctx, net->num_nodes >= 1, BCNN_INVALID_PARAMETER,
                       "Softmax layer can't be the first layer of the network\n");
    int src = bcnn_get_tensor_index_by_name(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Softmax layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const bcnn_tensor *s = &net->tensors[src];
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, s->n, s->c, s->h, s->w, dst_id));
    node.type = BCNN_LAYER_SOFTMAX;
    node.forward = bcnn_forward_softmax_layer;
    node.backward = bcnn_backward_softmax_layer;
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    return BCNN_SUCCESS;
}

void bcnn_forward_softmax_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    bcnn_cuda_check(bcnn_b200_softmax_forward(src->data_gpu, dst->data_gpu, src->n, src->c,
                                              src->h * src->w, bcnn_stream(net)));
}

void bcnn_backward_softmax_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    if (!src->grad_data_gpu) return;
    bcnn_net_grad_prepare_accumulate(net, node->src[0]);
    bcnn_cuda_check(bcnn_b200_axpy(src->grad_data_gpu, dst->grad_data_gpu,
                                   (size_t)bcnn_tensor_size(src), 1.0f, bcnn_stream(net)));
}

/* -------------------------------- cost -------------------------------- */

bcnn_status bcnn_add_cost_layer(bcnn_net *net, bcnn_loss loss, bcnn_loss_metric loss_metric,
                                float scale, const char *src_id, const char *label_id,
                                const char *dst_id) {
    (void)label_id; /* the reference always binds tensors[1] */
    bcnn_node node = {0};
    BCNN_CHECK_AND_LOG(net->log_ctx, net->num_nodes >= 1, BCNN_INVALID_PARAMETER,
                       "Cost layer can't be the first layer of the network\n");
    BCNN_CHECK_AND_LOG(net->log_ctx, loss == BCNN_LOSS_EUCLIDEAN, BCNN_INVALID_PARAMETER,
                       "Cost layer: only the euclidean loss is available on the B200 path\n");
    int src = bcnn_get_tensor_index_by_name(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Cost layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    node.type = BCNN_LAYER_COST;
    node.param_size = sizeof(bcnn_cost_param);
    bcnn_cost_param *param = (bcnn_cost_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    param->scale = scale;
    param->loss = loss;
    param->loss_metric = loss_metric;
    node.forward = bcnn_forward_cost_layer;
    node.backward = bcnn_backward_cost_layer;
    const bcnn_tensor *s = &net->tensors[src];
    const int n = s->n, c = s->c, h = s->h, w = s->w;
    bcnn_tensor_set_shape(&net->tensors[1], n, c, h, w, 0);
    BCNN_CHECK_STATUS(bcnn_tensor_allocate(&net->tensors[1], net->mode));
    BCNN_CHECK_STATUS(bcnn_tensor_ensure_host(&net->tensors[1]));
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, 1));
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, c, h, w, dst_id));
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    return BCNN_SUCCESS;
}

void bcnn_forward_cost_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_cost_param *param = (bcnn_cost_param *)node->param;
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    bcnn_tensor *label = &net->tensors[1];
    if (!label->data_gpu) return;
    bcnn_cuda_check(bcnn_b200_cost_forward(src->data_gpu, label->data_gpu, dst->grad_data_gpu,
                                           dst->data_gpu, src->n, src->c * src->h * src->w,
                                           (int)param->loss_metric, bcnn_stream(net)));
}

void bcnn_backward_cost_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_cost_param *param = (bcnn_cost_param *)node->param;
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    if (!src->grad_data_gpu || !dst->grad_data_gpu) return;
    bcnn_net_grad_prepare_accumulate(net, node->src[0]);
    bcnn_cuda_check(bcnn_b200_axpy(src->grad_data_gpu, dst->grad_data_gpu,
                                   (size_t)bcnn_tensor_size(src), param->scale, bcnn_stream(net)));
}

/* ------------------------------- eltwise ------------------------------- */

bcnn_status bcnn_add_eltwise_layer(bcnn_net *net, bcnn_activation activation, const char *src_id1,
                                   const char *src_id2, const char *dst_id) {
    bcnn_node node = {0};
    int found1 = 0, found2 = 0;
    for (int i = net->num_tensors - 1; i >= 0 && !(found1 && found2); --i) {
        const char *name = net->tensors[i].name;
        if (!name) continue;
        if (!found1 && strcmp(name, src_id1) == 0) {
            BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, i));
            found1 = 1;
        } else if (!found2 && strcmp(name, src_id2) == 0) {
            BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, i));
            found2 = 1;
        }
    }
    BCNN_CHECK_AND_LOG(net->log_ctx, found1, BCNN_INVALID_PARAMETER,
                       "Eltwise layer: invalid input node name %s\n", src_id1);
    BCNN_CHECK_AND_LOG(net->log_ctx, found2, BCNN_INVALID_PARAMETER,
                       "Eltwise layer: invalid input node name %s\n", src_id2);
    const bcnn_tensor *a = &net->tensors[node.src[0]], *b = &net->tensors[node.src[1]];
    BCNN_CHECK_AND_LOG(net->log_ctx, a->n == b->n && a->c == b->c && a->h == b->h && a->w == b->w,
                       BCNN_INVALID_PARAMETER,
                       "Eltwise layer: tensors %s and %s must have the same shape on the B200 "
                       "path\n", src_id1, src_id2);
    node.type = BCNN_LAYER_ELTWISE;
    node.param_size = sizeof(bcnn_eltwise_param);
    bcnn_eltwise_param *param = (bcnn_eltwise_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    param->activation = activation;
    param->stride[0] = param->stride[1] = 1;
    param->min_dim[0] = a->c;
    param->min_dim[1] = a->h;
    param->min_dim[2] = a->w;
    node.forward = bcnn_forward_eltwise_layer;
    node.backward = bcnn_backward_eltwise_layer;
    const int n = a->n, c = a->c, h = a->h, w = a->w;
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, c, h, w, dst_id));
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx, "[EltWiseAdd] %-8s , %-16s -> %-8s (%4d x%4d x%4d)\n", src_id1,
              src_id2, dst_id, w, h, c);
    return BCNN_SUCCESS;
}

void bcnn_forward_eltwise_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_eltwise_param *param = (bcnn_eltwise_param *)node->param;
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *dst = &t[node->dst[0]];
    const int sz = bcnn_tensor_size(dst);
    const int n_add = bcnn_ctx(net)->reference_quirks ? param->min_dim[0] * dst->h * dst->w : sz;
    if (bcnn_net_node_is_resident(net, node)) { /* sample 0 is the first C*H*W elements in NHWC too */
        /* operands whose batch norm was left to this node (conv + BN without activation read by
         * nobody else, see bcnn_conv_layer.c): normalise, add and activate in one pass */
        const void *x16[2];
        const float *mean[2] = {NULL, NULL}, *var[2] = {NULL, NULL}, *gamma[2] = {NULL, NULL},
                    *beta[2] = {NULL, NULL};
        int fused = 0;
        for (int i = 0; i < 2; ++i) {
            bcnn_resident *r = bcnn_net_res(net, node->src[i]);
            const bcnn_activation a = param->activation;
            if (r->data_at == BCNN_RES_DEFERRED && n_add == sz &&
                (a == BCNN_ACT_NONE || a == BCNN_ACT_RELU || a == BCNN_ACT_LRELU)) {
                bcnn_conv_layer_bn_operand(net, &net->nodes[r->producer], &x16[i], &mean[i], &var[i],
                                           &gamma[i], &beta[i]);
                fused = 1;
            } else {
                x16[i] = bcnn_net_data16_in(net, node->src[i]);
            }
        }
        void *y16 = bcnn_net_data16_out(net, node->dst[0]);
        if (fused)
            bcnn_cuda_check(bcnn_b200_bn_add_act_nhwc(x16[0], mean[0], var[0], gamma[0], beta[0], x16[1],
                                                      mean[1], var[1], gamma[1], beta[1], y16,
                                                      (size_t)dst->n * dst->h * dst->w, dst->c,
                                                      param->activation, bcnn_stream(net)));
        else
            bcnn_cuda_check(bcnn_b200_eltwise_forward_bf16(x16[0], x16[1], y16, (size_t)sz, (size_t)n_add,
                                                           param->activation, bcnn_stream(net)));
        return;
    }
    bcnn_cuda_check(bcnn_b200_eltwise_forward(t[node->src[0]].data_gpu, t[node->src[1]].data_gpu,
                                              dst->data_gpu, sz, n_add, param->activation,
                                              bcnn_stream(net)));
}

void bcnn_backward_eltwise_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_eltwise_param *param = (bcnn_eltwise_param *)node->param;
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *dst = &t[node->dst[0]];
    const int sz = bcnn_tensor_size(dst);
    const int n_add = bcnn_ctx(net)->reference_quirks ? param->min_dim[0] * dst->h * dst->w : sz;
    int flags = 0;
    if (t[node->src[0]].grad_data_gpu && bcnn_net_grad_accumulate(net, node->src[0])) flags |= 1;
    if (t[node->src[1]].grad_data_gpu && bcnn_net_grad_accumulate(net, node->src[1])) flags |= 2;
    if (bcnn_net_node_is_resident(net, node)) {
        const void *y16 = param->activation == BCNN_ACT_NONE ? NULL : bcnn_net_data16_in(net, node->dst[0]);
        void *dy16 = bcnn_net_grad16_in(net, node->dst[0]);
        void *g16[2] = {NULL, NULL};
        /* aliased branches that are conv + BN: this node also does the reduction pass of their
         * batch-norm backward (one extra read of the branch's raw result instead of a pass of its own) */
        const void *bn_x[2] = {NULL, NULL};
        const float *bn_mean[2] = {NULL, NULL};
        float *bn_partial[2] = {NULL, NULL};
        bcnn_node *bn_node[2] = {NULL, NULL};
        for (int i = 0; i < 2; ++i) {
            if (!t[node->src[i]].grad_data_gpu) continue;
            /* a source only this node reads, produced by a conv + BN without activation: its
             * gradient IS the masked dy this node leaves in dst.grad -- the producer reads it there
             * (bcnn_resident.grad_alias) instead of from a copy written here */
            int consumer = -1;
            bcnn_resident *r = bcnn_net_res(net, node->src[i]);
            if (!(flags & (1 << i)) && n_add == sz && r->producer >= 0 && r->producer < net->num_nodes &&
                bcnn_net_sole_eltwise_consumer(net, node->src[i], &consumer) &&
                &net->nodes[consumer] == node && bcnn_conv_layer_takes_grad_alias(net, node->src[i])) {
                r->grad_alias = node->dst[0] + 1;
                if (bcnn_conv_layer_bn_reduce_operand(net, &net->nodes[r->producer], &bn_x[i], &bn_mean[i],
                                                      &bn_partial[i]))
                    bn_node[i] = &net->nodes[r->producer];
                continue;
            }
            g16[i] = (flags & (1 << i)) ? bcnn_net_grad16_in(net, node->src[i])
                                        : bcnn_net_grad16_out(net, node->src[i]);
        }
        const bcnn_activation a = param->activation;
        if ((bn_node[0] || bn_node[1]) && n_add == sz &&
            (a == BCNN_ACT_NONE || a == BCNN_ACT_RELU || a == BCNN_ACT_LRELU)) {
            int rows = 0;
            bcnn_cuda_check(bcnn_b200_eltwise_backward_bn_reduce_bf16(
                y16 ? y16 : dy16, dy16, g16[0], g16[1], (size_t)dst->n * dst->h * dst->w, dst->c, a, flags,
                bn_node[0] ? bn_x[0] : NULL, bn_mean[0], bn_partial[0], bn_node[1] ? bn_x[1] : NULL, bn_mean[1],
                bn_partial[1], &rows, bcnn_stream(net)));
            for (int i = 0; i < 2; ++i)
                if (bn_node[i]) bcnn_conv_layer_bn_reduce_done(bn_node[i], rows);
        } else
        bcnn_cuda_check(bcnn_b200_eltwise_backward_bf16(y16 ? y16 : dy16, dy16, g16[0], g16[1], (size_t)sz,
                                                        (size_t)n_add, param->activation, flags,
                                                        bcnn_stream(net)));
        bcnn_net_grad16_modified(net, node->dst[0]);
        for (int i = 0; i < 2; ++i)
            if (g16[i]) bcnn_net_grad16_modified(net, node->src[i]);
        return;
    }
    bcnn_cuda_check(bcnn_b200_eltwise_backward(
        dst->data_gpu, dst->grad_data_gpu, t[node->src[0]].grad_data_gpu,
        t[node->src[1]].grad_data_gpu, sz, n_add, param->activation, flags, bcnn_stream(net)));
}

/* ------------------------------- concat ------------------------------- */
/* src/layers/bcnn_concat_layer.c:34-142: channel concatenation of num_src tensors with equal
 * spatial size; backward adds the output-gradient slices into the sources' gradients. */

bcnn_status bcnn_add_concat_layer(bcnn_net *net, int num_src, char *const *src_ids,
                                  const char *dst_id) {
    bcnn_node node = {0};
    BCNN_CHECK_AND_LOG(net->log_ctx, net->num_nodes >= 1, BCNN_INVALID_PARAMETER,
                       "Concat layer can't be the first layer of the network\n");
    BCNN_CHECK_AND_LOG(net->log_ctx, num_src >= 1, BCNN_INVALID_PARAMETER,
                       "Concat layer: no input\n");
    for (int i = 0; i < num_src; ++i) {
        int tid = bcnn_get_tensor_index_by_name(net, src_ids[i]);
        BCNN_CHECK_AND_LOG(net->log_ctx, tid >= 0, BCNN_INVALID_PARAMETER,
                           "Concat layer: invalid input node name %s\n", src_ids[i]);
        BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, tid));
    }
    const bcnn_tensor *first = &net->tensors[node.src[0]];
    int out_c = first->c;
    for (int i = 1; i < node.num_src; ++i) {
        const bcnn_tensor *t = &net->tensors[node.src[i]];
        BCNN_CHECK_AND_LOG(net->log_ctx, t->w == first->w && t->h == first->h && t->n == first->n,
                           BCNN_INVALID_PARAMETER,
                           "Concat layer: inconsistent sizes between node %s (%dx%d) and node %s "
                           "(%dx%d)\n", src_ids[0], first->w, first->h, src_ids[i], t->w, t->h);
        out_c += t->c;
    }
    const int n = first->n, h = first->h, w = first->w;
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, out_c, h, w, dst_id));
    node.type = BCNN_LAYER_CONCAT;
    node.forward = bcnn_forward_concat_layer;
    node.backward = bcnn_backward_concat_layer;
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    return BCNN_SUCCESS;
}

void bcnn_forward_concat_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_tensor *dst = &net->tensors[node->dst[0]];
    const int dst_sz = dst->c * dst->h * dst->w;
    int offset = 0;
    for (int i = 0; i < node->num_src; ++i) {
        bcnn_tensor *src = &net->tensors[node->src[i]];
        const int src_sz = src->c * src->h * src->w;
        bcnn_cuda_check(bcnn_b200_concat_forward(src->data_gpu, dst->data_gpu, src->n, src_sz, dst_sz,
                                                 offset, bcnn_stream(net)));
        offset += src_sz;
    }
}

void bcnn_backward_concat_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_tensor *dst = &net->tensors[node->dst[0]];
    const int dst_sz = dst->c * dst->h * dst->w;
    int offset = 0;
    for (int i = 0; i < node->num_src; ++i) {
        bcnn_tensor *src = &net->tensors[node->src[i]];
        const int src_sz = src->c * src->h * src->w;
        if (src->grad_data_gpu)
            bcnn_cuda_check(bcnn_b200_concat_backward(
                dst->grad_data_gpu, src->grad_data_gpu, src->n, src_sz, dst_sz, offset,
                bcnn_net_grad_accumulate(net, node->src[i]), bcnn_stream(net)));
        offset += src_sz;
    }
}

/* ------------------------------ upsample ------------------------------ */
/* src/layers/bcnn_upsample_layer.c:33-142: nearest-neighbour upsampling by `size`. */

static void bcnn_release_param_upsample_layer(bcnn_node *node) { (void)node; }

bcnn_status bcnn_add_upsample_layer(bcnn_net *net, int size, const char *src_id,
                                    const char *dst_id) {
    bcnn_node node = {0};
    BCNN_CHECK_AND_LOG(net->log_ctx, size >= 1, BCNN_INVALID_PARAMETER,
                       "Upsample layer: invalid factor %d\n", size);
    int src = 0;
    if (net->num_nodes > 0) {
        src = bcnn_get_tensor_index_by_name(net, src_id);
        BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                           "Upsample layer: invalid input node name %s\n", src_id);
    }
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const bcnn_tensor *s = &net->tensors[src];
    const int n = s->n, c = s->c, h = s->h, w = s->w;
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, c, h * size, w * size, dst_id));
    node.type = BCNN_LAYER_UPSAMPLE;
    node.param_size = sizeof(bcnn_upsample_param);
    bcnn_upsample_param *param = (bcnn_upsample_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    param->size = size;
    node.param = param;
    node.forward = bcnn_forward_upsample_layer;
    node.backward = bcnn_backward_upsample_layer;
    node.release_param = bcnn_release_param_upsample_layer;
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    return BCNN_SUCCESS;
}

void bcnn_forward_upsample_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_upsample_param *param = (bcnn_upsample_param *)node->param;
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    bcnn_cuda_check(bcnn_b200_upsample_forward(src->data_gpu, dst->data_gpu, src->n, src->c, src->h,
                                               src->w, param->size, bcnn_stream(net)));
}

void bcnn_backward_upsample_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_upsample_param *param = (bcnn_upsample_param *)node->param;
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    if (!src->grad_data_gpu) return;
    bcnn_cuda_check(bcnn_b200_upsample_backward(dst->grad_data_gpu, src->grad_data_gpu, src->n, src->c,
                                                src->h, src->w, param->size,
                                                bcnn_net_grad_accumulate(net, node->src[0]),
                                                bcnn_stream(net)));
}
