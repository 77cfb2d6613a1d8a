/*
 * bcnn_resident.c -- resident BF16 NHWC activations (BCNN_B200_MATH_TC_BF16).
 *
 * bcnn's tensors are FP32 NCHW (reference inc/bcnn/bcnn.h:242-255) and every layer of the
 * reference reads and writes that format. With FP32 storage a ResNet-50 step is bound by the
 * bytes the batch-norm / residual / pooling passes move, and every convolution needs a transposed
 * BF16 copy of its operand in front of the TMA. In the resident mode a tensor (and its gradient)
 * therefore lives as a BF16 NHWC twin between the layers that know the format -- convolution,
 * max / average pooling, the residual add -- and the FP32 NCHW buffers of the bcnn_tensor are
 * brought up to date only when somebody needs them:
 *   - bcnn_get_tensor_by_index / bcnn_get_tensor_by_name (the public read-out, reference
 *     src/bcnn_net.c:388-408),
 *   - in front of a node that is not format-aware (fully-connected, softmax, cost, concat,
 *     upsample, yolo, depthwise, stand-alone batch norm / activation): bcnn_forward / bcnn_backward
 *     call the *_f32_before_* / *_after_* hooks around it.
 * Per tensor the state says which copy holds the current value (BCNN_RES_F32 / _BF16 / _BOTH);
 * conversions are the two kernels at the top of csrc/nhwc_bf16.cu. Parameter tensors never get a
 * twin. All of it is host-side bookkeeping on the net's stream: no synchronisation.
 *
 * CUDA graphs: the state transitions of a step depend only on the node sequence, so a recorded
 * step replays the same conversions; the twins are allocated during the eager warm-up step.
 */
#include "bcnn_net.h"

#include "bcnn_conv_layer.h"
#include "bcnn_tensor.h"

int bcnn_net_resident(bcnn_net *net) { return bcnn_ctx(net)->conv_math == BCNN_B200_MATH_TC_BF16; }

static bcnn_resident *res_of(bcnn_net *net, int idx) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (idx >= ctx->res_count) {
        const int grown = net->num_tensors > idx + 1 ? net->num_tensors : idx + 1;
        bcnn_resident *r = (bcnn_resident *)realloc(ctx->res, (size_t)grown * sizeof(bcnn_resident));
        if (!r) {
            fprintf(stderr, "[ERROR] resident state: out of memory\n");
            exit(BCNN_FAILED_ALLOC);
        }
        memset(r + ctx->res_count, 0, (size_t)(grown - ctx->res_count) * sizeof(bcnn_resident));
        ctx->res = r;
        ctx->res_count = grown;
    }
    return &ctx->res[idx];
}

bcnn_resident *bcnn_net_res(bcnn_net *net, int idx) { return res_of(net, idx); }

int bcnn_net_sole_eltwise_consumer(bcnn_net *net, int idx, int *consumer_out) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (!bcnn_net_resident(net) || ctx->reference_quirks) return 0;
    int found = -1;
    for (int i = 0; i < net->num_nodes; ++i) {
        const bcnn_node *node = &net->nodes[i];
        for (int j = 0; j < node->num_src; ++j) {
            if (node->src[j] != idx) continue;
            if (found >= 0 || node->type != BCNN_LAYER_ELTWISE || j > 1) return 0;
            found = i;
        }
    }
    if (found < 0 || !bcnn_net_tensor_can16(net, net->nodes[found].dst[0])) return 0;
    if (net->nodes[found].src[0] == net->nodes[found].src[1]) return 0;
    if (consumer_out) *consumer_out = found;
    return 1;
}

static void materialize_deferred(bcnn_net *net, int idx) {
    bcnn_resident *r = res_of(net, idx);
    if (r->data_at != BCNN_RES_DEFERRED) return;
    bcnn_conv_layer_materialize(net, &net->nodes[r->producer]); /* leaves data_at = BCNN_RES_BF16 */
}

int bcnn_net_tensor_can16(bcnn_net *net, int idx) {
    const bcnn_tensor *t = &net->tensors[idx];
    return t->c % 8 == 0 && bcnn_tensor_size(t) > 0 && t->data_gpu != NULL;
}

static void *twin(bcnn_net *net, void **slot, int idx) {
    if (!*slot) {
        *slot = bcnn_b200_malloc((size_t)bcnn_tensor_size(&net->tensors[idx]) * 2);
        if (!*slot) {
            fprintf(stderr, "[ERROR] [CUDA] cannot allocate the BF16 twin of tensor %s\n",
                    net->tensors[idx].name ? net->tensors[idx].name : "?");
            exit(BCNN_CUDA_FAILED_ALLOC);
        }
    }
    return *slot;
}

static void to16(bcnn_net *net, const float *src, void *dst, const bcnn_tensor *t) {
    bcnn_cuda_check(bcnn_b200_f32nchw_to_bf16nhwc(src, dst, t->n, t->c, t->h * t->w, bcnn_stream(net)));
}
static void to32(bcnn_net *net, const void *src, float *dst, const bcnn_tensor *t) {
    bcnn_cuda_check(bcnn_b200_bf16nhwc_to_f32nchw(src, dst, t->n, t->c, t->h * t->w, bcnn_stream(net)));
}

void *bcnn_net_data16_in(bcnn_net *net, int idx) {
    materialize_deferred(net, idx);
    bcnn_resident *r = res_of(net, idx);
    void *p = twin(net, &r->data16, idx);
    if (r->data_at == BCNN_RES_F32) {
        to16(net, net->tensors[idx].data_gpu, p, &net->tensors[idx]);
        r->data_at = BCNN_RES_BOTH;
    }
    return p;
}

void *bcnn_net_grad16_in(bcnn_net *net, int idx) {
    bcnn_resident *r = res_of(net, idx);
    void *p = twin(net, &r->grad16, idx);
    if (r->grad_at == BCNN_RES_F32) {
        to16(net, net->tensors[idx].grad_data_gpu, p, &net->tensors[idx]);
        r->grad_at = BCNN_RES_BOTH;
    }
    return p;
}

void *bcnn_net_data16_out(bcnn_net *net, int idx) {
    bcnn_resident *r = res_of(net, idx);
    r->data_at = BCNN_RES_BF16;
    return twin(net, &r->data16, idx);
}

void *bcnn_net_grad16_out(bcnn_net *net, int idx) {
    bcnn_resident *r = res_of(net, idx);
    r->grad_at = BCNN_RES_BF16;
    return twin(net, &r->grad16, idx);
}

void bcnn_net_grad16_modified(bcnn_net *net, int idx) { res_of(net, idx)->grad_at = BCNN_RES_BF16; }

float *bcnn_net_data32_in(bcnn_net *net, int idx) {
    bcnn_tensor *t = &net->tensors[idx];
    if (idx < bcnn_ctx(net)->res_count) {
        materialize_deferred(net, idx);
        bcnn_resident *r = &bcnn_ctx(net)->res[idx];
        if (r->data_at == BCNN_RES_BF16 && t->data_gpu) {
            to32(net, r->data16, t->data_gpu, t);
            r->data_at = BCNN_RES_BOTH;
        }
    }
    return t->data_gpu;
}

float *bcnn_net_grad32_in(bcnn_net *net, int idx) {
    bcnn_tensor *t = &net->tensors[idx];
    if (idx < bcnn_ctx(net)->res_count) {
        bcnn_resident *r = &bcnn_ctx(net)->res[idx];
        if (r->grad_at == BCNN_RES_BF16 && t->grad_data_gpu) {
            to32(net, r->grad16, t->grad_data_gpu, t);
            r->grad_at = BCNN_RES_BOTH;
        }
    }
    return t->grad_data_gpu;
}

void bcnn_net_data32_written(bcnn_net *net, int idx) {
    if (idx < bcnn_ctx(net)->res_count) bcnn_ctx(net)->res[idx].data_at = BCNN_RES_F32;
}

void bcnn_net_grad32_written(bcnn_net *net, int idx) {
    if (idx < bcnn_ctx(net)->res_count) bcnn_ctx(net)->res[idx].grad_at = BCNN_RES_F32;
}

int bcnn_net_node_is_resident(bcnn_net *net, bcnn_node *node) {
    if (!bcnn_net_resident(net)) return 0;
    switch (node->type) {
        case BCNN_LAYER_CONV2D:
            return bcnn_conv_layer_is_resident(net, node);
        case BCNN_LAYER_MAXPOOL:
        case BCNN_LAYER_AVGPOOL:
            return bcnn_net_tensor_can16(net, node->src[0]);
        case BCNN_LAYER_ELTWISE:
            return bcnn_net_tensor_can16(net, node->dst[0]);
        default:
            return 0;
    }
}

void bcnn_net_node_f32_before_forward(bcnn_net *net, bcnn_node *node) {
    if (bcnn_ctx(net)->res_count == 0) return;
    for (int i = 0; i < node->num_src; ++i) (void)bcnn_net_data32_in(net, node->src[i]);
}

void bcnn_net_node_f32_after_forward(bcnn_net *net, bcnn_node *node) {
    if (bcnn_ctx(net)->res_count == 0) return;
    for (int i = 0; i < node->num_dst; ++i) bcnn_net_data32_written(net, node->dst[i]);
}

void bcnn_net_node_f32_before_backward(bcnn_net *net, bcnn_node *node) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (ctx->res_count == 0) return;
    for (int i = 0; i < node->num_src; ++i) {
        const int idx = node->src[i];
        (void)bcnn_net_data32_in(net, idx);
        /* a partial sum of this step that lives in the BF16 twin: the node will add to it */
        if (net->tensors[idx].grad_data_gpu && ctx->grad_fresh && idx < ctx->grad_state_tensors &&
            ctx->grad_fresh[idx])
            (void)bcnn_net_grad32_in(net, idx);
    }
    for (int i = 0; i < node->num_dst; ++i) {
        (void)bcnn_net_data32_in(net, node->dst[i]);
        (void)bcnn_net_grad32_in(net, node->dst[i]);
    }
}

void bcnn_net_node_f32_after_backward(bcnn_net *net, bcnn_node *node) {
    if (bcnn_ctx(net)->res_count == 0) return;
    for (int i = 0; i < node->num_src; ++i)
        if (net->tensors[node->src[i]].grad_data_gpu) bcnn_net_grad32_written(net, node->src[i]);
    /* in-place layers (activation backward, fused bias / batch-norm backward) rewrite dst.grad */
    for (int i = 0; i < node->num_dst; ++i) bcnn_net_grad32_written(net, node->dst[i]);
}

float *bcnn_net_nhwc_scratch(bcnn_net *net, int channels) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    const size_t need = bcnn_b200_nhwc_scratch_floats(channels);
    if (need > ctx->nhwc_scratch_floats) {
        /* grows only while the net warms up (eager steps); earlier users must be done with it */
        bcnn_cuda_check(bcnn_b200_stream_sync(ctx->stream));
        bcnn_b200_free(ctx->nhwc_scratch_gpu);
        ctx->nhwc_scratch_gpu = (float *)bcnn_b200_malloc(need * sizeof(float));
        if (!ctx->nhwc_scratch_gpu) {
            fprintf(stderr, "[ERROR] [CUDA] cannot allocate the NHWC reduction scratch\n");
            exit(BCNN_CUDA_FAILED_ALLOC);
        }
        ctx->nhwc_scratch_floats = need;
    }
    return ctx->nhwc_scratch_gpu;
}

void bcnn_net_resident_release(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    for (int i = 0; i < ctx->res_count; ++i) {
        bcnn_b200_free(ctx->res[i].data16);
        bcnn_b200_free(ctx->res[i].grad16);
    }
    free(ctx->res);
    ctx->res = NULL;
    ctx->res_count = 0;
    bcnn_b200_free(ctx->nhwc_scratch_gpu);
    ctx->nhwc_scratch_gpu = NULL;
    ctx->nhwc_scratch_floats = 0;
}
