/*
 * bcnn_conv_layer.c -- convolution node on the B200 kernels.
 *
 * Node layout is the reference's (jnbraun/bcnn src/layers/bcnn_conv_layer.c:45-365):
 * src[0] = x, src[1] = W [Cout, Cin/g, k, k] ("<src>_w"), src[2] = bias / beta ("<src>_b"),
 * with batch_norm: src[3] = running mean, src[4] = running var, src[5] = gamma
 * ("<src>_scales"); PReLU slopes follow at src[3 + 3*batch_norm]. Output extent
 * (H + 2p - k) / s + 1.
 *
 * Execution differs from the reference's GPU path (:590-790: per-image im2col + cuBLAS
 * SGEMM, then separate bias / BN / activation kernels): one implicit-GEMM launch covers
 * the whole batch; bias + activation are fused into its epilogue, or, with batch_norm,
 * the raw result goes to bn_workspace_gpu and one fused BN(+activation) pass writes dst.
 * Backward fuses activation-backward into the BN (or bias-gradient) reduction.
 * Semantics kept from the CPU path (:487-587): weight gradients accumulate (beta = 1),
 * the data gradient overwrites src.grad (beta = 0 + zero-filling col2im).
 */
#include "bcnn_conv_layer.h"

#include "bcnn_activation_layer.h"
#include "bcnn_batchnorm_layer.h"
#include "bcnn_learner.h"
#include "bcnn_tensor.h"

bcnn_status bcnn_add_convolutional_layer(bcnn_net *net, int n, int size, int stride, int pad,
                                         int num_groups, int batch_norm, bcnn_filler_type init,
                                         bcnn_activation activation, int quantize,
                                         const char *src_id, const char *dst_id) {
    (void)quantize;
    bcnn_node node = {0};
    int src = bcnn_net_find_src(net, src_id);
    if (net->num_nodes > 0) {
        BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                           "Convolution layer: invalid input node name %s\n", src_id);
    } else {
        BCNN_CHECK_AND_LOG(net->log_ctx, bcnn_tensor_size(&net->tensors[0]) > 0,
                           BCNN_INVALID_PARAMETER,
                           "Invalid input size of the network. Hint: Use 'bcnn_set_input_shape' "
                           "to set the network input size\n");
    }
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const int batch = net->tensors[src].n, cin = net->tensors[src].c, h = net->tensors[src].h,
              w = net->tensors[src].w;
    BCNN_CHECK_AND_LOG(net->log_ctx, num_groups > 0 && cin % num_groups == 0,
                       BCNN_INVALID_PARAMETER,
                       "Number of input channels has to be a multiple of the number of groups\n");
    BCNN_CHECK_AND_LOG(net->log_ctx, n % num_groups == 0, BCNN_INVALID_PARAMETER,
                       "Number of output channels has to be a multiple of the number of groups\n");
    BCNN_CHECK_AND_LOG(net->log_ctx, size > 0 && stride > 0 && pad >= 0, BCNN_INVALID_PARAMETER,
                       "Convolution layer: invalid kernel size / stride / pad\n");
    const int cin_g = cin / num_groups;
    bcnn_tensor_filler wfill = {.range = size * size * cin_g, .type = init};
    BCNN_CHECK_STATUS(
        bcnn_net_add_param_tensor(net, &node, n, cin_g, size, size, 1, src_id, "_w", &wfill));
    BCNN_CHECK_STATUS(bcnn_net_add_param_tensor(net, &node, 1, 1, 1, n, 1, src_id, "_b", NULL));

    node.type = BCNN_LAYER_CONV2D;
    node.param_size = sizeof(bcnn_conv_param);
    bcnn_conv_param *param = (bcnn_conv_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    param->activation = activation;
    param->pad = pad;
    param->num = n;
    param->size = size;
    param->stride = stride;
    param->num_groups = num_groups;
    node.forward = bcnn_forward_conv_layer;
    node.backward = bcnn_backward_conv_layer;
    node.update = bcnn_update_conv_layer;
    node.release_param = bcnn_release_param_conv_layer;

    const int ho = (h + 2 * pad - size) / stride + 1, wo = (w + 2 * pad - size) / stride + 1;
    BCNN_CHECK_AND_LOG(net->log_ctx, ho > 0 && wo > 0, BCNN_INVALID_PARAMETER,
                       "Convolution layer: empty output\n");
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, batch, n, ho, wo, dst_id));
    bcnn_b200_conv_desc desc = {batch, cin, h, w, n, ho, wo, size, stride, pad, num_groups};
    param->desc = desc;
    size_t ws_bytes = bcnn_b200_conv_workspace_bytes(&desc, BCNN_B200_MATH_TC);
    if (bcnn_b200_conv_nhwc_workspace_bytes(&desc) > ws_bytes)
        ws_bytes = bcnn_b200_conv_nhwc_workspace_bytes(&desc);
    param->workspace_size = (ws_bytes + sizeof(float) - 1) / sizeof(float);
    bcnn_net_require_workspace(net, param->workspace_size * sizeof(float));
    bcnn_net_require_dy_shadow(net, bcnn_b200_conv_dy_shadow_bytes(&desc, BCNN_B200_MATH_TC));
    param->reduce_scratch_gpu =
        (float *)bcnn_b200_malloc(bcnn_b200_bn_scratch_floats(n) * sizeof(float));
    BCNN_CHECK(param->reduce_scratch_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);

    if (batch_norm) {
        param->batch_norm = 1;
        char name[320];
        snprintf(name, sizeof(name), "%s_sav_mean", src_id);
        bcnn_tensor_create(&param->saved_mean, 1, 1, 1, n, 1, name, net->mode);
        snprintf(name, sizeof(name), "%s_sav_var", src_id);
        bcnn_tensor_create(&param->saved_variance, 1, 1, 1, n, 1, name, net->mode);
        bcnn_tensor_filler ones = {.value = 1.0f, .type = BCNN_FILLER_FIXED};
        BCNN_CHECK_STATUS(
            bcnn_net_add_param_tensor(net, &node, 1, 1, 1, n, 0, src_id, "_run_mean", NULL));
        BCNN_CHECK_STATUS(
            bcnn_net_add_param_tensor(net, &node, 1, 1, 1, n, 0, src_id, "_run_var", NULL));
        BCNN_CHECK_STATUS(
            bcnn_net_add_param_tensor(net, &node, 1, 1, 1, n, 1, src_id, "_scales", &ones));
        if (net->mode != BCNN_MODE_PREDICT) { /* raw conv output, needed by BN backward */
            param->bn_workspace_gpu =
                (float *)bcnn_b200_malloc((size_t)batch * n * ho * wo * sizeof(float));
            BCNN_CHECK(param->bn_workspace_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);
        }
    }
    if (activation == BCNN_ACT_PRELU)
        BCNN_CHECK_STATUS(
            bcnn_net_add_param_tensor(net, &node, 1, 1, 1, n, 0, src_id, "_prelu_slopes", NULL));

    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx,
              "[Conv2d]%s[%s] %-8s (%4d x%4d x%4d) -> %-8s (%4d x%4d x%4d) %5d (%d) %2d x %2d / "
              "%2d,%2d\n",
              batch_norm ? "[BN]" : "", bcnn_act2str(activation), src_id, w, h, cin, dst_id, wo,
              ho, n, n, num_groups, size, size, stride, pad);
    return BCNN_SUCCESS;
}

/* Does this node run on the resident (BF16 NHWC) kernels? All of its passes must: fprop, and when
 * the net trains, wgrad and (if the source has a gradient) dgrad; fused activations are the ones
 * the NHWC batch-norm kernels know. Decided once per node. */
int bcnn_conv_layer_is_resident(bcnn_net *net, bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    if (param->resident_state == 0) {
        const bcnn_tensor *src = &net->tensors[node->src[0]];
        const int mask = bcnn_b200_conv_nhwc_supported(&param->desc);
        int ok = (mask & 1) != 0;
        if (net->mode != BCNN_MODE_PREDICT) {
            ok = ok && (mask & 4);
            if (src->grad_data_gpu) ok = ok && (mask & 2);
        }
        const bcnn_activation a = param->activation;
        if (a == BCNN_ACT_PRELU) ok = 0;
        if (param->batch_norm && !(a == BCNN_ACT_NONE || a == BCNN_ACT_RELU || a == BCNN_ACT_LRELU)) ok = 0;
        param->resident_state = ok ? 1 : -1;
    }
    return param->resident_state > 0;
}

static int conv_reads_fp32_input(const bcnn_conv_param *param) {
    return bcnn_b200_conv_nhwc_x_keep_bytes(&param->desc) > 0; /* thin first layer: im2col route */
}

void bcnn_conv_layer_bn_operand(bcnn_net *net, bcnn_node *node, const void **raw, const float **mean,
                                const float **var, const float **gamma, const float **beta) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    bcnn_tensor *t = net->tensors;
    *raw = param->bn_raw16_gpu;
    *gamma = t[node->src[5]].data_gpu;
    *beta = t[node->src[2]].data_gpu;
    if (net->mode == BCNN_MODE_TRAIN) {
        *mean = param->saved_mean.data_gpu;
        *var = param->saved_variance.data_gpu;
    } else if (net->mode == BCNN_MODE_PREDICT) { /* folded at load time */
        *mean = *var = NULL;
    } else {
        *mean = t[node->src[3]].data_gpu;
        *var = t[node->src[4]].data_gpu;
    }
}

int bcnn_conv_layer_bn_reduce_operand(bcnn_net *net, bcnn_node *node, const void **raw, const float **mean,
                                      float **partial) {
    static int off = -1;
    if (off < 0) {
        const char *e = getenv("BCNN_B200_FUSED_BN_REDUCE");
        off = (e && e[0] == '0') ? 1 : 0;
    }
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    if (off || node->type != BCNN_LAYER_CONV2D || !param->batch_norm || net->mode != BCNN_MODE_TRAIN ||
        param->activation != BCNN_ACT_NONE || !param->bn_raw16_gpu)
        return 0;
    if (!param->bn_partial_gpu) {
        const int c = net->tensors[node->dst[0]].c;
        param->bn_partial_gpu = (float *)bcnn_b200_malloc(bcnn_b200_nhwc_scratch_floats(c) * sizeof(float));
        if (!param->bn_partial_gpu) return 0;
    }
    *raw = param->bn_raw16_gpu;
    *mean = param->saved_mean.data_gpu;
    *partial = param->bn_partial_gpu;
    return 1;
}

void bcnn_conv_layer_bn_reduce_done(bcnn_node *node, int rows) {
    ((bcnn_conv_param *)node->param)->bn_partial_rows = rows;
}

void bcnn_conv_layer_materialize(bcnn_net *net, bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    bcnn_tensor *dst = &net->tensors[node->dst[0]];
    const void *raw;
    const float *mean, *var, *gamma, *beta;
    bcnn_conv_layer_bn_operand(net, node, &raw, &mean, &var, &gamma, &beta);
    void *y16 = bcnn_net_data16_out(net, node->dst[0]); /* state: BF16 current */
    bcnn_cuda_check(bcnn_b200_bn_apply_nhwc(raw, y16, mean, var, gamma, beta,
                                            (size_t)dst->n * dst->h * dst->w, dst->c, param->activation,
                                            bcnn_stream(net)));
}

/* conv + BN without activation whose only reader is a resident residual add: that add applies the
 * normalisation (forward) and this node reads its incoming gradient from the add's output gradient
 * (backward). The node index of the add, or -1. */
static int conv_fused_into_eltwise(bcnn_net *net, bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    int consumer = -1;
    if (!param->batch_norm || param->activation != BCNN_ACT_NONE) return -1;
    if (!bcnn_net_sole_eltwise_consumer(net, node->dst[0], &consumer)) return -1;
    return consumer;
}

int bcnn_conv_layer_takes_grad_alias(bcnn_net *net, int idx) {
    for (int i = 0; i < net->num_nodes; ++i) {
        bcnn_node *node = &net->nodes[i];
        if (node->num_dst < 1 || node->dst[0] != idx) continue;
        if (node->type != BCNN_LAYER_CONV2D) return 0;
        const bcnn_conv_param *param = (const bcnn_conv_param *)node->param;
        return param->batch_norm && param->activation == BCNN_ACT_NONE && net->mode == BCNN_MODE_TRAIN &&
               bcnn_conv_layer_is_resident(net, node);
    }
    return 0;
}

static void conv_forward_resident(bcnn_net *net, bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *dst = &t[node->dst[0]];
    bcnn_tensor *weights = &t[node->src[1]], *biases = &t[node->src[2]];
    void *stream = ctx->stream;
    const size_t positions = (size_t)dst->n * dst->h * dst->w;
    const void *x;
    bcnn_b200_conv_shadows *sh = NULL;
    param->shadows.x_fmt = BCNN_B200_SHADOW_NONE; /* whatever was kept mirrors an older input */
    if (conv_reads_fp32_input(param)) {
        x = bcnn_net_data32_in(net, node->src[0]);
        if (net->mode == BCNN_MODE_TRAIN) { /* keep the im2col buffer for this step's wgrad */
            sh = &param->shadows;
            if (!sh->x) {
                const size_t bytes = bcnn_b200_conv_nhwc_x_keep_bytes(&param->desc);
                sh->x = bcnn_b200_malloc(bytes);
                sh->x_bytes = sh->x ? bytes : 0;
            }
            sh->x_fmt = BCNN_B200_SHADOW_NONE;
        }
    } else {
        x = bcnn_net_data16_in(net, node->src[0]);
    }
    const int deferred = conv_fused_into_eltwise(net, node) >= 0;
    void *y16 = bcnn_net_data16_out(net, node->dst[0]);
    if (!param->batch_norm) {
        bcnn_cuda_check(bcnn_b200_conv_forward_nhwc(&param->desc, x, weights->data_gpu, biases->data_gpu,
                                                    param->activation, y16, ctx->workspace_gpu,
                                                    ctx->workspace_bytes, sh, stream));
        return;
    }
    const float *gamma = t[node->src[5]].data_gpu;
    if (net->mode == BCNN_MODE_PREDICT && !deferred) { /* statistics folded into gamma / beta at load time */
        bcnn_cuda_check(bcnn_b200_conv_forward_nhwc(&param->desc, x, weights->data_gpu, NULL, BCNN_ACT_NONE,
                                                    y16, ctx->workspace_gpu, ctx->workspace_bytes, sh,
                                                    stream));
        bcnn_cuda_check(bcnn_b200_bn_apply_nhwc(y16, y16, NULL, NULL, gamma, biases->data_gpu, positions,
                                                dst->c, param->activation, stream));
        return;
    }
    if (!param->bn_raw16_gpu) {
        param->bn_raw16_gpu = bcnn_b200_malloc(positions * dst->c * 2);
        if (!param->bn_raw16_gpu) bcnn_cuda_check(2 /* cudaErrorMemoryAllocation */);
    }
    const float *mean = t[node->src[3]].data_gpu, *var = t[node->src[4]].data_gpu; /* VALID */
    if (net->mode == BCNN_MODE_TRAIN) {
        bcnn_cuda_check(bcnn_b200_conv_forward_bn_stats_nhwc(
            &param->desc, x, weights->data_gpu, param->bn_raw16_gpu, ctx->workspace_gpu,
            ctx->workspace_bytes, sh, param->saved_mean.data_gpu, param->saved_variance.data_gpu,
            t[node->src[3]].data_gpu, t[node->src[4]].data_gpu, bcnn_net_nhwc_scratch(net, dst->c),
            param->reduce_scratch_gpu, stream));
        mean = param->saved_mean.data_gpu;
        var = param->saved_variance.data_gpu;
    } else {
        bcnn_cuda_check(bcnn_b200_conv_forward_nhwc(&param->desc, x, weights->data_gpu, NULL, BCNN_ACT_NONE,
                                                    param->bn_raw16_gpu, ctx->workspace_gpu,
                                                    ctx->workspace_bytes, sh, stream));
    }
    if (deferred) { /* the residual add normalises; anybody else asking gets bcnn_conv_layer_materialize */
        bcnn_resident *r = bcnn_net_res(net, node->dst[0]);
        r->data_at = BCNN_RES_DEFERRED;
        r->producer = (int)(node - net->nodes);
        return;
    }
    bcnn_cuda_check(bcnn_b200_bn_apply_nhwc(param->bn_raw16_gpu, y16, mean, var, gamma, biases->data_gpu,
                                            positions, dst->c, param->activation, stream));
}

static void conv_backward_resident(bcnn_net *net, bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *src = &t[node->src[0]], *dst = &t[node->dst[0]];
    bcnn_tensor *weights = &t[node->src[1]], *biases = &t[node->src[2]];
    void *stream = ctx->stream;
    const size_t positions = (size_t)dst->n * dst->h * dst->w;
    /* incoming gradient: this tensor's own twin, or (fused residual add) the add's masked output
     * gradient; the batch-norm backward then writes its result into the own twin, so both tensors
     * read back as in the reference (in-place results, src/layers/bcnn_conv_layer.c:516-528) */
    bcnn_resident *rd = bcnn_net_res(net, node->dst[0]);
    void *dy16;
    void *dy_in;
    if (rd->grad_alias > 0) {
        dy_in = bcnn_net_grad16_in(net, rd->grad_alias - 1);
        dy16 = bcnn_net_grad16_out(net, node->dst[0]);
        rd->grad_alias = 0;
    } else {
        dy16 = dy_in = bcnn_net_grad16_in(net, node->dst[0]);
    }
    float *scratch = bcnn_net_nhwc_scratch(net, dst->c);
    if (param->batch_norm) {
        const int train = net->mode == BCNN_MODE_TRAIN;
        const float *mean = train ? param->saved_mean.data_gpu : t[node->src[3]].data_gpu;
        const float *var = train ? param->saved_variance.data_gpu : t[node->src[4]].data_gpu;
        if (param->bn_partial_rows > 0) {   /* the residual add behind this node did the reduction pass */
            bcnn_cuda_check(bcnn_b200_bn_backward_nhwc_partials(
                param->bn_raw16_gpu, dy_in, dy16, mean, var, t[node->src[5]].data_gpu, biases->data_gpu,
                t[node->src[5]].grad_data_gpu, biases->grad_data_gpu, param->saved_mean.grad_data_gpu,
                param->saved_variance.grad_data_gpu, positions, dst->c, param->bn_partial_gpu,
                param->bn_partial_rows, stream));
            param->bn_partial_rows = 0;
        } else
        bcnn_cuda_check(bcnn_b200_bn_backward_nhwc(
            param->bn_raw16_gpu, dy_in, dy16, mean, var, t[node->src[5]].data_gpu, biases->data_gpu,
            t[node->src[5]].grad_data_gpu, biases->grad_data_gpu, param->saved_mean.grad_data_gpu,
            param->saved_variance.grad_data_gpu, positions, dst->c, param->activation, scratch, stream));
    } else {
        const void *y16 = param->activation == BCNN_ACT_NONE ? NULL : bcnn_net_data16_in(net, node->dst[0]);
        bcnn_cuda_check(bcnn_b200_actbwd_grad_bias_nhwc(biases->grad_data_gpu, dy16, y16, param->activation,
                                                        positions, dst->c, scratch, stream));
    }
    bcnn_net_grad16_modified(net, node->dst[0]);
    const void *x = conv_reads_fp32_input(param) ? (const void *)bcnn_net_data32_in(net, node->src[0])
                                                 : (const void *)bcnn_net_data16_in(net, node->src[0]);
    bcnn_cuda_check(bcnn_b200_conv_backward_weights_nhwc(&param->desc, x, dy16, weights->grad_data_gpu,
                                                         ctx->workspace_gpu, ctx->workspace_bytes,
                                                         &param->shadows, stream));
    if (src->grad_data_gpu) {
        int accumulate = bcnn_net_grad_accumulate(net, node->src[0]);
        if (ctx->reference_quirks) accumulate = 0;
        void *dx16 = accumulate ? bcnn_net_grad16_in(net, node->src[0]) : bcnn_net_grad16_out(net, node->src[0]);
        bcnn_cuda_check(bcnn_b200_conv_backward_data_nhwc(&param->desc, weights->data_gpu, dy16, dx16,
                                                          accumulate, ctx->workspace_gpu,
                                                          ctx->workspace_bytes, stream));
        bcnn_net_grad16_modified(net, node->src[0]);
    }
}

void bcnn_forward_conv_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (bcnn_net_node_is_resident(net, node)) {
        conv_forward_resident(net, node);
        return;
    }
    const int conv_math = ctx->conv_math == BCNN_B200_MATH_FP32 ? BCNN_B200_MATH_FP32 : BCNN_B200_MATH_TC;
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *src = &t[node->src[0]], *dst = &t[node->dst[0]];
    bcnn_tensor *weights = &t[node->src[1]], *biases = &t[node->src[2]];
    void *stream = ctx->stream;
    param->conv_workspace_gpu = ctx->workspace_gpu;
    const bcnn_activation act = param->activation;
    /* PReLU needs per-channel slopes: run it as a separate pass after the fused part */
    const bcnn_activation fused_act = (act == BCNN_ACT_PRELU) ? BCNN_ACT_NONE : act;
    /* TRAIN: keep the NHWC shadow of the input for this step's wgrad */
    bcnn_b200_conv_shadows *sh = NULL;
    if (net->mode == BCNN_MODE_TRAIN && conv_math == BCNN_B200_MATH_TC) {
        sh = &param->shadows;
        if (!sh->x) {
            size_t bytes = bcnn_b200_conv_x_shadow_bytes(&param->desc, conv_math);
            if (bytes) {
                sh->x = bcnn_b200_malloc(bytes);
                sh->x_bytes = sh->x ? bytes : 0; /* without storage the passes transpose again */
            }
        }
        sh->x_fmt = BCNN_B200_SHADOW_NONE;
    }

    if (!param->batch_norm) {
        bcnn_cuda_check(bcnn_b200_conv_forward_sh(&param->desc, src->data_gpu, weights->data_gpu,
                                                  biases->data_gpu, fused_act, dst->data_gpu,
                                                  ctx->workspace_gpu, ctx->workspace_bytes,
                                                  conv_math, sh, stream));
    } else {
        float *raw = param->bn_workspace_gpu ? param->bn_workspace_gpu : dst->data_gpu;
        if (net->mode == BCNN_MODE_TRAIN) {
            /* batch statistics come out of the convolution epilogue; one pass normalises */
            bcnn_tensor *run_mean = &t[node->src[3]], *run_var = &t[node->src[4]];
            bcnn_cuda_check(bcnn_b200_conv_forward_bn_stats(
                &param->desc, src->data_gpu, weights->data_gpu, raw, ctx->workspace_gpu,
                ctx->workspace_bytes, conv_math, sh, param->saved_mean.data_gpu,
                param->saved_variance.data_gpu, run_mean->data_gpu, run_var->data_gpu,
                param->reduce_scratch_gpu, stream));
            bcnn_cuda_check(bcnn_b200_bn_apply(raw, dst->data_gpu, param->saved_mean.data_gpu,
                                               param->saved_variance.data_gpu,
                                               t[node->src[5]].data_gpu, biases->data_gpu, dst->n,
                                               dst->c, dst->h * dst->w, fused_act, stream));
        } else {
            bcnn_cuda_check(bcnn_b200_conv_forward_sh(&param->desc, src->data_gpu,
                                                      weights->data_gpu, NULL, BCNN_ACT_NONE, raw,
                                                      ctx->workspace_gpu, ctx->workspace_bytes,
                                                      conv_math, sh, stream));
            bcnn_b200_forward_batchnorm(net, raw, dst, &t[node->src[3]], &t[node->src[4]],
                                       &t[node->src[5]], biases, &param->saved_mean,
                                       &param->saved_variance, param->reduce_scratch_gpu,
                                       net->mode, fused_act);
        }
    }
    if (act == BCNN_ACT_PRELU) {
        bcnn_tensor *slopes = &t[node->src[3 + 3 * param->batch_norm]];
        bcnn_cuda_check(bcnn_b200_activation_forward(dst->data_gpu, bcnn_tensor_size(dst), act,
                                                     slopes->data_gpu, dst->w * dst->h, dst->c,
                                                     stream));
    }
}

void bcnn_backward_conv_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (bcnn_net_node_is_resident(net, node)) {
        conv_backward_resident(net, node);
        return;
    }
    const int conv_math = ctx->conv_math == BCNN_B200_MATH_FP32 ? BCNN_B200_MATH_FP32 : BCNN_B200_MATH_TC;
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *src = &t[node->src[0]], *dst = &t[node->dst[0]];
    bcnn_tensor *weights = &t[node->src[1]], *biases = &t[node->src[2]];
    void *stream = ctx->stream;
    bcnn_activation act = param->activation;

    if (act == BCNN_ACT_PRELU) {
        bcnn_tensor *slopes = &t[node->src[3 + 3 * param->batch_norm]];
        bcnn_cuda_check(bcnn_b200_activation_backward(
            dst->data_gpu, dst->grad_data_gpu, bcnn_tensor_size(dst), act, slopes->data_gpu,
            slopes->grad_data_gpu, dst->w * dst->h, dst->c, stream));
        act = BCNN_ACT_NONE;
    }
    if (param->batch_norm) {
        bcnn_b200_backward_batchnorm(net, param->bn_workspace_gpu, dst->data_gpu, dst,
                                    &t[node->src[3]], &t[node->src[4]], &t[node->src[5]], biases,
                                    &param->saved_mean, &param->saved_variance,
                                    param->reduce_scratch_gpu, net->mode, act);
    } else {
        bcnn_cuda_check(bcnn_b200_actbwd_grad_bias(biases->grad_data_gpu, dst->grad_data_gpu,
                                                   dst->data_gpu, act, dst->n, dst->c,
                                                   dst->h * dst->w, param->reduce_scratch_gpu,
                                                   stream));
    }
    /* the dy shadow written by wgrad serves dgrad below; the x shadow is the forward's */
    bcnn_b200_conv_shadows *sh = &param->shadows;
    sh->dy = ctx->dy_shadow_gpu;
    sh->dy_bytes = ctx->dy_shadow_gpu ? ctx->dy_shadow_bytes : 0;
    sh->dy_fmt = BCNN_B200_SHADOW_NONE;
    bcnn_cuda_check(bcnn_b200_conv_backward_weights_sh(
        &param->desc, src->data_gpu, dst->grad_data_gpu, weights->grad_data_gpu,
        ctx->workspace_gpu, ctx->workspace_bytes, conv_math, sh, stream));
    if (src->grad_data_gpu) {
        /* reference semantics: overwrite. With the quirks off, a source read by several
         * nodes (residual branches) accumulates instead: the first backward writer of the
         * step overwrites the stale buffer, every later consumer's contribution is summed. */
        int accumulate = bcnn_net_grad_accumulate(net, node->src[0]);
        if (ctx->reference_quirks) accumulate = 0;
        bcnn_cuda_check(bcnn_b200_conv_backward_data_sh(
            &param->desc, weights->data_gpu, dst->grad_data_gpu, src->grad_data_gpu, accumulate,
            ctx->workspace_gpu, ctx->workspace_bytes, conv_math, sh, stream));
    }
    sh->dy_fmt = BCNN_B200_SHADOW_NONE;
}

void bcnn_forward_conv_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_forward_conv_layer_gpu(net, node);
}

void bcnn_backward_conv_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_backward_conv_layer_gpu(net, node);
}

/* bcnn_update_conv_layer (reference :810-855): only W and the bias / beta are stepped;
 * gamma ("scales") receives gradients that no update ever applies. */
void bcnn_update_conv_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    bcnn_optimizer_step_gpu(net, &net->tensors[node->src[1]], &net->tensors[node->src[2]],
                            &param->adam_m_gpu, &param->adam_v_gpu);
}

void bcnn_release_param_conv_layer(bcnn_node *node) {
    bcnn_conv_param *param = (bcnn_conv_param *)node->param;
    bcnn_tensor_destroy(&param->saved_mean);
    bcnn_tensor_destroy(&param->saved_variance);
    bcnn_b200_free(param->bn_workspace_gpu);
    bcnn_b200_free(param->bn_raw16_gpu);
    bcnn_b200_free(param->bn_partial_gpu);
    bcnn_b200_free(param->shadows.x);
    bcnn_b200_free(param->reduce_scratch_gpu);
    bcnn_b200_free(param->adam_m_gpu);
    bcnn_b200_free(param->adam_v_gpu);
    /* conv_workspace_gpu belongs to the net */
}
