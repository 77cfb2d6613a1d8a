/*
 * bcnn_batchnorm_layer.c -- batchnorm node on the B200 kernels.
 *
 * Same node layout as jnbraun/bcnn src/layers/bcnn_batchnorm_layer.c:36-145: src[0] = x,
 * src[1] = running mean, src[2] = running var, src[3] = gamma ("scales", filled with 1),
 * src[4] = beta ("<src>_b"); saved_mean / saved_variance live in the param. The arithmetic
 * is the reference CPU path's (:147-332), executed as one statistics reduction + one
 * fused normalise/scale/shift(/activation) pass forward, and one reduction + one
 * elementwise pass backward, instead of the reference's ~7 and ~6 passes; x_hat and the
 * input copy are not materialised (the source tensor is still intact at backward time).
 */
#include "bcnn_batchnorm_layer.h"

#include <bcnn_b200_net.h>

#include "bcnn_tensor.h"

bcnn_status bcnn_add_batchnorm_layer(bcnn_net *net, const char *src_id, const char *dst_id) {
    bcnn_node node = {0};
    BCNN_CHECK_AND_LOG(net->log_ctx, net->num_nodes >= 1, BCNN_INVALID_PARAMETER,
                       "Batchnorm layer can't be the first layer of the network\n");
    int src = bcnn_net_find_src(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Batchnorm layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const int n = net->tensors[src].n, c = net->tensors[src].c, h = net->tensors[src].h,
              w = net->tensors[src].w;
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, c, h, w, dst_id));

    node.type = BCNN_LAYER_BATCHNORM;
    node.param_size = sizeof(bcnn_batchnorm_param);
    bcnn_batchnorm_param *param = (bcnn_batchnorm_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    node.forward = bcnn_forward_batchnorm_layer;
    node.backward = bcnn_backward_batchnorm_layer;
    node.release_param = bcnn_release_param_batchnorm_layer;

    char name[320];
    snprintf(name, sizeof(name), "%s_sav_mean", src_id);
    bcnn_tensor_create(&param->saved_mean, 1, 1, 1, c, 1, name, net->mode);
    snprintf(name, sizeof(name), "%s_sav_var", src_id);
    bcnn_tensor_create(&param->saved_variance, 1, 1, 1, c, 1, name, net->mode);
    bcnn_tensor_filler ones = {.value = 1.0f, .type = BCNN_FILLER_FIXED};
    BCNN_CHECK_STATUS(bcnn_net_add_param_tensor(net, &node, 1, 1, 1, c, 0, src_id, "_run_mean", NULL));
    BCNN_CHECK_STATUS(bcnn_net_add_param_tensor(net, &node, 1, 1, 1, c, 0, src_id, "_run_var", NULL));
    BCNN_CHECK_STATUS(bcnn_net_add_param_tensor(net, &node, 1, 1, 1, c, 1, src_id, "_scales", &ones));
    BCNN_CHECK_STATUS(bcnn_net_add_param_tensor(net, &node, 1, 1, 1, c, 1, src_id, "_b", NULL));
    param->reduce_scratch_gpu =
        (float *)bcnn_b200_malloc(bcnn_b200_bn_scratch_floats(c) * sizeof(float));
    BCNN_CHECK(param->reduce_scratch_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);

    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx, "[Batchnorm] %-8s (%4d x%4d x%4d) -> %-8s (%4d x%4d x%4d)\n", src_id,
              w, h, c, dst_id, w, h, c);
    return BCNN_SUCCESS;
}

void bcnn_b200_forward_batchnorm(bcnn_net *net, const float *x_gpu, bcnn_tensor *dst,
                                 bcnn_tensor *bn_mean, bcnn_tensor *bn_var, bcnn_tensor *bn_scales,
                                 bcnn_tensor *biases, bcnn_tensor *saved_mean,
                                 bcnn_tensor *saved_var, float *scratch_gpu, bcnn_mode mode,
                                 bcnn_activation act) {
    void *stream = bcnn_stream(net);
    const int n = dst->n, c = dst->c, hw = dst->h * dst->w;
    if (mode == BCNN_MODE_PREDICT) { /* statistics were folded into scales / biases */
        bcnn_cuda_check(bcnn_b200_scale_bias(x_gpu, dst->data_gpu, bn_scales->data_gpu,
                                             biases->data_gpu, n, c, hw, act, stream));
        return;
    }
    const float *mean = bn_mean->data_gpu, *var = bn_var->data_gpu; /* VALID: running stats */
    if (mode == BCNN_MODE_TRAIN) {
        bcnn_cuda_check(bcnn_b200_bn_stats(x_gpu, n, c, hw, saved_mean->data_gpu,
                                           saved_var->data_gpu, bn_mean->data_gpu,
                                           bn_var->data_gpu, scratch_gpu, stream));
        mean = saved_mean->data_gpu;
        var = saved_var->data_gpu;
    }
    bcnn_cuda_check(bcnn_b200_bn_apply(x_gpu, dst->data_gpu, mean, var, bn_scales->data_gpu,
                                       biases->data_gpu, n, c, hw, act, stream));
}

void bcnn_b200_backward_batchnorm(bcnn_net *net, const float *x_gpu, const float *y_gpu,
                                  bcnn_tensor *dst, bcnn_tensor *bn_mean, bcnn_tensor *bn_var,
                                  bcnn_tensor *bn_scales, bcnn_tensor *biases,
                                  bcnn_tensor *saved_mean, bcnn_tensor *saved_var,
                                  float *scratch_gpu, bcnn_mode mode, bcnn_activation act) {
    const int n = dst->n, c = dst->c, hw = dst->h * dst->w;
    /* outside TRAIN the reference differentiates through the running statistics
     * (bcnn_batchnorm_layer.c:308-311) */
    const float *mean = (mode == BCNN_MODE_TRAIN) ? saved_mean->data_gpu : bn_mean->data_gpu;
    const float *var = (mode == BCNN_MODE_TRAIN) ? saved_var->data_gpu : bn_var->data_gpu;
    bcnn_cuda_check(bcnn_b200_bn_backward(
        x_gpu, y_gpu, dst->grad_data_gpu, dst->grad_data_gpu, mean, var, bn_scales->data_gpu,
        biases->data_gpu, bn_scales->grad_data_gpu, biases->grad_data_gpu, saved_mean->grad_data_gpu,
        saved_var->grad_data_gpu, n, c, hw, act, scratch_gpu, bcnn_stream(net)));
}

/* ---- reference-signature entry points (see the header) ---- */
static float *compat_scratch(int c) {
    static float *buf = NULL;
    static size_t floats = 0;
    const size_t need = bcnn_b200_bn_scratch_floats(c);
    if (need > floats) {
        bcnn_b200_stream_sync(bcnn_b200_current_stream());
        bcnn_b200_free(buf);
        buf = (float *)bcnn_b200_malloc(need * sizeof(float));
        floats = buf ? need : 0;
    }
    return buf;
}

void bcnn_forward_batchnorm_gpu(bcnn_tensor *src, bcnn_tensor *dst, bcnn_tensor *bn_mean,
                                bcnn_tensor *bn_var, bcnn_tensor *bn_scales, bcnn_tensor *biases,
                                bcnn_tensor *saved_mean, bcnn_tensor *saved_var, float *x_norm_gpu,
                                float *workspace_gpu, bcnn_mode mode) {
    (void)x_norm_gpu;
    void *stream = bcnn_b200_current_stream();
    const int n = dst->n, c = dst->c, hw = dst->h * dst->w;
    const float *x = src->data_gpu;
    if (workspace_gpu && workspace_gpu != src->data_gpu && mode == BCNN_MODE_TRAIN) {
        bcnn_cuda_check(bcnn_b200_memcpy_d2d(workspace_gpu, src->data_gpu,
                                             (size_t)n * c * hw * sizeof(float), stream));
        x = workspace_gpu;
    }
    if (mode == BCNN_MODE_PREDICT) {
        bcnn_cuda_check(bcnn_b200_scale_bias(x, dst->data_gpu, bn_scales->data_gpu, biases->data_gpu,
                                             n, c, hw, BCNN_ACT_NONE, stream));
        return;
    }
    const float *mean = bn_mean->data_gpu, *var = bn_var->data_gpu;
    if (mode == BCNN_MODE_TRAIN) {
        float *scratch = compat_scratch(c);
        if (!scratch) bcnn_cuda_check(2 /* cudaErrorMemoryAllocation */);
        bcnn_cuda_check(bcnn_b200_bn_stats(x, n, c, hw, saved_mean->data_gpu, saved_var->data_gpu,
                                           bn_mean->data_gpu, bn_var->data_gpu, scratch, stream));
        mean = saved_mean->data_gpu;
        var = saved_var->data_gpu;
    }
    bcnn_cuda_check(bcnn_b200_bn_apply(x, dst->data_gpu, mean, var, bn_scales->data_gpu,
                                       biases->data_gpu, n, c, hw, BCNN_ACT_NONE, stream));
}

void bcnn_backward_batchnorm_gpu(bcnn_tensor *src, bcnn_tensor *dst, bcnn_tensor *bn_mean,
                                 bcnn_tensor *bn_var, bcnn_tensor *bn_scales, bcnn_tensor *biases,
                                 bcnn_tensor *saved_mean, bcnn_tensor *saved_var, float *x_norm_gpu,
                                 float *workspace_gpu, bcnn_mode mode) {
    (void)x_norm_gpu;
    void *stream = bcnn_b200_current_stream();
    const int n = dst->n, c = dst->c, hw = dst->h * dst->w;
    const float *x = workspace_gpu ? workspace_gpu : src->data_gpu;
    const float *mean = (mode == BCNN_MODE_TRAIN) ? saved_mean->data_gpu : bn_mean->data_gpu;
    const float *var = (mode == BCNN_MODE_TRAIN) ? saved_var->data_gpu : bn_var->data_gpu;
    float *scratch = compat_scratch(c);
    if (!scratch) bcnn_cuda_check(2 /* cudaErrorMemoryAllocation */);
    bcnn_cuda_check(bcnn_b200_bn_backward(
        x, NULL, dst->grad_data_gpu, dst->grad_data_gpu, mean, var, bn_scales->data_gpu, NULL,
        bn_scales->grad_data_gpu, biases->grad_data_gpu, saved_mean->grad_data_gpu,
        saved_var->grad_data_gpu, n, c, hw, BCNN_ACT_NONE, scratch, stream));
    if (src->grad_data_gpu && src->grad_data_gpu != dst->grad_data_gpu)
        bcnn_cuda_check(bcnn_b200_memcpy_d2d(src->grad_data_gpu, dst->grad_data_gpu,
                                             (size_t)n * c * hw * sizeof(float), stream));
}

void bcnn_forward_batchnorm_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_batchnorm_param *param = (bcnn_batchnorm_param *)node->param;
    bcnn_tensor *t = net->tensors;
    bcnn_b200_forward_batchnorm(net, t[node->src[0]].data_gpu, &t[node->dst[0]], &t[node->src[1]],
                               &t[node->src[2]], &t[node->src[3]], &t[node->src[4]],
                               &param->saved_mean, &param->saved_variance,
                               param->reduce_scratch_gpu, net->mode, BCNN_ACT_NONE);
}

void bcnn_backward_batchnorm_layer_gpu(bcnn_net *net, bcnn_node *node) {
    bcnn_batchnorm_param *param = (bcnn_batchnorm_param *)node->param;
    bcnn_tensor *t = net->tensors;
    bcnn_tensor *src = &t[node->src[0]], *dst = &t[node->dst[0]];
    bcnn_b200_backward_batchnorm(net, src->data_gpu, NULL, dst, &t[node->src[1]], &t[node->src[2]],
                                &t[node->src[3]], &t[node->src[4]], &param->saved_mean,
                                &param->saved_variance, param->reduce_scratch_gpu, net->mode,
                                BCNN_ACT_NONE);
    if (src->grad_data_gpu) { /* overwrite, as the reference's bcnn_copy_f32 (:327-330) */
        (void)bcnn_net_grad_accumulate(net, node->src[0]);
        bcnn_cuda_check(bcnn_b200_memcpy_d2d(src->grad_data_gpu, dst->grad_data_gpu,
                                             (size_t)bcnn_tensor_size(dst) * sizeof(float),
                                             bcnn_stream(net)));
    }
}

void bcnn_forward_batchnorm_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_forward_batchnorm_layer_gpu(net, node);
}

void bcnn_backward_batchnorm_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_backward_batchnorm_layer_gpu(net, node);
}

void bcnn_release_param_batchnorm_layer(bcnn_node *node) {
    bcnn_batchnorm_param *param = (bcnn_batchnorm_param *)node->param;
    bcnn_tensor_destroy(&param->saved_mean);
    bcnn_tensor_destroy(&param->saved_variance);
    bcnn_b200_free(param->reduce_scratch_gpu);
}

#include <bcnn_b200_net.h>

#include "bcnn_conv_layer.h"

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
    void *stream = bcnn_stream(net);
    bcnn_cuda_check(bcnn_b200_memcpy_d2h(mean_out, mean->data_gpu, (size_t)c * sizeof(float), stream));
    bcnn_cuda_check(bcnn_b200_memcpy_d2h(var_out, var->data_gpu, (size_t)c * sizeof(float), stream));
    bcnn_cuda_check(bcnn_b200_stream_sync(stream));
    return c;
}
