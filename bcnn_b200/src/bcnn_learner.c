/*
 * bcnn_learner.c -- learning-rate schedules, optimizer setters and the update loop.
 * Behaviour of jnbraun/bcnn src/bcnn_learner.c:29-65 (schedules), :67-103 (SGD with the
 * momentum kept in the gradient buffer), :106-164 (Adam), :167-225 (bcnn_update and the
 * setters, including the quirk that bcnn_set_adam_optimizer never switches `optimizer`,
 * Adam is reached through bcnn_net_set_param("optimizer", "adam") only).
 * The five BLAS-1 launches of bcnn_sgd_update_gpu and the nine of bcnn_adam_update_gpu are
 * one fused kernel per tensor here.
 */
#include "bcnn_learner.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "bcnn_dp.h"
#include "bcnn_net.h"

static void update_learning_rate(bcnn_net *net) {
    bcnn_learner *ln = net->learner;
    ln->seen += net->batch_size;
    int iter = ln->seen / net->batch_size;
    switch (ln->decay_type) {
        case BCNN_LR_DECAY_STEP:
            ln->learning_rate = ln->base_learning_rate * (float)pow(ln->scale, iter / ln->step);
            break;
        case BCNN_LR_DECAY_INV:
            ln->learning_rate =
                ln->base_learning_rate * (float)pow(1.0f + ln->gamma * iter, -ln->power);
            break;
        case BCNN_LR_DECAY_EXP:
            ln->learning_rate = ln->base_learning_rate * (float)pow(ln->gamma, iter);
            break;
        case BCNN_LR_DECAY_POLY:
            ln->learning_rate = ln->base_learning_rate *
                                (float)pow(1 - (float)iter / ln->max_batches, ln->power);
            break;
        case BCNN_LR_DECAY_SIGMOID:
            ln->learning_rate =
                ln->base_learning_rate *
                (1.0f / (1.0f + (float)exp(ln->gamma * (iter - ln->step))));
            break;
        case BCNN_LR_DECAY_CONSTANT:
        default:
            break;
    }
}

/* Launch what bcnn_update_nodes has collected so far. */
static void sgd_batch_flush(bcnn_net *net) {
    bcnn_b200_sgd_batch *b = bcnn_ctx(net)->sgd_batch;
    if (b && b->count > 0) {
        bcnn_cuda_check(bcnn_b200_sgd_update_multi(b, bcnn_stream(net)));
        b->count = 0;
        b->first_block[0] = 0;
    }
}

/* One SGD pass over a tensor: joins the step's batch inside bcnn_update_nodes, its own launch
 * otherwise (or when the tensor does not fit the batch's common scalars / alignment). */
static void sgd_pass(bcnn_net *net, float *w, float *g, int n, float wd_scale, float step,
                     float g_scale) {
    bcnn_b200_sgd_batch *b = bcnn_ctx(net)->sgd_batch;
    if (n <= 0) return;
    if (!b || (((uintptr_t)w | (uintptr_t)g) & 15) != 0) {
        bcnn_cuda_check(bcnn_b200_sgd_update(w, g, (size_t)n, wd_scale, step, g_scale, bcnn_stream(net)));
        return;
    }
    if (b->count > 0 && (b->step != step || b->g_scale != g_scale)) sgd_batch_flush(net);
    if (b->count == BCNN_B200_SGD_MULTI_MAX) sgd_batch_flush(net);
    const int i = b->count++;
    b->w[i] = w;
    b->g[i] = g;
    b->n[i] = (unsigned int)n;
    b->wd_scale[i] = wd_scale;
    b->step = step;
    b->g_scale = g_scale;
    b->first_block[i + 1] = b->first_block[i] + ((unsigned int)n + 4095u) / 4096u;
}

void bcnn_sgd_update_gpu(bcnn_net *net, float *weights, float *biases, float *weights_grad,
                         float *biases_grad, int weights_size, int biases_size, int batch_size,
                         float learning_rate, float momentum, float decay) {
    const float step = -learning_rate / batch_size;
    const float g_scale = bcnn_net_grad_post_scale(net, momentum);
    if (biases && biases_grad) sgd_pass(net, biases, biases_grad, biases_size, 0.0f, step, g_scale);
    if (weights && weights_grad)
        sgd_pass(net, weights, weights_grad, weights_size, decay * batch_size, step, g_scale);
}

/* Host side of bcnn_adam_update_gpu (reference :134-164). The bias branch is the SGD bias
 * branch (axpy + scal by `momentum`); the weight branch is one fused kernel. `iter` is what the
 * layers pass: learner->seen, i.e. samples, not steps. */
void bcnn_adam_update_gpu(bcnn_net *net, float *weights, float *biases, float *weights_grad,
                          float *biases_grad, float *adam_m, float *adam_v, int weights_size,
                          int biases_size, int batch_size, int iter, float beta1, float beta2,
                          float learning_rate, float momentum, float decay) {
    const float mu_correction = sqrtf(1.0f - powf(beta2, (float)iter + 1)) /
                                (1.0f - powf(beta1, (float)iter + 1));
    if (biases && biases_grad)
        bcnn_sgd_update_gpu(net, NULL, biases, NULL, biases_grad, 0, biases_size, batch_size,
                            learning_rate, momentum, 0.0f);
    if (weights && weights_grad && adam_m && adam_v) {
        const float alpha = -learning_rate / batch_size * mu_correction;
        bcnn_cuda_check(bcnn_b200_adam_update(weights, weights_grad, adam_m, adam_v,
                                              (size_t)weights_size, decay * batch_size, beta1,
                                              beta2, alpha, bcnn_stream(net)));
    }
}

static float *zeroed_device_floats(bcnn_net *net, size_t n) {
    float *p = (float *)bcnn_b200_malloc(n * sizeof(float));
    if (p) bcnn_cuda_check(bcnn_b200_fill_f32(p, n, 0.0f, bcnn_stream(net)));
    return p;
}

/* The optimizer switch every parametrised layer's update runs (reference
 * bcnn_conv_layer.c:810-855, bcnn_depthwise_conv_layer.c:563-608, bcnn_fc_layer.c:301-346).
 * The reference allocates the Adam moments when the layer is created and crashes on NULL
 * moments if the optimizer is switched afterwards; here they appear, zeroed, on the first
 * Adam step, which gives the same numbers in the first case and working code in the second. */
void bcnn_optimizer_step_gpu(bcnn_net *net, bcnn_tensor *weights, bcnn_tensor *biases,
                             float **adam_m_gpu, float **adam_v_gpu) {
    bcnn_learner *ln = net->learner;
    const int w_sz = bcnn_tensor_size(weights), b_sz = bcnn_tensor_size(biases);
    const int batch = bcnn_net_global_batch(net);
    switch (ln->optimizer) {
        case BCNN_OPTIM_ADAM:
            if (!*adam_m_gpu) *adam_m_gpu = zeroed_device_floats(net, (size_t)w_sz);
            if (!*adam_v_gpu) *adam_v_gpu = zeroed_device_floats(net, (size_t)w_sz);
            if (!*adam_m_gpu || !*adam_v_gpu) bcnn_cuda_check(2 /* cudaErrorMemoryAllocation */);
            bcnn_adam_update_gpu(net, weights->data_gpu, biases->data_gpu, weights->grad_data_gpu,
                                 biases->grad_data_gpu, *adam_m_gpu, *adam_v_gpu, w_sz, b_sz,
                                 batch, ln->seen, ln->beta1, ln->beta2, ln->learning_rate,
                                 ln->momentum, ln->decay);
            break;
        case BCNN_OPTIM_SGD:
            bcnn_sgd_update_gpu(net, weights->data_gpu, biases->data_gpu, weights->grad_data_gpu,
                                biases->grad_data_gpu, w_sz, b_sz, batch, ln->learning_rate,
                                ln->momentum, ln->decay);
            break;
        default:
            break;
    }
}

/* The two halves of bcnn_update (reference :167-175): the host bookkeeping (samples seen,
 * learning-rate schedule) and the per-node update kernels. The step graph of
 * bcnn_b200_train_step records the kernels and replays them; the bookkeeping runs every step. */
void bcnn_update_schedule(bcnn_net *net) { update_learning_rate(net); }

void bcnn_update_nodes(bcnn_net *net) {
    /* the SGD passes of the step travel together (BCNN_B200_SGD_MULTI=0: one launch per tensor).
     * Tensors are distinct buffers and every pass is element-wise, so the order of the passes --
     * and of Adam's weight kernels among them -- does not matter. */
    static int multi = -1;
    if (multi < 0) {
        const char *e = getenv("BCNN_B200_SGD_MULTI");
        multi = (e && e[0] == '0') ? 0 : 1;
    }
    bcnn_b200_sgd_batch *batch = multi ? (bcnn_b200_sgd_batch *)calloc(1, sizeof(*batch)) : NULL;
    bcnn_ctx(net)->sgd_batch = batch;
    for (int i = 0; i < net->num_nodes; ++i) {
        bcnn_node *node = &net->nodes[i];
        if (node->update) node->update(net, node);
    }
    sgd_batch_flush(net);
    bcnn_ctx(net)->sgd_batch = NULL;
    free(batch);
}

void bcnn_update(bcnn_net *net) {
    bcnn_update_schedule(net);
    bcnn_dp_before_update(net);
    bcnn_update_nodes(net);
}

static bcnn_learner *learner_of(bcnn_net *net) {
    if (!net->learner) net->learner = (bcnn_learner *)calloc(1, sizeof(bcnn_learner));
    return net->learner;
}

void bcnn_set_learning_rate_policy(bcnn_net *net, bcnn_lr_decay decay_type, float gamma,
                                   float scale, float power, int max_batches, int step) {
    bcnn_learner *ln = learner_of(net);
    ln->decay_type = decay_type;
    ln->gamma = gamma;
    ln->scale = scale;
    ln->power = power;
    ln->max_batches = max_batches;
    ln->step = step;
}

void bcnn_set_adam_optimizer(bcnn_net *net, float learning_rate, float beta1, float beta2) {
    bcnn_learner *ln = learner_of(net);
    ln->base_learning_rate = ln->learning_rate = learning_rate;
    ln->beta1 = beta1;
    ln->beta2 = beta2;
    ln->momentum = 0.9f; /* `optimizer` is left untouched, exactly like the reference */
}

void bcnn_set_sgd_optimizer(bcnn_net *net, float learning_rate, float momentum) {
    bcnn_learner *ln = learner_of(net);
    ln->base_learning_rate = ln->learning_rate = learning_rate;
    ln->momentum = momentum;
}

void bcnn_set_weight_regularizer(bcnn_net *net, float weight_decay) {
    learner_of(net)->decay = weight_decay;
}

/* The shape and learner keys of the reference's bcnn_net_set_param (src/bcnn_net.c:506-553);
 * like the reference, solver keys are ignored while the net has no learner (PREDICT) and the
 * data-augmentation keys (no loader here) always. */
void bcnn_net_set_param(bcnn_net *net, const char *name, const char *val) {
    static const struct { const char *text; bcnn_lr_decay decay; } policies[] = {
        {"sigmoid", BCNN_LR_DECAY_SIGMOID}, {"constant", BCNN_LR_DECAY_CONSTANT},
        {"exp", BCNN_LR_DECAY_EXP},         {"inv", BCNN_LR_DECAY_INV},
        {"step", BCNN_LR_DECAY_STEP},       {"poly", BCNN_LR_DECAY_POLY}};
    if (!net || !name || !val) return;
    /* input shape and batch (:507-518); these do not need a learner */
    if (!strcmp(name, "input_width") || !strcmp(name, "width")) {
        net->tensors[0].w = atoi(val);
        return;
    } else if (!strcmp(name, "input_height") || !strcmp(name, "height")) {
        net->tensors[0].h = atoi(val);
        return;
    } else if (!strcmp(name, "input_channels") || !strcmp(name, "channels")) {
        net->tensors[0].c = atoi(val);
        return;
    } else if (!strcmp(name, "batch_size") || !strcmp(name, "batch")) {
        net->batch_size = net->tensors[0].n = atoi(val);
        return;
    }
    bcnn_learner *ln = net->learner;
    if (!ln) return;
    if (!strcmp(name, "learning_policy") || !strcmp(name, "decay_type")) {
        ln->decay_type = BCNN_LR_DECAY_CONSTANT; /* unknown text falls back to constant */
        for (size_t i = 0; i < sizeof(policies) / sizeof(policies[0]); ++i)
            if (!strcmp(val, policies[i].text)) ln->decay_type = policies[i].decay;
    } else if (!strcmp(name, "optimizer")) {
        if (!strcmp(val, "sgd")) ln->optimizer = BCNN_OPTIM_SGD;
        if (!strcmp(val, "adam")) ln->optimizer = BCNN_OPTIM_ADAM;
    } else if (!strcmp(name, "max_batches")) {
        ln->max_batches = atoi(val);
    } else if (!strcmp(name, "step")) {
        ln->step = atoi(val);
    } else if (!strcmp(name, "learning_rate")) {
        ln->base_learning_rate = ln->learning_rate = (float)atof(val);
    } else if (!strcmp(name, "beta1")) {
        ln->beta1 = (float)atof(val);
    } else if (!strcmp(name, "beta2")) {
        ln->beta2 = (float)atof(val);
    } else if (!strcmp(name, "decay")) {
        ln->decay = (float)atof(val);
    } else if (!strcmp(name, "momentum")) {
        ln->momentum = (float)atof(val);
    } else if (!strcmp(name, "gamma")) {
        ln->gamma = (float)atof(val);
    }
}
