/*
 * bcnn_learner.c -- learning-rate schedules, optimizer setters and the update loop.
 * Behaviour of jnbraun/bcnn src/bcnn_learner.c:29-65 (schedules), :67-103 (SGD with the
 * momentum kept in the gradient buffer), :167-225 (bcnn_update and the setters, including
 * the quirk that bcnn_set_adam_optimizer never switches `optimizer`).
 * The five BLAS-1 launches of bcnn_sgd_update_gpu are one fused kernel per tensor here.
 */
#include "bcnn_learner.h"

#include <math.h>

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

void bcnn_sgd_update_gpu(bcnn_net *net, float *weights, float *biases, float *weights_grad,
                         float *biases_grad, int weights_size, int biases_size, int batch_size,
                         float learning_rate, float momentum, float decay) {
    void *stream = bcnn_stream(net);
    const float step = -learning_rate / batch_size;
    const float g_scale = bcnn_net_grad_post_scale(net, momentum);
    if (biases && biases_grad)
        bcnn_cuda_check(bcnn_b200_sgd_update(biases, biases_grad, (size_t)biases_size, 0.0f, step,
                                             g_scale, stream));
    if (weights && weights_grad)
        bcnn_cuda_check(bcnn_b200_sgd_update(weights, weights_grad, (size_t)weights_size,
                                             decay * batch_size, step, g_scale, stream));
}

void bcnn_update(bcnn_net *net) {
    update_learning_rate(net);
    bcnn_dp_before_update(net);
    for (int i = 0; i < net->num_nodes; ++i) {
        bcnn_node *node = &net->nodes[i];
        if (node->update) node->update(net, node);
    }
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
