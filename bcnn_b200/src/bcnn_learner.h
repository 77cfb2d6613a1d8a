/*
 * bcnn_learner.h -- optimizer state; layout of jnbraun/bcnn src/bcnn_learner.h:29-44.
 */
#ifndef BCNN_LEARNER_H
#define BCNN_LEARNER_H

#include <bcnn/bcnn.h>

typedef struct {
    int step;
    int seen;        /* samples seen so far */
    int max_batches;
    float momentum;
    float decay;
    float base_learning_rate;
    float learning_rate;
    float gamma;
    float scale;
    float power;
    float beta1;
    float beta2;
    bcnn_optimizer optimizer;
    bcnn_lr_decay decay_type;
} bcnn_learner;

#ifdef __cplusplus
extern "C" {
#endif
/* Device SGD-momentum step for one (weights, biases) pair; same argument meaning as
 * reference bcnn_sgd_update_gpu (src/bcnn_learner.c:86-103), plus the net, from which the
 * stream and the data-parallel scaling are taken. */
void bcnn_sgd_update_gpu(bcnn_net *net, float *weights, float *biases, float *weights_grad,
                         float *biases_grad, int weights_size, int biases_size, int batch_size,
                         float learning_rate, float momentum, float decay);
/* bcnn_update = bcnn_update_schedule (host: samples seen, learning-rate policy) + the
 * data-parallel join + bcnn_update_nodes (the per-node update kernels). */
void bcnn_update_schedule(bcnn_net *net);
void bcnn_update_nodes(bcnn_net *net);
/* Device Adam step; argument meaning of reference bcnn_adam_update_gpu (src/bcnn_learner.c:
 * 134-164) plus the net. */
void bcnn_adam_update_gpu(bcnn_net *net, float *weights, float *biases, float *weights_grad,
                          float *biases_grad, float *adam_m, float *adam_v, int weights_size,
                          int biases_size, int batch_size, int iter, float beta1, float beta2,
                          float learning_rate, float momentum, float decay);
/* SGD or Adam step of one (weights, biases) pair, by net->learner->optimizer; the Adam moment
 * buffers are created on first use. */
void bcnn_optimizer_step_gpu(bcnn_net *net, bcnn_tensor *weights, bcnn_tensor *biases,
                             float **adam_m_gpu, float **adam_v_gpu);
#ifdef __cplusplus
}
#endif
#endif /* BCNN_LEARNER_H */
