/*
 * bcnn_b200_net.h -- net-level extensions of libbcnn_b200.so beyond bcnn's public API.
 *
 * bcnn's GPU build moves data with synchronous cudaMemcpy hidden in its loader and
 * weight reader (src/bcnn_data.c:413-425, src/bcnn_net.c:1240-1299), which are out of
 * scope here; these calls are the explicit equivalents, plus the data-parallel hooks
 * (bcnn has no multi-GPU support at all) and read-only accessors the parity tests use.
 * Plain C ABI: pointers, ints, floats.
 */
#ifndef BCNN_B200_NET_H
#define BCNN_B200_NET_H

#include <bcnn/bcnn.h>
#include <bcnn_b200.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Convolution arithmetic for every conv node of `net` (BCNN_B200_MATH_*). Default: tensor cores
 * (BCNN_B200_MATH_TC); env BCNN_B200_CONV_MATH=fp32 or this call select the FP32 SIMT
 * verification path. */
BCNN_B200_API void bcnn_b200_set_conv_math(bcnn_net *net, int math);
BCNN_B200_API int bcnn_b200_get_conv_math(bcnn_net *net);
/* Reference-quirk mode (default OFF; env BCNN_B200_REFERENCE_QUIRKS=1 or this call turn it ON =
 * results identical to bcnn's CPU path, which the parity tests need):
 *   ON : the residual add touches only sample 0 of the second input (bcnn_eltwise_layer.c
 *        :119-121) and a convolution's data gradient always overwrites src.grad
 *        (bcnn_conv_layer.c:567-578), even when src feeds several nodes;
 *   OFF: batch-correct residual add, and data gradients accumulate into tensors that have
 *        more than one consumer. ResNet-style training needs OFF to be meaningful. */
BCNN_B200_API void bcnn_b200_set_reference_quirks(bcnn_net *net, int on);
BCNN_B200_API int bcnn_b200_get_reference_quirks(bcnn_net *net);
/* Solver parameters by name, as the reference's cfg reader sets them: the learner branch of
 * bcnn_net_set_param (src/bcnn_net.h:75, src/bcnn_net.c:506-553; internal but non-static there,
 * same name and meaning here). Keys: max_batches, learning_policy | decay_type
 * (sigmoid|constant|exp|inv|step|poly), optimizer (sgd|adam), step, learning_rate, beta1,
 * beta2, decay, momentum, gamma. This is the only way to reach Adam, in the reference too:
 * bcnn_set_adam_optimizer never switches the optimizer. Also the shape keys
 * input_width | width, input_height | height, input_channels | channels, batch_size | batch.
 * Augmentation keys belong to the file loader and are ignored. */
BCNN_B200_API void bcnn_net_set_param(bcnn_net *net, const char *name, const char *val);
/* CUDA-graph replay (default on; env BCNN_B200_GRAPHS=0 or on = 0 turns it off).
 *   PREDICT: the first bcnn_forward of a configuration launches its kernels one by one, the
 *     second is captured, later ones are one graph launch.
 *   TRAIN: bcnn_b200_train_step / bcnn_train_on_batch do the same with forward + backward (one
 *     graph per input buffer of the pipeline) and add the SGD update kernels while the learning
 *     rate is constant. Plain bcnn_forward / bcnn_backward calls are never captured in TRAIN
 *     mode; data-parallel nets, nets being profiled and nets with extra inputs run eagerly.
 * Graphs are rebuilt when nodes / tensors / conv math / buffers / solver scalars change.
 * Results are identical to the eager path (same kernels, same order).
 * get: 0 = off, 1 = on, 2 = on and a graph is live. */
BCNN_B200_API void bcnn_b200_set_graphs(bcnn_net *net, int on);
BCNN_B200_API int bcnn_b200_get_graphs(bcnn_net *net);
/* The CUDA stream (cudaStream_t) every kernel of this net is launched on. */
BCNN_B200_API void *bcnn_b200_get_stream(bcnn_net *net);
/* The process-current stream: what the entry points that keep the reference's net-less
 * signatures (bcnn_forward_activation_gpu, bcnn_forward_batchnorm_gpu, bcnn_cuda_*) launch on.
 * bcnn_forward / bcnn_backward / bcnn_update set it to their net's stream for the duration of the
 * loop; outside of them it is whatever was set last (initially NULL = the legacy default stream,
 * which is what the reference's CUDA path uses everywhere). */
BCNN_B200_API void *bcnn_b200_current_stream(void);
BCNN_B200_API void bcnn_b200_set_current_stream(void *stream);
/* Block the host until the net's stream (and its comm stream) are idle. */
BCNN_B200_API void bcnn_b200_sync(bcnn_net *net);

/* Host -> device copy of tensor `index` (data, and grad when both mirrors exist).
 * Equivalent of the H2D copies in bcnn_load_weights / bcnn_loader_next. Synchronous. */
BCNN_B200_API bcnn_status bcnn_b200_upload_tensor(bcnn_net *net, int index);
/* Asynchronous upload of the input tensors and the label from their pinned host
 * mirrors on the net's stream (what bcnn_loader_next does every step). Returns the
 * number of bytes queued. */
BCNN_B200_API size_t bcnn_b200_upload_inputs(bcnn_net *net);
/* Mean loss over the cost nodes, as bcnn_get_loss (src/bcnn_net.c:431-450): a
 * device -> host read of one float per cost node; synchronises the stream. */
BCNN_B200_API float bcnn_b200_get_loss(bcnn_net *net);
/* One training step = optional input upload + bcnn_forward + bcnn_backward +
 * bcnn_update, the body of bcnn_train_on_batch (src/bcnn_net.c:452-463) without the
 * file loader. Returns the loss when fetch_loss != 0 (which synchronises), else 0.
 * upload_inputs: 0 = inputs already on the device; 1 = upload the host mirrors on the
 * compute stream, then step (bcnn_loader_next's synchronous cudaMemcpy, src/bcnn_data.c:
 * 402-427, made asynchronous); 2 = input pipeline: the step consumes the batch staged by
 * the previous call (the first call stages its own) and the batch now in the host mirrors
 * is uploaded to a second set of device buffers on a copy stream while the step computes,
 * to be consumed by the next call -- fill the mirrors with batch i+1 before step i. With
 * fetch_loss the mirrors may be refilled when the call returns; without it after
 * bcnn_b200_sync. */
BCNN_B200_API float bcnn_b200_train_step(bcnn_net *net, int upload_inputs, int fetch_loss);
/* Prologue of the input pipeline: stage the batch now in the host mirrors (batch 0) for the
 * first bcnn_b200_train_step(net, 2, ...). Returns the bytes uploaded; the mirrors may be
 * refilled on return. */
BCNN_B200_API size_t bcnn_b200_prefetch_inputs(bcnn_net *net);

/* Per-node CUDA-event timers: when enabled, bcnn_forward / bcnn_backward bracket every
 * node with events on the net's stream; read the last step's durations per node. */
BCNN_B200_API void bcnn_b200_profile(bcnn_net *net, int enable);
BCNN_B200_API int bcnn_b200_profile_node_ms(bcnn_net *net, int node, float *fwd_ms,
                                            float *bwd_ms);

/* ---- introspection for tests ---- */
BCNN_B200_API int bcnn_b200_num_nodes(bcnn_net *net);
BCNN_B200_API int bcnn_b200_num_tensors(bcnn_net *net);
BCNN_B200_API int bcnn_b200_node_type(bcnn_net *net, int node);
BCNN_B200_API int bcnn_b200_node_src(bcnn_net *net, int node, int i); /* -1 if out of range */
BCNN_B200_API int bcnn_b200_node_dst(bcnn_net *net, int node, int i);
/* dims[4] = {n, c, h, w} of tensor `index` without touching its buffers (bcnn_get_tensor_by_index
 * refreshes the host copies, a full device -> host transfer). Returns 0, or -1 for a bad index. */
BCNN_B200_API int bcnn_b200_tensor_dims(bcnn_net *net, int index, int *dims);
/* Copies the max-pool argmax of node `node` (param->indexes_gpu) into host_out
 * (count = size of the node's dst tensor). Returns the count or -1. */
BCNN_B200_API int bcnn_b200_maxpool_indexes(bcnn_net *net, int node, int *host_out);
/* Copies saved_mean / saved_variance of a conv(+BN) or batchnorm node. Returns the
 * channel count or -1. */
BCNN_B200_API int bcnn_b200_bn_saved_stats(bcnn_net *net, int node, float *mean_out,
                                           float *var_out);

/* Cross-check of the yolo loss kernels: evaluates the detection loss of yolo node `node` with
 * the reference's host loops on the head tensor now in its dst buffer (device -> host, loss,
 * gradient -> device) and returns the cost; -1 when `node` is not a yolo node of a net with
 * gradients. bcnn_forward never takes this route: in TRAIN mode the loss runs on the device. */
BCNN_B200_API float bcnn_b200_yolo_loss_on_host(bcnn_net *net, int node);

/* ---- data parallelism (one process per GPU, NCCL all-reduce of weight grads) ---- */
#define BCNN_B200_DP_ID_BYTES 128
/* Rank 0 creates the NCCL unique id; the launcher broadcasts the 128 bytes to every
 * rank (torch.distributed / MPI / a file -- plumbing, not part of this library). */
BCNN_B200_API int bcnn_b200_dp_get_unique_id(char id[BCNN_B200_DP_ID_BYTES]);
/* Join the communicator. After this, bcnn_backward all-reduces each node's parameter
 * gradients on a dedicated comm stream as soon as that node's backward is queued, and
 * bcnn_update waits for them, uses batch_size * world as divisor and keeps the
 * momentum-in-gradient-buffer invariant. */
BCNN_B200_API int bcnn_b200_dp_init(bcnn_net *net, int rank, int world,
                                    const char id[BCNN_B200_DP_ID_BYTES]);
BCNN_B200_API void bcnn_b200_dp_shutdown(bcnn_net *net);
BCNN_B200_API int bcnn_b200_dp_world(bcnn_net *net);
/* Bytes all-reduced per step (sum over the parameter gradient tensors an update consumes: weights
 * and bias / beta; batch-norm scale gradients are never applied, SURVEY.md H6, and stay local). */
BCNN_B200_API size_t bcnn_b200_dp_bytes_per_step(bcnn_net *net);
/* NCCL launches per step: gradients travel in ~25 MB buckets, one aggregated group each. */
BCNN_B200_API int bcnn_b200_dp_groups_per_step(bcnn_net *net);
/* Milliseconds per all-reduce of the whole gradient set with nothing else running (bench.py's
 * bus-bandwidth figure). Sums the gradient buffers repeatedly: call it after the measurements. */
BCNN_B200_API float bcnn_b200_dp_allreduce_probe_ms(bcnn_net *net, int iters);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_B200_NET_H */
