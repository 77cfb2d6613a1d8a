/*
 * bcnn_net.h -- the net object. Field order of the reference's CUDA flavour
 * (jnbraun/bcnn src/bcnn_net.h:34-67) so code that peeks at net->tensors[] /
 * net->nodes[] keeps working; the B200 state hangs off cuda_ctx.
 */
#ifndef BCNN_NET_H
#define BCNN_NET_H

#include <bcnn/bcnn.h>

#include "bcnn_learner.h"
#include "bcnn_node.h"
#include "bcnn_tensor.h"
#include "bcnn_utils.h"

#ifdef __cplusplus
extern "C" {
#endif

struct bcnn_dp_state; /* bcnn_dp.c */

typedef struct bcnn_cuda_context {
    int workspace_size;   /* floats; kept for source compatibility */
    float *workspace_gpu; /* conv workspace shared by all conv nodes */
    /* --- B200 extensions --- */
    size_t workspace_bytes;   /* max over conv nodes, sized at construction */
    /* input pipeline (bcnn_b200_train_step, upload_inputs == 2): the next batch travels to a
     * second set of device buffers on copy_stream while the current step computes */
    void *copy_stream;
    void *evt_uploaded;       /* copy -> compute: staged batch is on the device */
    void *evt_consumed;       /* compute -> copy: last step reading the staging buffers is done */
    float **stage_gpu;        /* per uploaded tensor (inputs[], then the label) */
    int stage_valid;
    void *dy_shadow_gpu;      /* NHWC shadow of the output gradient of the conv node running
                                 backward (shared by its wgrad and dgrad), max over nodes */
    size_t dy_shadow_bytes;
    void *stream;             /* compute stream (cudaStream_t) */
    int conv_math;            /* BCNN_B200_MATH_* */
    int conv_math_explicit;   /* chosen by the caller / environment: bcnn_compile_net keeps it */
    int reference_quirks;     /* see bcnn_b200_set_reference_quirks; default 1 */
    struct bcnn_dp_state *dp; /* NULL unless bcnn_b200_dp_init succeeded */
    /* bcnn_update_nodes collects the SGD passes of all parameter tensors of a step here and launches
     * them together (bcnn_b200_sgd_update_multi); NULL outside bcnn_update_nodes */
    struct bcnn_b200_sgd_batch *sgd_batch;
    /* Packed weight images of every resident convolution node (fprop, dgrad, the classes of strided
     * dgrads), kept here and re-packed by ONE launch at the start of every TRAIN-mode forward pass; the
     * kernel library finds them through its (weights, pass) registry while packs_on. Built at the end
     * of the first eager TRAIN forward (when every node has decided whether it is resident), dropped by
     * bcnn_compile_net / math changes / bcnn_end_net. packs_state: 0 not built, 1 ready, -1 off
     * (BCNN_B200_PACK_TABLE=0 or nothing to pack). */
    int packs_state;
    int packs_jobs;
    unsigned int packs_grid;
    void *packs_jobs_host, *packs_jobs_gpu, *packs_images_gpu;
    int capturing; /* a CUDA-graph capture is open on the compute stream: no allocations */
    /* per-node CUDA-event timers (the reference only has commented-out bh_timer calls in
     * bcnn_forward, src/bcnn_net.c:416-420); 4 events per node: fwd begin/end, bwd begin/end */
    int profile;
    int profile_nodes;
    void **profile_events;
    /* Lazy gradient reset. bcnn_reset_gradients (reference src/bcnn_net.c:361-375) zero-fills
     * every output gradient before each node's forward so that backward writers can `+=`.
     * Here bcnn_forward only marks those buffers stale; the first backward writer of a step
     * overwrites (no fill, no read), later writers accumulate: same values, two memory
     * passes fewer per activation. grad_fresh[t] != 0 <=> tensor t's gradient holds a partial
     * sum of the current step. */
    /* PREDICT-mode forward as a CUDA graph (batch-1 inference is launch-bound: ~60 kernels of
     * a few microseconds each). State machine: the first forward of a configuration runs eagerly
     * (one-time initialisation inside the launchers), the second is captured and replayed, every
     * later one is a single cudaGraphLaunch. The key (nodes, tensors, math, input buffer)
     * invalidates the graph when the net changes. graphs: 1 = on (default), 0 = off. */
    int graphs;
    /* TRAIN: forward + backward of bcnn_b200_train_step / bcnn_train_on_batch as a graph (small
     * nets -- the reference's own mnist / cifar examples -- are launch-bound: ~100 kernels in
     * well under a millisecond). The update kernels join the graph when their scalars cannot
     * change (SGD with a constant learning rate); otherwise (schedules, Adam's bias correction)
     * the update stays eager. Two slots, keyed by the input / label buffers,
     * because the input pipeline alternates two sets of them. Not used with per-node profiling
     * or extra inputs. Data parallelism: NCCL calls stay outside the graphs (see train_graph). */
    struct {
        void *exec;
        /* data parallelism: the backward pass of the first `split` nodes as a second graph, so that
         * the gradients of everything above it travel while it runs (bcnn_net.c:train_graph) */
        void *exec_tail;
        unsigned long long kernels_tail;
        int split;
        const void *input, *label;
        unsigned long long kernels;
        /* with_update: the SGD update kernels are part of the graph (constant learning rate
         * only: their scalars are then the same every step); lr / momentum / decay are the
         * values they were recorded with and part of the key */
        int with_update;
        float lr, momentum, decay;
    } step_graph[2];
    /* data parallelism: the update kernels as their own graph, replayed behind the all-reduce */
    struct {
        void *exec;
        unsigned long long kernels;
        float lr, momentum, decay;
    } update_graph;
    unsigned long long fwd_graph_kernels;
    int step_graph_warm, step_graph_next;
    void *fwd_graph;
    int fwd_graph_warm;
    int fwd_graph_nodes, fwd_graph_tensors, fwd_graph_math;
    const void *fwd_graph_input;
    unsigned char *grad_fresh;
    int *consumers; /* activation consumers per tensor (bcnn_net_num_consumers) */
    int grad_state_tensors, grad_state_nodes;
    /* Resident BF16 NHWC activations (conv_math == BCNN_B200_MATH_TC_BF16, bcnn_resident.c):
     * per tensor a BF16 NHWC twin of data / grad, made on first use, and which copy is current. */
    struct bcnn_resident *res;
    int res_count;
    float *nhwc_scratch_gpu; /* per-channel reduction partials of the NHWC kernels (max over nodes) */
    size_t nhwc_scratch_floats;
} bcnn_cuda_context;

/* Where the current value of a tensor's data / gradient lives. */
enum {
    BCNN_RES_F32 = 0,
    BCNN_RES_BF16 = 1,
    BCNN_RES_BOTH = 2,
    /* data only: the tensor is the not-yet-applied batch norm of its producer's raw convolution
     * result (the residual add that consumes it applies the normalisation itself); anybody else
     * who asks for it gets it materialised first (bcnn_conv_layer_materialize) */
    BCNN_RES_DEFERRED = 3
};
typedef struct bcnn_resident {
    void *data16, *grad16;
    unsigned char data_at, grad_at; /* BCNN_RES_* */
    int producer;   /* BCNN_RES_DEFERRED: index of the producing node */
    /* backward: this tensor's incoming gradient was not written; it is the (masked) gradient of
     * tensor grad_alias - 1, left there by the residual add that consumes this tensor. 0 = none. */
    int grad_alias;
} bcnn_resident;

struct bcnn_net {
    int batch_size;
    int num_nodes;
    int num_tensors;
    int num_inputs;
    int *inputs; /* indexes of the input tensors in tensors[] */
    bcnn_mode mode;
    bcnn_log_context log_ctx;
    bcnn_node *nodes;
    bcnn_tensor *tensors;
    bcnn_learner *learner;
    void *data_loader; /* unused: file loaders are out of scope */
    void *data_aug;    /* unused */
    void *gemm_ctx;    /* unused: no CPU GEMM on this path */
    void *cuda_ctx;    /* bcnn_cuda_context* */
    int num_threads;
};

bcnn_status bcnn_net_create_cuda_context(bcnn_net *net);
bcnn_status bcnn_net_add_node(bcnn_net *net, bcnn_node node);
bcnn_status bcnn_net_add_tensor(bcnn_net *net, bcnn_tensor tensor);

/* -- helpers shared by the layer files -- */
static inline bcnn_cuda_context *bcnn_ctx(bcnn_net *net) {
    return (bcnn_cuda_context *)net->cuda_ctx;
}
static inline void *bcnn_stream(bcnn_net *net) { return bcnn_ctx(net)->stream; }
/* Resolve the source tensor of a new node: by name (reverse scan, last definition wins),
 * or tensor 0 when the net is still empty. Returns the index or -1. */
int bcnn_net_find_src(bcnn_net *net, const char *src_id);
/* Create a named parameter tensor, append it to the net and to node->src[]. */
bcnn_status bcnn_net_add_param_tensor(bcnn_net *net, bcnn_node *node, int n, int c, int h, int w,
                                      int has_grad, const char *prefix, const char *suffix,
                                      const bcnn_tensor_filler *filler);
/* Allocate the output tensor of a node (device buffers), name it, append it. */
bcnn_status bcnn_net_add_dst_tensor(bcnn_net *net, bcnn_node *node, int n, int c, int h, int w,
                                    const char *dst_id);
/* Number of nodes that read tensor `index` as an activation input (src[0], or both
 * inputs of an eltwise node). */
int bcnn_net_num_consumers(bcnn_net *net, int index);
/* For a backward writer that can overwrite: returns 1 when tensor `index`'s gradient already
 * holds a partial sum of this step (accumulate into it), 0 when it is stale (overwrite it);
 * either way the buffer counts as fresh afterwards. */
int bcnn_net_grad_accumulate(bcnn_net *net, int index);
/* For a backward writer that can only accumulate: zero-fill the buffer now if it is stale. */
void bcnn_net_grad_prepare_accumulate(bcnn_net *net, int index);
/* Grow the shared conv workspace requirement. */
void bcnn_net_require_workspace(bcnn_net *net, size_t bytes);
void bcnn_net_require_dy_shadow(bcnn_net *net, size_t bytes);
/* Global batch (local batch x data-parallel world) and the factor gradients are scaled
 * with after an SGD step (momentum, or momentum / world under data parallelism). */
int bcnn_net_global_batch(bcnn_net *net);
float bcnn_net_grad_post_scale(bcnn_net *net, float momentum);

/* ---- resident BF16 NHWC activations (bcnn_resident.c) ----
 * In BCNN_B200_MATH_TC_BF16 the layers between two convolutions keep activations and their
 * gradients as BF16 NHWC. bcnn_tensor.data_gpu / grad_data_gpu (FP32 NCHW, the reference's
 * contract, inc/bcnn/bcnn.h:242-255) are brought up to date on demand: by
 * bcnn_get_tensor_by_index and in front of every node that is not format-aware. */
int bcnn_net_resident(bcnn_net *net);              /* the net runs in the resident mode */
int bcnn_net_tensor_can16(bcnn_net *net, int idx); /* shape allows a BF16 NHWC twin (c % 8 == 0) */
/* current value as BF16 NHWC (converted from FP32 if that copy is the current one) */
void *bcnn_net_data16_in(bcnn_net *net, int idx);
void *bcnn_net_grad16_in(bcnn_net *net, int idx);
/* buffer about to be overwritten with BF16 NHWC content: afterwards it is the only current copy */
void *bcnn_net_data16_out(bcnn_net *net, int idx);
void *bcnn_net_grad16_out(bcnn_net *net, int idx);
/* the BF16 copy was modified in place: the FP32 copy is stale */
void bcnn_net_grad16_modified(bcnn_net *net, int idx);
/* current value as FP32 NCHW in data_gpu / grad_data_gpu (converted if needed) */
float *bcnn_net_data32_in(bcnn_net *net, int idx);
float *bcnn_net_grad32_in(bcnn_net *net, int idx);
/* data_gpu / grad_data_gpu was (or is about to be) written directly */
void bcnn_net_data32_written(bcnn_net *net, int idx);
void bcnn_net_grad32_written(bcnn_net *net, int idx);
/* what a node that only knows FP32 NCHW needs before / after its forward and backward */
void bcnn_net_node_f32_before_forward(bcnn_net *net, bcnn_node *node);
void bcnn_net_node_f32_after_forward(bcnn_net *net, bcnn_node *node);
void bcnn_net_node_f32_before_backward(bcnn_net *net, bcnn_node *node);
void bcnn_net_node_f32_after_backward(bcnn_net *net, bcnn_node *node);
/* 1 when the node's own forward / backward handle the resident format */
int bcnn_net_node_is_resident(bcnn_net *net, bcnn_node *node);
float *bcnn_net_nhwc_scratch(bcnn_net *net, int channels);
/* The resident residual-add node `consumer` is the only reader of tensor idx (1), so the batch norm
 * of the convolution producing idx may be applied inside that add and idx's gradient read from the
 * add's output gradient; 0 otherwise. *consumer_out receives the node index. */
int bcnn_net_sole_eltwise_consumer(bcnn_net *net, int idx, int *consumer_out);
bcnn_resident *bcnn_net_res(bcnn_net *net, int idx);
void bcnn_net_resident_release(bcnn_net *net);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_NET_H */
