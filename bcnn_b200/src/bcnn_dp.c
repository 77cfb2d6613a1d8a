/*
 * bcnn_dp.c -- data-parallel training: one bcnn_net replica per GPU (one process each),
 * NCCL sum all-reduce of every parameter-gradient tensor over NVLink, queued on a
 * dedicated comm stream as soon as the owning node's backward has been enqueued, so the
 * transfer overlaps the rest of the backward pass; bcnn_update joins it.
 *
 * NCCL is bound at run time (dlopen) so the library has no link-time dependency and
 * re-uses the copy torch already loaded in a bench/test process.
 *
 * Momentum-in-gradient-buffer (reference src/bcnn_learner.c:74,80): the
 * gradient buffer holds m*v_prev + g_local when backward ends. Every rank holds the same
 * m*v_prev, so the SGD kernel leaves (m / world) * v behind instead of m * v: the sum
 * over ranks then restores exactly m*v_prev + sum_r g_r with no extra pass or buffer.
 */
#include "bcnn_dp.h"

#include <dlfcn.h>

#include <bcnn_b200_net.h>

#include "bcnn_tensor.h"

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(void **comm, int nranks, nccl_uid id, int rank);
typedef int (*fn_allreduce)(const void *send, void *recv, size_t count, int dtype, int op,
                            void *comm, void *stream);
typedef int (*fn_void)(void);
typedef int (*fn_comm)(void *comm);
typedef const char *(*fn_errstr)(int);

enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0 };

static struct {
    void *lib;
    fn_get_uid get_uid;
    fn_init_rank init_rank;
    fn_allreduce allreduce;
    fn_void group_start, group_end;
    fn_comm destroy;
    fn_errstr errstr;
} nccl;

/* Gradients are all-reduced in buckets: tensors of consecutive nodes (in backward order) are
 * collected until BUCKET_BYTES are pending, then issued as ONE NCCL group (NCCL aggregates the
 * collectives of a group into a single launch) behind one event. ResNet-50: 102 MB in 5 launches
 * instead of one per node. */
#define BUCKET_BYTES ((size_t)25 << 20)
#define BUCKET_MAX_TENSORS 256

struct bcnn_dp_state {
    int rank, world;
    void *comm;
    void *comm_stream;
    void *evt_ready; /* compute -> comm */
    void *evt_done;  /* comm -> compute */
    size_t bytes_per_step, bytes_this_step;
    float *pending[BUCKET_MAX_TENSORS];
    size_t pending_count[BUCKET_MAX_TENSORS];
    int num_pending;
    size_t pending_bytes;
    int groups_per_step, groups_this_step;
    /* set while a step graph is recorded / replayed: backward does not issue transfers (NCCL calls
     * stay out of the graph), bcnn_dp_allreduce_all issues them behind the replay */
    int defer;
};

static int nccl_bind(void) {
    if (nccl.lib) return 0;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        fprintf(stderr, "[ERROR] [NCCL] cannot load libnccl.so.2: %s\n", dlerror());
        return -1;
    }
    nccl.get_uid = (fn_get_uid)dlsym(lib, "ncclGetUniqueId");
    nccl.init_rank = (fn_init_rank)dlsym(lib, "ncclCommInitRank");
    nccl.allreduce = (fn_allreduce)dlsym(lib, "ncclAllReduce");
    nccl.group_start = (fn_void)dlsym(lib, "ncclGroupStart");
    nccl.group_end = (fn_void)dlsym(lib, "ncclGroupEnd");
    nccl.destroy = (fn_comm)dlsym(lib, "ncclCommDestroy");
    nccl.errstr = (fn_errstr)dlsym(lib, "ncclGetErrorString");
    if (!nccl.get_uid || !nccl.init_rank || !nccl.allreduce || !nccl.group_start ||
        !nccl.group_end || !nccl.destroy) {
        fprintf(stderr, "[ERROR] [NCCL] libnccl.so.2 lacks an expected symbol\n");
        return -1;
    }
    nccl.lib = lib;
    return 0;
}

#define nccl_check(RET)                                                              \
    do { int r_ = (RET);                                                             \
         if (r_ != 0) {                                                              \
             fprintf(stderr, "[ERROR] [NCCL] %s (%s:%d)\n",                          \
                     nccl.errstr ? nccl.errstr(r_) : "error", __FILE__, __LINE__);   \
             exit(r_); } } while (0)

int bcnn_b200_dp_get_unique_id(char id[BCNN_B200_DP_ID_BYTES]) {
    if (nccl_bind() != 0) return -1;
    nccl_uid uid;
    memset(&uid, 0, sizeof(uid));
    int r = nccl.get_uid(&uid);
    memcpy(id, uid.internal, sizeof(uid.internal));
    return r;
}

int bcnn_b200_dp_init(bcnn_net *net, int rank, int world, const char id[BCNN_B200_DP_ID_BYTES]) {
    if (world <= 1) return 0;
    if (nccl_bind() != 0) return -1;
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    struct bcnn_dp_state *dp = (struct bcnn_dp_state *)calloc(1, sizeof(*dp));
    if (!dp) return -1;
    nccl_uid uid;
    memcpy(uid.internal, id, sizeof(uid.internal));
    nccl_check(nccl.init_rank(&dp->comm, world, uid, rank));
    dp->rank = rank;
    dp->world = world;
    dp->comm_stream = bcnn_b200_stream_create();
    dp->evt_ready = bcnn_b200_event_create();
    dp->evt_done = bcnn_b200_event_create();
    ctx->dp = dp;
    return 0;
}

int bcnn_dp_world_size(bcnn_net *net) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    return dp ? dp->world : 1;
}

int bcnn_b200_dp_world(bcnn_net *net) { return bcnn_dp_world_size(net); }

size_t bcnn_b200_dp_bytes_per_step(bcnn_net *net) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    return dp ? dp->bytes_per_step : 0;
}

static void flush_bucket(bcnn_net *net, struct bcnn_dp_state *dp) {
    if (dp->num_pending == 0) return;
    /* everything enqueued on the compute stream so far (the backward passes that produced the
     * pending gradients) precedes the transfer; under stream capture this forks the comm stream
     * into the graph */
    bcnn_cuda_check(bcnn_b200_event_record(dp->evt_ready, bcnn_stream(net)));
    bcnn_cuda_check(bcnn_b200_stream_wait_event(dp->comm_stream, dp->evt_ready));
    nccl_check(nccl.group_start());
    for (int i = 0; i < dp->num_pending; ++i)
        nccl_check(nccl.allreduce(dp->pending[i], dp->pending[i], dp->pending_count[i], NCCL_FLOAT32,
                                  NCCL_SUM, dp->comm, dp->comm_stream));
    nccl_check(nccl.group_end());
    dp->num_pending = 0;
    dp->pending_bytes = 0;
    dp->groups_this_step++;
}

/* Parameter gradients an optimizer step consumes: W and bias / beta of the nodes that have an
 * update function (src[1], src[2]; the PReLU slopes of an activation node are src[1]). Batch-norm
 * scale gradients are written by backward but never applied: they stay local. */
void bcnn_dp_set_deferred(bcnn_net *net, int on) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    if (dp) dp->defer = on;
}

static void collect_node(bcnn_net *net, struct bcnn_dp_state *dp, bcnn_node *node);

void bcnn_dp_after_node_backward(bcnn_net *net, bcnn_node *node) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    if (!dp || dp->defer) return;
    collect_node(net, dp, node);
}

/* Every gradient bucket of the step, behind whatever the compute stream holds (a replayed step
 * graph): same buckets in the same order as the overlapped path. */
void bcnn_dp_allreduce_range(bcnn_net *net, int first, int end) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    if (!dp) return;
    /* one group for the whole range (fewer, larger messages) */
    dp->defer = 2;
    for (int i = end - 1; i >= first; --i) collect_node(net, dp, &net->nodes[i]);
    dp->defer = 0;
    flush_bucket(net, dp);
}

void bcnn_dp_allreduce_all(bcnn_net *net) { bcnn_dp_allreduce_range(net, 0, net->num_nodes); }

/* Where to cut the backward pass of a replayed step in two: the node index `split` such that the
 * nodes below it hold at most ~8 % of the gradient bytes (ResNet-50: the stem and the first two
 * stages -- most of the backward TIME, because their activations are the large ones). */
int bcnn_dp_backward_split(bcnn_net *net) {
    size_t total = 0, below = 0;
    for (int i = 0; i < net->num_nodes; ++i) {
        const bcnn_node *node = &net->nodes[i];
        if (!node->update) continue;
        for (int j = 1; j <= 2 && j < node->num_src; ++j) total += (size_t)bcnn_tensor_size(&net->tensors[node->src[j]]);
    }
    int split = 0;
    for (int i = 0; i < net->num_nodes; ++i) {
        const bcnn_node *node = &net->nodes[i];
        if (node->update)
            for (int j = 1; j <= 2 && j < node->num_src; ++j) below += (size_t)bcnn_tensor_size(&net->tensors[node->src[j]]);
        if (below * 12 > total) break;
        split = i + 1;
    }
    return split < net->num_nodes ? split : 0;
}

static void collect_node(bcnn_net *net, struct bcnn_dp_state *dp, bcnn_node *node) {
    if (node->update) {
        const int last = node->type == BCNN_LAYER_ACTIVATION ? 1 : 2;
        for (int i = 1; i <= last && i < node->num_src; ++i) {
            bcnn_tensor *t = &net->tensors[node->src[i]];
            if (!t->has_grad || !t->grad_data_gpu) continue;
            const size_t count = (size_t)bcnn_tensor_size(t);
            if (dp->num_pending == BUCKET_MAX_TENSORS) flush_bucket(net, dp);
            dp->pending[dp->num_pending] = t->grad_data_gpu;
            dp->pending_count[dp->num_pending++] = count;
            dp->pending_bytes += count * sizeof(float);
            dp->bytes_this_step += count * sizeof(float);
        }
    }
    /* a full bucket goes out now and overlaps the rest of backward; the remainder when the first
     * node's backward is done */
    if (dp->defer == 2) return;
    if (dp->pending_bytes >= BUCKET_BYTES || node == &net->nodes[0]) flush_bucket(net, dp);
}

/* Join: the compute stream waits for every transfer issued so far. Idempotent. */
void bcnn_dp_before_update(bcnn_net *net) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    if (!dp) return;
    flush_bucket(net, dp);
    bcnn_cuda_check(bcnn_b200_event_record(dp->evt_done, dp->comm_stream));
    bcnn_cuda_check(bcnn_b200_stream_wait_event(bcnn_stream(net), dp->evt_done));
    if (dp->bytes_this_step) {
        dp->bytes_per_step = dp->bytes_this_step;
        dp->groups_per_step = dp->groups_this_step;
    }
    dp->bytes_this_step = 0;
    dp->groups_this_step = 0;
}

int bcnn_b200_dp_groups_per_step(bcnn_net *net) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    return dp ? dp->groups_per_step : 0;
}

/* Time of `iters` all-reduces of the whole gradient set (the step's buckets, back to back on the
 * comm stream, nothing else running): milliseconds per set, for the bus-bandwidth figure of the
 * bench line. The gradient buffers are summed over and over: call it after the measurements. */
float bcnn_b200_dp_allreduce_probe_ms(bcnn_net *net, int iters) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    if (!dp || iters < 1) return 0.f;
    bcnn_b200_sync(net);
    void *e0 = bcnn_b200_event_create(), *e1 = bcnn_b200_event_create();
    float ms = 0.f;
    for (int it = -1; it < iters; ++it) { /* one untimed round first */
        if (it == 0) bcnn_cuda_check(bcnn_b200_event_record(e0, dp->comm_stream));
        bcnn_dp_allreduce_all(net);
    }
    bcnn_cuda_check(bcnn_b200_event_record(e1, dp->comm_stream));
    bcnn_cuda_check(bcnn_b200_stream_sync(dp->comm_stream));
    ms = bcnn_b200_event_elapsed_ms(e0, e1) / (float)iters;
    bcnn_b200_event_destroy(e0);
    bcnn_b200_event_destroy(e1);
    dp->bytes_this_step = 0;
    dp->groups_this_step = 0;
    return ms;
}

void bcnn_dp_sync(bcnn_net *net) {
    struct bcnn_dp_state *dp = bcnn_ctx(net)->dp;
    if (dp) bcnn_cuda_check(bcnn_b200_stream_sync(dp->comm_stream));
}

void bcnn_dp_release(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    struct bcnn_dp_state *dp = ctx ? ctx->dp : NULL;
    if (!dp) return;
    bcnn_b200_stream_sync(dp->comm_stream);
    if (dp->comm) nccl.destroy(dp->comm);
    bcnn_b200_event_destroy(dp->evt_ready);
    bcnn_b200_event_destroy(dp->evt_done);
    bcnn_b200_stream_destroy(dp->comm_stream);
    free(dp);
    ctx->dp = NULL;
}

void bcnn_b200_dp_shutdown(bcnn_net *net) { bcnn_dp_release(net); }
