/*
 * bcnn_net.c -- graph runtime of the hot path: net lifetime, tensor / node tables, the
 * forward / backward loops and tensor getters.
 *
 * Mirrors the behaviour of jnbraun/bcnn src/bcnn_net.c:61-463 for the functions on the
 * path (bcnn_init_net :61, bcnn_net_add_node/_tensor :236-258, bcnn_set_input_shape :280,
 * bcnn_compile_net :356, bcnn_reset_gradients :361, getters :377-408, bcnn_forward :410,
 * bcnn_backward :424, bcnn_train_on_batch / bcnn_predict_on_batch :452-483, bcnn_add_input
 * :260). Weight files: bcnn_model.c. The cfg parser and data loaders of that file are out
 * of scope. Everything runs on one CUDA stream per net; nothing synchronises unless a
 * getter needs host data.
 */
#include "bcnn_net.h"

#include "bcnn_dp.h"
#include "bcnn_tensor.h"
#include "bcnn_yolo.h"
#include <bcnn_b200_net.h>

static void forward_graph_drop(bcnn_cuda_context *ctx);
static void packs_drop(bcnn_net *net);

static void *g_current_stream = NULL;
void *bcnn_b200_current_stream(void) { return g_current_stream; }
void bcnn_b200_set_current_stream(void *stream) { g_current_stream = stream; }

bcnn_status bcnn_net_create_cuda_context(bcnn_net *net) {
    bcnn_cuda_context *ctx = (bcnn_cuda_context *)calloc(1, sizeof(bcnn_cuda_context));
    BCNN_CHECK(ctx != NULL, BCNN_FAILED_ALLOC);
    ctx->stream = bcnn_b200_stream_create();
    if (!ctx->stream) {
        fprintf(stderr,
                "[ERROR] [CUDA] bcnn_b200 needs a CUDA device (sm_100a); there is no CPU path\n");
        free(ctx);
        return BCNN_CUDA_FAILED_ALLOC;
    }
    const char *graphs = getenv("BCNN_B200_GRAPHS");
    ctx->graphs = !(graphs && graphs[0] == '0');
    /* Default: tensor-core math with resident BF16 NHWC activations (BCNN_B200_MATH_TC_BF16).
     * BCNN_B200_CONV_MATH=tc keeps FP32 NCHW tensors between the layers (tensor-core convolutions on
     * transposed shadows), =fp32 selects the FP32 SIMT verification path; bcnn_b200_set_conv_math
     * does the same per net. */
    const char *math = getenv("BCNN_B200_CONV_MATH");
    ctx->conv_math = BCNN_B200_MATH_TC_BF16;
    if (math && (math[0] == 'f' || math[0] == 'F' || math[0] == '0')) ctx->conv_math = BCNN_B200_MATH_FP32;
    else if (math && (math[0] == 't' || math[0] == 'T' || math[0] == '1')) ctx->conv_math = BCNN_B200_MATH_TC;
    ctx->conv_math_explicit = math && math[0];
    /* Batch-correct residual semantics by default; the reference's two residual bugs (H2, H3) are
     * replicated only on request (parity tests): BCNN_B200_REFERENCE_QUIRKS=1 or
     * bcnn_b200_set_reference_quirks(net, 1). */
    const char *quirks = getenv("BCNN_B200_REFERENCE_QUIRKS");
    ctx->reference_quirks = (quirks && quirks[0] && quirks[0] != '0') ? 1 : 0;
    net->cuda_ctx = ctx;
    return BCNN_SUCCESS;
}

bcnn_status bcnn_init_net(bcnn_net **net, bcnn_mode mode) {
    bcnn_net *p = (bcnn_net *)calloc(1, sizeof(bcnn_net));
    BCNN_CHECK(p != NULL, BCNN_FAILED_ALLOC);
    p->mode = mode;
    /* tensors[0] = "input", tensors[1] = "label", shaped later */
    bcnn_tensor input = {0}, label = {0};
    input.name = bcnn_strdup_("input");
    label.name = bcnn_strdup_("label");
    bcnn_net_add_tensor(p, input);
    bcnn_net_add_tensor(p, label);
    if (mode != BCNN_MODE_PREDICT) p->learner = (bcnn_learner *)calloc(1, sizeof(bcnn_learner));
    p->num_inputs = 1;
    p->inputs = (int *)calloc(1, sizeof(int));
    p->num_threads = 1;
    bcnn_status st = bcnn_net_create_cuda_context(p);
    if (st != BCNN_SUCCESS) {
        bcnn_end_net(&p);
        return st;
    }
    *net = p;
    return BCNN_SUCCESS;
}

void bcnn_end_net(bcnn_net **net) {
    bcnn_net *p = *net;
    if (!p) return;
    bcnn_cuda_context *ctx = bcnn_ctx(p);
    if (ctx) {
        bcnn_b200_stream_sync(ctx->stream);
        packs_drop(p);
        bcnn_dp_release(p);
        if (g_current_stream == ctx->stream) g_current_stream = NULL; /* about to be destroyed */
    }
    for (int i = 0; i < p->num_nodes; ++i) {
        bcnn_node *node = &p->nodes[i];
        if (node->release_param) node->release_param(node);
        free(node->src);
        free(node->dst);
        free(node->param);
    }
    free(p->nodes);
    for (int i = 0; i < p->num_tensors; ++i) bcnn_tensor_destroy(&p->tensors[i]);
    free(p->tensors);
    free(p->learner);
    free(p->inputs);
    if (ctx) {
        for (int i = 0; i < 4 * ctx->profile_nodes; ++i) bcnn_b200_event_destroy(ctx->profile_events[i]);
        free(ctx->profile_events);
        free(ctx->grad_fresh);
        free(ctx->consumers);
        forward_graph_drop(ctx);
        bcnn_net_resident_release(p);
        bcnn_b200_free(ctx->workspace_gpu);
        bcnn_b200_free(ctx->dy_shadow_gpu);
        if (ctx->stage_gpu) {
            for (int i = 0; i <= p->num_inputs; ++i) bcnn_b200_free(ctx->stage_gpu[i]);
            free(ctx->stage_gpu);
        }
        bcnn_b200_event_destroy(ctx->evt_uploaded);
        bcnn_b200_event_destroy(ctx->evt_consumed);
        bcnn_b200_stream_destroy(ctx->copy_stream);
        bcnn_b200_stream_destroy(ctx->stream);
        free(ctx);
    }
    free(p); /* like the reference, the caller's pointer is left dangling */
}

void bcnn_set_log_context(bcnn_net *net, bcnn_log_callback fct, bcnn_log_level level) {
    net->log_ctx.fct = fct;
    net->log_ctx.lvl = level;
}

bcnn_status bcnn_set_num_threads(bcnn_net *net, int num_threads, const int *cpu_ids) {
    (void)cpu_ids; /* host threads do no arithmetic on this path */
    net->num_threads = num_threads < 1 ? 1 : num_threads;
    return BCNN_SUCCESS;
}

int bcnn_get_num_threads(bcnn_net *net) { return net->num_threads; }

bcnn_status bcnn_net_add_node(bcnn_net *net, bcnn_node node) {
    bcnn_node *grown = (bcnn_node *)realloc(net->nodes, (size_t)(net->num_nodes + 1) * sizeof(bcnn_node));
    BCNN_CHECK_AND_LOG(net->log_ctx, grown != NULL, BCNN_FAILED_ALLOC, "Internal allocation error\n");
    grown[net->num_nodes++] = node;
    net->nodes = grown;
    return BCNN_SUCCESS;
}

bcnn_status bcnn_net_add_tensor(bcnn_net *net, bcnn_tensor tensor) {
    bcnn_tensor *grown =
        (bcnn_tensor *)realloc(net->tensors, (size_t)(net->num_tensors + 1) * sizeof(bcnn_tensor));
    BCNN_CHECK_AND_LOG(net->log_ctx, grown != NULL, BCNN_FAILED_ALLOC, "Internal allocation error\n");
    grown[net->num_tensors++] = tensor;
    net->tensors = grown;
    return BCNN_SUCCESS;
}

void bcnn_set_input_shape(bcnn_net *net, int width, int height, int channels, int batch_size) {
    net->batch_size = batch_size;
    bcnn_tensor_set_shape(&net->tensors[0], batch_size, channels, height, width, 0);
}

int bcnn_get_batch_size(bcnn_net *net) { return net->batch_size; }

bcnn_status bcnn_set_mode(bcnn_net *net, bcnn_mode mode) {
    /* TRAIN <-> VALID switches are free; gradient buffers exist only if the net was
     * created in TRAIN / VALID mode (as in the reference, bcnn_tensor.c:111). */
    if (net->mode != mode) forward_graph_drop(bcnn_ctx(net));
    net->mode = mode;
    return BCNN_SUCCESS;
}

/* Resident tensors pay when the layers between convolutions know the format. A net with nodes that
 * only read FP32 NCHW activations at full spatial size (depthwise convolution, stand-alone batch
 * norm / activation, concat, upsample, yolo: MobileNet, YOLOv3-tiny, the mnist example) would convert
 * around every one of them, so unless the caller chose a mode such a net stays on FP32-tensor
 * tensor-core math (the round-1 layout). */
static void choose_default_math(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (ctx->conv_math_explicit || ctx->conv_math != BCNN_B200_MATH_TC_BF16) return;
    for (int i = 0; i < net->num_nodes; ++i) {
        const bcnn_node *node = &net->nodes[i];
        const bcnn_tensor *src = node->num_src > 0 ? &net->tensors[node->src[0]] : NULL;
        const int spatial = src && src->h * src->w > 1;
        switch (node->type) {
            case BCNN_LAYER_CONV2D:
            case BCNN_LAYER_MAXPOOL:
            case BCNN_LAYER_AVGPOOL:
            case BCNN_LAYER_ELTWISE:
                break;
            default:
                if (spatial) {
                    ctx->conv_math = BCNN_B200_MATH_TC;
                    return;
                }
        }
    }
}

bcnn_status bcnn_compile_net(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    choose_default_math(net);
    forward_graph_drop(ctx); /* buffers are reallocated below */
    packs_drop(net);
    /* (re)allocate the input tensor, with an eager pinned host mirror the caller fills */
    BCNN_CHECK_STATUS(bcnn_tensor_allocate(&net->tensors[0], net->mode));
    BCNN_CHECK_STATUS(bcnn_tensor_ensure_host(&net->tensors[0]));
    if (net->tensors[1].data_gpu) BCNN_CHECK_STATUS(bcnn_tensor_ensure_host(&net->tensors[1]));
    /* shared conv workspace = max requirement over the conv nodes */
    bcnn_b200_free(ctx->workspace_gpu);
    ctx->workspace_gpu = NULL;
    if (ctx->workspace_bytes) {
        ctx->workspace_gpu = (float *)bcnn_b200_malloc(ctx->workspace_bytes);
        BCNN_CHECK(ctx->workspace_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);
    }
    ctx->workspace_size = (int)(ctx->workspace_bytes / sizeof(float));
    bcnn_b200_free(ctx->dy_shadow_gpu);
    ctx->dy_shadow_gpu = NULL;
    if (ctx->dy_shadow_bytes && net->mode == BCNN_MODE_TRAIN) {
        ctx->dy_shadow_gpu = bcnn_b200_malloc(ctx->dy_shadow_bytes);
        BCNN_CHECK(ctx->dy_shadow_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);
    }
    return BCNN_SUCCESS;
}

int bcnn_get_tensor_index_by_name(bcnn_net *net, const char *name) {
    for (int i = net->num_tensors - 1; i >= 0; --i)
        if (net->tensors[i].name && strcmp(net->tensors[i].name, name) == 0) return i;
    return -1;
}

bcnn_tensor *bcnn_get_tensor_by_index(bcnn_net *net, int index) {
    if (index < 0 || index >= net->num_tensors) return NULL;
    bcnn_tensor *t = &net->tensors[index];
    size_t bytes = (size_t)bcnn_tensor_size(t) * sizeof(float);
    if (bytes == 0 || !t->data_gpu) return t;
    void *stream = bcnn_stream(net);
    if (bcnn_tensor_ensure_host(t) != BCNN_SUCCESS) return NULL;
    /* resident mode: the FP32 NCHW buffers are materialised from the BF16 NHWC twins here */
    (void)bcnn_net_data32_in(net, index);
    if (t->grad_data_gpu) (void)bcnn_net_grad32_in(net, index);
    bcnn_cuda_check(bcnn_b200_memcpy_d2h(t->data, t->data_gpu, bytes, stream));
    if (t->grad_data_gpu && t->grad_data)
        bcnn_cuda_check(bcnn_b200_memcpy_d2h(t->grad_data, t->grad_data_gpu, bytes, stream));
    bcnn_cuda_check(bcnn_b200_stream_sync(stream));
    return t;
}

bcnn_tensor *bcnn_get_tensor_by_name(bcnn_net *net, const char *name) {
    return bcnn_get_tensor_by_index(net, bcnn_get_tensor_index_by_name(net, name));
}

/* (Re)build the per-tensor gradient state tables after the graph changed. */
static void grad_state_sync(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (ctx->grad_state_tensors == net->num_tensors && ctx->grad_state_nodes == net->num_nodes &&
        ctx->grad_fresh)
        return;
    free(ctx->grad_fresh);
    free(ctx->consumers);
    ctx->grad_fresh = (unsigned char *)malloc((size_t)net->num_tensors);
    ctx->consumers = (int *)malloc((size_t)net->num_tensors * sizeof(int));
    memset(ctx->grad_fresh, 1, (size_t)net->num_tensors); /* outside TRAIN: always accumulate */
    for (int i = 0; i < net->num_tensors; ++i) ctx->consumers[i] = bcnn_net_num_consumers(net, i);
    ctx->grad_state_tensors = net->num_tensors;
    ctx->grad_state_nodes = net->num_nodes;
}

/* The reference zero-fills the gradient of a node's outputs before its forward (TRAIN only,
 * bcnn_reset_gradients); weight gradients are never reset (momentum lives there). Here the
 * buffers are only marked stale (see bcnn_cuda_context.grad_fresh); a tensor nobody consumes
 * has no backward writer and is really filled. */
static void reset_output_gradients(bcnn_net *net, bcnn_node *node) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    for (int i = 0; i < node->num_dst; ++i) {
        const int idx = node->dst[i];
        bcnn_tensor *t = &net->tensors[idx];
        if (!t->grad_data_gpu) continue;
        bcnn_net_grad32_written(net, idx); /* stale either way; a resident writer claims it again */
        if (ctx->consumers[idx] == 0) {
            bcnn_cuda_check(bcnn_b200_fill_f32(t->grad_data_gpu, (size_t)bcnn_tensor_size(t), 0.0f,
                                               bcnn_stream(net)));
            ctx->grad_fresh[idx] = 1;
        } else {
            ctx->grad_fresh[idx] = 0;
        }
    }
}

int bcnn_net_grad_accumulate(bcnn_net *net, int index) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    grad_state_sync(net);
    const int fresh = ctx->grad_fresh[index];
    ctx->grad_fresh[index] = 1;
    return fresh;
}

void bcnn_net_grad_prepare_accumulate(bcnn_net *net, int index) {
    if (bcnn_net_grad_accumulate(net, index)) return;
    bcnn_tensor *t = &net->tensors[index];
    if (t->grad_data_gpu) {
        bcnn_cuda_check(bcnn_b200_fill_f32(t->grad_data_gpu, (size_t)bcnn_tensor_size(t), 0.0f,
                                           bcnn_stream(net)));
        bcnn_net_grad32_written(net, index);
    }
}

static inline void profile_mark(bcnn_net *net, int node, int slot) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (ctx->profile && node < ctx->profile_nodes)
        bcnn_cuda_check(bcnn_b200_event_record(ctx->profile_events[4 * node + slot], ctx->stream));
}

/* ---- packed weight images (bcnn_net.h: packs_state) ---- */
#include "bcnn_conv_layer.h"

static void packs_drop(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (ctx->packs_jobs_host) { /* unregister: entry = {w, dst, dgrad, ...} (csrc/conv_tma.cu PackJob) */
        const size_t jb = bcnn_b200_conv_pack_job_bytes();
        for (int i = 0; i < ctx->packs_jobs; ++i) {
            const char *job = (const char *)ctx->packs_jobs_host + (size_t)i * jb;
            const float *w = *(const float *const *)job;
            const int dgrad = *(const int *)(job + 2 * sizeof(void *));
            bcnn_b200_conv_prepacked_set(w, dgrad, NULL);
        }
    }
    free(ctx->packs_jobs_host);
    bcnn_b200_free(ctx->packs_jobs_gpu);
    bcnn_b200_free(ctx->packs_images_gpu);
    ctx->packs_jobs_host = ctx->packs_jobs_gpu = ctx->packs_images_gpu = NULL;
    ctx->packs_jobs = 0;
    ctx->packs_grid = 0;
    if (ctx->packs_state > 0) ctx->packs_state = 0;
}

static void packs_prepare(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    const char *e = getenv("BCNN_B200_PACK_TABLE");
    ctx->packs_state = -1;
    if ((e && e[0] == '0') || !bcnn_net_resident(net)) return;
    const size_t jb = bcnn_b200_conv_pack_job_bytes();
    const int max_jobs = 8 * net->num_nodes + 8;
    char *jobs = (char *)calloc((size_t)max_jobs, jb);
    if (!jobs) return;
    /* pass 0 sizes the images, pass 1 lays them out */
    size_t total = 0;
    char *images = NULL;
    int count = 0;
    for (int pass = 0; pass < 2; ++pass) {
        size_t off = 0;
        count = 0;
        for (int i = 0; i < net->num_nodes; ++i) {
            bcnn_node *node = &net->nodes[i];
            if (node->type != BCNN_LAYER_CONV2D || !bcnn_conv_layer_is_resident(net, node)) continue;
            const bcnn_conv_param *param = (const bcnn_conv_param *)node->param;
            const float *w = net->tensors[node->src[1]].data_gpu;
            const int passes = net->tensors[node->src[0]].grad_data_gpu ? 2 : 1;
            for (int dgrad = 0; dgrad < passes; ++dgrad) {
                size_t bytes = 0;
                const int n = bcnn_b200_conv_nhwc_pack_jobs(&param->desc, dgrad, w, pass ? images + off : NULL,
                                                            jobs + (size_t)count * jb, max_jobs - count, &bytes);
                if (n <= 0) continue; /* this pass of this layer keeps packing per call */
                if (pass) bcnn_b200_conv_prepacked_set(w, dgrad, images + off);
                count += n;
                off += (bytes + 255) & ~(size_t)255;
            }
        }
        if (!pass) {
            total = off;
            if (!count || !total) break;
            images = (char *)bcnn_b200_malloc(total);
            if (!images) { count = 0; break; }
        }
    }
    if (!count || !images) {
        bcnn_b200_free(images);
        free(jobs);
        return;
    }
    ctx->packs_grid = bcnn_b200_conv_pack_table_finish(jobs, count);
    ctx->packs_jobs_gpu = bcnn_b200_malloc((size_t)count * jb);
    ctx->packs_jobs_host = jobs;
    ctx->packs_images_gpu = images;
    ctx->packs_jobs = count;
    if (!ctx->packs_jobs_gpu ||
        bcnn_b200_memcpy_h2d(ctx->packs_jobs_gpu, jobs, (size_t)count * jb, bcnn_stream(net)) != 0) {
        packs_drop(net);
        ctx->packs_state = -1;
        return;
    }
    /* the backward pass of this very step already reads the images */
    if (bcnn_b200_conv_pack_run(ctx->packs_jobs_gpu, ctx->packs_jobs, ctx->packs_grid, bcnn_stream(net)) != 0) {
        packs_drop(net);
        ctx->packs_state = -1;
        return;
    }
    bcnn_b200_stream_sync(bcnn_stream(net));
    ctx->packs_state = 1;
}

static void forward_nodes(bcnn_net *net) {
    g_current_stream = bcnn_stream(net);
    {   /* every packed weight image of the step in one launch (TRAIN; other modes pack per call) */
        bcnn_cuda_context *ctx = bcnn_ctx(net);
        const int on = ctx->packs_state == 1 && net->mode == BCNN_MODE_TRAIN;
        bcnn_b200_conv_prepacked_enable(on);
        if (on)
            bcnn_cuda_check(bcnn_b200_conv_pack_run(ctx->packs_jobs_gpu, ctx->packs_jobs, ctx->packs_grid,
                                                    ctx->stream));
    }
    for (int i = 0; i < net->num_nodes; ++i) {
        bcnn_node *node = &net->nodes[i];
        profile_mark(net, i, 0);
        if (net->mode == BCNN_MODE_TRAIN) reset_output_gradients(net, node);
        if (bcnn_net_node_is_resident(net, node)) {
            node->forward(net, node);
        } else { /* the node reads and writes the FP32 NCHW buffers */
            bcnn_net_node_f32_before_forward(net, node);
            node->forward(net, node);
            bcnn_net_node_f32_after_forward(net, node);
        }
        profile_mark(net, i, 1);
    }
    {   /* first eager TRAIN forward: every node knows its route now */
        bcnn_cuda_context *ctx = bcnn_ctx(net);
        if (ctx->packs_state == 0 && net->mode == BCNN_MODE_TRAIN && !ctx->capturing) packs_prepare(net);
    }
}

static void forward_graph_drop(bcnn_cuda_context *ctx) {
    bcnn_b200_graph_destroy(ctx->fwd_graph);
    ctx->fwd_graph = NULL;
    ctx->fwd_graph_warm = 0;
    for (int i = 0; i < 2; ++i) {
        bcnn_b200_graph_destroy(ctx->step_graph[i].exec);
        ctx->step_graph[i].exec = NULL;
        bcnn_b200_graph_destroy(ctx->step_graph[i].exec_tail);
        ctx->step_graph[i].exec_tail = NULL;
    }
    bcnn_b200_graph_destroy(ctx->update_graph.exec);
    ctx->update_graph.exec = NULL;
    ctx->step_graph_warm = 0;
}

/* Same configuration as the one the live graphs were recorded for? Otherwise drop them. */
static void graph_key_check(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (ctx->fwd_graph_nodes == net->num_nodes && ctx->fwd_graph_tensors == net->num_tensors &&
        ctx->fwd_graph_math == ctx->conv_math)
        return;
    forward_graph_drop(ctx);
    ctx->fwd_graph_nodes = net->num_nodes;
    ctx->fwd_graph_tensors = net->num_tensors;
    ctx->fwd_graph_math = ctx->conv_math;
}

/* PREDICT-mode forward through a CUDA graph; see bcnn_cuda_context.graphs. Returns 1 when the
 * step was run (replayed), 0 when the caller has to run it eagerly. */
static int forward_graph(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (!ctx->graphs || net->mode != BCNN_MODE_PREDICT || ctx->profile || net->num_nodes == 0)
        return 0;
    graph_key_check(net);
    if (ctx->fwd_graph_input != (const void *)net->tensors[0].data_gpu) {
        forward_graph_drop(ctx);
        ctx->fwd_graph_input = net->tensors[0].data_gpu;
    }
    int recorded_now = 0; /* the capture itself already counted its launches once */
    if (!ctx->fwd_graph) {
        recorded_now = 1;
        if (!ctx->fwd_graph_warm) { /* first forward of this configuration: eager */
            ctx->fwd_graph_warm = 1;
            return 0;
        }
        if (bcnn_b200_graph_begin(ctx->stream) != 0) {
            ctx->graphs = 0;
            return 0;
        }
        const unsigned long long before = bcnn_b200_launch_count();
        forward_nodes(net);
        ctx->fwd_graph_kernels = bcnn_b200_launch_count() - before;
        ctx->fwd_graph = bcnn_b200_graph_end(ctx->stream);
        if (!ctx->fwd_graph) { /* capture refused: stay eager for the rest of this net's life */
            BCNN_WARNING(net->log_ctx, "CUDA graph capture of the forward pass failed; running eagerly\n");
            ctx->graphs = 0;
            return 0;
        }
    }
    bcnn_cuda_check(bcnn_b200_graph_launch(ctx->fwd_graph, recorded_now ? 0 : ctx->fwd_graph_kernels,
                                           ctx->stream));
    return 1;
}

void bcnn_forward(bcnn_net *net) {
    grad_state_sync(net);
    if (forward_graph(net)) return;
    forward_nodes(net);
}

void bcnn_b200_set_graphs(bcnn_net *net, int on) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    ctx->graphs = on != 0;
    if (!on) forward_graph_drop(ctx);
}

int bcnn_b200_get_graphs(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    const int live = ctx->fwd_graph || ctx->step_graph[0].exec || ctx->step_graph[1].exec;
    return !ctx->graphs ? 0 : (live ? 2 : 1);
}

static void backward_range(bcnn_net *net, int first, int end);

void bcnn_backward(bcnn_net *net) { backward_range(net, 0, net->num_nodes); }

/* backward of nodes end - 1 .. first */
static void backward_range(bcnn_net *net, int first, int end) {
    g_current_stream = bcnn_stream(net);
    bcnn_b200_conv_prepacked_enable(bcnn_ctx(net)->packs_state == 1 && net->mode == BCNN_MODE_TRAIN);
    for (int i = end - 1; i >= first; --i) {
        bcnn_node *node = &net->nodes[i];
        profile_mark(net, i, 2);
        if (bcnn_net_node_is_resident(net, node)) {
            node->backward(net, node);
        } else {
            bcnn_net_node_f32_before_backward(net, node);
            node->backward(net, node);
            bcnn_net_node_f32_after_backward(net, node);
        }
        profile_mark(net, i, 3);
        bcnn_dp_after_node_backward(net, node); /* no-op without data parallelism */
    }
}

void bcnn_b200_profile(bcnn_net *net, int enable) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (enable && ctx->profile_nodes != net->num_nodes) {
        for (int i = 0; i < 4 * ctx->profile_nodes; ++i) bcnn_b200_event_destroy(ctx->profile_events[i]);
        free(ctx->profile_events);
        ctx->profile_nodes = net->num_nodes;
        ctx->profile_events = (void **)calloc((size_t)4 * net->num_nodes, sizeof(void *));
        for (int i = 0; i < 4 * net->num_nodes; ++i) ctx->profile_events[i] = bcnn_b200_event_create();
    }
    ctx->profile = enable;
}

int bcnn_b200_profile_node_ms(bcnn_net *net, int node, float *fwd_ms, float *bwd_ms) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (!ctx->profile_events || node < 0 || node >= ctx->profile_nodes) return -1;
    bcnn_cuda_check(bcnn_b200_stream_sync(ctx->stream));
    void **e = ctx->profile_events + 4 * node;
    if (fwd_ms) *fwd_ms = bcnn_b200_event_elapsed_ms(e[0], e[1]);
    if (bwd_ms) *bwd_ms = bcnn_b200_event_elapsed_ms(e[2], e[3]);
    return 0;
}

/* ---------------- helpers for the layer constructors ---------------- */

int bcnn_net_find_src(bcnn_net *net, const char *src_id) {
    if (net->num_nodes == 0) return 0;
    return bcnn_get_tensor_index_by_name(net, src_id);
}

bcnn_status bcnn_net_add_param_tensor(bcnn_net *net, bcnn_node *node, int n, int c, int h, int w,
                                      int has_grad, const char *prefix, const char *suffix,
                                      const bcnn_tensor_filler *filler) {
    char name[320];
    snprintf(name, sizeof(name), "%s%s", prefix, suffix);
    bcnn_tensor t = {0};
    bcnn_tensor_create(&t, n, c, h, w, has_grad, name, net->mode);
    BCNN_CHECK_AND_LOG(net->log_ctx, t.data_gpu != NULL, BCNN_CUDA_FAILED_ALLOC,
                       "Failed to allocate tensor %s\n", name);
    if (filler) bcnn_tensor_fill(&t, *filler);
    BCNN_CHECK_STATUS(bcnn_net_add_tensor(net, t));
    return bcnn_node_add_input(net, node, net->num_tensors - 1);
}

bcnn_status bcnn_net_add_dst_tensor(bcnn_net *net, bcnn_node *node, int n, int c, int h, int w,
                                    const char *dst_id) {
    bcnn_tensor t = {0};
    bcnn_tensor_set_shape(&t, n, c, h, w, 1);
    BCNN_CHECK_STATUS(bcnn_tensor_allocate(&t, net->mode));
    t.name = bcnn_strdup_(dst_id);
    BCNN_CHECK_STATUS(bcnn_net_add_tensor(net, t));
    return bcnn_node_add_output(net, node, net->num_tensors - 1);
}

int bcnn_net_num_consumers(bcnn_net *net, int index) {
    int count = 0;
    for (int i = 0; i < net->num_nodes; ++i) {
        const bcnn_node *node = &net->nodes[i];
        if (node->type == BCNN_LAYER_ACTIVATION) continue; /* in place: not a new reader */
        int inputs = node->type == BCNN_LAYER_ELTWISE ? 2 : 1;
        for (int j = 0; j < inputs && j < node->num_src; ++j)
            if (node->src[j] == index) ++count;
    }
    return count;
}

void bcnn_net_require_dy_shadow(bcnn_net *net, size_t bytes) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (bytes > ctx->dy_shadow_bytes) ctx->dy_shadow_bytes = bytes;
}

void bcnn_net_require_workspace(bcnn_net *net, size_t bytes) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (bytes > ctx->workspace_bytes) ctx->workspace_bytes = bytes;
}

int bcnn_net_global_batch(bcnn_net *net) { return net->batch_size * bcnn_dp_world_size(net); }

float bcnn_net_grad_post_scale(bcnn_net *net, float momentum) {
    return momentum / (float)bcnn_dp_world_size(net);
}

/* ---------------- extension API (include/bcnn_b200_net.h) ---------------- */

void bcnn_b200_set_conv_math(bcnn_net *net, int math) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    ctx->conv_math_explicit = 1;
    if (ctx->conv_math == math) return;
    /* leaving the resident mode: every tensor's current value goes back to the FP32 buffers */
    for (int i = 0; i < ctx->res_count && i < net->num_tensors; ++i) {
        (void)bcnn_net_data32_in(net, i);
        if (net->tensors[i].grad_data_gpu) (void)bcnn_net_grad32_in(net, i);
        ctx->res[i].data_at = ctx->res[i].grad_at = BCNN_RES_F32;
    }
    forward_graph_drop(ctx);
    packs_drop(net);
    ctx->packs_state = 0;
    ctx->conv_math = math;
}
void bcnn_b200_set_reference_quirks(bcnn_net *net, int on) {
    if (bcnn_ctx(net)->reference_quirks != on) forward_graph_drop(bcnn_ctx(net));
    bcnn_ctx(net)->reference_quirks = on;
}
int bcnn_b200_get_reference_quirks(bcnn_net *net) { return bcnn_ctx(net)->reference_quirks; }
int bcnn_b200_get_conv_math(bcnn_net *net) { return bcnn_ctx(net)->conv_math; }
void *bcnn_b200_get_stream(bcnn_net *net) { return bcnn_stream(net); }

void bcnn_b200_sync(bcnn_net *net) {
    bcnn_dp_sync(net);
    if (bcnn_ctx(net)->copy_stream)
        bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_ctx(net)->copy_stream));
    bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_stream(net)));
}

bcnn_status bcnn_b200_upload_tensor(bcnn_net *net, int index) {
    if (index < 0 || index >= net->num_tensors) return BCNN_INVALID_PARAMETER;
    bcnn_tensor *t = &net->tensors[index];
    size_t bytes = (size_t)bcnn_tensor_size(t) * sizeof(float);
    void *stream = bcnn_stream(net);
    if (t->data && t->data_gpu) bcnn_cuda_check(bcnn_b200_memcpy_h2d(t->data_gpu, t->data, bytes, stream));
    if (t->grad_data && t->grad_data_gpu)
        bcnn_cuda_check(bcnn_b200_memcpy_h2d(t->grad_data_gpu, t->grad_data, bytes, stream));
    bcnn_net_data32_written(net, index);
    bcnn_net_grad32_written(net, index);
    bcnn_cuda_check(bcnn_b200_stream_sync(stream));
    return BCNN_SUCCESS;
}

size_t bcnn_b200_upload_inputs(bcnn_net *net) {
    size_t queued = 0;
    void *stream = bcnn_stream(net);
    for (int i = 0; i <= net->num_inputs; ++i) { /* inputs[], then the label (tensor 1) */
        bcnn_tensor *t = &net->tensors[i < net->num_inputs ? net->inputs[i] : 1];
        size_t bytes = (size_t)bcnn_tensor_size(t) * sizeof(float);
        if (!t->data || !t->data_gpu || bytes == 0) continue;
        bcnn_cuda_check(bcnn_b200_memcpy_h2d(t->data_gpu, t->data, bytes, stream));
        bcnn_net_data32_written(net, i < net->num_inputs ? net->inputs[i] : 1);
        queued += bytes;
    }
    return queued;
}

float bcnn_b200_get_loss(bcnn_net *net) {
    float loss = 0.f;
    int count = 0;
    void *stream = bcnn_stream(net);
    for (int i = 0; i < net->num_nodes; ++i) {
        if (net->nodes[i].type == BCNN_LAYER_YOLOV3) { /* reference :437-443 */
            loss += bcnn_yolo_cost(net, &net->nodes[i]);
            ++count;
            continue;
        }
        if (net->nodes[i].type != BCNN_LAYER_COST) continue;
        float v = 0.f;
        bcnn_cuda_check(bcnn_b200_memcpy_d2h(&v, net->tensors[net->nodes[i].dst[0]].data_gpu,
                                             sizeof(float), stream));
        bcnn_cuda_check(bcnn_b200_stream_sync(stream));
        loss += v;
        ++count;
    }
    return count ? loss / count : 0.f;
}

/* ---- input pipeline: double-buffered uploads on a copy stream ---- */
static bcnn_tensor *pipeline_tensor(bcnn_net *net, int i) {
    return &net->tensors[i < net->num_inputs ? net->inputs[i] : 1]; /* inputs[], then the label */
}

/* Queue the upload of the host mirrors into the staging buffers. The copy waits for the last
 * step that read those buffers (they were the live inputs of the previous step), nothing else:
 * it overlaps whatever the compute stream is running. */
static size_t pipeline_prefetch(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    size_t queued = 0;
    if (!ctx->copy_stream) {
        ctx->copy_stream = bcnn_b200_stream_create();
        ctx->evt_uploaded = bcnn_b200_event_create();
        ctx->evt_consumed = bcnn_b200_event_create();
        ctx->stage_gpu = (float **)calloc((size_t)net->num_inputs + 1, sizeof(float *));
        if (!ctx->copy_stream || !ctx->evt_uploaded || !ctx->evt_consumed || !ctx->stage_gpu) {
            fprintf(stderr, "[ERROR] input pipeline: cannot create stream / events\n");
            exit(BCNN_CUDA_FAILED_ALLOC);
        }
        bcnn_cuda_check(bcnn_b200_event_record(ctx->evt_consumed, ctx->stream));
    }
    bcnn_cuda_check(bcnn_b200_stream_wait_event(ctx->copy_stream, ctx->evt_consumed));
    for (int i = 0; i <= net->num_inputs; ++i) {
        bcnn_tensor *t = pipeline_tensor(net, i);
        size_t bytes = (size_t)bcnn_tensor_size(t) * sizeof(float);
        if (!t->data || !t->data_gpu || bytes == 0) continue;
        if (!ctx->stage_gpu[i]) {
            ctx->stage_gpu[i] = (float *)bcnn_b200_malloc(bytes);
            if (!ctx->stage_gpu[i]) exit(BCNN_CUDA_FAILED_ALLOC);
        }
        bcnn_cuda_check(bcnn_b200_memcpy_h2d(ctx->stage_gpu[i], t->data, bytes, ctx->copy_stream));
        queued += bytes;
    }
    bcnn_cuda_check(bcnn_b200_event_record(ctx->evt_uploaded, ctx->copy_stream));
    ctx->stage_valid = 1;
    return queued;
}

size_t bcnn_b200_prefetch_inputs(bcnn_net *net) {
    size_t queued = pipeline_prefetch(net);
    bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_ctx(net)->copy_stream));
    return queued;
}

/* Make the staged batch the live one: swap the device buffers once the copy has landed. */
static void pipeline_swap_in(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (!ctx->stage_valid) pipeline_prefetch(net); /* first step: nothing was staged yet */
    bcnn_cuda_check(bcnn_b200_stream_wait_event(ctx->stream, ctx->evt_uploaded));
    for (int i = 0; i <= net->num_inputs; ++i) {
        bcnn_tensor *t = pipeline_tensor(net, i);
        if (!ctx->stage_gpu[i]) continue;
        float *live = t->data_gpu;
        t->data_gpu = ctx->stage_gpu[i];
        ctx->stage_gpu[i] = live;
        bcnn_net_data32_written(net, i < net->num_inputs ? net->inputs[i] : 1);
    }
    ctx->stage_valid = 0;
    /* completes when every earlier step has: from then on the old live buffers are free */
    bcnn_cuda_check(bcnn_b200_event_record(ctx->evt_consumed, ctx->stream));
}

/* forward + backward (+ the update kernels) of a training step through a CUDA graph; see
 * bcnn_cuda_context.step_graph. Returns 0 when the caller runs the step eagerly, 1 when forward
 * and backward were replayed, 2 when the update kernels were replayed with them. */
static int train_graph(bcnn_net *net) {
    bcnn_cuda_context *ctx = bcnn_ctx(net);
    if (!ctx->graphs || net->mode != BCNN_MODE_TRAIN || ctx->profile || net->num_inputs != 1 ||
        net->num_nodes == 0)
        return 0;
    if (ctx->dp) {
        /* Data parallelism: forward + backward are replayed as one graph and the gradient buckets
         * are all-reduced behind it (4 NCCL launches for ResNet-50, ~0.45 ms on NVLink), then the
         * update runs eagerly. Capturing the NCCL calls on a forked comm stream inside the graph
         * deadlocked both ranks on this stack (gpurun r2n), and the eager step it replaced cost
         * 1.8 ms of launch overhead per step -- more than the overlap was worth.
         * BCNN_B200_DP_GRAPH=0: fully eager steps with the transfers overlapping backward. */
        static int dp_graph = -1;
        if (dp_graph < 0) {
            const char *e = getenv("BCNN_B200_DP_GRAPH");
            dp_graph = !(e && e[0] == '0');
        }
        if (!dp_graph) return 0;
    }
    graph_key_check(net);
    if (!ctx->step_graph_warm) { /* first step of this configuration: eager (lazy allocations) */
        ctx->step_graph_warm = 1;
        return 0;
    }
    const void *input = net->tensors[0].data_gpu, *label = net->tensors[1].data_gpu;
    const bcnn_learner *ln = net->learner;
    const int with_update = ln && ln->optimizer == BCNN_OPTIM_SGD &&
                            ln->decay_type == BCNN_LR_DECAY_CONSTANT && !ctx->dp;
    int slot = -1, recorded_now = 0;
    for (int i = 0; i < 2; ++i)
        if (ctx->step_graph[i].exec && ctx->step_graph[i].input == input &&
            ctx->step_graph[i].label == label && ctx->step_graph[i].with_update == with_update &&
            (!with_update || (ctx->step_graph[i].lr == ln->learning_rate &&
                              ctx->step_graph[i].momentum == ln->momentum &&
                              ctx->step_graph[i].decay == ln->decay)))
            slot = i;
    if (slot < 0) {
        slot = ctx->step_graph_next;
        ctx->step_graph_next ^= 1;
        bcnn_b200_graph_destroy(ctx->step_graph[slot].exec);
        ctx->step_graph[slot].exec = NULL;
        grad_state_sync(net);
        if (bcnn_b200_graph_begin(ctx->stream) != 0) {
            ctx->graphs = 0;
            return 0;
        }
        const unsigned long long before = bcnn_b200_launch_count();
        ctx->capturing = 1;
        bcnn_dp_set_deferred(net, 1); /* no NCCL call inside the capture */
        const int split = ctx->dp ? bcnn_dp_backward_split(net) : 0;
        bcnn_b200_graph_destroy(ctx->step_graph[slot].exec_tail);
        ctx->step_graph[slot].exec_tail = NULL;
        ctx->step_graph[slot].split = split;
        forward_nodes(net);
        backward_range(net, split, net->num_nodes);
        if (split > 0) { /* data parallelism: the rest of backward is a second graph */
            ctx->step_graph[slot].kernels = bcnn_b200_launch_count() - before;
            ctx->step_graph[slot].exec = bcnn_b200_graph_end(ctx->stream);
            if (ctx->step_graph[slot].exec && bcnn_b200_graph_begin(ctx->stream) == 0) {
                const unsigned long long before_tail = bcnn_b200_launch_count();
                backward_range(net, 0, split);
                ctx->step_graph[slot].kernels_tail = bcnn_b200_launch_count() - before_tail;
                ctx->step_graph[slot].exec_tail = bcnn_b200_graph_end(ctx->stream);
            }
            if (!ctx->step_graph[slot].exec_tail) { /* half a step is no step: run eagerly */
                bcnn_b200_graph_destroy(ctx->step_graph[slot].exec);
                ctx->step_graph[slot].exec = NULL;
            }
            bcnn_dp_set_deferred(net, 0);
        } else {
        bcnn_dp_set_deferred(net, 0);
        if (with_update) bcnn_update_nodes(net);
        ctx->step_graph[slot].kernels = bcnn_b200_launch_count() - before;
        ctx->step_graph[slot].exec = bcnn_b200_graph_end(ctx->stream);
        }
        ctx->capturing = 0;
        if (!ctx->step_graph[slot].exec) {
            BCNN_WARNING(net->log_ctx, "CUDA graph capture of the training step failed; running eagerly\n");
            ctx->graphs = 0;
            return 0;
        }
        ctx->step_graph[slot].input = input;
        ctx->step_graph[slot].label = label;
        ctx->step_graph[slot].with_update = with_update;
        if (with_update) {
            ctx->step_graph[slot].lr = ln->learning_rate;
            ctx->step_graph[slot].momentum = ln->momentum;
            ctx->step_graph[slot].decay = ln->decay;
        }
        recorded_now = 1;
    }
    bcnn_cuda_check(bcnn_b200_graph_launch(ctx->step_graph[slot].exec,
                                           recorded_now ? 0 : ctx->step_graph[slot].kernels, ctx->stream));
    if (ctx->dp) {
        /* gradients of the nodes whose backward is done travel while the tail graph (the stem and
         * the early stages: few parameters, most of the backward time) runs */
        const int split = ctx->step_graph[slot].split;
        bcnn_dp_allreduce_range(net, split, net->num_nodes);
        if (ctx->step_graph[slot].exec_tail) {
            bcnn_cuda_check(bcnn_b200_graph_launch(ctx->step_graph[slot].exec_tail,
                                                   recorded_now ? 0 : ctx->step_graph[slot].kernels_tail,
                                                   ctx->stream));
            bcnn_dp_allreduce_range(net, 0, split);
        }
        /* constant learning rate: the update kernels replay as a second graph behind the join */
        if (ln && ln->optimizer == BCNN_OPTIM_SGD && ln->decay_type == BCNN_LR_DECAY_CONSTANT) {
            bcnn_dp_before_update(net);
            if (ctx->update_graph.exec &&
                (ctx->update_graph.lr != ln->learning_rate || ctx->update_graph.momentum != ln->momentum ||
                 ctx->update_graph.decay != ln->decay)) {
                bcnn_b200_graph_destroy(ctx->update_graph.exec);
                ctx->update_graph.exec = NULL;
            }
            int fresh = 0;
            if (!ctx->update_graph.exec && bcnn_b200_graph_begin(ctx->stream) == 0) {
                const unsigned long long before = bcnn_b200_launch_count();
                bcnn_update_nodes(net);
                ctx->update_graph.kernels = bcnn_b200_launch_count() - before;
                ctx->update_graph.exec = bcnn_b200_graph_end(ctx->stream);
                ctx->update_graph.lr = ln->learning_rate;
                ctx->update_graph.momentum = ln->momentum;
                ctx->update_graph.decay = ln->decay;
                fresh = 1;
            }
            if (ctx->update_graph.exec) {
                bcnn_cuda_check(bcnn_b200_graph_launch(ctx->update_graph.exec,
                                                       fresh ? 0 : ctx->update_graph.kernels, ctx->stream));
                return 2;
            }
        }
        return 1; /* bcnn_update joins the transfers and steps eagerly */
    }
    return with_update ? 2 : 1;
}

float bcnn_b200_train_step(bcnn_net *net, int upload_inputs, int fetch_loss) {
    if (upload_inputs == 2) pipeline_swap_in(net);
    else if (upload_inputs) bcnn_b200_upload_inputs(net);
    const int replayed = train_graph(net);
    if (!replayed) {
        bcnn_forward(net);
        bcnn_backward(net);
    }
    if (replayed == 2) bcnn_update_schedule(net); /* the kernels ran inside the graph */
    else bcnn_update(net);
    if (upload_inputs == 2) pipeline_prefetch(net);
    float loss = fetch_loss ? bcnn_b200_get_loss(net) : 0.f;
    /* the host mirrors may be refilled once the call returns */
    if (upload_inputs == 2 && fetch_loss)
        bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_ctx(net)->copy_stream));
    return loss;
}

/* bcnn_train_on_batch (reference src/bcnn_net.c:452-463): next batch, forward, backward,
 * update, loss. The reference pulls the batch from its file loader (bcnn_loader_next,
 * src/bcnn_data.c:402-427: decode, augment, synchronous cudaMemcpy) -- out of scope here, and a
 * NULL dereference there when no loader is set. Here the batch is what the caller left in the
 * pinned host mirrors of the inputs and the label; it is uploaded on the compute stream. Use
 * bcnn_b200_train_step(net, 2, ...) for the overlapped input pipeline. */
float bcnn_train_on_batch(bcnn_net *net) { return bcnn_b200_train_step(net, 1, 1); }

/* bcnn_predict_on_batch (reference :465-483): forward on the batch in the host mirrors; *out is
 * the last node's output, or the cost node's prediction input, with its host copy refreshed. */
float bcnn_predict_on_batch(bcnn_net *net, bcnn_tensor **out) {
    if (out) *out = NULL;
    if (net->num_nodes < 1) return 0.f;
    bcnn_b200_upload_inputs(net);
    bcnn_forward(net);
    const bcnn_node *last = &net->nodes[net->num_nodes - 1];
    const int out_id = last->type == BCNN_LAYER_COST ? last->src[0] : last->dst[0];
    if (out) *out = bcnn_get_tensor_by_index(net, out_id);
    return bcnn_b200_get_loss(net);
}

/* bcnn_add_input (reference :260-278): one more input tensor [batch, c, h, w] without
 * gradient, registered in net->inputs[] so the upload calls and the input pipeline carry it.
 * Call it after bcnn_set_input_shape (it takes the batch size from the net) and before the
 * first pipelined step. */
bcnn_status bcnn_add_input(bcnn_net *net, int w, int h, int c, const char *name) {
    BCNN_CHECK_AND_LOG(net->log_ctx, w > 0 && h > 0 && c > 0 && net->batch_size > 0 && name,
                       BCNN_INVALID_PARAMETER, "bcnn_add_input: invalid shape or name\n");
    BCNN_CHECK_AND_LOG(net->log_ctx, bcnn_ctx(net)->stage_gpu == NULL, BCNN_INVALID_PARAMETER,
                       "bcnn_add_input: the input pipeline is already running\n");
    bcnn_tensor input = {0};
    bcnn_tensor_set_shape(&input, net->batch_size, c, h, w, 0);
    BCNN_CHECK_STATUS(bcnn_tensor_allocate(&input, net->mode));
    BCNN_CHECK_STATUS(bcnn_tensor_ensure_host(&input));
    input.name = bcnn_strdup_(name);
    BCNN_CHECK_STATUS(bcnn_net_add_tensor(net, input));
    int *grown = (int *)realloc(net->inputs, (size_t)(net->num_inputs + 1) * sizeof(int));
    BCNN_CHECK_AND_LOG(net->log_ctx, grown != NULL, BCNN_FAILED_ALLOC,
                       "Internal allocation error\n");
    net->inputs = grown;
    net->inputs[net->num_inputs++] = net->num_tensors - 1;
    return BCNN_SUCCESS;
}

int bcnn_b200_num_nodes(bcnn_net *net) { return net->num_nodes; }
int bcnn_b200_num_tensors(bcnn_net *net) { return net->num_tensors; }
int bcnn_b200_node_type(bcnn_net *net, int node) {
    return (node >= 0 && node < net->num_nodes) ? (int)net->nodes[node].type : -1;
}
int bcnn_b200_node_src(bcnn_net *net, int node, int i) {
    if (node < 0 || node >= net->num_nodes || i < 0 || i >= net->nodes[node].num_src) return -1;
    return net->nodes[node].src[i];
}
int bcnn_b200_node_dst(bcnn_net *net, int node, int i) {
    if (node < 0 || node >= net->num_nodes || i < 0 || i >= net->nodes[node].num_dst) return -1;
    return net->nodes[node].dst[i];
}
int bcnn_b200_tensor_dims(bcnn_net *net, int index, int *dims) {
    if (index < 0 || index >= net->num_tensors || !dims) return -1;
    const bcnn_tensor *t = &net->tensors[index];
    dims[0] = t->n; dims[1] = t->c; dims[2] = t->h; dims[3] = t->w;
    return 0;
}
