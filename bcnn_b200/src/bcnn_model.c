/*
 * bcnn_model.c -- weight files: bcnn_save_weights / bcnn_load_weights.
 *
 * File layouts are those of jnbraun/bcnn src/bcnn_net.c:
 *   .bcnnmodel (anything that is not *.weights / *.onnx), writer :597-681, reader :1485-1558
 *       "BCNN" | u32 major | u32 minor | u32 patch | per node, in graph order:
 *         conv / depthwise : bias, weights [, run_mean, run_var, scales   (conv with BN)]
 *                            [, prelu slopes  -- READ by :1305-1321 but never WRITTEN by :597-681]
 *         prelu activation : slopes
 *         batchnorm        : run_mean, run_var, scales, bias
 *         fully connected  : bias, weights
 *   Darknet *.weights, reader only (:1508-1527, :1232-1300, :1365-1372, :1439-1466)
 *       i32 major | i32 minor | i32 revision | seen (u64 when major*10+minor >= 2, else i32) |
 *         conv / depthwise : bias [, scales, run_mean, run_var], weights [, prelu slopes]
 *         batchnorm        : scales, run_mean, run_var
 *         fully connected  : bias, weights (transposed in place when major or minor > 1000)
 * In PREDICT mode the reader folds the running statistics into scales / bias
 * (:1278-1289, :1394-1404) -- the form bcnn_b200_scale_bias applies at inference.
 *
 * Organisation here: one table per node lists its records (tensor, element count) in file
 * order; saving and loading are the same walk over that table with the transfer direction
 * swapped. The host mirrors of the parameter tensors are the staging buffers (pinned), the
 * device buffers stay the truth: save = D2H + fwrite, load = fread [+ fold / transpose] + H2D.
 *
 * Deviation from the reference, on purpose: a short read makes bcnn_load_weights return
 * BCNN_INVALID_MODEL. The reference logs the error inside its per-layer readers but drops their
 * status (:1539-1548) and reports BCNN_SUCCESS for a truncated file. Tensors read before the
 * truncation point are loaded in both.
 */
#include <math.h>
#include <stdint.h>

#include "bcnn_activation_layer.h"
#include "bcnn_conv_layer.h"
#include "bcnn_net.h"
#include "bcnn_tensor.h"

#define MODEL_MAGIC "BCNN"
enum { MODEL_FMT_BCNN = 0, MODEL_FMT_DARKNET = 1, MODEL_FMT_ONNX = 2 };
enum { MODEL_MAX_RECORDS = 6 };

typedef struct {
    bcnn_tensor *tensor;
    int count; /* floats */
} model_record;

typedef struct {
    model_record rec[MODEL_MAX_RECORDS];
    int num;
    /* post-processing of a load */
    bcnn_tensor *fold_bias, *fold_scales, *fold_mean, *fold_var; /* BN fold (PREDICT) */
    bcnn_tensor *transpose;                                      /* darknet fc weights */
    int rows, cols;
} model_plan;

static void plan_add(model_plan *p, bcnn_tensor *t, int count) {
    p->rec[p->num].tensor = t;
    p->rec[p->num].count = count;
    ++p->num;
}

/* Records of `node` in file order. `loading` adds what only the reader touches. */
static void plan_node(bcnn_net *net, bcnn_node *node, int format, int loading, model_plan *p) {
    bcnn_tensor *t = net->tensors;
    memset(p, 0, sizeof(*p));
    switch (node->type) {
        case BCNN_LAYER_CONV2D:
        case BCNN_LAYER_DEPTHWISE_CONV2D: {
            bcnn_tensor *w = &t[node->src[1]], *b = &t[node->src[2]];
            bcnn_conv_param *cp =
                node->type == BCNN_LAYER_CONV2D ? (bcnn_conv_param *)node->param : NULL;
            const int bn = cp && cp->batch_norm == 1;
            plan_add(p, b, bcnn_tensor_size(b));
            if (format == MODEL_FMT_BCNN) plan_add(p, w, bcnn_tensor_size(w));
            if (bn) {
                bcnn_tensor *m = &t[node->src[3]], *v = &t[node->src[4]], *s = &t[node->src[5]];
                if (format == MODEL_FMT_DARKNET) plan_add(p, s, bcnn_tensor_size(s));
                plan_add(p, m, bcnn_tensor_size(m));
                plan_add(p, v, bcnn_tensor_size(v));
                if (format == MODEL_FMT_BCNN) plan_add(p, s, bcnn_tensor_size(s));
                p->fold_bias = b;
                p->fold_scales = s;
                p->fold_mean = m;
                p->fold_var = v;
            }
            if (format == MODEL_FMT_DARKNET) plan_add(p, w, bcnn_tensor_size(w));
            /* The reference's saver never writes the slopes its reader expects (:597-681 vs
             * :1305-1321), so it cannot read back its own conv+PReLU files. Here both walks carry
             * them: a file written by this library loads in this library AND in the reference. */
            (void)loading;
            if (cp && cp->activation == BCNN_ACT_PRELU) {
                bcnn_tensor *slopes = &t[node->src[3 + 3 * cp->batch_norm]];
                plan_add(p, slopes, bcnn_tensor_size(slopes));
            }
            break;
        }
        case BCNN_LAYER_ACTIVATION: {
            bcnn_activation_param *ap = (bcnn_activation_param *)node->param;
            if (ap->activation == BCNN_ACT_PRELU && format == MODEL_FMT_BCNN) {
                bcnn_tensor *slopes = &t[node->src[1]];
                plan_add(p, slopes, bcnn_tensor_size(slopes));
            }
            break;
        }
        case BCNN_LAYER_BATCHNORM: {
            const int c = t[node->dst[0]].c;
            bcnn_tensor *m = &t[node->src[1]], *v = &t[node->src[2]], *s = &t[node->src[3]],
                        *b = &t[node->src[4]];
            if (format == MODEL_FMT_DARKNET) plan_add(p, s, c);
            plan_add(p, m, c);
            plan_add(p, v, c);
            if (format == MODEL_FMT_BCNN) {
                plan_add(p, s, c);
                plan_add(p, b, c);
            }
            p->fold_bias = b;
            p->fold_scales = s;
            p->fold_mean = m;
            p->fold_var = v;
            break;
        }
        case BCNN_LAYER_FULL_CONNECTED: {
            bcnn_tensor *w = &t[node->src[1]], *b = &t[node->src[2]];
            plan_add(p, b, bcnn_tensor_size(b));
            plan_add(p, w, bcnn_tensor_size(w));
            p->transpose = w;
            p->rows = bcnn_tensor_size3d(&t[node->src[0]]);
            p->cols = bcnn_tensor_size3d(&t[node->dst[0]]);
            break;
        }
        default:
            break;
    }
}

static bcnn_status to_host(bcnn_net *net, bcnn_tensor *t, int count) {
    BCNN_CHECK_STATUS(bcnn_tensor_ensure_host(t));
    BCNN_CHECK(t->data != NULL, BCNN_FAILED_ALLOC);
    if (t->data_gpu)
        bcnn_cuda_check(bcnn_b200_memcpy_d2h(t->data, t->data_gpu, (size_t)count * sizeof(float),
                                             bcnn_stream(net)));
    return BCNN_SUCCESS;
}

static void to_device(bcnn_net *net, bcnn_tensor *t, int count) {
    if (t->data_gpu)
        bcnn_cuda_check(bcnn_b200_memcpy_h2d(t->data_gpu, t->data, (size_t)count * sizeof(float),
                                             bcnn_stream(net)));
}

bcnn_status bcnn_save_weights(bcnn_net *net, const char *filename) {
    FILE *fp = filename ? fopen(filename, "wb") : NULL;
    BCNN_CHECK_AND_LOG(net->log_ctx, fp, BCNN_INVALID_PARAMETER,
                       "Could not open model file %s\n", filename ? filename : "(null)");
    const uint32_t version[3] = {BCNN_VERSION_MAJOR, BCNN_VERSION_MINOR, BCNN_VERSION_PATCH};
    bcnn_status st = BCNN_SUCCESS;
    int ok = fwrite(MODEL_MAGIC, 1, 4, fp) == 4 && fwrite(version, sizeof(uint32_t), 3, fp) == 3;
    for (int i = 0; ok && st == BCNN_SUCCESS && i < net->num_nodes; ++i) {
        model_plan plan;
        plan_node(net, &net->nodes[i], MODEL_FMT_BCNN, /*loading=*/0, &plan);
        for (int r = 0; r < plan.num && st == BCNN_SUCCESS; ++r)
            st = to_host(net, plan.rec[r].tensor, plan.rec[r].count);
        if (st != BCNN_SUCCESS || plan.num == 0) continue;
        bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_stream(net)));
        for (int r = 0; ok && r < plan.num; ++r)
            ok = fwrite(plan.rec[r].tensor->data, sizeof(float), (size_t)plan.rec[r].count, fp) ==
                 (size_t)plan.rec[r].count;
    }
    ok = (fclose(fp) == 0) && ok;
    BCNN_CHECK_STATUS(st);
    BCNN_CHECK_AND_LOG(net->log_ctx, ok, BCNN_INVALID_DATA, "Could not write model file %s\n",
                       filename);
    return BCNN_SUCCESS;
}

/* The reference decides by the text after the last '.' of the whole path (:1468-1483). */
static int model_format_of(const char *filename) {
    const char *dot = strrchr(filename, '.');
    const char *ext = dot ? dot + 1 : filename;
    if (strcmp(ext, "weights") == 0) return MODEL_FMT_DARKNET;
    if (strcmp(ext, "onnx") == 0) return MODEL_FMT_ONNX;
    return MODEL_FMT_BCNN;
}

/* [rows x cols] row-major -> [cols x rows] row-major, in place through a scratch copy. */
static bcnn_status transpose_host(float *a, int rows, int cols) {
    float *tmp = (float *)malloc((size_t)rows * cols * sizeof(float));
    BCNN_CHECK(tmp != NULL, BCNN_FAILED_ALLOC);
    memcpy(tmp, a, (size_t)rows * cols * sizeof(float));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) a[(size_t)c * rows + r] = tmp[(size_t)r * cols + c];
    free(tmp);
    return BCNN_SUCCESS;
}

/* beta' = beta - gamma * mean / sqrt(var + 1e-6), gamma' = gamma / sqrt(var + 1e-6): the same
 * float operations, in the same order, as the reference's loop (:1281-1288). */
static void fold_batchnorm(float *bias, float *scales, const float *mean, const float *var,
                           int c) {
    for (int i = 0; i < c; ++i) {
        const float sd = sqrtf(var[i] + 0.000001f);
        const float sm = scales[i] * mean[i];
        bias[i] = bias[i] - sm / sd;
        scales[i] = scales[i] / sd;
    }
}

static int read_header(bcnn_net *net, FILE *fp, int format, const char *filename,
                       int *need_transpose) {
    *need_transpose = 0;
    if (format == MODEL_FMT_BCNN) {
        char magic[4] = {0};
        uint32_t version[3] = {0};
        size_t got = fread(magic, 1, 4, fp);
        got += fread(version, sizeof(uint32_t), 3, fp);
        if (got != 7 || memcmp(magic, MODEL_MAGIC, 4) != 0) {
            bcnn_log(net->log_ctx, BCNN_LOG_ERROR, "Invalid format for model file %s\n", filename);
            return 0;
        }
        BCNN_INFO(net->log_ctx, "BCNN version %d.%d.%d used for model %s\n", (int)version[0],
                  (int)version[1], (int)version[2], filename);
        return 1;
    }
    int32_t head[3] = {0};
    if (fread(head, sizeof(int32_t), 3, fp) != 3) return 0;
    uint64_t seen = 0;
    if (head[0] * 10 + head[1] >= 2 && head[0] < 1000 && head[1] < 1000) {
        if (fread(&seen, sizeof(uint64_t), 1, fp) != 1) return 0;
    } else {
        int32_t seen32 = 0;
        if (fread(&seen32, sizeof(int32_t), 1, fp) != 1) return 0;
        seen = (uint64_t)seen32;
    }
    BCNN_INFO(net->log_ctx, "Darknet version %d.%d seen %lu\n", (int)head[0], (int)head[1],
              (unsigned long)seen);
    *need_transpose = head[0] > 1000 || head[1] > 1000;
    return 1;
}

bcnn_status bcnn_load_weights(bcnn_net *net, const char *filename) {
    BCNN_CHECK_AND_LOG(net->log_ctx, filename != NULL, BCNN_INVALID_PARAMETER,
                       "Can not open file %s\n", "(null)");
    const int format = model_format_of(filename);
    FILE *fp = fopen(filename, "rb");
    BCNN_CHECK_AND_LOG(net->log_ctx, fp, BCNN_INVALID_PARAMETER, "Can not open file %s\n",
                       filename);
    if (format == MODEL_FMT_ONNX) {
        fclose(fp);
        BCNN_ERROR(net->log_ctx, BCNN_INVALID_MODEL, "Model file %s format is not yet supported\n",
                   filename);
    }
    int need_transpose = 0;
    if (!read_header(net, fp, format, filename, &need_transpose)) {
        fclose(fp);
        return BCNN_INVALID_MODEL;
    }
    bcnn_status st = BCNN_SUCCESS;
    for (int i = 0; i < net->num_nodes && st == BCNN_SUCCESS; ++i) {
        model_plan plan;
        plan_node(net, &net->nodes[i], format, /*loading=*/1, &plan);
        int complete = 1;
        for (int r = 0; r < plan.num; ++r) {
            bcnn_tensor *t = plan.rec[r].tensor;
            const int count = plan.rec[r].count;
            st = bcnn_tensor_ensure_host(t);
            if (st != BCNN_SUCCESS) break;
            const size_t got = fread(t->data, sizeof(float), (size_t)count, fp);
            if (got != (size_t)count) {
                bcnn_log(net->log_ctx, BCNN_LOG_ERROR,
                         "Inconsistent size for %s: expected %d but found %lu\n",
                         t->name ? t->name : "?", count, (unsigned long)got);
                st = BCNN_INVALID_MODEL;
                complete = 0;
                break;
            }
        }
        if (plan.num == 0 || !complete) continue;
        if (plan.fold_scales && net->mode == BCNN_MODE_PREDICT)
            fold_batchnorm(plan.fold_bias->data, plan.fold_scales->data, plan.fold_mean->data,
                           plan.fold_var->data, plan.rec[0].count);
        if (plan.transpose && need_transpose)
            st = transpose_host(plan.transpose->data, plan.rows, plan.cols);
        for (int r = 0; r < plan.num; ++r) to_device(net, plan.rec[r].tensor, plan.rec[r].count);
        /* a standalone batchnorm read from a Darknet file keeps its bias but still folds into it */
        if (plan.fold_bias && net->mode == BCNN_MODE_PREDICT)
            to_device(net, plan.fold_bias, plan.rec[0].count);
        bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_stream(net)));
    }
    fclose(fp);
    if (st == BCNN_SUCCESS) BCNN_INFO(net->log_ctx, "Model %s loaded succesfully\n", filename);
    return st;
}
