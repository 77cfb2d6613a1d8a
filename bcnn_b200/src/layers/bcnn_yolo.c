/*
 * bcnn_yolo.c -- YOLOv3 output node, inference side.
 *
 * Node layout and checks of jnbraun/bcnn src/layers/bcnn_yolo.c:15-107: src[0] = the head's
 * [N, num*(coords+classes+1), H, W] map, dst[0] the same shape, param = mask + anchor sizes.
 *   * forward, every mode (:226-250: copy, logistic on centre offsets, objectness and class
 *     scores): one device kernel; the reference copies the head to the host for it even in its
 *     CUDA build (:418-431).
 *   * TRAIN: the detection loss (:251-416) is host code in the reference, also in its CUDA build
 *     (D2H of the head, loss on the CPU, H2D of the gradient: a pipeline stall in every step).
 *     Here it is three kernels on the net's stream (csrc/yolo.cu) and nothing leaves the device
 *     until somebody asks for the loss. bcnn_yolo_loss_host keeps the reference's loops as host
 *     code on the host mirrors -- bit-identical to the reference for identical head tensors --
 *     and serves as the cross-check of the kernels (bcnn_b200_yolo_loss_on_host); the forward
 *     pass never calls it.
 *   * bcnn_yolo_get_detections (:470-639: box decoding, letterbox correction, objectness NMS) is
 *     host post-processing in the reference and is host code here too, on the heads' host mirrors
 *     after one D2H refresh per head.
 */
#include "bcnn_yolo.h"

#include <math.h>

#include "bcnn_tensor.h"
#include <bcnn_b200_net.h>

bcnn_status bcnn_add_yolo_layer(bcnn_net *net, int num_boxes_per_cell, int classes, int coords,
                                int total, int *mask, float *anchors, const char *src_id,
                                const char *dst_id) {
    bcnn_node node = {0};
    BCNN_CHECK_AND_LOG(net->log_ctx, net->num_nodes >= 1, BCNN_INVALID_PARAMETER,
                       "Yolo layer can't be the first layer of the network\n");
    int src = bcnn_net_find_src(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Yolo layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx,
                       num_boxes_per_cell > 0 && classes >= 0 && coords >= 2 && total > 0 && mask,
                       BCNN_INVALID_PARAMETER, "Yolo layer: invalid box / class / anchor counts\n");
    for (int k = 0; k < num_boxes_per_cell; ++k) /* the kernels index the anchor table with these */
        BCNN_CHECK_AND_LOG(net->log_ctx, mask[k] >= 0 && mask[k] < total, BCNN_INVALID_PARAMETER,
                           "Yolo layer: mask entry %d outside the %d anchors\n", mask[k], total);
    BCNN_CHECK_STATUS(bcnn_node_add_input(net, &node, src));
    const int n = net->tensors[src].n, c = net->tensors[src].c, h = net->tensors[src].h,
              w = net->tensors[src].w;
    BCNN_CHECK_AND_LOG(net->log_ctx, num_boxes_per_cell * (classes + coords + 1) == c,
                       BCNN_INVALID_PARAMETER, "Yolo layer: inconsistent number of channels %d\n",
                       num_boxes_per_cell * (classes + coords + 1));

    node.type = BCNN_LAYER_YOLOV3;
    node.param_size = sizeof(bcnn_yolo_param);
    bcnn_yolo_param *param = (bcnn_yolo_param *)calloc(1, node.param_size);
    BCNN_CHECK(param != NULL, BCNN_FAILED_ALLOC);
    node.param = param;
    param->num = num_boxes_per_cell;
    param->total = total;
    param->classes = classes;
    param->coords = coords;
    param->max_boxes = BCNN_DETECTION_MAX_BOXES;
    param->truths = param->max_boxes * (coords + 1);
    param->mask = (int *)calloc((size_t)num_boxes_per_cell, sizeof(int));
    param->cost = (float *)calloc(1, sizeof(float));
    BCNN_CHECK(param->mask != NULL && param->cost != NULL, BCNN_FAILED_ALLOC);
    memcpy(param->mask, mask, (size_t)num_boxes_per_cell * sizeof(int));
    /* anchor sizes: 0.5 unless given (reference :55-63) */
    char name[320];
    snprintf(name, sizeof(name), "%s_b", src_id);
    bcnn_tensor_create(&param->biases, 1, 1, 1, total * 2, 0, name, net->mode);
    BCNN_CHECK(param->biases.data != NULL, BCNN_FAILED_ALLOC);
    for (int i = 0; i < total * 2; ++i) param->biases.data[i] = anchors ? anchors[i] : 0.5f;
    bcnn_cuda_check(bcnn_b200_memcpy_h2d(param->biases.data_gpu, param->biases.data,
                                         (size_t)total * 2 * sizeof(float), bcnn_stream(net)));
    bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_stream(net)));
    param->mask_gpu = (int *)bcnn_b200_malloc((size_t)num_boxes_per_cell * sizeof(int));
    param->cost_gpu =
        (float *)bcnn_b200_malloc((size_t)bcnn_b200_yolo_cost_scratch_floats() * sizeof(float));
    BCNN_CHECK(param->mask_gpu != NULL && param->cost_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);
    bcnn_cuda_check(bcnn_b200_memcpy_h2d(param->mask_gpu, param->mask,
                                         (size_t)num_boxes_per_cell * sizeof(int), bcnn_stream(net)));
    bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_stream(net)));
    node.forward = bcnn_forward_yolo_layer;
    node.backward = bcnn_backward_yolo_layer;
    node.release_param = bcnn_release_param_yolo_layer;

    /* TRAIN / VALID: the label holds up to 50 boxes x (x, y, w, h, class) per sample (:69-74) */
    if (net->mode != BCNN_MODE_PREDICT && net->tensors[1].data_gpu == NULL) {
        bcnn_tensor_set_shape(&net->tensors[1], n, 1, 1, BCNN_DETECTION_MAX_BOXES * 5, 0);
        BCNN_CHECK_STATUS(bcnn_tensor_allocate(&net->tensors[1], net->mode));
        BCNN_CHECK_STATUS(bcnn_tensor_ensure_host(&net->tensors[1]));
    }
    BCNN_CHECK_STATUS(bcnn_net_add_dst_tensor(net, &node, n, c, h, w, dst_id));
    BCNN_CHECK_STATUS(bcnn_net_add_node(net, node));
    BCNN_INFO(net->log_ctx, "[Yolo] %-8s (%4d x%4d x%4d) -> %-8s (%4d x%4d x%4d) %5d\n", src_id,
              w, h, c, dst_id, w, h, c, classes);
    return BCNN_SUCCESS;
}

typedef struct { float x, y, w, h; } yolo_box;

static float span_overlap(float c1, float e1, float c2, float e2) {
    const float lo1 = c1 - e1 / 2, lo2 = c2 - e2 / 2, hi1 = c1 + e1 / 2, hi2 = c2 + e2 / 2;
    return (hi1 < hi2 ? hi1 : hi2) - (lo1 > lo2 ? lo1 : lo2);
}

static float yolo_iou(yolo_box a, yolo_box b) {
    const float w = span_overlap(a.x, a.w, b.x, b.w), h = span_overlap(a.y, a.h, b.y, b.h);
    const float inter = (w < 0 || h < 0) ? 0 : w * h;
    const float uni = a.w * a.h + b.w * b.h - inter;
    return inter / uni;
}

/* Box predicted at cell (col, row) by the anchor of size (aw, ah); `e` points at the cell's
 * first entry, entries are `hw` floats apart (get_yolo_box, reference :137-146). */
static yolo_box decode_box(const float *e, int hw, float aw, float ah, int col, int row, int lw,
                           int lh, int netw, int neth) {
    yolo_box b;
    b.x = (col + e[0]) / lw;
    b.y = (row + e[hw]) / lh;
    b.w = expf(e[2 * (size_t)hw]) * aw / netw;
    b.h = expf(e[3 * (size_t)hw]) * ah / neth;
    return b;
}

/* The truth boxes of one sample: x, y, w, h, class per box, the list ends at x == 0. */
static int truth_at(const float *label, int t, int stride, yolo_box *box) {
    const float *f = label + (size_t)t * stride;
    box->x = f[0]; box->y = f[1]; box->w = f[2]; box->h = f[3];
    return f[0] != 0;
}

/* Detection loss of one yolo node on the host mirrors (reference :251-416): fills dst->grad_data
 * (the gradient of the loss w.r.t. the activated head) and param->cost. Two sweeps per sample:
 * every predicted box pays for its objectness unless it overlaps some truth by more than 0.5;
 * every truth box then claims the cell it falls in, for the anchor that fits its size best when
 * that anchor belongs to this head, and sets box, objectness and class targets there. */
void bcnn_yolo_loss_host(bcnn_net *net, bcnn_node *node) {
    bcnn_yolo_param *param = (bcnn_yolo_param *)node->param;
    const bcnn_tensor *dst = &net->tensors[node->dst[0]], *label = &net->tensors[1];
    const int lw = dst->w, lh = dst->h, hw = lw * lh, coords = param->coords;
    const int group = coords + param->classes + 1, chw = param->num * group * hw;
    const int netw = net->tensors[0].w, neth = net->tensors[0].h;
    const float *anchors = param->biases.data;
    const float *out = dst->data;
    float *delta = dst->grad_data;
    memset(delta, 0, (size_t)dst->n * chw * sizeof(float));
    float sum_iou = 0, sum_class = 0, sum_obj = 0, sum_anyobj = 0, recall50 = 0, recall75 = 0;
    int count = 0;
    for (int b = 0; b < dst->n; ++b) {
        const float *truths = label->data + (size_t)b * param->truths;
        const size_t base = (size_t)b * chw;
        for (int row = 0; row < lh; ++row)
            for (int col = 0; col < lw; ++col)
                for (int a = 0; a < param->num; ++a) {
                    const size_t cell = base + (size_t)a * group * hw + (size_t)row * lw + col;
                    const int anchor = param->mask[a];
                    const yolo_box pred = decode_box(out + cell, hw, anchors[2 * anchor],
                                                     anchors[2 * anchor + 1], col, row, lw, lh,
                                                     netw, neth);
                    float best_iou = 0;
                    yolo_box truth;
                    for (int t = 0; t < param->max_boxes && truth_at(truths, t, coords + 1, &truth); ++t) {
                        const float iou = yolo_iou(pred, truth);
                        if (iou > best_iou) best_iou = iou;
                    }
                    const size_t obj = cell + (size_t)coords * hw;
                    delta[obj] = out[obj] - 0;
                    if (best_iou > 0.5) delta[obj] = 0;
                    sum_anyobj += out[obj];
                }
        yolo_box truth;
        for (int t = 0; t < param->max_boxes && truth_at(truths, t, coords + 1, &truth); ++t) {
            const int col = (int)(truth.x * lw), row = (int)(truth.y * lh);
            if (col < 0 || col >= lw || row < 0 || row >= lh) continue; /* the reference writes out of bounds here */
            yolo_box centred = truth;
            centred.x = centred.y = 0;
            float best_iou = 0;
            int best = 0;
            for (int n = 0; n < param->total; ++n) {
                const yolo_box prior = {0, 0, anchors[2 * n] / netw, anchors[2 * n + 1] / neth};
                const float iou = yolo_iou(prior, centred);
                if (iou > best_iou) { best_iou = iou; best = n; }
            }
            int a = -1;
            for (int k = 0; k < param->num && a < 0; ++k)
                if (param->mask[k] == best) a = k;
            if (a < 0) continue; /* another head owns this anchor */
            const size_t cell = base + (size_t)a * group * hw + (size_t)row * lw + col;
            const float aw = anchors[2 * best], ah = anchors[2 * best + 1];
            const float iou = yolo_iou(decode_box(out + cell, hw, aw, ah, col, row, lw, lh, netw, neth),
                                       truth);
            const float scale = (2 - truth.w * truth.h);
            const float target[4] = {truth.x * lw - col, truth.y * lh - row,
                                     logf(truth.w * netw / aw), logf(truth.h * neth / ah)};
            for (int e = 0; e < 4; ++e)
                delta[cell + (size_t)e * hw] = -scale * (target[e] - out[cell + (size_t)e * hw]);
            const size_t obj = cell + (size_t)coords * hw;
            sum_obj += out[obj];
            delta[obj] = out[obj] - 1;
            const int cls = (int)truths[(size_t)t * (coords + 1) + coords];
            const size_t first_class = obj + hw;
            if (delta[first_class]) { /* cell already claimed: only this class is pushed up */
                delta[first_class + (size_t)hw * cls] = out[first_class + (size_t)hw * cls] - 1;
                sum_class += out[first_class + (size_t)hw * cls];
            } else {
                for (int n = 0; n < param->classes; ++n) {
                    delta[first_class + (size_t)hw * n] = out[first_class + (size_t)hw * n] - ((n == cls) ? 1 : 0);
                    if (n == cls) sum_class += out[first_class + (size_t)hw * n];
                }
            }
            ++count;
            if (iou > 0.5f) recall50 += 1;
            if (iou > 0.75f) recall75 += 1;
            sum_iou += iou;
        }
    }
    double sq = 0; /* the reference sums in float lanes; the cost is a report, not an operand */
    for (size_t i = 0; i < (size_t)dst->n * chw; ++i) sq += (double)delta[i] * delta[i];
    *param->cost = powf(sqrtf((float)sq), 2);
    BCNN_INFO(net->log_ctx,
              "Yolo Avg IOU: %f Class: %f Obj: %f No Obj: %f .5R: %f, .75R: %f num_boxes: %d "
              "cost: %f\n", sum_iou / count, sum_class / count, sum_obj / count,
              sum_anyobj / ((float)hw * param->num * dst->n), recall50 / count, recall75 / count,
              count, *param->cost);
}

/* The reference's TRAIN data flow (:418-431) as a cross-check of the kernels: dst.data (device)
 * -> host, loss on the host, gradient -> dst.grad (device). Returns the cost, or -1 when the node
 * has no gradient buffer. */
static float yolo_loss_round_trip(bcnn_net *net, bcnn_node *node) {
    bcnn_tensor *dst = &net->tensors[node->dst[0]];
    void *stream = bcnn_stream(net);
    const size_t bytes = (size_t)bcnn_tensor_size(dst) * sizeof(float);
    if (!dst->grad_data_gpu || bcnn_tensor_ensure_host(dst) != BCNN_SUCCESS || !net->tensors[1].data)
        return -1.f;
    bcnn_cuda_check(bcnn_b200_memcpy_d2h(dst->data, dst->data_gpu, bytes, stream));
    bcnn_cuda_check(bcnn_b200_stream_sync(stream));
    bcnn_yolo_loss_host(net, node);
    bcnn_cuda_check(bcnn_b200_memcpy_h2d(dst->grad_data_gpu, dst->grad_data, bytes, stream));
    return *((bcnn_yolo_param *)node->param)->cost;
}

float bcnn_b200_yolo_loss_on_host(bcnn_net *net, int node_index) {
    if (node_index < 0 || node_index >= net->num_nodes ||
        net->nodes[node_index].type != BCNN_LAYER_YOLOV3)
        return -1.f;
    return yolo_loss_round_trip(net, &net->nodes[node_index]);
}

void bcnn_forward_yolo_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_yolo_param *param = (bcnn_yolo_param *)node->param;
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    void *stream = bcnn_stream(net);
    bcnn_cuda_check(bcnn_b200_yolo_activate(src->data_gpu, dst->data_gpu, src->n, param->num,
                                            param->classes, param->coords, src->h * src->w, stream));
    if (net->mode != BCNN_MODE_TRAIN || !dst->grad_data_gpu || !net->tensors[1].data_gpu) return;
    bcnn_cuda_check(bcnn_b200_yolo_loss_forward(
        dst->data_gpu, net->tensors[1].data_gpu, param->biases.data_gpu, param->mask_gpu,
        dst->grad_data_gpu, param->cost_gpu, dst->n, param->num, param->classes, param->coords,
        dst->w, dst->h, net->tensors[0].w, net->tensors[0].h, param->total, param->max_boxes,
        stream));
}

/* This step's loss of a yolo node: one float from the device (synchronises the stream). */
float bcnn_yolo_cost(bcnn_net *net, bcnn_node *node) {
    bcnn_yolo_param *param = (bcnn_yolo_param *)node->param;
    if (net->mode == BCNN_MODE_TRAIN && param->cost_gpu) {
        bcnn_cuda_check(bcnn_b200_memcpy_d2h(param->cost, param->cost_gpu, sizeof(float),
                                             bcnn_stream(net)));
        bcnn_cuda_check(bcnn_b200_stream_sync(bcnn_stream(net)));
    }
    return *param->cost;
}

/* src.grad += dst.grad (reference :432-447) */
void bcnn_backward_yolo_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    if (!src->grad_data_gpu || !dst->grad_data_gpu) return;
    bcnn_net_grad_prepare_accumulate(net, node->src[0]);
    bcnn_cuda_check(bcnn_b200_axpy(src->grad_data_gpu, dst->grad_data_gpu,
                                   (size_t)bcnn_tensor_size(src), 1.0f, bcnn_stream(net)));
}

void bcnn_release_param_yolo_layer(bcnn_node *node) {
    bcnn_yolo_param *param = (bcnn_yolo_param *)node->param;
    free(param->cost);
    free(param->mask);
    bcnn_b200_free(param->mask_gpu);
    bcnn_b200_free(param->cost_gpu);
    bcnn_tensor_destroy(&param->biases);
}

/* ---- detections: host post-processing, semantics of reference :470-639 ---- */
static int by_objectness_desc(const void *pa, const void *pb) {
    const float diff = ((const bcnn_output_detection *)pa)->objectness -
                       ((const bcnn_output_detection *)pb)->objectness;
    return diff < 0 ? 1 : (diff > 0 ? -1 : 0);
}

/* Visit every (cell, anchor) of every yolo head whose objectness exceeds `thresh`, in the
 * reference's order (heads in node order, cells row-major, anchors innermost). With dets == NULL
 * it only counts. */
static int yolo_collect(bcnn_net *net, int batch, float thresh, bcnn_output_detection *dets) {
    int count = 0;
    for (int k = 0; k < net->num_nodes; ++k) {
        if (net->nodes[k].type != BCNN_LAYER_YOLOV3) continue;
        const bcnn_yolo_param *param = (const bcnn_yolo_param *)net->nodes[k].param;
        const bcnn_tensor *dst = &net->tensors[net->nodes[k].dst[0]];
        const int hw = dst->w * dst->h, group = param->coords + param->classes + 1;
        const float *head = dst->data + (size_t)batch * dst->c * hw;
        for (int cell = 0; cell < hw; ++cell)
            for (int a = 0; a < param->num; ++a) {
                const float *entry = head + (size_t)a * group * hw + cell; /* stride hw per entry */
                const float objectness = entry[(size_t)param->coords * hw];
                if (!(objectness > thresh)) continue;
                if (dets) {
                    bcnn_output_detection *d = &dets[count];
                    const int col = cell % dst->w, row = cell / dst->w, anchor = param->mask[a];
                    d->x = (col + entry[0]) / dst->w;
                    d->y = (row + entry[hw]) / dst->h;
                    d->w = expf(entry[2 * (size_t)hw]) * param->biases.data[2 * anchor] /
                           net->tensors[0].w;
                    d->h = expf(entry[3 * (size_t)hw]) * param->biases.data[2 * anchor + 1] /
                           net->tensors[0].h;
                    d->objectness = objectness;
                    d->num_classes = param->classes;
                    for (int j = 0; j < param->classes; ++j) {
                        const float prob = objectness * entry[(size_t)(param->coords + 1 + j) * hw];
                        d->prob[j] = prob > thresh ? prob : 0;
                    }
                }
                ++count;
            }
    }
    return count;
}

bcnn_output_detection *bcnn_yolo_get_detections(bcnn_net *net, int batch, int w, int h, int netw,
                                                int neth, float thresh, int relative,
                                                int *num_dets) {
    *num_dets = 0;
    int num_classes = 0, max_classes = 0, extra_coords = 0;
    for (int k = 0; k < net->num_nodes; ++k) { /* one D2H refresh per head */
        if (net->nodes[k].type != BCNN_LAYER_YOLOV3) continue;
        const bcnn_yolo_param *param = (const bcnn_yolo_param *)net->nodes[k].param;
        if (!bcnn_get_tensor_by_index(net, net->nodes[k].dst[0])) return NULL;
        if (batch < 0 || batch >= net->tensors[net->nodes[k].dst[0]].n) return NULL;
        num_classes = param->classes; /* the last head's, as the reference */
        if (param->classes > max_classes) max_classes = param->classes;
        if (param->coords - 4 > extra_coords) extra_coords = param->coords - 4;
    }
    const int count = yolo_collect(net, batch, thresh, NULL);
    if (count == 0) return NULL;
    bcnn_output_detection *dets =
        (bcnn_output_detection *)calloc((size_t)count, sizeof(bcnn_output_detection));
    if (!dets) return NULL;
    for (int i = 0; i < count; ++i) { /* owned by the caller, released with free() */
        dets[i].prob = (float *)calloc((size_t)(max_classes > 0 ? max_classes : 1), sizeof(float));
        if (extra_coords > 0) dets[i].mask = (float *)calloc((size_t)extra_coords, sizeof(float));
    }
    yolo_collect(net, batch, thresh, dets);

    /* undo the letterbox the image was fitted into the net input with (:470-496) */
    int new_w, new_h;
    if (((float)netw / w) < ((float)neth / h)) {
        new_w = netw;
        new_h = (h * netw) / w;
    } else {
        new_h = neth;
        new_w = (w * neth) / h;
    }
    for (int i = 0; i < count; ++i) {
        dets[i].x = (dets[i].x - (netw - new_w) / 2. / netw) / ((float)new_w / netw);
        dets[i].y = (dets[i].y - (neth - new_h) / 2. / neth) / ((float)new_h / neth);
        dets[i].w *= (float)netw / new_w;
        dets[i].h *= (float)neth / new_h;
        if (!relative) {
            dets[i].x *= w;
            dets[i].w *= w;
            dets[i].y *= h;
            dets[i].h *= h;
        }
    }

    /* greedy NMS on objectness at IoU 0.45 (:511-546); suppressed boxes keep their slot with
     * objectness and class scores zeroed */
    int live = count; /* boxes with zero objectness (possible for thresh < 0) go to the back */
    for (int i = 0; i < live; ++i)
        if (dets[i].objectness == 0) {
            const bcnn_output_detection tmp = dets[i];
            dets[i--] = dets[--live];
            dets[live] = tmp;
        }
    qsort(dets, (size_t)live, sizeof(bcnn_output_detection), by_objectness_desc);
    for (int i = 0; i < live; ++i) {
        if (dets[i].objectness == 0) continue;
        const yolo_box a = {dets[i].x, dets[i].y, dets[i].w, dets[i].h};
        for (int j = i + 1; j < live; ++j) {
            if (dets[j].objectness == 0) continue;
            const yolo_box b = {dets[j].x, dets[j].y, dets[j].w, dets[j].h};
            if (yolo_iou(a, b) > 0.45f) {
                dets[j].objectness = 0;
                for (int c = 0; c < num_classes; ++c) dets[j].prob[c] = 0;
            }
        }
    }
    *num_dets = count;
    return dets;
}
