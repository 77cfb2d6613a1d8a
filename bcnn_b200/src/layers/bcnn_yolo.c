/*
 * bcnn_yolo.c -- YOLOv3 output node, inference side.
 *
 * Node layout and checks of jnbraun/bcnn src/layers/bcnn_yolo.c:15-107: src[0] = the head's
 * [N, num*(coords+classes+1), H, W] map, dst[0] the same shape, param = mask + anchor sizes. The
 * forward of PREDICT / VALID mode (:226-250: copy, logistic on centre offsets, objectness and
 * class scores) is one device kernel here; the reference copies the head to the host for it even
 * in its CUDA build (:418-431). The TRAIN-mode detection loss (:251-416) is host code in the
 * reference and outside this path: a TRAIN-mode net refuses the layer instead of
 * silently training without a loss. bcnn_yolo_get_detections (:470-639: box decoding, letterbox
 * correction, objectness NMS) is host post-processing in the reference and is host code here
 * too, on the heads' host mirrors after one D2H refresh per head.
 */
#include "bcnn_yolo.h"

#include <math.h>

#include "bcnn_tensor.h"

bcnn_status bcnn_add_yolo_layer(bcnn_net *net, int num_boxes_per_cell, int classes, int coords,
                                int total, int *mask, float *anchors, const char *src_id,
                                const char *dst_id) {
    bcnn_node node = {0};
    BCNN_CHECK_AND_LOG(net->log_ctx, net->num_nodes >= 1, BCNN_INVALID_PARAMETER,
                       "Yolo layer can't be the first layer of the network\n");
    BCNN_CHECK_AND_LOG(net->log_ctx, net->mode != BCNN_MODE_TRAIN, BCNN_INVALID_PARAMETER,
                       "Yolo layer: the detection loss (TRAIN mode) is not part of the B200 path\n");
    int src = bcnn_net_find_src(net, src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx, src >= 0, BCNN_INVALID_PARAMETER,
                       "Yolo layer: invalid input node name %s\n", src_id);
    BCNN_CHECK_AND_LOG(net->log_ctx,
                       num_boxes_per_cell > 0 && classes >= 0 && coords >= 2 && total > 0 && mask,
                       BCNN_INVALID_PARAMETER, "Yolo layer: invalid box / class / anchor counts\n");
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
    node.forward = bcnn_forward_yolo_layer;
    node.backward = bcnn_backward_yolo_layer;
    node.release_param = bcnn_release_param_yolo_layer;

    /* VALID: the label holds up to 50 boxes x (x, y, w, h, class) per sample (:69-74) */
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

void bcnn_forward_yolo_layer(bcnn_net *net, bcnn_node *node) {
    bcnn_yolo_param *param = (bcnn_yolo_param *)node->param;
    bcnn_tensor *src = &net->tensors[node->src[0]], *dst = &net->tensors[node->dst[0]];
    bcnn_cuda_check(bcnn_b200_yolo_activate(src->data_gpu, dst->data_gpu, src->n, param->num,
                                            param->classes, param->coords, src->h * src->w,
                                            bcnn_stream(net)));
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
    bcnn_tensor_destroy(&param->biases);
}

/* ---- detections: host post-processing, semantics of reference :470-639 ---- */
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
