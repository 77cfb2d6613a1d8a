/*
 * bcnn_yolo.c -- YOLOv3 output node, inference side.
 *
 * Node layout and checks of jnbraun/bcnn src/layers/bcnn_yolo.c:15-107: src[0] = the head's
 * [N, num*(coords+classes+1), H, W] map, dst[0] the same shape, param = mask + anchor sizes. The
 * forward of PREDICT / VALID mode (:226-250: copy, logistic on centre offsets, objectness and
 * class scores) is one device kernel here; the reference copies the head to the host for it even
 * in its CUDA build (:418-431). The TRAIN-mode detection loss (:251-416) is host code in the
 * reference and outside this path: a TRAIN-mode net refuses the layer instead of
 * silently training without a loss. Box decoding / NMS (bcnn_yolo_get_detections, :470-639) is
 * the caller's post-processing and is not provided.
 */
#include "bcnn_yolo.h"

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
