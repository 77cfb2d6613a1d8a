/*
 * bcnn_yolo.h -- YOLOv3 output node; param layout of jnbraun/bcnn src/layers/bcnn_yolo.h:11-21.
 */
#ifndef BCNN_YOLO_H
#define BCNN_YOLO_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BCNN_DETECTION_MAX_BOXES 50 /* reference inc/bcnn/bcnn.h:233 */

typedef struct bcnn_yolo_param {
    int num;     /* boxes per cell (= number of masked anchors) */
    int classes;
    int coords;
    int truths;  /* label floats per sample */
    int max_boxes;
    int total;   /* anchors in the whole model */
    bcnn_tensor biases; /* anchor sizes, [1,1,1,2*total] */
    int *mask;
    float *cost;
    /* ---- B200 additions ---- */
    int *mask_gpu;   /* device copy of mask */
    float *cost_gpu; /* [0] = this step's loss, then the reduction partials */
} bcnn_yolo_param;

void bcnn_forward_yolo_layer(bcnn_net *net, bcnn_node *node);
void bcnn_backward_yolo_layer(bcnn_net *net, bcnn_node *node);
void bcnn_release_param_yolo_layer(bcnn_node *node);
/* TRAIN-mode detection loss on the host mirrors of dst (activated head, in), the label (in) and
 * dst's gradient (out); param->cost receives the loss. */
void bcnn_yolo_loss_host(bcnn_net *net, bcnn_node *node);
/* This step's loss of a yolo node (TRAIN: one float read back from the device). */
float bcnn_yolo_cost(bcnn_net *net, bcnn_node *node);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_YOLO_H */
