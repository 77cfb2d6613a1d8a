/*
 * bcnn.h -- public C99 API of bcnn_b200, the B200-native CNN layer hot path.
 *
 * ABI contract: every enum value, the bcnn_tensor layout (BCNN_USE_CUDA flavour)
 * and every function signature below is binary compatible with jnbraun/bcnn's
 * inc/bcnn/bcnn.h (enums :90-235, bcnn_tensor :242-255, functions :285-1043), so
 * a program written against bcnn links against libbcnn_b200.so unchanged for the
 * layer hot path and for the files either side of it (weight files, cfg files,
 * yolo detections). The rest of the reference API (file loaders, augmentation,
 * deconv / dropout / lrn layers, ...) is out of scope -- see DESIGN.md.
 *
 * This library is always the CUDA flavour: BCNN_USE_CUDA is forced on, so
 * bcnn_tensor carries data_gpu / grad_data_gpu. Host mirrors of layer outputs
 * are materialised lazily by bcnn_get_tensor_by_index / _by_name.
 */
#ifndef BCNN_H
#define BCNN_H

#include <stddef.h>
#include <stdint.h>

#ifndef BCNN_USE_CUDA
#define BCNN_USE_CUDA 1
#endif

#if defined(__GNUC__)
#define BCNN_API __attribute__((visibility("default")))
#else
#define BCNN_API
#endif

#define BCNN_VERSION_MAJOR 0
#define BCNN_VERSION_MINOR 2
#define BCNN_VERSION_PATCH 0

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcnn_net bcnn_net;       /* opaque: src/bcnn_net.h */
typedef struct bcnn_tensor bcnn_tensor; /* defined below */

/* -- enumerations (values are part of the ABI; order must not change) -- */

typedef enum {
    BCNN_SUCCESS, BCNN_INVALID_PARAMETER, BCNN_INVALID_DATA, BCNN_INVALID_MODEL,
    BCNN_FAILED_ALLOC, BCNN_INTERNAL_ERROR, BCNN_CUDA_FAILED_ALLOC,
    BCNN_UNKNOWN_ERROR
} bcnn_status;

typedef enum {
    BCNN_MODE_PREDICT, /* inference: no gradients, BN folded into scale/bias */
    BCNN_MODE_TRAIN,   /* forward + backward + update */
    BCNN_MODE_VALID    /* forward against ground truth, running BN statistics */
} bcnn_mode;

typedef enum {
    BCNN_LOAD_MNIST, BCNN_LOAD_CIFAR10, BCNN_LOAD_CLASSIFICATION_LIST,
    BCNN_LOAD_REGRESSION_LIST, BCNN_LOAD_DETECTION_LIST, BCNN_NUM_LOADERS
} bcnn_loader_type;

typedef enum {
    BCNN_LR_DECAY_CONSTANT, BCNN_LR_DECAY_STEP, BCNN_LR_DECAY_INV,
    BCNN_LR_DECAY_EXP, BCNN_LR_DECAY_POLY, BCNN_LR_DECAY_SIGMOID
} bcnn_lr_decay;

typedef enum {
    BCNN_LAYER_CONV2D, BCNN_LAYER_TRANSPOSE_CONV2D, BCNN_LAYER_DEPTHWISE_CONV2D,
    BCNN_LAYER_ACTIVATION, BCNN_LAYER_FULL_CONNECTED, BCNN_LAYER_MAXPOOL,
    BCNN_LAYER_AVGPOOL, BCNN_LAYER_SOFTMAX, BCNN_LAYER_DROPOUT,
    BCNN_LAYER_BATCHNORM, BCNN_LAYER_LRN, BCNN_LAYER_CONCAT, BCNN_LAYER_ELTWISE,
    BCNN_LAYER_UPSAMPLE, BCNN_LAYER_YOLOV3, BCNN_LAYER_RESHAPE, BCNN_LAYER_COST
} bcnn_layer_type;

typedef enum {
    BCNN_ACT_NONE, BCNN_ACT_TANH, BCNN_ACT_RELU, BCNN_ACT_RAMP,
    BCNN_ACT_SOFTPLUS,
    BCNN_ACT_LRELU, /* negative slope is 0.1, as computed by the reference */
    BCNN_ACT_ABS, BCNN_ACT_CLAMP, BCNN_ACT_PRELU, BCNN_ACT_LOGISTIC
} bcnn_activation;

typedef enum { BCNN_LOSS_EUCLIDEAN, BCNN_LOSS_LIFTED_STRUCT } bcnn_loss;

typedef enum {
    BCNN_METRIC_ERROR_RATE, BCNN_METRIC_LOGLOSS, BCNN_METRIC_SSE,
    BCNN_METRIC_MSE, BCNN_METRIC_CRPS, BCNN_METRIC_DICE
} bcnn_loss_metric;

/* Pooling padding policy (convolutions take an explicit pad). */
typedef enum { BCNN_PADDING_SAME, BCNN_PADDING_VALID, BCNN_PADDING_CAFFE } bcnn_padding;

typedef enum { BCNN_OPTIM_SGD, BCNN_OPTIM_ADAM } bcnn_optimizer;

typedef enum {
    BCNN_LOG_INFO = 0, BCNN_LOG_WARNING = 1, BCNN_LOG_ERROR = 2, BCNN_LOG_SILENT = 3
} bcnn_log_level;

typedef enum bcnn_filler_type {
    BCNN_FILLER_FIXED, BCNN_FILLER_XAVIER, BCNN_FILLER_MSRA
} bcnn_filler_type;

typedef void (*bcnn_log_callback)(const char *fmt, ...);

/* One detection of bcnn_yolo_get_detections (reference inc/bcnn/bcnn.h:260-266). */
typedef struct bcnn_output_detection {
    int num_classes;
    float x, y, w, h;
    float *prob;
    float *mask;
    float objectness;
} bcnn_output_detection;

/* NCHW float32 tensor. Element count is an int (< 2^31). */
struct bcnn_tensor {
    int n, c, h, w;
    int has_grad;
    char *name;
    float *data;          /* host mirror (may be NULL until first fetched) */
    float *grad_data;     /* host mirror of the gradient */
    float *data_gpu;      /* device buffer: the one the kernels work on */
    float *grad_data_gpu; /* device gradient buffer */
};

/* -- net lifetime / configuration -- */
BCNN_API bcnn_status bcnn_init_net(bcnn_net **net, bcnn_mode mode);
BCNN_API void bcnn_end_net(bcnn_net **net);
BCNN_API void bcnn_set_log_context(bcnn_net *net, bcnn_log_callback fct,
                                   bcnn_log_level level);
BCNN_API bcnn_status bcnn_set_num_threads(bcnn_net *net, int num_threads,
                                          const int *cpu_ids);
BCNN_API int bcnn_get_num_threads(bcnn_net *net);
BCNN_API void bcnn_set_input_shape(bcnn_net *net, int width, int height,
                                   int channels, int batch_size);
BCNN_API int bcnn_get_batch_size(bcnn_net *net);
/* One more input tensor (reference inc/bcnn/bcnn.h:368, src/bcnn_net.c:260-278). */
BCNN_API bcnn_status bcnn_add_input(bcnn_net *net, int width, int height, int channels,
                                    const char *name);
BCNN_API bcnn_status bcnn_compile_net(bcnn_net *net);
BCNN_API bcnn_status bcnn_set_mode(bcnn_net *net, bcnn_mode mode);

/* -- optimizer -- */
BCNN_API void bcnn_set_sgd_optimizer(bcnn_net *net, float learning_rate,
                                     float momentum);
BCNN_API void bcnn_set_adam_optimizer(bcnn_net *net, float learning_rate,
                                      float beta1, float beta2);
BCNN_API void bcnn_set_learning_rate_policy(bcnn_net *net, bcnn_lr_decay decay_type,
                                            float gamma, float scale, float power,
                                            int max_batches, int step);
BCNN_API void bcnn_set_weight_regularizer(bcnn_net *net, float weight_decay);

/* -- weight files (reference inc/bcnn/bcnn.h:421, :448; src/bcnn_net.c:597-681, :1485-1558) --
 * bcnn_save_weights writes the reference's .bcnnmodel layout byte for byte.
 * bcnn_load_weights reads .bcnnmodel and Darknet *.weights files; in BCNN_MODE_PREDICT it
 * folds the batch-norm running statistics into scales / bias as the reference's CPU build
 * does. A truncated file returns BCNN_INVALID_MODEL (the reference logs and returns success). */
/* Build the net from a .cfg / .conf file (bcnn dialect, or Darknet when model_path ends in
 * .weights), then load model_path if given (reference inc/bcnn/bcnn.h:437,
 * src/bcnn_net.c:1114-1218). Sections for layers this path does not have ([deconv], [lrn],
 * [dropout]) fail with BCNN_INVALID_PARAMETER. */
BCNN_API bcnn_status bcnn_load_net(bcnn_net *net, const char *config_path,
                                   const char *model_path);
BCNN_API bcnn_status bcnn_load_weights(bcnn_net *net, const char *model_path);
BCNN_API bcnn_status bcnn_save_weights(bcnn_net *net, const char *filename);

/* -- the three loops -- */
BCNN_API void bcnn_forward(bcnn_net *net);
BCNN_API void bcnn_backward(bcnn_net *net);
BCNN_API void bcnn_update(bcnn_net *net);
/* The train / predict calls of the reference (inc/bcnn/bcnn.h:683, :699; src/bcnn_net.c:
 * 452-483) without its file loader: the batch is taken from the pinned host mirrors of the
 * input tensors and the label (bcnn_get_tensor_by_name(net, "input")->data, ...), uploaded,
 * and the loss (mean over the cost nodes) is read back. */
BCNN_API float bcnn_train_on_batch(bcnn_net *net);
BCNN_API float bcnn_predict_on_batch(bcnn_net *net, bcnn_tensor **out);

/* -- tensor access (refreshes the host mirrors of data and grad) -- */
BCNN_API int bcnn_get_tensor_index_by_name(bcnn_net *net, const char *name);
BCNN_API bcnn_tensor *bcnn_get_tensor_by_index(bcnn_net *net, int index);
BCNN_API bcnn_tensor *bcnn_get_tensor_by_name(bcnn_net *net, const char *name);

/* -- layers on the hot path -- */
BCNN_API bcnn_status bcnn_add_convolutional_layer(
    bcnn_net *net, int num_filters, int size, int stride, int pad, int num_groups,
    int batch_norm, bcnn_filler_type init, bcnn_activation activation, int quantize,
    const char *src_id, const char *dst_id);
BCNN_API bcnn_status bcnn_add_depthwise_conv_layer(
    bcnn_net *net, int size, int stride, int pad, int batch_norm,
    bcnn_filler_type init, bcnn_activation activation, const char *src_id,
    const char *dst_id);
BCNN_API bcnn_status bcnn_add_batchnorm_layer(bcnn_net *net, const char *src_id,
                                              const char *dst_id);
BCNN_API bcnn_status bcnn_add_activation_layer(bcnn_net *net, bcnn_activation type,
                                               const char *id);
BCNN_API bcnn_status bcnn_add_maxpool_layer(bcnn_net *net, int size, int stride,
                                            bcnn_padding padding, const char *src_id,
                                            const char *dst_id);
BCNN_API bcnn_status bcnn_add_avgpool_layer(bcnn_net *net, const char *src_id,
                                            const char *dst_id);

/* -- glue layers either side of the path -- */
BCNN_API bcnn_status bcnn_add_fullc_layer(bcnn_net *net, int output_size,
                                          bcnn_filler_type init,
                                          bcnn_activation activation, int quantize,
                                          const char *src_id, const char *dst_id);
BCNN_API bcnn_status bcnn_add_softmax_layer(bcnn_net *net, const char *src_id,
                                            const char *dst_id);
BCNN_API bcnn_status bcnn_add_eltwise_layer(bcnn_net *net, bcnn_activation activation,
                                            const char *src_id1, const char *src_id2,
                                            const char *dst_id);
/* Concatenation along the channel axis (reference inc/bcnn/bcnn.h:956) and nearest upsampling
 * by an integer factor (:999): the glue of YOLOv3-tiny's second head. */
BCNN_API bcnn_status bcnn_add_concat_layer(bcnn_net *net, int num_src, char *const *src_ids,
                                           const char *dst_id);
BCNN_API bcnn_status bcnn_add_upsample_layer(bcnn_net *net, int size, const char *src_id,
                                             const char *dst_id);
/* YOLOv3 output layer (reference inc/bcnn/bcnn.h:1039, src/layers/bcnn_yolo.c:15-107).
 * The head activation runs on the device in every mode. The detection loss of TRAIN mode is host
 * code in the reference even in its CUDA build (:418-431) and is host code here: head to the
 * host, loss against the [N,1,1,250] box label, gradient back. */
BCNN_API bcnn_status bcnn_add_yolo_layer(bcnn_net *net, int num_boxes_per_cell, int classes,
                                         int coords, int total, int *mask, float *anchors,
                                         const char *src_id, const char *dst_id);
/* Boxes above `thresh` of sample `batch` from every yolo head, letterbox-corrected for a
 * width x height image and NMS-filtered (reference inc/bcnn/bcnn.h:718, src/layers/bcnn_yolo.c:
 * 548-639). Host post-processing after one device -> host copy per head. The caller frees
 * dets[i].prob, dets[i].mask and the array. NULL and *num_dets = 0 when nothing passes. */
BCNN_API bcnn_output_detection *bcnn_yolo_get_detections(bcnn_net *net, int batch, int width,
                                                         int height, int netw, int neth,
                                                         float thresh, int relative,
                                                         int *num_dets);
BCNN_API bcnn_status bcnn_add_cost_layer(bcnn_net *net, bcnn_loss loss,
                                         bcnn_loss_metric loss_metric, float scale,
                                         const char *src_id, const char *label_id,
                                         const char *dst_id);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_H */
