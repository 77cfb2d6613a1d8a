/*
 * bcnn_tensor.h -- tensor helpers of the host runtime; same entry points as
 * jnbraun/bcnn src/bcnn_tensor.h:41-62. Device buffers are the primary storage;
 * host mirrors are pinned and, for layer outputs, allocated on first fetch.
 */
#ifndef BCNN_TENSOR_H
#define BCNN_TENSOR_H

#include <bcnn/bcnn.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tensor_filler {
    int range;
    float value;
    bcnn_filler_type type;
} bcnn_tensor_filler;

/* Parameter-style tensor: shape + device buffers + eager host mirror + name. */
void bcnn_tensor_create(bcnn_tensor *t, int n, int c, int h, int w, int has_grad,
                        const char *name, int net_state);
void bcnn_tensor_fill(bcnn_tensor *t, bcnn_tensor_filler filler);
void bcnn_tensor_destroy(bcnn_tensor *t);
void bcnn_tensor_set_shape(bcnn_tensor *t, int n, int c, int h, int w, int has_grad);
/* Layer-output tensor: device buffers only; host mirror comes lazily. */
bcnn_status bcnn_tensor_allocate(bcnn_tensor *t, int net_state);
bcnn_status bcnn_tensor_allocate_buffer(bcnn_tensor *t, int net_state, size_t size);
/* Make sure host mirrors exist (pinned, zero-filled on creation). */
bcnn_status bcnn_tensor_ensure_host(bcnn_tensor *t);
void bcnn_tensor_free(bcnn_tensor *t);
int bcnn_tensor_size(const bcnn_tensor *t);
int bcnn_tensor_size3d(const bcnn_tensor *t);
int bcnn_tensor_size2d(const bcnn_tensor *t);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_TENSOR_H */
