/*
 * bcnn_tensor.c -- NCHW float32 tensors with device-resident storage.
 * Behaviour follows jnbraun/bcnn src/bcnn_tensor.c:40-145 (shape bookkeeping, fillers
 * with libc rand() for Xavier, gradient buffers only outside PREDICT mode); the storage
 * policy is B200-first: cudaMalloc'd buffers are the truth, host mirrors are pinned and
 * lazily created for layer outputs (a ResNet-50 batch-256 replica holds > 40 GB of
 * activations that should never be duplicated in host RAM).
 */
#include "bcnn_tensor.h"

#include <math.h>

#include "bcnn_utils.h"

void bcnn_tensor_set_shape(bcnn_tensor *t, int n, int c, int h, int w, int has_grad) {
    t->n = n; t->c = c; t->h = h; t->w = w;
    t->has_grad = has_grad;
}

int bcnn_tensor_size(const bcnn_tensor *t) { return t->w * t->h * t->c * t->n; }
int bcnn_tensor_size3d(const bcnn_tensor *t) { return t->w * t->h * t->c; }
int bcnn_tensor_size2d(const bcnn_tensor *t) { return t->w * t->h; }

static float *host_zeros(size_t n) {
    float *p = (float *)bcnn_b200_malloc_host(n * sizeof(float));
    if (p) memset(p, 0, n * sizeof(float));
    return p;
}

bcnn_status bcnn_tensor_ensure_host(bcnn_tensor *t) {
    size_t size = (size_t)bcnn_tensor_size(t);
    if (size == 0) return BCNN_SUCCESS;
    if (!t->data && t->data_gpu) {
        t->data = host_zeros(size);
        BCNN_CHECK(t->data != NULL, BCNN_FAILED_ALLOC);
    }
    if (!t->grad_data && t->grad_data_gpu) {
        t->grad_data = host_zeros(size);
        BCNN_CHECK(t->grad_data != NULL, BCNN_FAILED_ALLOC);
    }
    return BCNN_SUCCESS;
}

bcnn_status bcnn_tensor_allocate_buffer(bcnn_tensor *t, int net_state, size_t size) {
    bcnn_tensor_free(t);
    if (size == 0) return BCNN_INVALID_PARAMETER;
    t->data_gpu = (float *)bcnn_b200_malloc(size * sizeof(float));
    BCNN_CHECK(t->data_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);
    if (t->has_grad && net_state != BCNN_MODE_PREDICT) {
        t->grad_data_gpu = (float *)bcnn_b200_malloc(size * sizeof(float));
        BCNN_CHECK(t->grad_data_gpu != NULL, BCNN_CUDA_FAILED_ALLOC);
    }
    return BCNN_SUCCESS;
}

bcnn_status bcnn_tensor_allocate(bcnn_tensor *t, int net_state) {
    return bcnn_tensor_allocate_buffer(t, net_state, (size_t)t->n * t->c * t->h * t->w);
}

void bcnn_tensor_create(bcnn_tensor *t, int n, int c, int h, int w, int has_grad,
                        const char *name, int net_state) {
    bcnn_tensor_set_shape(t, n, c, h, w, has_grad);
    if (bcnn_tensor_allocate(t, net_state) == BCNN_SUCCESS) bcnn_tensor_ensure_host(t);
    free(t->name);
    t->name = bcnn_strdup_(name);
}

void bcnn_tensor_fill(bcnn_tensor *t, bcnn_tensor_filler filler) {
    if (!t->data) return;
    int sz = bcnn_tensor_size(t);
    switch (filler.type) {
        case BCNN_FILLER_XAVIER: {
            float amp = sqrtf(3.0f / filler.range);
            for (int i = 0; i < sz; ++i) t->data[i] = amp * (2 * ((float)rand() / RAND_MAX) - 1);
            break;
        }
        case BCNN_FILLER_MSRA: { /* gaussian via Box-Muller on libc rand() */
            float amp = sqrtf(2.0f / filler.range);
            for (int i = 0; i < sz; ++i) {
                float u1 = ((float)rand() + 1.0f) / ((float)RAND_MAX + 2.0f);
                float u2 = (float)rand() / RAND_MAX;
                t->data[i] = amp * sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
            }
            break;
        }
        case BCNN_FILLER_FIXED:
            for (int i = 0; i < sz; ++i) t->data[i] = filler.value;
            break;
    }
    bcnn_cuda_check(bcnn_b200_memcpy_h2d(t->data_gpu, t->data, (size_t)sz * sizeof(float), NULL));
    bcnn_cuda_check(bcnn_b200_stream_sync(NULL));
}

void bcnn_tensor_free(bcnn_tensor *t) {
    bcnn_b200_free_host(t->data);
    t->data = NULL;
    bcnn_b200_free_host(t->grad_data);
    t->grad_data = NULL;
    bcnn_b200_free(t->data_gpu);
    t->data_gpu = NULL;
    bcnn_b200_free(t->grad_data_gpu);
    t->grad_data_gpu = NULL;
}

void bcnn_tensor_destroy(bcnn_tensor *t) {
    bcnn_tensor_free(t);
    bcnn_tensor_set_shape(t, 0, 0, 0, 0, 0);
    free(t->name);
    t->name = NULL;
}
