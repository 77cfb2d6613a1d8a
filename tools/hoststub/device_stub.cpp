// device_stub.cpp -- DEVELOPMENT AID, not part of the product and never shipped or loaded by it.
//
// Host-memory stand-ins for csrc/device.cu and csrc/optim.cu so that the C host runtime
// (bcnn_b200/src/**/*.c: graph construction, weight files, optimizer dispatch) can be exercised
// in a container without a GPU: "device" buffers are malloc'd, copies are memcpy, the two
// optimizer kernels are restated in scalar C with the same operation order. Every other kernel
// still lives in its .cu object and fails at launch, so forward / backward cannot run here.
// tools/hoststub/build.sh links it into tools/hoststub/libbcnn_hoststub.so (git-ignored,
// gpurun-ignored).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace b200 {
unsigned long long g_launch_count = 0;
int sm_count() { return 148; }
}  // namespace b200

extern "C" {
int bcnn_b200_set_device(int) { return 0; }
int bcnn_b200_device_count(void) { return 1; }
int bcnn_b200_sm_count(void) { return 148; }
void *bcnn_b200_malloc(size_t bytes) { return calloc(1, bytes ? bytes : 4); }
void bcnn_b200_free(void *p) { free(p); }
void *bcnn_b200_malloc_host(size_t bytes) { return malloc(bytes ? bytes : 4); }
void bcnn_b200_free_host(void *p) { free(p); }
int bcnn_b200_memcpy_h2d(void *d, const void *s, size_t n, void *) { memcpy(d, s, n); return 0; }
int bcnn_b200_memcpy_d2h(void *d, const void *s, size_t n, void *) { memcpy(d, s, n); return 0; }
int bcnn_b200_memcpy_d2d(void *d, const void *s, size_t n, void *) { memmove(d, s, n); return 0; }
void *bcnn_b200_stream_create(void) { return malloc(4); }
void bcnn_b200_stream_destroy(void *s) { free(s); }
int bcnn_b200_stream_sync(void *) { return 0; }
void *bcnn_b200_event_create(void) { return malloc(4); }
void bcnn_b200_event_destroy(void *e) { free(e); }
int bcnn_b200_event_record(void *, void *) { return 0; }
int bcnn_b200_stream_wait_event(void *, void *) { return 0; }
float bcnn_b200_event_elapsed_ms(void *, void *) { return 0.f; }
int bcnn_b200_graph_begin(void *) { return 1; }  // no capture on the host: the runtime stays eager
void *bcnn_b200_graph_end(void *) { return nullptr; }
int bcnn_b200_graph_launch(void *, unsigned long long, void *) { return 1; }
void bcnn_b200_graph_destroy(void *) {}
const char *bcnn_b200_error_string(int) { return "host stub: no device"; }
uint64_t bcnn_b200_launch_count(void) { return b200::g_launch_count; }
int bcnn_b200_fill_f32(float *x, size_t n, float v, void *) {
    for (size_t i = 0; i < n; ++i) x[i] = v;
    return 0;
}
int bcnn_b200_axpy(float *y, const float *x, size_t n, float a, void *) {
    for (size_t i = 0; i < n; ++i) y[i] = fmaf(a, x[i], y[i]);
    return 0;
}
// scalar restatements of sgd_kernel / adam_kernel (csrc/optim.cu), same operation order
int bcnn_b200_sgd_update(float *w, float *g, size_t n, float wd_scale, float step, float g_scale,
                         void *) {
    for (size_t j = 0; j < n; ++j) {
        volatile float t = wd_scale * w[j];
        float gv = g[j] + t;
        t = step * gv;
        w[j] = w[j] + t;
        g[j] = gv * g_scale;
    }
    return 0;
}
// layout of bcnn_b200_sgd_batch (include/bcnn_b200.h)
struct stub_sgd_batch {
    float *w[96];
    float *g[96];
    unsigned int n[96];
    float wd_scale[96];
    unsigned int first_block[97];
    int count;
    float step, g_scale;
};
int bcnn_b200_sgd_update_multi(const stub_sgd_batch *b, void *stream) {
    for (int i = 0; b && i < b->count; ++i)
        bcnn_b200_sgd_update(b->w[i], b->g[i], b->n[i], b->wd_scale[i], b->step, b->g_scale, stream);
    return 0;
}
int bcnn_b200_adam_update(float *w, float *g, float *m, float *v, size_t n, float wd_scale,
                          float beta1, float beta2, float alpha, void *) {
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    const size_t tail = n & ~(size_t)7;
    for (size_t j = 0; j < n; ++j) {
        volatile float t = wd_scale * w[j];
        volatile float gv = g[j] + t;
        volatile float a = gv * omb1, b = m[j] * beta1;
        m[j] = a + b;
        volatile float g2 = gv * gv;
        a = g2 * omb2;
        b = v[j] * beta2;
        v[j] = a + b;
        volatile float den = sqrtf(v[j]) + 0.0000001f;
        float q = (j >= tail && !(fabsf(den) > 0.00001f)) ? 0.0f : m[j] / den;
        t = alpha * q;
        w[j] = w[j] + t;
        g[j] = 0.0f;
    }
    return 0;
}
int bcnn_b200_softmax_forward(const float *, float *, int, int, int, void *) { return 999; }
int bcnn_b200_cost_forward(const float *, const float *, float *, float *, int, int, int, void *) {
    return 999;
}
}
