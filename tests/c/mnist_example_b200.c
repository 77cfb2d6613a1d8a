/* A C99 program written against bcnn's public API, compiled with
 *     gcc -std=gnu99 -DBCNN_USE_CUDA -Iinclude tests/c/mnist_example_b200.c -Lbcnn_b200 -lbcnn_b200
 * and run by tests/test_c_program.py. create_network() is the network of the reference's
 * examples/mnist/mnist_example.c:30-55 call for call (batch 64 as BASELINE.json names it); the
 * training loop is the example's bcnn_train_on_batch loop on synthetic images (the file loader
 * is out of scope). The second half calls the net-less entry points with the reference's own
 * prototypes (src/layers/bcnn_activation_layer.h:48-51, src/kernels/bcnn_mat.h:258-309), declared
 * here exactly as the reference declares them, to prove they link and compute. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <bcnn/bcnn.h>
#include <bcnn_b200.h>

/* prototypes as in the reference's internal headers */
void bcnn_forward_activation_gpu(float *x, int sz, bcnn_activation a);
void bcnn_backward_activation_gpu(float *x, float *dx, int sz, bcnn_activation a);
void bcnn_cuda_fill_f32(int n, float alpha, float *x, int incx);
void bcnn_cuda_copy_f32(int n, float *x, int incx, float *y, int incy);
void bcnn_cuda_axpy(int n, float alpha, float *x, int incx, float *y, int incy);
void bcnn_cuda_scal(int n, float alpha, float *x, int incx);
void bcnn_cuda_add_bias(float *output, float *bias, int batch_size, int num_channels, int spatial_size);
void bcnn_cuda_grad_bias(float *grad_bias, float *grad_data, int batch_size, int num_channels,
                         int spatial_size);
void bcnn_cuda_gemm(int trans_a, int trans_b, int m, int n, int k, float alpha, float *a, int lda,
                    float *b, int ldb, float beta, float *c, int ldc);

static int create_network(bcnn_net *net) {
    bcnn_set_input_shape(net, 28, 28, 1, 64);
    bcnn_add_convolutional_layer(net, 32, 3, 1, 1, 1, 0, BCNN_FILLER_XAVIER, BCNN_ACT_RELU, 0, "input",
                                 "conv1");
    bcnn_add_batchnorm_layer(net, "conv1", "bn1");
    bcnn_add_maxpool_layer(net, 2, 2, BCNN_PADDING_SAME, "bn1", "pool1");
    bcnn_add_convolutional_layer(net, 32, 3, 1, 1, 1, 0, BCNN_FILLER_XAVIER, BCNN_ACT_RELU, 0, "pool1",
                                 "conv2");
    bcnn_add_batchnorm_layer(net, "conv2", "bn2");
    bcnn_add_maxpool_layer(net, 2, 2, BCNN_PADDING_SAME, "bn2", "pool2");
    bcnn_add_fullc_layer(net, 256, BCNN_FILLER_XAVIER, BCNN_ACT_RELU, 0, "pool2", "fc1");
    bcnn_add_batchnorm_layer(net, "fc1", "bn3");
    bcnn_add_fullc_layer(net, 10, BCNN_FILLER_XAVIER, BCNN_ACT_RELU, 0, "bn3", "fc2");
    bcnn_add_softmax_layer(net, "fc2", "softmax");
    bcnn_add_cost_layer(net, BCNN_LOSS_EUCLIDEAN, BCNN_METRIC_ERROR_RATE, 1.0f, "softmax", "label",
                        "cost");
    return 0;
}

static unsigned lcg(unsigned *s) { return *s = *s * 1664525u + 1013904223u; }

static int close_to(const char *what, float got, float want) {
    if (fabsf(got - want) <= 1e-5f * fmaxf(1.0f, fabsf(want))) return 0;
    fprintf(stderr, "%s: got %g, want %g\n", what, got, want);
    return 1;
}

int main(void) {
    bcnn_net *net = NULL;
    if (bcnn_init_net(&net, BCNN_MODE_TRAIN) != BCNN_SUCCESS) return 2;
    bcnn_set_log_context(net, NULL, BCNN_LOG_SILENT);
    create_network(net);
    bcnn_set_sgd_optimizer(net, 0.003f, 0.9f); /* mnist_example.c:136-139 */
    bcnn_set_weight_regularizer(net, 0.0005f);
    if (bcnn_compile_net(net) != BCNN_SUCCESS) return 3;

    bcnn_tensor *in = bcnn_get_tensor_by_name(net, "input");
    bcnn_tensor *lab = bcnn_get_tensor_by_name(net, "label");
    if (!in || !lab || !in->data || !lab->data || !in->data_gpu) return 4;
    const int batch = in->n, img = in->c * in->h * in->w, classes = lab->c;
    unsigned seed = 12345u;
    float first = 0.f, last = 0.f;
    for (int step = 0; step < 3; ++step) {
        /* a class-dependent pattern plus noise, so three steps can lower the loss */
        for (int b = 0; b < batch; ++b) {
            const int cls = b % classes;
            for (int i = 0; i < img; ++i)
                in->data[b * img + i] = ((i + 3 * cls) % 10 < 3 ? 0.8f : -0.2f) +
                                        0.1f * ((float)(lcg(&seed) >> 8) / 8388608.0f - 1.0f);
            for (int c = 0; c < classes; ++c) lab->data[b * classes + c] = (c == cls) ? 1.0f : 0.0f;
        }
        const float loss = bcnn_train_on_batch(net);
        printf("step %d loss %f\n", step, loss);
        if (!isfinite(loss)) return 5;
        if (step == 0) first = loss;
        last = loss;
    }
    bcnn_tensor *w = bcnn_get_tensor_by_name(net, "input_w"); /* refreshed host copy */
    float wsum = 0.f;
    for (int i = 0; i < w->n * w->c * w->h * w->w; ++i) wsum += fabsf(w->data[i]);
    printf("first %f last %f sum|conv1 w| %f\n", first, last, wsum);
    if (!(wsum > 0.f) || !isfinite(wsum)) return 6;
    bcnn_end_net(&net);

    /* ---- net-less helpers with the reference's prototypes, on raw device buffers ---- */
    enum { N = 1000 };
    float host[N], out[N];
    float *x = (float *)bcnn_b200_malloc(N * sizeof(float));
    float *y = (float *)bcnn_b200_malloc(N * sizeof(float));
    int bad = 0;
    for (int i = 0; i < N; ++i) host[i] = (float)(i % 7) - 3.0f;
    bcnn_b200_memcpy_h2d(x, host, sizeof(host), NULL);
    bcnn_forward_activation_gpu(x, N, BCNN_ACT_RELU);
    bcnn_b200_memcpy_d2h(out, x, sizeof(out), NULL);
    bcnn_b200_stream_sync(NULL);
    for (int i = 0; i < N; ++i) bad += close_to("relu", out[i], host[i] > 0 ? host[i] : 0.f);
    bcnn_cuda_fill_f32(N, 2.0f, y, 1);
    bcnn_backward_activation_gpu(x, y, N, BCNN_ACT_RELU); /* y *= (x > 0) */
    bcnn_cuda_axpy(N, 0.5f, x, 1, y, 1);                  /* y += 0.5 x */
    bcnn_cuda_scal(N, 3.0f, y, 1);
    bcnn_b200_memcpy_d2h(out, y, sizeof(out), NULL);
    bcnn_b200_stream_sync(NULL);
    for (int i = 0; i < N; ++i) {
        const float r = host[i] > 0 ? host[i] : 0.f;
        bad += close_to("fill/actbwd/axpy/scal", out[i], 3.0f * ((r > 0 ? 2.0f : 0.0f) + 0.5f * r));
    }
    bcnn_cuda_copy_f32(N / 2, x, 2, y, 1);                /* strided gather */
    bcnn_b200_memcpy_d2h(out, y, (N / 2) * sizeof(float), NULL);
    bcnn_b200_stream_sync(NULL);
    for (int i = 0; i < N / 2; ++i) bad += close_to("copy", out[i], host[2 * i] > 0 ? host[2 * i] : 0.f);
    /* gemm: C[3x5] = A^T[3x4] * B[4x5] + 1 * C, against a host loop */
    float a[12], b[20], c[15], cref[15];
    for (int i = 0; i < 12; ++i) a[i] = 0.25f * (float)(i - 5);
    for (int i = 0; i < 20; ++i) b[i] = 0.5f * (float)((i * 3) % 7 - 3);
    for (int i = 0; i < 15; ++i) c[i] = cref[i] = (float)i;
    for (int m = 0; m < 3; ++m)
        for (int n = 0; n < 5; ++n)
            for (int k = 0; k < 4; ++k) cref[m * 5 + n] += 2.0f * a[k * 3 + m] * b[k * 5 + n];
    float *da = (float *)bcnn_b200_malloc(sizeof(a)), *db = (float *)bcnn_b200_malloc(sizeof(b)),
          *dc = (float *)bcnn_b200_malloc(sizeof(c));
    bcnn_b200_memcpy_h2d(da, a, sizeof(a), NULL);
    bcnn_b200_memcpy_h2d(db, b, sizeof(b), NULL);
    bcnn_b200_memcpy_h2d(dc, c, sizeof(c), NULL);
    bcnn_cuda_gemm(1, 0, 3, 5, 4, 2.0f, da, 3, db, 5, 1.0f, dc, 5);
    bcnn_b200_memcpy_d2h(c, dc, sizeof(c), NULL);
    bcnn_b200_stream_sync(NULL);
    for (int i = 0; i < 15; ++i) bad += close_to("gemm", c[i], cref[i]);
    /* add_bias / grad_bias on [2, 3, 4] */
    float t[24], bias[3] = {1.f, -2.f, 0.5f}, gb[3] = {10.f, 20.f, 30.f};
    for (int i = 0; i < 24; ++i) t[i] = (float)i;
    float *dt = (float *)bcnn_b200_malloc(sizeof(t)), *dbias = (float *)bcnn_b200_malloc(sizeof(bias)),
          *dgb = (float *)bcnn_b200_malloc(sizeof(gb));
    bcnn_b200_memcpy_h2d(dt, t, sizeof(t), NULL);
    bcnn_b200_memcpy_h2d(dbias, bias, sizeof(bias), NULL);
    bcnn_b200_memcpy_h2d(dgb, gb, sizeof(gb), NULL);
    bcnn_cuda_add_bias(dt, dbias, 2, 3, 4);
    bcnn_cuda_grad_bias(dgb, dt, 2, 3, 4);
    float t2[24], gb2[3];
    bcnn_b200_memcpy_d2h(t2, dt, sizeof(t2), NULL);
    bcnn_b200_memcpy_d2h(gb2, dgb, sizeof(gb2), NULL);
    bcnn_b200_stream_sync(NULL);
    for (int ch = 0; ch < 3; ++ch) {
        float want = gb[ch];
        for (int bb = 0; bb < 2; ++bb)
            for (int i = 0; i < 4; ++i) {
                const int idx = (bb * 3 + ch) * 4 + i;
                bad += close_to("add_bias", t2[idx], t[idx] + bias[ch]);
                want += t[idx] + bias[ch];
            }
        bad += close_to("grad_bias", gb2[ch], want);
    }
    bcnn_b200_free(x); bcnn_b200_free(y); bcnn_b200_free(da); bcnn_b200_free(db); bcnn_b200_free(dc);
    bcnn_b200_free(dt); bcnn_b200_free(dbias); bcnn_b200_free(dgb);
    if (bad) return 7;
    printf("OK\n");
    return 0;
}
