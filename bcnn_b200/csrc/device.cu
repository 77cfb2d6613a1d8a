// device.cu -- device, stream, event and memory helpers behind the C ABI
// (include/bcnn_b200.h). Replaces the helpers of reference src/bcnn_utils.c:124-201.
#include "common.cuh"

namespace b200 {
unsigned long long g_launch_count = 0;

int sm_count() {
    static int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
        if (cached <= 0) cached = 148;
    }
    return cached;
}
}  // namespace b200

using namespace b200;

extern "C" {

int bcnn_b200_set_device(int device) { return (int)cudaSetDevice(device); }

int bcnn_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int bcnn_b200_sm_count(void) { return sm_count(); }

void *bcnn_b200_malloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) bytes = 4;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        fprintf(stderr, "[ERROR] [CUDA] cudaMalloc(%zu) failed: %s\n", bytes,
                cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    // the net's streams are non-blocking (no implicit ordering with the legacy stream this
    // memset runs on): finish it before anyone can launch work on the buffer
    cudaMemset(p, 0, bytes);
    cudaStreamSynchronize(0);
    return p;
}

void bcnn_b200_free(void *p) {
    if (p) cudaFree(p);
}

void *bcnn_b200_malloc_host(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) bytes = 4;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}

void bcnn_b200_free_host(void *p) {
    if (p) cudaFreeHost(p);
}

int bcnn_b200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream));
}
int bcnn_b200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream));
}
int bcnn_b200_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream));
}

void *bcnn_b200_stream_create(void) {
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    return (void *)s;
}
void bcnn_b200_stream_destroy(void *s) {
    if (s) cudaStreamDestroy(as_stream(s));
}
int bcnn_b200_stream_sync(void *s) { return (int)cudaStreamSynchronize(as_stream(s)); }

void *bcnn_b200_event_create(void) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    return (void *)e;
}
void bcnn_b200_event_destroy(void *e) {
    if (e) cudaEventDestroy((cudaEvent_t)e);
}
int bcnn_b200_event_record(void *e, void *s) {
    return (int)cudaEventRecord((cudaEvent_t)e, as_stream(s));
}
int bcnn_b200_stream_wait_event(void *s, void *e) {
    return (int)cudaStreamWaitEvent(as_stream(s), (cudaEvent_t)e, 0);
}
float bcnn_b200_event_elapsed_ms(void *a, void *b) {
    float ms = -1.f;
    cudaEventSynchronize((cudaEvent_t)b);
    cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b);
    return ms;
}
// ---- CUDA graphs: capture what a stream is given between begin and end, replay it later ----
int bcnn_b200_graph_begin(void *s) {
    return (int)cudaStreamBeginCapture(as_stream(s), cudaStreamCaptureModeThreadLocal);
}
// Ends the capture and instantiates the graph. Returns the executable graph, or nullptr when the
// capture was invalidated or empty (the captured work was NOT executed either way).
void *bcnn_b200_graph_end(void *s) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaError_t err = cudaStreamEndCapture(as_stream(s), &graph);
    if (err == cudaSuccess && graph) err = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (err != cudaSuccess) {
        cudaGetLastError();  // clear the sticky capture error: the caller falls back to eager launches
        return nullptr;
    }
    return (void *)exec;
}
// `kernels` = the launches that were recorded into the graph: a replay executes that many kernels,
// and bcnn_b200_launch_count() keeps counting kernels, not graph launches.
int bcnn_b200_graph_launch(void *exec, unsigned long long kernels, void *s) {
    g_launch_count += kernels;
    return (int)cudaGraphLaunch((cudaGraphExec_t)exec, as_stream(s));
}
void bcnn_b200_graph_destroy(void *exec) {
    if (exec) cudaGraphExecDestroy((cudaGraphExec_t)exec);
}

const char *bcnn_b200_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }

uint64_t bcnn_b200_launch_count(void) { return (uint64_t)g_launch_count; }

}  // extern "C"

// ---------------------------------------------------------------------------
// BLAS-1 class kernels
// ---------------------------------------------------------------------------
namespace {

__global__ void fill_kernel(float *__restrict__ x, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t n4 = n >> 2;
    float4 v4 = make_float4(v, v, v, v);
    float4 *x4 = reinterpret_cast<float4 *>(x);
    for (size_t j = i; j < n4; j += stride) x4[j] = v4;
    for (size_t j = (n4 << 2) + i; j < n; j += stride) x[j] = v;
}

__global__ void axpy_kernel(float *__restrict__ y, const float *__restrict__ x, size_t n,
                            float a) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = i; j < n; j += stride) y[j] = fmaf(a, x[j], y[j]);
}

}  // namespace

extern "C" int bcnn_b200_fill_f32(float *x, size_t n, float value, void *stream) {
    if (n == 0) return 0;
    if (value == 0.0f) {
        ++g_launch_count;
        return (int)cudaMemsetAsync(x, 0, n * sizeof(float), as_stream(stream));
    }
    // the vector body needs 16-byte alignment (cudaMalloc'd tensors always have it);
    // peel an unaligned head with the scalar tail loop of a zero-length vector body
    size_t head = 0;
    uintptr_t mis = reinterpret_cast<uintptr_t>(x) & 15;
    if (mis) {
        head = (16 - mis) / sizeof(float);
        if (head > n) head = n;
        fill_kernel<<<1, 32, 0, as_stream(stream)>>>(x, head < 4 ? head : 3, value);
        ++g_launch_count;
    }
    if (n > head)
        fill_kernel<<<stream_grid((n - head) / 4 + 1, 256), 256, 0, as_stream(stream)>>>(
            x + head, n - head, value);
    return launched();
}

extern "C" int bcnn_b200_axpy(float *y, const float *x, size_t n, float a, void *stream) {
    if (n == 0) return 0;
    axpy_kernel<<<stream_grid(n, 256), 256, 0, as_stream(stream)>>>(y, x, n, a);
    return launched();
}
