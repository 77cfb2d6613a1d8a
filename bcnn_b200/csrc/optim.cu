// optim.cu -- fused SGD-momentum and Adam updates, softmax and euclidean cost for sm_100a.
//
// SGD: the reference spends five BLAS-1 launches per parameter tensor
// (bcnn_sgd_update_gpu, src/bcnn_learner.c:86-103: axpy, scal, axpy, axpy, scal);
// here it is one read-modify-write pass over (w, g): 16 B/element instead of 40.
// Momentum stays in the gradient buffer exactly as in the reference.
// Adam: nine BLAS-1 passes over (w, g, m, v) in bcnn_adam_update_cpu / _gpu
// (src/bcnn_learner.c:106-164) become one pass: 32 B/element instead of 96.
#include <limits.h>
#include <float.h>

#include "common.cuh"

using namespace b200;

namespace {

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(256)
sgd_kernel(float *__restrict__ w, float *__restrict__ g, size_t n, float wd_scale, float step,
           float g_scale, bool vec) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t done = 0;
    if (vec) {
        size_t n4 = n >> 2;
        for (size_t j = tid; j < n4; j += gstride) {
            float4 wv = reinterpret_cast<float4 *>(w)[j];
            float4 gv = reinterpret_cast<float4 *>(g)[j];
            // separate mul / add roundings, as the reference's axpy (no FMA contraction)
            gv.x = __fadd_rn(gv.x, __fmul_rn(wd_scale, wv.x));
            gv.y = __fadd_rn(gv.y, __fmul_rn(wd_scale, wv.y));
            gv.z = __fadd_rn(gv.z, __fmul_rn(wd_scale, wv.z));
            gv.w = __fadd_rn(gv.w, __fmul_rn(wd_scale, wv.w));
            wv.x = __fadd_rn(wv.x, __fmul_rn(step, gv.x));
            wv.y = __fadd_rn(wv.y, __fmul_rn(step, gv.y));
            wv.z = __fadd_rn(wv.z, __fmul_rn(step, gv.z));
            wv.w = __fadd_rn(wv.w, __fmul_rn(step, gv.w));
            gv.x *= g_scale; gv.y *= g_scale; gv.z *= g_scale; gv.w *= g_scale;
            reinterpret_cast<float4 *>(w)[j] = wv;
            reinterpret_cast<float4 *>(g)[j] = gv;
        }
        done = n4 << 2;
    }
    for (size_t j = done + tid; j < n; j += gstride) {
        float gv = __fadd_rn(g[j], __fmul_rn(wd_scale, w[j]));
        w[j] = __fadd_rn(w[j], __fmul_rn(step, gv));
        g[j] = gv * g_scale;
    }
}

// One Adam step for one element, in the operation order and with the roundings of
// bcnn_adam_update_cpu (src/bcnn_learner.c:118-129): axpy, axpby, vmul, axpby, pow(.,0.5),
// add_scalar(1e-7), vdiv, axpy, zero-fill. sqrt stands for powf(v, 0.5f) (correctly rounded
// here; glibc's powf is within 0.82 ulp of it). `guard`: the elements bcnn_vdiv handles in its
// scalar tail (n % 8, src/kernels/bcnn_mat.c:277-310 with AVX) get 0 instead of a quotient
// when the denominator is <= 1e-5; the vector body divides unconditionally.
__device__ __forceinline__ void adam_element(float &w, float &g, float &m, float &v, float wd_scale,
                                             float one_minus_b1, float b1, float one_minus_b2,
                                             float b2, float alpha, bool guard) {
    g = __fadd_rn(g, __fmul_rn(wd_scale, w));
    m = __fadd_rn(__fmul_rn(g, one_minus_b1), __fmul_rn(m, b1));
    const float g2 = __fmul_rn(g, g);
    v = __fadd_rn(__fmul_rn(g2, one_minus_b2), __fmul_rn(v, b2));
    const float den = __fadd_rn(__fsqrt_rn(v), 0.0000001f);
    const float q = (guard && !(fabsf(den) > 0.00001f)) ? 0.0f : __fdiv_rn(m, den);
    w = __fadd_rn(w, __fmul_rn(alpha, q));
    g = 0.0f;
}

__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ w, float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, size_t n, float wd_scale, float one_minus_b1, float b1,
            float one_minus_b2, float b2, float alpha, bool vec) {
    const size_t gstride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t tail = n & ~(size_t)7;  // first element of bcnn_vdiv's scalar tail
    size_t done = 0;
    if (vec) {
        const size_t n4 = n >> 2;
        for (size_t j = tid; j < n4; j += gstride) {
            float4 wv = reinterpret_cast<float4 *>(w)[j];
            float4 gv = reinterpret_cast<float4 *>(g)[j];
            float4 mv = reinterpret_cast<float4 *>(m)[j];
            float4 vv = reinterpret_cast<float4 *>(v)[j];
            const bool guard = (j << 2) >= tail;  // tail is a multiple of 4: whole vector or none
            adam_element(wv.x, gv.x, mv.x, vv.x, wd_scale, one_minus_b1, b1, one_minus_b2, b2, alpha, guard);
            adam_element(wv.y, gv.y, mv.y, vv.y, wd_scale, one_minus_b1, b1, one_minus_b2, b2, alpha, guard);
            adam_element(wv.z, gv.z, mv.z, vv.z, wd_scale, one_minus_b1, b1, one_minus_b2, b2, alpha, guard);
            adam_element(wv.w, gv.w, mv.w, vv.w, wd_scale, one_minus_b1, b1, one_minus_b2, b2, alpha, guard);
            reinterpret_cast<float4 *>(w)[j] = wv;
            reinterpret_cast<float4 *>(g)[j] = gv;
            reinterpret_cast<float4 *>(m)[j] = mv;
            reinterpret_cast<float4 *>(v)[j] = vv;
        }
        done = n4 << 2;
    }
    for (size_t j = done + tid; j < n; j += gstride)
        adam_element(w[j], g[j], m[j], v[j], wd_scale, one_minus_b1, b1, one_minus_b2, b2, alpha,
                     j >= tail);
}

// One warp per (sample, spatial position): softmax over channels in the
// log-sum-exp form of src/layers/bcnn_softmax_layer.c:88-155 (double exp/log rounded
// to float at each step, as the reference's casts do).
__global__ void __launch_bounds__(256)
softmax_kernel(const float *__restrict__ x, float *__restrict__ y, int rows, int c, int hw,
               FastDiv d_hw) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        uint32_t b, i;
        d_hw.divmod(r, b, i);
        const float *p = x + (size_t)b * c * hw + i;
        float *q = y + (size_t)b * c * hw + i;
        float vmax = -FLT_MAX;
        for (int j = lane; j < c; j += 32) vmax = fmaxf(vmax, p[(size_t)j * hw]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        float sum = 0.f;
        for (int j = lane; j < c; j += 32) sum += (float)exp((double)(p[(size_t)j * hw] - vmax));
        sum = warp_sum(sum);
        float lse = sum ? vmax + (float)log((double)sum) : vmax - 100.0f;
        for (int j = lane; j < c; j += 32)
            q[(size_t)j * hw] = (float)exp((double)(p[(size_t)j * hw] - lse));
    }
}

// sgd_kernel over a table of tensors: a CTA finds its tensor by bisection of first_block and owns
// 4096 consecutive elements of it (four float4 per thread, loads first).
__global__ void __launch_bounds__(256)
sgd_multi_kernel(const __grid_constant__ bcnn_b200_sgd_batch b) {
    int lo = 0, hi = b.count;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (blockIdx.x >= b.first_block[mid]) lo = mid;
        else hi = mid;
    }
    float *__restrict__ w = b.w[lo];
    float *__restrict__ g = b.g[lo];
    const unsigned int n = b.n[lo];
    const float wd_scale = b.wd_scale[lo], step = b.step, g_scale = b.g_scale;
    const unsigned int base = (blockIdx.x - b.first_block[lo]) * 4096u + threadIdx.x * 4u;
    float4 wv[4], gv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const unsigned int j = base + u * 1024u;
        if (j + 3 < n) {
            wv[u] = *reinterpret_cast<const float4 *>(w + j);
            gv[u] = *reinterpret_cast<const float4 *>(g + j);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const unsigned int j = base + u * 1024u;
        if (j + 3 < n) {
            float4 x = wv[u], y = gv[u];
            // separate mul / add roundings, as the reference's axpy (no FMA contraction)
            y.x = __fadd_rn(y.x, __fmul_rn(wd_scale, x.x));
            y.y = __fadd_rn(y.y, __fmul_rn(wd_scale, x.y));
            y.z = __fadd_rn(y.z, __fmul_rn(wd_scale, x.z));
            y.w = __fadd_rn(y.w, __fmul_rn(wd_scale, x.w));
            x.x = __fadd_rn(x.x, __fmul_rn(step, y.x));
            x.y = __fadd_rn(x.y, __fmul_rn(step, y.y));
            x.z = __fadd_rn(x.z, __fmul_rn(step, y.z));
            x.w = __fadd_rn(x.w, __fmul_rn(step, y.w));
            y.x *= g_scale; y.y *= g_scale; y.z *= g_scale; y.w *= g_scale;
            *reinterpret_cast<float4 *>(w + j) = x;
            *reinterpret_cast<float4 *>(g + j) = y;
        } else {
            for (unsigned int e = j; e < n && e < j + 4; ++e) {   // the tensor's last, partial vector
                const float y = __fadd_rn(g[e], __fmul_rn(wd_scale, w[e]));
                w[e] = __fadd_rn(w[e], __fmul_rn(step, y));
                g[e] = y * g_scale;
            }
        }
    }
}

// grad = pred - label: plain stream over the n x classes tensor.
__global__ void __launch_bounds__(256)
cost_grad_kernel(const float *__restrict__ pred, const float *__restrict__ label, float *__restrict__ grad,
                 int total) {
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) grad[i] = pred[i] - label[i];
}

// The scalar metric in one single-CTA pass (fixed summation order: deterministic). metric kinds =
// bcnn_loss_metric. Per-sample scans (error rate, dice) take one warp per sample, lanes striding the
// classes: with one thread per sample the 1000-class scan of a batch-256 classifier was 0.1 ms of
// dependent uncoalesced loads.
__global__ void __launch_bounds__(1024)
cost_metric_kernel(const float *__restrict__ pred, const float *__restrict__ label,
                   float *__restrict__ metric, int n, int input_size, int kind) {
    __shared__ float red[32];
    const int total = n * input_size;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[1] = {0.f};
    if (kind == 1 || kind == 2 || kind == 3 || kind == 4) {
        for (int i = threadIdx.x; i < total; i += 1024) {
            const float e = pred[i] - label[i];
            if (kind != 1) acc[0] += e * e;                           // SSE / MSE / CRPS
            else if (label[i] > 0.0f) {                               // LOGLOSS
                float pv = fminf(fmaxf(pred[i], 1e-8f), 1.0f - 1e-8f);
                acc[0] += (float)-log((double)pv);
            }
        }
    }
    if (kind == 0) {  // ERROR_RATE: argmax with strict '>' from FLT_MIN, first max wins
        for (int b = warp; b < n; b += 32) {
            const float *row = pred + (size_t)b * input_size;
            float pmax = FLT_MIN;
            int best = INT_MAX;   // no class above FLT_MIN yet
            for (int j = lane; j < input_size; j += 32) {
                const float v = row[j];
                if (v > pmax) { pmax = v; best = j; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float v2 = __shfl_xor_sync(0xffffffffu, pmax, o);
                const int b2 = __shfl_xor_sync(0xffffffffu, best, o);
                if (v2 > pmax || (v2 == pmax && b2 < best)) { pmax = v2; best = b2; }
            }
            if (best == INT_MAX) best = 0;
            if (lane == 0 && label[(size_t)b * input_size + best] == 0) acc[0] += 1.0f;
        }
    }
    if (kind == 5) {  // DICE
        for (int b = warp; b < n; b += 32) {
            int num = 0, den = 0;
            for (int j = lane; j < input_size; j += 32) {
                float l = label[(size_t)b * input_size + j];
                float hit = (float)(pred[(size_t)b * input_size + j] > 0.5f);
                num += (int)(l * hit);
                den += (int)(l + hit);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                num += __shfl_xor_sync(0xffffffffu, num, o);
                den += __shfl_xor_sync(0xffffffffu, den, o);
            }
            if (lane == 0) acc[0] += (float)(2.0f * num + 1.0f) / (den + 1.0f);
        }
    }
    block_sum<1, 1024>(acc, red);
    if (threadIdx.x == 0) {
        float m = acc[0];
        if (kind == 3) m /= input_size;
        metric[0] = m;
    }
}

}  // namespace

extern "C" int bcnn_b200_sgd_update(float *w, float *g, size_t n, float wd_scale, float step,
                                    float g_scale, void *stream) {
    if (n == 0) return 0;
    bool vec = aligned16(w) && aligned16(g);
    sgd_kernel<<<stream_grid(vec ? n / 4 + 1 : n, 256), 256, 0, as_stream(stream)>>>(
        w, g, n, wd_scale, step, g_scale, vec);
    return launched();
}

extern "C" int bcnn_b200_sgd_update_multi(const bcnn_b200_sgd_batch *batch, void *stream) {
    if (!batch || batch->count <= 0) return 0;
    if (batch->count > BCNN_B200_SGD_MULTI_MAX) return (int)cudaErrorInvalidValue;
    const unsigned int grid = batch->first_block[batch->count];
    if (grid == 0) return 0;
    sgd_multi_kernel<<<grid, 256, 0, as_stream(stream)>>>(*batch);
    return launched();
}

extern "C" int bcnn_b200_adam_update(float *w, float *g, float *m, float *v, size_t n,
                                     float wd_scale, float beta1, float beta2, float alpha,
                                     void *stream) {
    if (n == 0) return 0;
    bool vec = aligned16(w) && aligned16(g) && aligned16(m) && aligned16(v);
    adam_kernel<<<stream_grid(vec ? n / 4 + 1 : n, 256), 256, 0, as_stream(stream)>>>(
        w, g, m, v, n, wd_scale, 1.0f - beta1, beta1, 1.0f - beta2, beta2, alpha, vec);
    return launched();
}

extern "C" int bcnn_b200_softmax_forward(const float *x, float *y, int n, int c, int hw,
                                         void *stream) {
    int rows = n * hw;
    if (rows <= 0 || c <= 0) return 0;
    softmax_kernel<<<stream_grid((size_t)rows * 32, 256), 256, 0, as_stream(stream)>>>(
        x, y, rows, c, hw, FastDiv(hw));
    return launched();
}

extern "C" int bcnn_b200_cost_forward(const float *pred, const float *label, float *grad,
                                      float *metric, int n, int input_size, int metric_kind,
                                      void *stream) {
    if (n <= 0 || input_size <= 0) return 0;
    const int total = n * input_size;
    if (grad) {
        cost_grad_kernel<<<stream_grid((size_t)total, 256), 256, 0, as_stream(stream)>>>(pred, label, grad, total);
        int err = launched();
        if (err) return err;
    }
    cost_metric_kernel<<<1, 1024, 0, as_stream(stream)>>>(pred, label, metric, n, input_size, metric_kind);
    return launched();
}
