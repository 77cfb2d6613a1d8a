// batchnorm.cu -- batch normalisation forward / backward for sm_100a.
//
// HBM-bound. Per-channel reductions run as grid = (channels, splits) CTAs with
// warp-shuffle + shared-memory block sums; the last CTA of a channel (atomic ticket)
// folds the per-split partials in split order, so results are deterministic. The
// elementwise passes are flat float4 streams with the channel recovered by two
// multiply-shift divisions per 4 elements.
//
// Arithmetic follows the reference CPU path (NOT its .cu, which uses an unbiased
// variance and different epsilons -- SURVEY.md H4/H10):
//   forward : src/layers/bcnn_batchnorm_layer.c:147-242  (eps 1e-6, biased variance
//             E[x^2] - mean^2, running = 0.9 running + 0.1 batch)
//   backward: src/layers/bcnn_batchnorm_layer.c:263-332 + src/kernels/bcnn_mat.c:692-727
//             (eps 1e-5, var*sqrt(var) + 1e-5 in the variance term)
#include "common.cuh"

using namespace b200;

namespace {

constexpr int RT = 256;          // threads of the reduction kernels
constexpr int MAX_SPLITS = 64;   // scratch layout assumes this (bcnn_b200_bn_scratch_floats)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int reduce_splits(int n, int c) {
    int target = 2 * sm_count();
    int s = ceil_div(target, c);
    if (s > n) s = n;
    if (s > MAX_SPLITS) s = MAX_SPLITS;
    return s < 1 ? 1 : s;
}

// Images [b0, b1) handled by split `s` of `splits` (contiguous ranges).
__device__ __forceinline__ void split_range(int n, int s, int splits, int &b0, int &b1) {
    b0 = (int)(((long long)n * s) / splits);
    b1 = (int)(((long long)n * (s + 1)) / splits);
}

// ---- forward statistics ---------------------------------------------------------
__global__ void __launch_bounds__(RT)
bn_stats_kernel(const float *__restrict__ x, int n, int c, int hw, float *__restrict__ saved_mean,
                float *__restrict__ saved_var, float *__restrict__ run_mean,
                float *__restrict__ run_var, float *__restrict__ partial,
                unsigned int *__restrict__ tickets, FastDiv div_hw, FastDiv div_hw4, bool vec) {
    __shared__ float red[2 * RT / 32];
    __shared__ bool last;
    const int ch = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
    int b0, b1;
    split_range(n, split, splits, b0, b1);
    float acc[2] = {0.f, 0.f};
    if (vec) {
        const int hw4 = hw >> 2;
        const uint32_t total4 = (uint32_t)(b1 - b0) * hw4;
        for (uint32_t j = threadIdx.x; j < total4; j += RT) {
            uint32_t b, i4;
            div_hw4.divmod(j, b, i4);
            float4 v = ld_stream4(x + ((size_t)(b0 + b) * c + ch) * hw + (i4 << 2));
            acc[0] += (v.x + v.y) + (v.z + v.w);
            acc[1] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
    } else {
        const uint32_t total = (uint32_t)(b1 - b0) * hw;
        for (uint32_t j = threadIdx.x; j < total; j += RT) {
            uint32_t b, i;
            div_hw.divmod(j, b, i);
            float v = __ldg(x + ((size_t)(b0 + b) * c + ch) * hw + i);
            acc[0] += v;
            acc[1] += v * v;
        }
    }
    block_sum<2, RT>(acc, red);
    if (threadIdx.x == 0) {
        partial[((size_t)ch * MAX_SPLITS + split) * 2 + 0] = acc[0];
        partial[((size_t)ch * MAX_SPLITS + split) * 2 + 1] = acc[1];
        __threadfence();
        last = (atomicAdd(tickets + ch, 1u) == (unsigned)splits - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < splits; ++i) {
            s1 += __ldcg(partial + ((size_t)ch * MAX_SPLITS + i) * 2 + 0);
            s2 += __ldcg(partial + ((size_t)ch * MAX_SPLITS + i) * 2 + 1);
        }
        const float scale = 1.0f / (n * hw);
        const float mean = s1 * scale;
        const float var = s2 * scale - mean * mean;
        saved_mean[ch] = mean;
        saved_var[ch] = var;
        if (run_mean) {
            run_mean[ch] = run_mean[ch] * 0.9f + 0.1f * mean;
            run_var[ch] = run_var[ch] * 0.9f + 0.1f * var;
        }
        tickets[ch] = 0;
    }
}

// ---- forward apply: y = act(gamma * (x - mean) / sqrt(var + eps) + beta) -----------
// NORMALISE=false gives the PREDICT-mode y = act(gamma * x + beta).
template <bool NORMALISE>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ mean,
                const float *__restrict__ var, const float *__restrict__ gamma,
                const float *__restrict__ beta, size_t total, int act, FastDiv div_hw,
                FastDiv div_c, bool vec) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        size_t n4 = total >> 2;
        for (size_t j = tid; j < n4; j += gstride) {
            uint32_t q, ch;
            div_c.divmod(div_hw.div((uint32_t)(j << 2)), q, ch);
            float g = __ldg(gamma + ch), b = __ldg(beta + ch);
            float m = 0.f, inv = 1.f;
            if (NORMALISE) {
                m = __ldg(mean + ch);
                inv = 1.0f / sqrtf(__ldg(var + ch) + 0.000001f);
            }
            // in-place use (x == y) must not go through the non-coherent path
            float4 v = (x == y) ? reinterpret_cast<const float4 *>(x)[j] : ld_stream4(x + (j << 2));
            v.x = act_fwd((v.x - m) * inv * g + b, act, 0.f);
            v.y = act_fwd((v.y - m) * inv * g + b, act, 0.f);
            v.z = act_fwd((v.z - m) * inv * g + b, act, 0.f);
            v.w = act_fwd((v.w - m) * inv * g + b, act, 0.f);
            reinterpret_cast<float4 *>(y)[j] = v;
        }
    } else {
        for (size_t j = tid; j < total; j += gstride) {
            uint32_t q, ch;
            div_c.divmod(div_hw.div((uint32_t)j), q, ch);
            float g = __ldg(gamma + ch), b = __ldg(beta + ch);
            float m = 0.f, inv = 1.f;
            if (NORMALISE) {
                m = __ldg(mean + ch);
                inv = 1.0f / sqrtf(__ldg(var + ch) + 0.000001f);
            }
            y[j] = act_fwd((x[j] - m) * inv * g + b, act, 0.f);
        }
    }
}

// ---- backward reduction: S1 = sum dy', S2 = sum dy' * (x - mean), dy' = dy * act'(y) ----
__global__ void __launch_bounds__(RT)
bn_bwd_reduce_kernel(const float *__restrict__ x, const float *__restrict__ y,
                     const float *__restrict__ dy, const float *__restrict__ mean,
                     const float *__restrict__ var, const float *__restrict__ gamma,
                     float *__restrict__ g_gamma, float *__restrict__ g_beta,
                     float *__restrict__ d_mean, float *__restrict__ d_var, int n, int c, int hw,
                     int act, float *__restrict__ partial, unsigned int *__restrict__ tickets,
                     FastDiv div_hw, FastDiv div_hw4, bool vec) {
    __shared__ float red[2 * RT / 32];
    __shared__ bool last;
    const int ch = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
    int b0, b1;
    split_range(n, split, splits, b0, b1);
    const float m = mean[ch];
    float acc[2] = {0.f, 0.f};
    if (vec) {
        const int hw4 = hw >> 2;
        const uint32_t total4 = (uint32_t)(b1 - b0) * hw4;
        for (uint32_t j = threadIdx.x; j < total4; j += RT) {
            uint32_t b, i4;
            div_hw4.divmod(j, b, i4);
            size_t off = ((size_t)(b0 + b) * c + ch) * hw + (i4 << 2);
            float4 xv = ld_stream4(x + off);
            float4 g = ld_stream4(dy + off);
            if (act != ACT_NONE) {
                float4 yv = ld_stream4(y + off);
                g.x *= act_bwd_factor(yv.x, act, 0.f);
                g.y *= act_bwd_factor(yv.y, act, 0.f);
                g.z *= act_bwd_factor(yv.z, act, 0.f);
                g.w *= act_bwd_factor(yv.w, act, 0.f);
            }
            acc[0] += (g.x + g.y) + (g.z + g.w);
            acc[1] += (g.x * (xv.x - m) + g.y * (xv.y - m)) + (g.z * (xv.z - m) + g.w * (xv.w - m));
        }
    } else {
        const uint32_t total = (uint32_t)(b1 - b0) * hw;
        for (uint32_t j = threadIdx.x; j < total; j += RT) {
            uint32_t b, i;
            div_hw.divmod(j, b, i);
            size_t off = ((size_t)(b0 + b) * c + ch) * hw + i;
            float g = __ldg(dy + off);
            if (act != ACT_NONE) g *= act_bwd_factor(__ldg(y + off), act, 0.f);
            acc[0] += g;
            acc[1] += g * (__ldg(x + off) - m);
        }
    }
    block_sum<2, RT>(acc, red);
    if (threadIdx.x == 0) {
        partial[((size_t)ch * MAX_SPLITS + split) * 2 + 0] = acc[0];
        partial[((size_t)ch * MAX_SPLITS + split) * 2 + 1] = acc[1];
        __threadfence();
        last = (atomicAdd(tickets + ch, 1u) == (unsigned)splits - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < splits; ++i) {
            s1 += __ldcg(partial + ((size_t)ch * MAX_SPLITS + i) * 2 + 0);
            s2 += __ldcg(partial + ((size_t)ch * MAX_SPLITS + i) * 2 + 1);
        }
        const float v = var[ch], g = gamma[ch];
        g_beta[ch] += s1;                                  // bcnn_grad_bias
        g_gamma[ch] += s2 / sqrtf(v + 0.000001f);          // bcnn_grad_scales on x_norm (eps 1e-6)
        // after dy *= gamma: sums scale by gamma
        d_mean[ch] = (g * s1) * (-1.0f / sqrtf(v + 0.00001f));
        d_var[ch] = (g * s2) * (-0.5f / (v * sqrtf(v) + 0.00001f));
        tickets[ch] = 0;
    }
}

// ---- backward apply: dx = dy'*gamma/sqrt(var+1e-5) + d_var*2(x-mean)/m + d_mean/m -------
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ y,
                    const float *dy, float *dx, const float *__restrict__ mean,
                    const float *__restrict__ var, const float *__restrict__ gamma,
                    const float *__restrict__ d_mean, const float *__restrict__ d_var,
                    size_t total, int count /* n*hw */, int act, FastDiv div_hw, FastDiv div_c,
                    bool vec) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float inv_count = 1.0f / (float)count;
    if (vec) {
        size_t n4 = total >> 2;
        for (size_t j = tid; j < n4; j += gstride) {
            uint32_t q, ch;
            div_c.divmod(div_hw.div((uint32_t)(j << 2)), q, ch);
            const float m = __ldg(mean + ch);
            const float k1 = __ldg(gamma + ch) / sqrtf(__ldg(var + ch) + 0.00001f);
            const float k2 = __ldg(d_var + ch) * 2.0f * inv_count;
            const float k3 = __ldg(d_mean + ch) * inv_count;
            float4 xv = ld_stream4(x + (j << 2));
            float4 g = reinterpret_cast<const float4 *>(dy)[j];
            if (act != ACT_NONE) {
                float4 yv = ld_stream4(y + (j << 2));
                g.x *= act_bwd_factor(yv.x, act, 0.f);
                g.y *= act_bwd_factor(yv.y, act, 0.f);
                g.z *= act_bwd_factor(yv.z, act, 0.f);
                g.w *= act_bwd_factor(yv.w, act, 0.f);
            }
            float4 r;
            r.x = g.x * k1 + k2 * (xv.x - m) + k3;
            r.y = g.y * k1 + k2 * (xv.y - m) + k3;
            r.z = g.z * k1 + k2 * (xv.z - m) + k3;
            r.w = g.w * k1 + k2 * (xv.w - m) + k3;
            reinterpret_cast<float4 *>(dx)[j] = r;
        }
    } else {
        for (size_t j = tid; j < total; j += gstride) {
            uint32_t q, ch;
            div_c.divmod(div_hw.div((uint32_t)j), q, ch);
            const float m = __ldg(mean + ch);
            const float k1 = __ldg(gamma + ch) / sqrtf(__ldg(var + ch) + 0.00001f);
            const float k2 = __ldg(d_var + ch) * 2.0f * inv_count;
            const float k3 = __ldg(d_mean + ch) * inv_count;
            float g = dy[j];
            if (act != ACT_NONE) g *= act_bwd_factor(__ldg(y + j), act, 0.f);
            dx[j] = g * k1 + k2 * (__ldg(x + j) - m) + k3;
        }
    }
}

}  // namespace

extern "C" int bcnn_b200_bn_stats(const float *x, int n, int c, int hw, float *saved_mean,
                                  float *saved_var, float *run_mean, float *run_var,
                                  float *scratch, void *stream) {
    if ((size_t)n * c * hw == 0) return 0;
    int splits = reduce_splits(n, c);
    unsigned int *tickets = reinterpret_cast<unsigned int *>(scratch + (size_t)c * MAX_SPLITS * 4);
    dim3 grid(c, splits);
    bn_stats_kernel<<<grid, RT, 0, as_stream(stream)>>>(x, n, c, hw, saved_mean, saved_var,
                                                        run_mean, run_var, scratch, tickets,
                                                        FastDiv(hw), FastDiv(hw >> 2 ? hw >> 2 : 1),
                                                        (hw % 4) == 0 && aligned16(x));
    return launched();
}

extern "C" int bcnn_b200_bn_apply(const float *x, float *y, const float *mean, const float *var,
                                  const float *gamma, const float *beta, int n, int c, int hw,
                                  int act, void *stream) {
    size_t total = (size_t)n * c * hw;
    if (total == 0) return 0;
    bool vec = (hw % 4) == 0 && aligned16(x) && aligned16(y);
    bn_apply_kernel<true><<<stream_grid(vec ? total / 4 : total, 256), 256, 0, as_stream(stream)>>>(
        x, y, mean, var, gamma, beta, total, act, FastDiv(hw), FastDiv(c), vec);
    return launched();
}

extern "C" int bcnn_b200_scale_bias(const float *x, float *y, const float *gamma,
                                    const float *beta, int n, int c, int hw, int act,
                                    void *stream) {
    size_t total = (size_t)n * c * hw;
    if (total == 0) return 0;
    bool vec = (hw % 4) == 0 && aligned16(x) && aligned16(y);
    bn_apply_kernel<false><<<stream_grid(vec ? total / 4 : total, 256), 256, 0,
                             as_stream(stream)>>>(x, y, nullptr, nullptr, gamma, beta, total, act,
                                                  FastDiv(hw), FastDiv(c), vec);
    return launched();
}

extern "C" int bcnn_b200_bn_backward(const float *x, const float *y, float *dy, float *dx_out,
                                     const float *mean, const float *var, const float *gamma,
                                     float *g_gamma, float *g_beta, float *d_mean, float *d_var,
                                     int n, int c, int hw, int act, float *scratch, void *stream) {
    size_t total = (size_t)n * c * hw;
    if (total == 0) return 0;
    cudaStream_t st = as_stream(stream);
    int splits = reduce_splits(n, c);
    unsigned int *tickets = reinterpret_cast<unsigned int *>(scratch + (size_t)c * MAX_SPLITS * 4);
    dim3 grid(c, splits);
    bn_bwd_reduce_kernel<<<grid, RT, 0, st>>>(x, y, dy, mean, var, gamma, g_gamma, g_beta, d_mean,
                                              d_var, n, c, hw, act, scratch, tickets, FastDiv(hw),
                                              FastDiv(hw >> 2 ? hw >> 2 : 1),
                                              (hw % 4) == 0 && aligned16(x) && aligned16(dy) &&
                                                  (y == nullptr || aligned16(y)));
    int err = launched();
    if (err) return err;
    bool vec = (hw % 4) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx_out) &&
               (y == nullptr || aligned16(y));
    bn_bwd_apply_kernel<<<stream_grid(vec ? total / 4 : total, 256), 256, 0, st>>>(
        x, y, dy, dx_out, mean, var, gamma, d_mean, d_var, total, n * hw, act, FastDiv(hw),
        FastDiv(c), vec);
    return launched();
}
