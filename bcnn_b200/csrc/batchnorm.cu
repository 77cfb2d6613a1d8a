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
#include <cstdlib>

#include "common.cuh"
#include "conv_impl.cuh"

using namespace b200;

namespace {

constexpr int RT = 256;          // threads of the reduction kernels
constexpr int MAX_SPLITS = 64;   // scratch layout assumes this (bcnn_b200_bn_scratch_floats)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// The lean elementwise streams walk their tensor from the END: the producer (a convolution or a
// per-channel reduction, both ascending) has just left the tail of the tensor in the 126 MB L2, and
// what these streams write last -- the head -- is what the next ascending consumer reads first.
inline bool reverse_streams() {
    static const bool on = !getenv("BCNN_B200_NO_REVERSE");
    return on;
}

inline bool l2_hints() {
    static const bool on = !getenv("BCNN_B200_NO_L2_HINTS");
    return on;
}

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

// gamma * (x - mean) * inv + beta, every step rounded separately (the reference scales and adds
// the bias in separate passes). Forward and backward share it so that the backward can rebuild
// the sign of the pre-activation bit for bit instead of re-reading y.
__device__ __forceinline__ float bn_affine(float v, float m, float inv, float g, float b) {
    return __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v, m), inv), g), b);
}
// ReLU / leaky-ReLU derivative from the pre-activation (y > 0 <=> pre > 0 for both).
__device__ __forceinline__ float relu_factor(float pre, float neg_slope) {
    return pre > 0.f ? 1.0f : neg_slope;
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
        // UNROLL independent 16-byte loads in flight per thread: a channel's reduction runs on
        // few CTAs, so memory-level parallelism has to come from inside the thread
        constexpr int UNROLL = 4;
        for (uint32_t j0 = threadIdx.x; j0 < total4; j0 += RT * UNROLL) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t j = j0 + u * RT;
                if (j < total4) {
                    uint32_t b, i4;
                    div_hw4.divmod(j, b, i4);
                    v[u] = ld_stream4(x + ((size_t)(b0 + b) * c + ch) * hw + (i4 << 2));
                } else {
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                acc[0] += (v[u].x + v[u].y) + (v[u].z + v[u].w);
                acc[1] += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
            }
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

// ---- forward statistics from the partial sums of a convolution epilogue ----------------
// partial[(row * 2 + k) * c + ch]; grid = (channel groups of 32, row segments). A warp reads 128
// contiguous bytes per row; warps, segments and rows are folded in a fixed order.
__global__ void __launch_bounds__(256)
bn_stats_finalize_kernel(const float *__restrict__ partial, int rows, int c, float inv_count,
                         float *__restrict__ saved_mean, float *__restrict__ saved_var,
                         float *__restrict__ run_mean, float *__restrict__ run_var,
                         float *__restrict__ seg_sums, unsigned int *__restrict__ tickets) {
    __shared__ float red[8][32][2];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + lane;
    const int seg = blockIdx.y, segs = gridDim.y;
    const int r0 = (int)(((long long)rows * seg) / segs), r1 = (int)(((long long)rows * (seg + 1)) / segs);
    float s1 = 0.f, s2 = 0.f;
    if (ch < c) {
        const size_t pitch = (size_t)2 * c;
        int r = r0 + w;
        for (; r + 24 < r1; r += 32) {  // four independent rows in flight per thread
            const float *q = partial + (size_t)r * pitch + ch;
            const float a0 = __ldcs(q), b0 = __ldcs(q + c);
            const float a1 = __ldcs(q + 8 * pitch), b1 = __ldcs(q + 8 * pitch + c);
            const float a2 = __ldcs(q + 16 * pitch), b2 = __ldcs(q + 16 * pitch + c);
            const float a3 = __ldcs(q + 24 * pitch), b3 = __ldcs(q + 24 * pitch + c);
            s1 += (a0 + a1) + (a2 + a3);
            s2 += (b0 + b1) + (b2 + b3);
        }
        for (; r < r1; r += 8) {
            const float *q = partial + (size_t)r * pitch + ch;
            s1 += __ldcs(q);
            s2 += __ldcs(q + c);
        }
    }
    red[w][lane][0] = s1;
    red[w][lane][1] = s2;
    __syncthreads();
    if (w == 0) {
        for (int i = 1; i < 8; ++i) {
            s1 += red[i][lane][0];
            s2 += red[i][lane][1];
        }
        if (ch < c) {
            seg_sums[((size_t)ch * MAX_SPLITS + seg) * 2 + 0] = s1;
            seg_sums[((size_t)ch * MAX_SPLITS + seg) * 2 + 1] = s2;
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) last = (atomicAdd(tickets + blockIdx.x, 1u) == (unsigned)segs - 1);
    }
    __syncthreads();
    if (last && w == 0) {
        __threadfence();
        if (ch < c) {
            float t1 = 0.f, t2 = 0.f;
            for (int i = 0; i < segs; ++i) {
                t1 += __ldcg(seg_sums + ((size_t)ch * MAX_SPLITS + i) * 2 + 0);
                t2 += __ldcg(seg_sums + ((size_t)ch * MAX_SPLITS + i) * 2 + 1);
            }
            const float mean = t1 * inv_count;
            const float var = t2 * inv_count - mean * mean;
            saved_mean[ch] = mean;
            saved_var[ch] = var;
            if (run_mean) {
                run_mean[ch] = run_mean[ch] * 0.9f + 0.1f * mean;
                run_var[ch] = run_var[ch] * 0.9f + 0.1f * var;
            }
        }
        if (lane == 0) tickets[blockIdx.x] = 0;
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
        const size_t n4 = total >> 2;
        constexpr int UNROLL = 2;  // independent 16-byte loads in flight per thread
        for (size_t j0 = tid; j0 < n4; j0 += gstride * UNROLL) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t j = j0 + u * gstride;
                // in-place use (x == y) must not go through the non-coherent path
                if (j < n4) v[u] = (x == y) ? reinterpret_cast<const float4 *>(x)[j] : ld_stream4(x + (j << 2));
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t j = j0 + u * gstride;
                if (j >= n4) break;
                uint32_t q, ch;
                div_c.divmod(div_hw.div((uint32_t)(j << 2)), q, ch);
                const float g = __ldg(gamma + ch), b = __ldg(beta + ch);
                float m = 0.f, inv = 1.f;
                if (NORMALISE) {
                    m = __ldg(mean + ch);
                    inv = 1.0f / sqrtf(__ldg(var + ch) + 0.000001f);
                }
                float4 r;
                r.x = act_fwd(bn_affine(v[u].x, m, inv, g, b), act, 0.f);
                r.y = act_fwd(bn_affine(v[u].y, m, inv, g, b), act, 0.f);
                r.z = act_fwd(bn_affine(v[u].z, m, inv, g, b), act, 0.f);
                r.w = act_fwd(bn_affine(v[u].w, m, inv, g, b), act, 0.f);
                reinterpret_cast<float4 *>(y)[j] = r;
            }
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
            y[j] = act_fwd(bn_affine(x[j], m, inv, g, b), act, 0.f);
        }
    }
}

// ---- backward reduction: S1 = sum dy', S2 = sum dy' * (x - mean), dy' = dy * act'(y) ----
__global__ void __launch_bounds__(RT)
bn_bwd_reduce_kernel(const float *__restrict__ x, const float *__restrict__ y,
                     const float *__restrict__ dy, const float *__restrict__ mean,
                     const float *__restrict__ var, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float *__restrict__ g_gamma,
                     float *__restrict__ g_beta,
                     float *__restrict__ d_mean, float *__restrict__ d_var, int n, int c, int hw,
                     int act, float *__restrict__ partial, unsigned int *__restrict__ tickets,
                     FastDiv div_hw, FastDiv div_hw4, bool vec) {
    __shared__ float red[2 * RT / 32];
    __shared__ bool last;
    const int ch = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
    int b0, b1;
    split_range(n, split, splits, b0, b1);
    const float m = mean[ch];
    // ReLU family with beta known: the mask comes from the recomputed pre-activation, y is not read
    const bool remask = beta != nullptr && (act == ACT_RELU || act == ACT_LRELU);
    const float inv6 = remask ? 1.0f / sqrtf(var[ch] + 0.000001f) : 0.f;
    const float gch = remask ? gamma[ch] : 0.f, bch = remask ? beta[ch] : 0.f;
    const float neg = act == ACT_LRELU ? 0.1f : 0.f;
    float acc[2] = {0.f, 0.f};
    if (vec) {
        const int hw4 = hw >> 2;
        const uint32_t total4 = (uint32_t)(b1 - b0) * hw4;
        constexpr int UNROLL = 4;
        for (uint32_t j0 = threadIdx.x; j0 < total4; j0 += RT * UNROLL) {
            float4 xv[UNROLL], g[UNROLL], yv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t j = j0 + u * RT;
                xv[u] = g[u] = yv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < total4) {
                    uint32_t b, i4;
                    div_hw4.divmod(j, b, i4);
                    size_t off = ((size_t)(b0 + b) * c + ch) * hw + (i4 << 2);
                    xv[u] = ld_stream4(x + off);
                    g[u] = ld_stream4(dy + off);
                    if (act != ACT_NONE && !remask) yv[u] = ld_stream4(y + off);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (remask) {
                    g[u].x *= relu_factor(bn_affine(xv[u].x, m, inv6, gch, bch), neg);
                    g[u].y *= relu_factor(bn_affine(xv[u].y, m, inv6, gch, bch), neg);
                    g[u].z *= relu_factor(bn_affine(xv[u].z, m, inv6, gch, bch), neg);
                    g[u].w *= relu_factor(bn_affine(xv[u].w, m, inv6, gch, bch), neg);
                } else if (act != ACT_NONE) {
                    g[u].x *= act_bwd_factor(yv[u].x, act, 0.f);
                    g[u].y *= act_bwd_factor(yv[u].y, act, 0.f);
                    g[u].z *= act_bwd_factor(yv[u].z, act, 0.f);
                    g[u].w *= act_bwd_factor(yv[u].w, act, 0.f);
                }
                acc[0] += (g[u].x + g[u].y) + (g[u].z + g[u].w);
                acc[1] += (g[u].x * (xv[u].x - m) + g[u].y * (xv[u].y - m)) +
                          (g[u].z * (xv[u].z - m) + g[u].w * (xv[u].w - m));
            }
        }
    } else {
        const uint32_t total = (uint32_t)(b1 - b0) * hw;
        for (uint32_t j = threadIdx.x; j < total; j += RT) {
            uint32_t b, i;
            div_hw.divmod(j, b, i);
            size_t off = ((size_t)(b0 + b) * c + ch) * hw + i;
            float g = __ldg(dy + off);
            const float xv = __ldg(x + off);
            if (remask) g *= relu_factor(bn_affine(xv, m, inv6, gch, bch), neg);
            else if (act != ACT_NONE) g *= act_bwd_factor(__ldg(y + off), act, 0.f);
            acc[0] += g;
            acc[1] += g * (xv - m);
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
                    const float *__restrict__ beta, const float *__restrict__ d_mean,
                    const float *__restrict__ d_var,
                    size_t total, int count /* n*hw */, int act, FastDiv div_hw, FastDiv div_c,
                    bool vec) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float inv_count = 1.0f / (float)count;
    const bool remask = beta != nullptr && (act == ACT_RELU || act == ACT_LRELU);
    const float neg = act == ACT_LRELU ? 0.1f : 0.f;
    if (vec) {
        const size_t n4 = total >> 2;
        constexpr int UNROLL = 2;
        for (size_t j0 = tid; j0 < n4; j0 += gstride * UNROLL) {
            float4 xv[UNROLL], gv[UNROLL], yv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t j = j0 + u * gstride;
                if (j < n4) {
                    xv[u] = ld_stream4(x + (j << 2));
                    gv[u] = reinterpret_cast<const float4 *>(dy)[j];
                    if (act != ACT_NONE && !remask) yv[u] = ld_stream4(y + (j << 2));
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t j = j0 + u * gstride;
                if (j >= n4) break;
                uint32_t q, ch;
                div_c.divmod(div_hw.div((uint32_t)(j << 2)), q, ch);
                const float m = __ldg(mean + ch);
                const float k1 = __ldg(gamma + ch) / sqrtf(__ldg(var + ch) + 0.00001f);
                const float k2 = __ldg(d_var + ch) * 2.0f * inv_count;
                const float k3 = __ldg(d_mean + ch) * inv_count;
                float4 g = gv[u];
                if (remask) {
                    const float inv6 = 1.0f / sqrtf(__ldg(var + ch) + 0.000001f);
                    const float gch = __ldg(gamma + ch), bch = __ldg(beta + ch);
                    g.x *= relu_factor(bn_affine(xv[u].x, m, inv6, gch, bch), neg);
                    g.y *= relu_factor(bn_affine(xv[u].y, m, inv6, gch, bch), neg);
                    g.z *= relu_factor(bn_affine(xv[u].z, m, inv6, gch, bch), neg);
                    g.w *= relu_factor(bn_affine(xv[u].w, m, inv6, gch, bch), neg);
                } else if (act != ACT_NONE) {
                    g.x *= act_bwd_factor(yv[u].x, act, 0.f);
                    g.y *= act_bwd_factor(yv[u].y, act, 0.f);
                    g.z *= act_bwd_factor(yv[u].z, act, 0.f);
                    g.w *= act_bwd_factor(yv[u].w, act, 0.f);
                }
                float4 r;
                r.x = g.x * k1 + k2 * (xv[u].x - m) + k3;
                r.y = g.y * k1 + k2 * (xv[u].y - m) + k3;
                r.z = g.z * k1 + k2 * (xv[u].z - m) + k3;
                r.w = g.w * k1 + k2 * (xv[u].w - m) + k3;
                reinterpret_cast<float4 *>(dx)[j] = r;
            }
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
            const float xv = __ldg(x + j);
            if (remask)
                g *= relu_factor(bn_affine(xv, m, 1.0f / sqrtf(__ldg(var + ch) + 0.000001f),
                                           __ldg(gamma + ch), __ldg(beta + ch)), neg);
            else if (act != ACT_NONE) g *= act_bwd_factor(__ldg(y + j), act, 0.f);
            dx[j] = g * k1 + k2 * (xv - m) + k3;
        }
    }
}


// ---- lean float4 streams with per-channel constants tabulated in shared memory ------------
// The generic kernels above rebuild 1/sqrt(var + eps) and friends for every float4 (~140
// instructions per warp-level float4: issue-bound at 60 % of HBM); here each CTA tabulates them
// once per channel, the stream costs one 16-byte shared-memory read per float4 and the activation
// is a template parameter. Same arithmetic, same rounding order as the generic kernels.
template <int ACT>   // ACT_NONE, ACT_RELU or ACT_LRELU
__global__ void __launch_bounds__(256)
bn_apply_fast_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ mean,
                     const float *__restrict__ var, const float *__restrict__ gamma,
                     const float *__restrict__ beta, uint32_t n4, int c, FastDiv div_hw4, FastDiv div_c,
                     bool reverse, bool hints) {
    extern __shared__ float4 bn_tab[];   // {mean, 1/sqrt(var + 1e-6), gamma, beta}
    const uint64_t pol_in = l2_policy_evict_first(), pol_out = l2_policy_evict_last();
    for (int i = threadIdx.x; i < c; i += 256)
        bn_tab[i] = make_float4(__ldg(mean + i), 1.0f / sqrtf(__ldg(var + i) + 0.000001f), __ldg(gamma + i),
                                __ldg(beta + i));
    __syncthreads();
    const uint32_t gstride = gridDim.x * 256u;
    constexpr int UNROLL = 4;   // independent 16-byte loads in flight per thread
    for (uint32_t j0 = blockIdx.x * 256u + threadIdx.x; j0 < n4; j0 += gstride * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t i = j0 + u * gstride;
            if (i < n4) {
                const float *src = x + ((size_t)(reverse ? n4 - 1 - i : i) << 2);
                v[u] = hints ? ld_stream4_hint(src, pol_in) : ld_stream4(src);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t i = j0 + u * gstride;
            if (i < n4) {
                const uint32_t j = reverse ? n4 - 1 - i : i;
                uint32_t q, ch;
                div_c.divmod(div_hw4.div(j), q, ch);
                const float4 t = bn_tab[ch];
                float4 r;
                r.x = bn_affine(v[u].x, t.x, t.y, t.z, t.w);
                r.y = bn_affine(v[u].y, t.x, t.y, t.z, t.w);
                r.z = bn_affine(v[u].z, t.x, t.y, t.z, t.w);
                r.w = bn_affine(v[u].w, t.x, t.y, t.z, t.w);
                if (ACT == ACT_RELU) {
                    r.x *= (float)(r.x > 0); r.y *= (float)(r.y > 0); r.z *= (float)(r.z > 0); r.w *= (float)(r.w > 0);
                } else if (ACT == ACT_LRELU) {
                    r.x = r.x > 0 ? r.x : 0.1f * r.x; r.y = r.y > 0 ? r.y : 0.1f * r.y;
                    r.z = r.z > 0 ? r.z : 0.1f * r.z; r.w = r.w > 0 ? r.w : 0.1f * r.w;
                }
                if (hints) st_stream4_hint(y + ((size_t)j << 2), r, pol_out);
                else st_stream4(y + ((size_t)j << 2), r);
            }
        }
    }
}

// REMASK: ReLU / leaky-ReLU mask rebuilt from x (beta known); otherwise no activation.
template <bool REMASK>
__global__ void __launch_bounds__(256)
bn_bwd_apply_fast_kernel(const float *__restrict__ x, const float *dy, float *dx,
                         const float *__restrict__ mean, const float *__restrict__ var,
                         const float *__restrict__ gamma, const float *__restrict__ beta,
                         const float *__restrict__ d_mean, const float *__restrict__ d_var, uint32_t n4,
                         int c, int count, float neg, FastDiv div_hw4, FastDiv div_c, bool reverse,
                         bool hints) {
    const uint64_t pol_in = l2_policy_evict_first(), pol_out = l2_policy_evict_last();
    extern __shared__ float4 bn_tab[];   // [c] {mean, k1, k2, k3}, then [c] {1/sqrt(var + 1e-6), gamma, beta, -}
    const float inv_count = 1.0f / (float)count;
    for (int i = threadIdx.x; i < c; i += 256) {
        const float v = __ldg(var + i), g = __ldg(gamma + i);
        bn_tab[i] = make_float4(__ldg(mean + i), g / sqrtf(v + 0.00001f), __ldg(d_var + i) * 2.0f * inv_count,
                                __ldg(d_mean + i) * inv_count);
        if (REMASK) bn_tab[c + i] = make_float4(1.0f / sqrtf(v + 0.000001f), g, __ldg(beta + i), 0.f);
    }
    __syncthreads();
    const uint32_t gstride = gridDim.x * 256u;
    constexpr int UNROLL = 2;
    for (uint32_t j0 = blockIdx.x * 256u + threadIdx.x; j0 < n4; j0 += gstride * UNROLL) {
        float4 xv[UNROLL], gv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t i = j0 + u * gstride;
            if (i < n4) {
                const uint32_t j = reverse ? n4 - 1 - i : i;
                // dx may alias dy: coherent load for dy
                if (hints) {
                    xv[u] = ld_stream4_hint(x + ((size_t)j << 2), pol_in);
                    gv[u] = ld4_hint(dy + ((size_t)j << 2), pol_in);
                } else {
                    xv[u] = ld_stream4(x + ((size_t)j << 2));
                    gv[u] = reinterpret_cast<const float4 *>(dy)[j];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t i = j0 + u * gstride;
            if (i < n4) {
                const uint32_t j = reverse ? n4 - 1 - i : i;
                uint32_t q, ch;
                div_c.divmod(div_hw4.div(j), q, ch);
                const float4 t = bn_tab[ch];
                float4 g = gv[u];
                if (REMASK) {
                    const float4 r = bn_tab[c + ch];
                    g.x *= relu_factor(bn_affine(xv[u].x, t.x, r.x, r.y, r.z), neg);
                    g.y *= relu_factor(bn_affine(xv[u].y, t.x, r.x, r.y, r.z), neg);
                    g.z *= relu_factor(bn_affine(xv[u].z, t.x, r.x, r.y, r.z), neg);
                    g.w *= relu_factor(bn_affine(xv[u].w, t.x, r.x, r.y, r.z), neg);
                }
                float4 o;
                o.x = g.x * t.y + t.z * (xv[u].x - t.x) + t.w;
                o.y = g.y * t.y + t.z * (xv[u].y - t.x) + t.w;
                o.z = g.z * t.y + t.z * (xv[u].z - t.x) + t.w;
                o.w = g.w * t.y + t.z * (xv[u].w - t.x) + t.w;
                if (hints) st_stream4_hint(dx + ((size_t)j << 2), o, pol_out);
                else reinterpret_cast<float4 *>(dx)[j] = o;
            }
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

int b200::bn_stats_from_partials(const float *partial, int rows, int c, double count,
                                 float *saved_mean, float *saved_var, float *run_mean,
                                 float *run_var, float *scratch, cudaStream_t st) {
    if (rows <= 0 || c <= 0) return 0;
    unsigned int *tickets = reinterpret_cast<unsigned int *>(scratch + (size_t)c * MAX_SPLITS * 4);
    const int groups = ceil_div(c, 32);
    int segs = ceil_div(2 * sm_count(), groups);
    if (segs > rows / 16) segs = rows / 16;
    if (segs > MAX_SPLITS) segs = MAX_SPLITS;
    if (segs < 1) segs = 1;
    dim3 grid(groups, segs);
    bn_stats_finalize_kernel<<<grid, 256, 0, st>>>(partial, rows, c, (float)(1.0 / count), saved_mean,
                                                   saved_var, run_mean, run_var, scratch, tickets);
    return launched();
}

extern "C" int bcnn_b200_bn_apply(const float *x, float *y, const float *mean, const float *var,
                                  const float *gamma, const float *beta, int n, int c, int hw,
                                  int act, void *stream) {
    size_t total = (size_t)n * c * hw;
    if (total == 0) return 0;
    bool vec = (hw % 4) == 0 && aligned16(x) && aligned16(y);
    static const bool lean = !getenv("BCNN_B200_NO_LEAN_BN");
    if (lean && vec && x != y && c <= 3072 && total / 4 < (1ull << 31) &&
        (act == ACT_NONE || act == ACT_RELU || act == ACT_LRELU)) {
        const uint32_t n4 = (uint32_t)(total / 4);
        const int grid = stream_grid(ceil_div_sz(n4, 4), 256);
        const size_t smem = (size_t)c * sizeof(float4);
        const FastDiv dhw4(hw / 4), dc(c);
        cudaStream_t st = as_stream(stream);
        const bool rev = reverse_streams();
        if (act == ACT_RELU)
            bn_apply_fast_kernel<ACT_RELU><<<grid, 256, smem, st>>>(x, y, mean, var, gamma, beta, n4, c, dhw4, dc, rev, l2_hints());
        else if (act == ACT_LRELU)
            bn_apply_fast_kernel<ACT_LRELU><<<grid, 256, smem, st>>>(x, y, mean, var, gamma, beta, n4, c, dhw4, dc, rev, l2_hints());
        else
            bn_apply_fast_kernel<ACT_NONE><<<grid, 256, smem, st>>>(x, y, mean, var, gamma, beta, n4, c, dhw4, dc, rev, l2_hints());
        return launched();
    }
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
                                     const float *beta, float *g_gamma, float *g_beta,
                                     float *d_mean, float *d_var, int n, int c, int hw, int act,
                                     float *scratch, void *stream) {
    size_t total = (size_t)n * c * hw;
    if (total == 0) return 0;
    cudaStream_t st = as_stream(stream);
    int splits = reduce_splits(n, c);
    unsigned int *tickets = reinterpret_cast<unsigned int *>(scratch + (size_t)c * MAX_SPLITS * 4);
    dim3 grid(c, splits);
    bn_bwd_reduce_kernel<<<grid, RT, 0, st>>>(x, y, dy, mean, var, gamma, beta, g_gamma, g_beta, d_mean,
                                              d_var, n, c, hw, act, scratch, tickets, FastDiv(hw),
                                              FastDiv(hw >> 2 ? hw >> 2 : 1),
                                              (hw % 4) == 0 && aligned16(x) && aligned16(dy) &&
                                                  (y == nullptr || aligned16(y)));
    int err = launched();
    if (err) return err;
    bool vec = (hw % 4) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx_out) &&
               (y == nullptr || aligned16(y));
    const bool remask = beta != nullptr && (act == ACT_RELU || act == ACT_LRELU);
    static const bool lean = !getenv("BCNN_B200_NO_LEAN_BN");
    if (lean && vec && c <= 1536 && total / 4 < (1ull << 31) && (remask || act == ACT_NONE)) {
        const uint32_t n4 = (uint32_t)(total / 4);
        const int grid = stream_grid(ceil_div_sz(n4, 2), 256);
        const FastDiv dhw4(hw / 4), dc(c);
        const float neg = act == ACT_LRELU ? 0.1f : 0.f;
        const bool rev = reverse_streams();
        if (remask)
            bn_bwd_apply_fast_kernel<true><<<grid, 256, (size_t)2 * c * sizeof(float4), st>>>(
                x, dy, dx_out, mean, var, gamma, beta, d_mean, d_var, n4, c, n * hw, neg, dhw4, dc, rev, l2_hints());
        else
            bn_bwd_apply_fast_kernel<false><<<grid, 256, (size_t)c * sizeof(float4), st>>>(
                x, dy, dx_out, mean, var, gamma, beta, d_mean, d_var, n4, c, n * hw, neg, dhw4, dc, rev, l2_hints());
        return launched();
    }
    bn_bwd_apply_kernel<<<stream_grid(vec ? total / 4 : total, 256), 256, 0, st>>>(
        x, y, dy, dx_out, mean, var, gamma, beta, d_mean, d_var, total, n * hw, act, FastDiv(hw),
        FastDiv(c), vec);
    return launched();
}
