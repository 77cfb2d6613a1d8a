// nhwc_bf16.cu -- the HBM-bound layers of the path on BF16 NHWC tensors (sm_100a).
//
// In the tensor-core math mode with resident activations (BCNN_B200_MATH_TC_BF16, DESIGN.md
// section 2) a convolution writes its result once as BF16 NHWC from the tcgen05 epilogue, and
// everything between two convolutions -- batch-norm apply (+activation), batch-norm backward,
// the residual add, max / average pooling -- reads and writes that format directly: half the
// bytes of the FP32 NCHW kernels of batchnorm.cu / pool.cu / activation.cu, and the operand of
// the next convolution's TMA loads with no transposition pass. FP32 NCHW copies (what
// bcnn_get_tensor_by_index hands out, reference inc/bcnn/bcnn.h:242-255) are materialised on
// demand by the two converters at the top.
//
// Element (n, h, w, c) lives at ((n * H + h) * W + w) * C + c; every kernel takes P = N * H * W
// positions x C channels with C % 8 == 0, one thread per 16-byte vector of 8 channels.
// A thread's channel group is fixed for its whole loop (the launchers choose grids whose stride
// is a multiple of C / 8), so per-channel constants live in registers.
//
// Arithmetic: the reference's formulas (src/layers/bcnn_batchnorm_layer.c:147-332, eps 1e-6
// forward / 1e-5 backward, var * sqrt(var) form; bcnn_maxpool_layer.c:145-191 first-max-wins with
// int32 flat NCHW argmax; bcnn_avgpool_layer.c:82-125; bcnn_eltwise_layer.c:111-161) evaluated in
// FP32 on BF16-rounded storage: the 2e-2 tensor-core tolerance class. The forward normalisation
// is folded to y = fma(x, a, b) with a = gamma / sqrt(var + 1e-6), b = beta - mean * a; backward
// rebuilds ReLU masks from the same fma, so mask and output agree bit for bit.
#include <cuda_bf16.h>
#include <float.h>

#include "common.cuh"
#include "conv_impl.cuh"

using namespace b200;

namespace {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_u4(const void *p) {   // coherent (in-place streams)
    return *reinterpret_cast<const uint4 *>(p);
}
__device__ __forceinline__ void st_u4(void *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LRELU) return v > 0 ? v : 0.1f * v;
    if (act == ACT_NONE) return v;
    return act_fwd(v, act, 0.f);
}

// ---- grid so that (grid * threads) % cg == 0 -----------------------------------------------
int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
int fixed_group_grid(size_t vectors, int cg, int threads, int per_thread) {
    const int unit = cg / gcd_i(cg, threads);   // grid must be a multiple of this
    size_t want = ceil_div_sz(vectors, (size_t)threads * per_thread);
    size_t cap = (size_t)sm_count() * 8;
    if (want > cap) want = cap;
    size_t g = ceil_div_sz(want ? want : 1, unit) * unit;
    return (int)g;
}

// ------------------------------------------------------------------ layout converters
// out[n][p][c] (bf16) = in[n][c][p] (f32): tiles of 64 channels x 32 positions.
__global__ void __launch_bounds__(256)
f32nchw_to_bf16nhwc_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, int C, int P) {
    __shared__ float tile[64][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *src = in + (size_t)n * C * P;
    __nv_bfloat16 *dst = out + (size_t)n * C * P;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + ty + 8 * i, p = p0 + tx;
        tile[ty + 8 * i][tx] = (c < C && p < P) ? __ldg(src + (size_t)c * P + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + ty + 8 * i, c = c0 + 2 * tx;
        if (p < P && c < C)
            *reinterpret_cast<uint32_t *>(dst + (size_t)p * C + c) =
                pack2(tile[2 * tx][ty + 8 * i], tile[2 * tx + 1][ty + 8 * i]);
    }
}
// out[n][c][p] (f32) = in[n][p][c] (bf16)
__global__ void __launch_bounds__(256)
bf16nhwc_to_f32nchw_kernel(const __nv_bfloat16 *__restrict__ in, float *__restrict__ out, int C, int P) {
    __shared__ float tile[32][65];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const __nv_bfloat16 *src = in + (size_t)n * C * P;
    float *dst = out + (size_t)n * C * P;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + ty + 8 * i, c = c0 + 2 * tx;
        float lo = 0.f, hi = 0.f;
        if (p < P && c < C) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)p * C + c));
            lo = __uint_as_float(v << 16); hi = __uint_as_float(v & 0xffff0000u);
        }
        tile[ty + 8 * i][2 * tx] = lo;
        tile[ty + 8 * i][2 * tx + 1] = hi;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + ty + 8 * i, p = p0 + tx;
        if (c < C && p < P) dst[(size_t)c * P + p] = tile[tx][ty + 8 * i];
    }
}

// ------------------------------------------------------------------ batch-norm forward apply
// y = act(fma(x, a, b)); NORMALISE: a = gamma / sqrt(var + 1e-6), b = beta - mean * a; otherwise
// (PREDICT, statistics folded at load time) a = gamma, b = beta.
template <int ACT, bool NORMALISE>
__global__ void __launch_bounds__(256)
bn_apply_nhwc_kernel(const __nv_bfloat16 *__restrict__ x, __nv_bfloat16 *__restrict__ y,
                     const float *__restrict__ mean, const float *__restrict__ var,
                     const float *__restrict__ gamma, const float *__restrict__ beta, size_t vectors,
                     int cg) {
    const size_t stride = (size_t)gridDim.x * 256;
    const size_t first = (size_t)blockIdx.x * 256 + threadIdx.x;
    const int ch0 = (int)(first % (size_t)cg) * 8;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float g = __ldg(gamma + ch0 + j), be = __ldg(beta + ch0 + j);
        if (NORMALISE) {
            a[j] = g / sqrtf(__ldg(var + ch0 + j) + 0.000001f);
            b[j] = be - __ldg(mean + ch0 + j) * a[j];
        } else {
            a[j] = g; b[j] = be;
        }
    }
    constexpr int UNROLL = 4;
    for (size_t i0 = first; i0 < vectors; i0 += stride * UNROLL) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = i0 + u * stride;
            if (i < vectors) v[u] = ld_stream_u4(x + i * 8);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = i0 + u * stride;
            if (i < vectors) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = apply_act(fmaf(f[j], a[j], b[j]), ACT);
                st_u4(y + i * 8, pack8(f));
            }
        }
    }
}

// ------------------------------------------------------------------ per-channel backward reductions
// MODE 0 (batch norm): dy' = dy * relu'(fma(x, a, b));  S1 = sum dy',  S2 = sum dy' * (x - mean)
// MODE 1 (bias):       dy' = dy * act'(y), written back; S1 = sum dy'
// MODE 2 (statistics): S1 = sum x, S2 = sum x * x (dy unused)
// Block = lanes x cgb threads (cgb channel groups, lanes position lanes); partial[(block * 2 + k) * C + ch]
template <int MODE>
__global__ void __launch_bounds__(256)
reduce_nhwc_kernel(const __nv_bfloat16 *__restrict__ x, __nv_bfloat16 *dy,
                   const float *__restrict__ mean, const float *__restrict__ var,
                   const float *__restrict__ gamma, const float *__restrict__ beta, int act, size_t P,
                   int C, int cgb, int lanes, float *__restrict__ partial) {
    extern __shared__ float red[];   // [lanes][cgb * 16]
    const int t = threadIdx.x;
    const int g = t % cgb, lane = t / cgb;
    const int cg0 = blockIdx.y * cgb;
    const int ch0 = (cg0 + g) * 8;
    const bool active = lane < lanes && ch0 < C;
    float a[8], b[8], m[8], s1[8], s2[8];
    const float neg = act == ACT_LRELU ? 0.1f : 0.f;
    const bool relu = act == ACT_RELU || act == ACT_LRELU;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = s2[j] = 0.f; a[j] = b[j] = m[j] = 0.f; }
    if (active && MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            m[j] = __ldg(mean + ch0 + j);
            a[j] = __ldg(gamma + ch0 + j) / sqrtf(__ldg(var + ch0 + j) + 0.000001f);
            b[j] = __ldg(beta + ch0 + j) - m[j] * a[j];
        }
    }
    if (active) {
        const size_t step = (size_t)gridDim.x * lanes;
        constexpr int UNROLL = 4;
        for (size_t p0 = (size_t)blockIdx.x * lanes + lane; p0 < P; p0 += step * UNROLL) {
            uint4 xv[UNROLL], gv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t p = p0 + u * step;
                if (p < P) {
                    xv[u] = ld_stream_u4(x + p * C + ch0);
                    if (MODE != 2) gv[u] = MODE == 0 ? ld_stream_u4(dy + p * C + ch0) : ld_u4(dy + p * C + ch0);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t p = p0 + u * step;
                if (p < P) {
                    float xf[8], gf[8];
                    unpack8(xv[u], xf);
                    if (MODE != 2) unpack8(gv[u], gf);
                    if (MODE == 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            s1[j] += xf[j];
                            s2[j] += xf[j] * xf[j];
                        }
                    } else if (MODE == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float d = gf[j];
                            if (relu) d *= (fmaf(xf[j], a[j], b[j]) > 0.f ? 1.0f : neg);
                            s1[j] += d;
                            s2[j] += d * (xf[j] - m[j]);
                        }
                    } else {   // x is the post-activation output y
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            gf[j] *= act_bwd_factor(xf[j], act, 0.f);
                            s1[j] += gf[j];
                        }
                        if (act != ACT_NONE) st_u4(dy + p * C + ch0, pack8(gf));
                    }
                }
            }
        }
    }
    // fold the position lanes of this block
    if (lane < lanes) {
        float *row = red + (size_t)lane * cgb * 16 + g * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) { row[j] = s1[j]; row[8 + j] = s2[j]; }
    }
    __syncthreads();
    for (int i = t; i < cgb * 16; i += 256) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[(size_t)l * cgb * 16 + i];
        const int gg = i / 16, k = (i % 16) / 8, j = i % 8;
        const int ch = (cg0 + gg) * 8 + j;
        if (ch < C) partial[((size_t)blockIdx.x * 2 + k) * C + ch] = s;
    }
}

// Fold the per-block partial rows of a reduction: block = 32 channels x 32 row lanes, lane ty sums
// rows ty, ty + 32, ... (coalesced over channels), shared memory folds the 32 lanes in lane order.
// Deterministic: the order depends only on the row count.
__device__ __forceinline__ void fold_partials(const float *__restrict__ partial, int blocks, int C, int ch,
                                              float &s1, float &s2, float (*red)[32][33]) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    float a = 0.f, b = 0.f;
    if (ch < C)
        for (int r = ty; r < blocks; r += 32) {
            a += __ldcs(partial + ((size_t)r * 2 + 0) * C + ch);
            b += __ldcs(partial + ((size_t)r * 2 + 1) * C + ch);
        }
    red[0][ty][tx] = a;
    red[1][ty][tx] = b;
    __syncthreads();
    s1 = s2 = 0.f;
    if (ty == 0) {
#pragma unroll
        for (int l = 0; l < 32; ++l) { s1 += red[0][l][tx]; s2 += red[1][l][tx]; }
    }
}

// Batch-norm backward, per channel: fold the block partials, then
//   g_beta += S1;  g_gamma += S2 / sqrt(var + 1e-6)                    (bcnn_grad_bias / _scales)
//   d_mean = gamma S1 (-1 / sqrt(var + 1e-5));  d_var = gamma S2 (-0.5 / (var sqrt(var) + 1e-5))
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_nhwc_kernel(const float *__restrict__ partial, int blocks, int C,
                            const float *__restrict__ var, const float *__restrict__ gamma,
                            float *g_gamma, float *g_beta, float *d_mean, float *d_var) {
    __shared__ float red[2][32][33];
    const int ch = blockIdx.x * 32 + threadIdx.x;
    float s1, s2;
    fold_partials(partial, blocks, C, ch, s1, s2, red);
    if (threadIdx.y != 0 || ch >= C) return;
    const float v = var[ch], g = gamma[ch];
    if (g_beta) g_beta[ch] += s1;
    if (g_gamma) g_gamma[ch] += s2 / sqrtf(v + 0.000001f);
    d_mean[ch] = (g * s1) * (-1.0f / sqrtf(v + 0.00001f));
    d_var[ch] = (g * s2) * (-0.5f / (v * sqrtf(v) + 0.00001f));
}
__global__ void __launch_bounds__(1024)
bias_bwd_finalize_nhwc_kernel(const float *__restrict__ partial, int blocks, int C, float *g_bias) {
    __shared__ float red[2][32][33];
    const int ch = blockIdx.x * 32 + threadIdx.x;
    float s1, s2;
    fold_partials(partial, blocks, C, ch, s1, s2, red);
    if (threadIdx.y != 0 || ch >= C) return;
    g_bias[ch] += s1;
}

// dx = dy' k1 + k2 (x - mean) + k3, k1 = gamma / sqrt(var + 1e-5), k2 = d_var 2 / count,
// k3 = d_mean / count (_normalize_backward, bcnn_batchnorm_layer.c:283-299). dx may alias dy.
__global__ void __launch_bounds__(256)
bn_bwd_apply_nhwc_kernel(const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *dy, __nv_bfloat16 *dx,
                         const float *__restrict__ mean, const float *__restrict__ var,
                         const float *__restrict__ gamma, const float *__restrict__ beta,
                         const float *__restrict__ d_mean, const float *__restrict__ d_var, int act,
                         float inv_count, size_t vectors, int cg) {
    const size_t stride = (size_t)gridDim.x * 256;
    const size_t first = (size_t)blockIdx.x * 256 + threadIdx.x;
    const int ch0 = (int)(first % (size_t)cg) * 8;
    const float neg = act == ACT_LRELU ? 0.1f : 0.f;
    const bool relu = act == ACT_RELU || act == ACT_LRELU;
    float a[8], b[8], m[8], k1[8], k2[8], k3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float v = __ldg(var + ch0 + j), g = __ldg(gamma + ch0 + j);
        m[j] = __ldg(mean + ch0 + j);
        a[j] = g / sqrtf(v + 0.000001f);
        b[j] = __ldg(beta + ch0 + j) - m[j] * a[j];
        k1[j] = g / sqrtf(v + 0.00001f);
        k2[j] = __ldg(d_var + ch0 + j) * 2.0f * inv_count;
        k3[j] = __ldg(d_mean + ch0 + j) * inv_count;
    }
    // four vectors per pass at two blocks per SM: fewer in flight with more blocks measured slower
    // (22.34 / 22.48 / 22.65 ms per ResNet-50 step for 4 / 2 / 1, gpurun r2zu)
    constexpr int UNROLL = 4;
    for (size_t i0 = first; i0 < vectors; i0 += stride * UNROLL) {
        uint4 xv[UNROLL], gv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = i0 + u * stride;
            if (i < vectors) { xv[u] = ld_stream_u4(x + i * 8); gv[u] = ld_u4(dy + i * 8); }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = i0 + u * stride;
            if (i < vectors) {
                float xf[8], gf[8];
                unpack8(xv[u], xf);
                unpack8(gv[u], gf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float d = gf[j];
                    if (relu) d *= (fmaf(xf[j], a[j], b[j]) > 0.f ? 1.0f : neg);
                    gf[j] = d * k1[j] + k2[j] * (xf[j] - m[j]) + k3[j];
                }
                st_u4(dx + i * 8, pack8(gf));
            }
        }
    }
}

// ------------------------------------------------------------------ batch norm + residual add
// y = act(bn_a(xa) + bn_b(xb)) in one pass: the block's last convolution (and its projection
// shortcut) never materialise their normalised outputs. Per operand: KIND 0 plain tensor, 1 batch
// norm (a = gamma / sqrt(var + 1e-6), b = beta - mean a), 2 folded scale / shift (PREDICT). The
// branches are added in FP32 (the unfused path rounds each to BF16 first: this is the closer one).
struct BnOperand { const __nv_bfloat16 *x; const float *mean, *var, *gamma, *beta; int kind; };
__device__ __forceinline__ void bn_coeffs(const BnOperand &o, int ch0, float (&a)[8], float (&b)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (o.kind == 0) { a[j] = 1.f; b[j] = 0.f; continue; }
        const float g = __ldg(o.gamma + ch0 + j), be = __ldg(o.beta + ch0 + j);
        if (o.kind == 1) {
            a[j] = g / sqrtf(__ldg(o.var + ch0 + j) + 0.000001f);
            b[j] = be - __ldg(o.mean + ch0 + j) * a[j];
        } else {
            a[j] = g; b[j] = be;
        }
    }
}
__global__ void __launch_bounds__(256)
bn_add_act_nhwc_kernel(const BnOperand oa, const BnOperand ob, __nv_bfloat16 *__restrict__ y, size_t vectors,
                       int cg, int act) {
    const size_t stride = (size_t)gridDim.x * 256;
    const size_t first = (size_t)blockIdx.x * 256 + threadIdx.x;
    const int ch0 = (int)(first % (size_t)cg) * 8;
    float aa[8], ab[8], ba[8], bb[8];
    bn_coeffs(oa, ch0, aa, ab);
    bn_coeffs(ob, ch0, ba, bb);
    constexpr int UNROLL = 4;
    for (size_t i0 = first; i0 < vectors; i0 += stride * UNROLL) {
        uint4 va[UNROLL], vb[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = i0 + u * stride;
            if (i < vectors) { va[u] = ld_stream_u4(oa.x + i * 8); vb[u] = ld_stream_u4(ob.x + i * 8); }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = i0 + u * stride;
            if (i < vectors) {
                float fa[8], fb[8];
                unpack8(va[u], fa);
                unpack8(vb[u], fb);
                // a plain operand has a = 1, b = 0: fma(x, 1, 0) is x exactly, so no per-kind branches
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    fa[j] = apply_act(fmaf(fa[j], aa[j], ab[j]) + fmaf(fb[j], ba[j], bb[j]), act);
                st_u4(y + i * 8, pack8(fa));
            }
        }
    }
}

// ------------------------------------------------------------------ residual add
__global__ void __launch_bounds__(256)
eltwise_fwd_bf16_kernel(const __nv_bfloat16 *__restrict__ a, const __nv_bfloat16 *__restrict__ b,
                        __nv_bfloat16 *__restrict__ y, size_t vectors, size_t add_vectors, int act) {
    const size_t stride = (size_t)gridDim.x * 256;
    constexpr int UNROLL = 4;
    for (size_t i0 = (size_t)blockIdx.x * 256 + threadIdx.x; i0 < vectors; i0 += stride * UNROLL) {
        uint4 av[UNROLL], bv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = i0 + u * stride;
            if (i < vectors) {
                av[u] = ld_stream_u4(a + i * 8);
                bv[u] = i < add_vectors ? ld_stream_u4(b + i * 8) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = i0 + u * stride;
            if (i < vectors) {
                float af[8], bf[8];
                unpack8(av[u], af);
                unpack8(bv[u], bf);
#pragma unroll
                for (int j = 0; j < 8; ++j) af[j] = apply_act(af[j] + bf[j], act);
                st_u4(y + i * 8, pack8(af));
            }
        }
    }
}
// dy' = dy * act'(y) (written back); da (+)= dy'; db[:add] (+)= dy', db beyond add = 0 when overwriting
__global__ void __launch_bounds__(256)
eltwise_bwd_bf16_kernel(const __nv_bfloat16 *__restrict__ y, __nv_bfloat16 *dy, __nv_bfloat16 *da,
                        __nv_bfloat16 *db, size_t vectors, size_t add_vectors, int act, int flags) {
    // one vector per iteration: a 4-way unrolled variant with every load hoisted measured slower
    // (0.738 vs 0.772 of HBM, gpurun r2t: 117 registers)
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < vectors; i += stride) {
        float yf[8], gf[8];
        unpack8(ld_stream_u4(y + i * 8), yf);
        unpack8(ld_u4(dy + i * 8), gf);
        if (act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) gf[j] *= act_bwd_factor(yf[j], act, 0.f);
            st_u4(dy + i * 8, pack8(gf));
        }
        if (da) {
            if (flags & 1) {
                float of[8];
                unpack8(ld_u4(da + i * 8), of);
#pragma unroll
                for (int j = 0; j < 8; ++j) of[j] += gf[j];
                st_u4(da + i * 8, pack8(of));
            } else {
                st_u4(da + i * 8, pack8(gf));
            }
        }
        if (db) {
            if (i < add_vectors) {
                if (flags & 2) {
                    float of[8];
                    unpack8(ld_u4(db + i * 8), of);
#pragma unroll
                    for (int j = 0; j < 8; ++j) of[j] += gf[j];
                    st_u4(db + i * 8, pack8(of));
                } else {
                    st_u4(db + i * 8, pack8(gf));
                }
            } else if (!(flags & 2)) {
                st_u4(db + i * 8, make_uint4(0, 0, 0, 0));
            }
        }
    }
}

// Residual-add backward fused with the first pass of the batch-norm backward of the branches it feeds
// (a branch = conv + BN without activation read only by this add: its incoming gradient IS the masked
// dy left here): dy' = dy * act'(y) written back, da / db copies or accumulations as above, and per
// channel S1 = sum dy', S2 = sum dy' (x_branch - mean_branch) for up to two branches, in the partial-row
// layout of reduce_nhwc_kernel<0> -- the branch's own reduction pass (a read of x and of dy') becomes
// one extra read of x here. Block = lanes x cgb threads as in reduce_nhwc_kernel.
__global__ void __launch_bounds__(256, 3)
eltwise_bwd_bn_reduce_kernel(const __nv_bfloat16 *__restrict__ y, __nv_bfloat16 *dy, __nv_bfloat16 *da,
                             __nv_bfloat16 *db, int act, int flags, size_t P, int C, int cgb, int lanes,
                             const __nv_bfloat16 *__restrict__ xa, const float *__restrict__ mean_a,
                             float *__restrict__ partial_a, const __nv_bfloat16 *__restrict__ xb,
                             const float *__restrict__ mean_b, float *__restrict__ partial_b) {
    extern __shared__ float red[];   // [lanes][cgb * 16]
    const int t = threadIdx.x;
    const int g = t % cgb, lane = t / cgb;
    const int cg0 = blockIdx.y * cgb;
    const int ch0 = (cg0 + g) * 8;
    const bool active = lane < lanes && ch0 < C;
    float ma[8], mb[8], s1[8], s2a[8], s2b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = s2a[j] = s2b[j] = 0.f; ma[j] = mb[j] = 0.f; }
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (xa) ma[j] = __ldg(mean_a + ch0 + j);
            if (xb) mb[j] = __ldg(mean_b + ch0 + j);
        }
        const size_t step = (size_t)gridDim.x * lanes;
        // one position per pass and three blocks per SM: with two positions in flight per thread the
        // kernel needed 117 registers (two blocks per SM) and ran at 0.75 of HBM, which cost what the
        // fusion saves (23.02 vs 22.38 ms per ResNet-50 step, gpurun r2zt)
        constexpr int UNROLL = 1;
        for (size_t p0 = (size_t)blockIdx.x * lanes + lane; p0 < P; p0 += step * UNROLL) {
            // every load of the pass goes out before the first use, the old value of an accumulated
            // branch gradient included (a fused branch has no da / db: at most one of ov's two uses)
            uint4 yv[UNROLL], gv[UNROLL], av[UNROLL], bv[UNROLL], ov[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t p = p0 + u * step;
                if (p < P) {
                    const size_t o = p * C + ch0;
                    yv[u] = ld_stream_u4(y + o);
                    gv[u] = ld_u4(dy + o);
                    if (xa) av[u] = ld_stream_u4(xa + o);
                    if (xb) bv[u] = ld_stream_u4(xb + o);
                    if (da && (flags & 1)) ov[u] = ld_u4(da + o);
                    else if (db && (flags & 2)) ov[u] = ld_u4(db + o);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t p = p0 + u * step;
                if (p < P) {
                    const size_t o = p * C + ch0;
                    float yf[8], gf[8], xf[8];
                    unpack8(yv[u], yf);
                    unpack8(gv[u], gf);
                    if (act != ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) gf[j] *= act_bwd_factor(yf[j], act, 0.f);
                        const uint4 packed = pack8(gf);
                        st_u4(dy + o, packed);
                        unpack8(packed, gf);   // the sums are those of the stored (BF16) gradient
                    }
                    if (da) {
                        if (flags & 1) {
                            float of[8];
                            unpack8(ov[u], of);
#pragma unroll
                            for (int j = 0; j < 8; ++j) of[j] += gf[j];
                            st_u4(da + o, pack8(of));
                        } else {
                            st_u4(da + o, pack8(gf));
                        }
                    }
                    if (db) {
                        if (flags & 2) {
                            float of[8];
                            if (da && (flags & 1)) unpack8(ld_u4(db + o), of);   // both branches accumulate: rare
                            else unpack8(ov[u], of);
#pragma unroll
                            for (int j = 0; j < 8; ++j) of[j] += gf[j];
                            st_u4(db + o, pack8(of));
                        } else {
                            st_u4(db + o, pack8(gf));
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) s1[j] += gf[j];
                    if (xa) {
                        unpack8(av[u], xf);
#pragma unroll
                        for (int j = 0; j < 8; ++j) s2a[j] += gf[j] * (xf[j] - ma[j]);
                    }
                    if (xb) {
                        unpack8(bv[u], xf);
#pragma unroll
                        for (int j = 0; j < 8; ++j) s2b[j] += gf[j] * (xf[j] - mb[j]);
                    }
                }
            }
        }
    }
    // fold the position lanes of this block, branch a then branch b
    for (int br = 0; br < 2; ++br) {
        float *partial = br == 0 ? partial_a : partial_b;
        if ((br == 0 ? xa : xb) == nullptr) continue;   // uniform over the block
        if (br == 1) __syncthreads();
        if (lane < lanes) {
            float *row = red + (size_t)lane * cgb * 16 + g * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) { row[j] = s1[j]; row[8 + j] = br == 0 ? s2a[j] : s2b[j]; }
        }
        __syncthreads();
        for (int i = t; i < cgb * 16; i += 256) {
            float s = 0.f;
            for (int l = 0; l < lanes; ++l) s += red[(size_t)l * cgb * 16 + i];
            const int gg = i / 16, k = (i % 16) / 8, j = i % 8;
            const int ch = (cg0 + gg) * 8 + j;
            if (ch < C) partial[((size_t)blockIdx.x * 2 + k) * C + ch] = s;
        }
    }
}

// ------------------------------------------------------------------ max pooling
// Window rows oh*s .. oh*s+k-1 (no leading pad, out-of-range taps skipped), strict > in (row,
// column) scan order from -FLT_MAX, so the first maximum wins; idx = flat NCHW index of the
// winner, -1 when nothing beat -FLT_MAX (bcnn_maxpool_layer.c:145-191). The index BUFFER is laid
// out like y (NHWC); its VALUES are NCHW-flat, as the reference's.
// KT: compile-time window size (2, 3) so that every tap is loaded before the first compare; 0 = any
template <int KT>
__global__ void __launch_bounds__(256)
maxpool_fwd_nhwc_kernel(const __nv_bfloat16 *__restrict__ x, __nv_bfloat16 *__restrict__ y,
                        int *__restrict__ idx, int N, int C, int H, int W, int k, int s, int Ho, int Wo,
                        size_t vectors, int cg) {
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < vectors; i += stride) {
        const int g = (int)(i % (size_t)cg);
        size_t pos = i / (size_t)cg;
        const int ow = (int)(pos % Wo); pos /= Wo;
        const int oh = (int)(pos % Ho);
        const int n = (int)(pos / Ho);
        float best[8];
        int arg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -FLT_MAX; arg[j] = -1; }
        if (KT > 0) {
            uint4 v[KT > 0 ? KT * KT : 1];
#pragma unroll
            for (int t = 0; t < KT * KT; ++t) {
                const int ih = oh * s + t / KT, iw = ow * s + t % KT;
                if (ih < H && iw < W) v[t] = ld_stream_u4(x + (((size_t)n * H + ih) * W + iw) * C + g * 8);
            }
#pragma unroll
            for (int t = 0; t < KT * KT; ++t) {
                const int ih = oh * s + t / KT, iw = ow * s + t % KT;
                if (ih < H && iw < W) {
                    float f[8];
                    unpack8(v[t], f);
                    const int base = ((n * C + g * 8) * H + ih) * W + iw;   // channel j adds j * H * W
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (f[j] > best[j]) { best[j] = f[j]; arg[j] = base + j * H * W; }
                }
            }
        } else
        for (int kh = 0; kh < k; ++kh) {
            const int ih = oh * s + kh;
            if (ih >= H) break;
            for (int kw = 0; kw < k; ++kw) {
                const int iw = ow * s + kw;
                if (iw >= W) break;
                float f[8];
                unpack8(ld_stream_u4(x + (((size_t)n * H + ih) * W + iw) * C + g * 8), f);
                const int base = ((n * C + g * 8) * H + ih) * W + iw;   // channel j adds j * H * W
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (f[j] > best[j]) { best[j] = f[j]; arg[j] = base + j * H * W; }
            }
        }
        st_u4(y + i * 8, pack8(best));
        int4 *ip = reinterpret_cast<int4 *>(idx + i * 8);
        ip[0] = make_int4(arg[0], arg[1], arg[2], arg[3]);
        ip[1] = make_int4(arg[4], arg[5], arg[6], arg[7]);
    }
}
// dx[winner] += dy, gather form: every input element sums the windows that elected it, in the
// window order of the reference's scatter loop (bcnn_maxpool_layer.c:258-273).
__global__ void __launch_bounds__(256)
maxpool_bwd_nhwc_kernel(__nv_bfloat16 *dx, const __nv_bfloat16 *__restrict__ dy,
                        const int *__restrict__ idx, int N, int C, int H, int W, int k, int s, int Ho,
                        int Wo, size_t vectors, int cg, int accumulate) {
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < vectors; i += stride) {
        const int g = (int)(i % (size_t)cg);
        size_t pos = i / (size_t)cg;
        const int iw = (int)(pos % W); pos /= W;
        const int ih = (int)(pos % H);
        const int n = (int)(pos / H);
        float acc[8];
        if (accumulate) unpack8(ld_u4(dx + i * 8), acc);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        }
        const int base = ((n * C + g * 8) * H + ih) * W + iw;
        int oh0 = ih - k + 1; oh0 = oh0 > 0 ? (oh0 + s - 1) / s : 0;
        int ow0 = iw - k + 1; ow0 = ow0 > 0 ? (ow0 + s - 1) / s : 0;
        const int oh1 = min(ih / s, Ho - 1), ow1 = min(iw / s, Wo - 1);
        if (k <= 2 * s) {
            // an element sits in at most 2 x 2 windows: issue every load before the first use
            int4 i0[4], i1[4];
            uint4 dv[4];
            bool live[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int oh = oh0 + (q >> 1), ow = ow0 + (q & 1);
                live[q] = oh <= oh1 && ow <= ow1;
                if (live[q]) {
                    const size_t o = ((((size_t)n * Ho + oh) * Wo + ow) * C + g * 8);
                    i0[q] = __ldg(reinterpret_cast<const int4 *>(idx + o));
                    i1[q] = __ldg(reinterpret_cast<const int4 *>(idx + o) + 1);
                    dv[q] = ld_stream_u4(dy + o);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (live[q]) {
                    float f[8];
                    unpack8(dv[q], f);
                    const int id[8] = {i0[q].x, i0[q].y, i0[q].z, i0[q].w, i1[q].x, i1[q].y, i1[q].z, i1[q].w};
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (id[j] == base + j * H * W) acc[j] += f[j];
                }
            }
        } else
        for (int oh = oh0; oh <= oh1; ++oh)
            for (int ow = ow0; ow <= ow1; ++ow) {
                const size_t o = ((((size_t)n * Ho + oh) * Wo + ow) * C + g * 8);
                const int4 i0 = __ldg(reinterpret_cast<const int4 *>(idx + o));
                const int4 i1 = __ldg(reinterpret_cast<const int4 *>(idx + o) + 1);
                float f[8];
                unpack8(ld_stream_u4(dy + o), f);
                const int id[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (id[j] == base + j * H * W) acc[j] += f[j];
            }
        st_u4(dx + i * 8, pack8(acc));
    }
}

// ------------------------------------------------------------------ global average pooling
// y[n][c] (f32) = sum over the plane / HW; one CTA per (image, 32 channel groups)
__global__ void __launch_bounds__(256)
avgpool_fwd_nhwc_kernel(const __nv_bfloat16 *__restrict__ x, float *__restrict__ y, int C, int HW) {
    __shared__ float red[8][33][8];
    const int n = blockIdx.y, t = threadIdx.x;
    const int g = blockIdx.x * 32 + (t & 31), lane = t >> 5;   // 8 position lanes
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g * 8 < C) {
        for (int p = lane; p < HW; p += 8) {
            float f[8];
            unpack8(ld_stream_u4(x + ((size_t)n * HW + p) * C + g * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += f[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[lane][t & 31][j] = s[j];
    __syncthreads();
    if (lane == 0 && g * 8 < C) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float tot = 0.f;
            for (int l = 0; l < 8; ++l) tot += red[l][t & 31][j];
            y[(size_t)n * C + g * 8 + j] = tot / (float)HW;
        }
    }
}
// dx[n][p][c] (+)= dy[n][c] / HW
__global__ void __launch_bounds__(256)
avgpool_bwd_nhwc_kernel(__nv_bfloat16 *dx, const float *__restrict__ dy, int C, int HW, size_t vectors,
                        int cg, int accumulate) {
    const size_t stride = (size_t)gridDim.x * 256;
    const float inv = 1.0f / (float)HW;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < vectors; i += stride) {
        const int g = (int)(i % (size_t)cg);
        const size_t n = i / ((size_t)cg * HW);
        const float4 d0 = __ldg(reinterpret_cast<const float4 *>(dy + n * C + g * 8));
        const float4 d1 = __ldg(reinterpret_cast<const float4 *>(dy + n * C + g * 8) + 1);
        float f[8] = {d0.x * inv, d0.y * inv, d0.z * inv, d0.w * inv, d1.x * inv, d1.y * inv, d1.z * inv, d1.w * inv};
        if (accumulate) {
            float o[8];
            unpack8(ld_u4(dx + i * 8), o);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += o[j];
        }
        st_u4(dx + i * 8, pack8(f));
    }
}

// Stride-2 windows of size 2 or 3: one thread owns the 2 x 2 input block (2i .. 2i+1, 2j .. 2j+1) of
// one channel group. Only the windows (i-1 .. i) x (j-1 .. j) touch it (k = 2: just (i, j)), so each
// window's index / gradient vectors are loaded once per block instead of once per input element: a
// quarter of the L2 traffic of the per-element gather above (which bound it: 0.58 ms for ResNet-50's
// stem pool at batch 256 against 0.11 ms of HBM time). Same accumulation order per element.
template <int KT>
__global__ void __launch_bounds__(256)
maxpool_bwd_s2_nhwc_kernel(__nv_bfloat16 *dx, const __nv_bfloat16 *__restrict__ dy,
                           const int *__restrict__ idx, int N, int C, int H, int W, int Ho, int Wo,
                           int Hb, int Wb, size_t blocks, int cg, int accumulate) {
    const size_t stride = (size_t)gridDim.x * 256;
    constexpr int NW = KT == 3 ? 2 : 1;   // windows per axis that can touch the block
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < blocks; i += stride) {
        const int g = (int)(i % (size_t)cg);
        size_t pos = i / (size_t)cg;
        const int bj = (int)(pos % Wb); pos /= Wb;
        const int bi = (int)(pos % Hb);
        const int n = (int)(pos / Hb);
        int4 i0[NW * NW], i1[NW * NW];
        uint4 dv[NW * NW];
        bool live[NW * NW];
#pragma unroll
        for (int q = 0; q < NW * NW; ++q) {
            const int oh = bi - (NW - 1) + q / NW, ow = bj - (NW - 1) + q % NW;
            live[q] = oh >= 0 && oh < Ho && ow >= 0 && ow < Wo;
            if (live[q]) {
                const size_t o = ((((size_t)n * Ho + oh) * Wo + ow) * C + g * 8);
                i0[q] = __ldg(reinterpret_cast<const int4 *>(idx + o));
                i1[q] = __ldg(reinterpret_cast<const int4 *>(idx + o) + 1);
                dv[q] = ld_stream_u4(dy + o);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ih = 2 * bi + (e >> 1), iw = 2 * bj + (e & 1);
            if (ih >= H || iw >= W) continue;
            const size_t xo = ((((size_t)n * H + ih) * W + iw) * C + g * 8);
            float acc[8];
            if (accumulate) unpack8(ld_u4(dx + xo), acc);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            }
            const int base = ((n * C + g * 8) * H + ih) * W + iw;
#pragma unroll
            for (int q = 0; q < NW * NW; ++q) {
                if (live[q]) {
                    float f[8];
                    unpack8(dv[q], f);
                    const int id[8] = {i0[q].x, i0[q].y, i0[q].z, i0[q].w, i1[q].x, i1[q].y, i1[q].z, i1[q].w};
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (id[j] == base + j * H * W) acc[j] += f[j];
                }
            }
            st_u4(dx + xo, pack8(acc));
        }
    }
}

struct ReducePlan { int cgb, lanes, gx, gy; size_t smem; };
ReducePlan plan_reduce(size_t P, int C, int blocks_per_sm = 2) {
    ReducePlan r;
    const int cg = C / 8;
    r.cgb = cg < 256 ? cg : 256;
    r.lanes = 256 / r.cgb;
    r.gy = ceil_div(cg, r.cgb);
    size_t want = ceil_div_sz(P, (size_t)r.lanes * 8);
    size_t cap = (size_t)blocks_per_sm * sm_count() / r.gy;
    if (cap < 1) cap = 1;
    r.gx = (int)(want < cap ? (want ? want : 1) : cap);
    r.smem = (size_t)r.lanes * r.cgb * 16 * sizeof(float);
    return r;
}

}  // namespace

// ---------------------------------------------------------------------------------- C ABI
extern "C" int bcnn_b200_f32nchw_to_bf16nhwc(const float *in, void *out, int n, int c, int hw, void *stream) {
    if ((size_t)n * c * hw == 0) return 0;
    if (c % 2) return (int)cudaErrorInvalidValue;
    dim3 grid(ceil_div(hw, 32), ceil_div(c, 64), n);
    f32nchw_to_bf16nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, reinterpret_cast<__nv_bfloat16 *>(out), c, hw);
    return launched();
}
extern "C" int bcnn_b200_bf16nhwc_to_f32nchw(const void *in, float *out, int n, int c, int hw, void *stream) {
    if ((size_t)n * c * hw == 0) return 0;
    if (c % 2) return (int)cudaErrorInvalidValue;
    dim3 grid(ceil_div(hw, 32), ceil_div(c, 64), n);
    bf16nhwc_to_f32nchw_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16 *>(in), out, c, hw);
    return launched();
}

extern "C" size_t bcnn_b200_nhwc_scratch_floats(int c) { return (size_t)2 * 4 * sm_count() * c + 64; }   // up to 4 partial rows per SM

extern "C" int bcnn_b200_bn_stats_nhwc(const void *x, size_t positions, int c, float *saved_mean,
                                       float *saved_var, float *run_mean, float *run_var,
                                       float *nhwc_scratch, float *scratch, void *stream) {
    if (positions == 0 || c == 0) return 0;
    if (c % 8) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    const ReducePlan r = plan_reduce(positions, c);
    reduce_nhwc_kernel<2><<<dim3(r.gx, r.gy), 256, r.smem, st>>>(
        reinterpret_cast<const __nv_bfloat16 *>(x), nullptr, nullptr, nullptr, nullptr, nullptr, ACT_NONE,
        positions, c, r.cgb, r.lanes, nhwc_scratch);
    int err = launched();
    if (err) return err;
    return b200::bn_stats_from_partials(nhwc_scratch, r.gx, c, (double)positions, saved_mean, saved_var,
                                        run_mean, run_var, scratch, st);
}

extern "C" int bcnn_b200_bn_apply_nhwc(const void *x, void *y, const float *mean, const float *var,
                                       const float *gamma, const float *beta, size_t positions, int c,
                                       int act, void *stream) {
    if (positions == 0 || c == 0) return 0;
    if (c % 8) return (int)cudaErrorInvalidValue;
    const int cg = c / 8;
    const size_t vectors = positions * cg;
    const int grid = fixed_group_grid(vectors, cg, 256, 4);
    cudaStream_t st = as_stream(stream);
    const __nv_bfloat16 *xb = reinterpret_cast<const __nv_bfloat16 *>(x);
    __nv_bfloat16 *yb = reinterpret_cast<__nv_bfloat16 *>(y);
    const bool norm = mean != nullptr;
#define BN_APPLY(A, N) bn_apply_nhwc_kernel<A, N><<<grid, 256, 0, st>>>(xb, yb, mean, var, gamma, beta, vectors, cg)
    if (norm) {
        if (act == ACT_RELU) BN_APPLY(ACT_RELU, true);
        else if (act == ACT_LRELU) BN_APPLY(ACT_LRELU, true);
        else if (act == ACT_NONE) BN_APPLY(ACT_NONE, true);
        else return (int)cudaErrorInvalidValue;
    } else {
        if (act == ACT_RELU) BN_APPLY(ACT_RELU, false);
        else if (act == ACT_LRELU) BN_APPLY(ACT_LRELU, false);
        else if (act == ACT_NONE) BN_APPLY(ACT_NONE, false);
        else return (int)cudaErrorInvalidValue;
    }
#undef BN_APPLY
    return launched();
}

extern "C" int bcnn_b200_bn_backward_nhwc(const void *x, void *dy, void *dx, const float *mean,
                                          const float *var, const float *gamma, const float *beta,
                                          float *g_gamma, float *g_beta, float *d_mean, float *d_var,
                                          size_t positions, int c, int act, float *scratch, void *stream) {
    if (positions == 0 || c == 0) return 0;
    if (c % 8 || !(act == ACT_NONE || act == ACT_RELU || act == ACT_LRELU)) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    const ReducePlan r = plan_reduce(positions, c);
    const __nv_bfloat16 *xb = reinterpret_cast<const __nv_bfloat16 *>(x);
    __nv_bfloat16 *gb = reinterpret_cast<__nv_bfloat16 *>(dy);
    reduce_nhwc_kernel<0><<<dim3(r.gx, r.gy), 256, r.smem, st>>>(xb, gb, mean, var, gamma, beta, act,
                                                                 positions, c, r.cgb, r.lanes, scratch);
    int err = launched();
    if (err) return err;
    bn_bwd_finalize_nhwc_kernel<<<ceil_div(c, 32), dim3(32, 32), 0, st>>>(scratch, r.gx, c, var, gamma, g_gamma,
                                                                            g_beta, d_mean, d_var);
    err = launched();
    if (err) return err;
    const int cg = c / 8;
    const size_t vectors = positions * cg;
    bn_bwd_apply_nhwc_kernel<<<fixed_group_grid(vectors, cg, 256, 4), 256, 0, st>>>(
        xb, gb, reinterpret_cast<__nv_bfloat16 *>(dx), mean, var, gamma, beta, d_mean, d_var, act,
        1.0f / (float)positions, vectors, cg);
    return launched();
}

// Second half of bcnn_b200_bn_backward_nhwc for a layer whose reduction pass was fused into the
// residual add behind it (bcnn_b200_eltwise_backward_bn_reduce_bf16): finalize from its partial rows,
// then apply. act must be NONE (the fused layers have no activation of their own).
extern "C" int bcnn_b200_bn_backward_nhwc_partials(const void *x, void *dy, void *dx, const float *mean,
                                                   const float *var, const float *gamma, const float *beta,
                                                   float *g_gamma, float *g_beta, float *d_mean, float *d_var,
                                                   size_t positions, int c, const float *partial, int rows,
                                                   void *stream) {
    if (positions == 0 || c == 0) return 0;
    if (c % 8 || rows <= 0 || !partial) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    bn_bwd_finalize_nhwc_kernel<<<ceil_div(c, 32), dim3(32, 32), 0, st>>>(partial, rows, c, var, gamma, g_gamma,
                                                                            g_beta, d_mean, d_var);
    int err = launched();
    if (err) return err;
    const int cg = c / 8;
    const size_t vectors = positions * cg;
    bn_bwd_apply_nhwc_kernel<<<fixed_group_grid(vectors, cg, 256, 4), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16 *>(x), reinterpret_cast<__nv_bfloat16 *>(dy),
        reinterpret_cast<__nv_bfloat16 *>(dx), mean, var, gamma, beta, d_mean, d_var, ACT_NONE,
        1.0f / (float)positions, vectors, cg);
    return launched();
}

extern "C" int bcnn_b200_eltwise_backward_bn_reduce_bf16(const void *y, void *dy, void *da, void *db,
                                                         size_t positions, int c, int act, int accumulate_flags,
                                                         const void *xa, const float *mean_a, float *partial_a,
                                                         const void *xb, const float *mean_b, float *partial_b,
                                                         int *rows, void *stream) {
    if (rows) *rows = 0;
    if (positions == 0 || c == 0) return 0;
    if (c % 8 || !(act == ACT_NONE || act == ACT_RELU || act == ACT_LRELU) || (!xa && !xb) ||
        (xa && (!mean_a || !partial_a)) || (xb && (!mean_b || !partial_b)))
        return (int)cudaErrorInvalidValue;
    const ReducePlan r = plan_reduce(positions, c, 3);   // three blocks per SM
    eltwise_bwd_bn_reduce_kernel<<<dim3(r.gx, r.gy), 256, r.smem, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16 *>(y), reinterpret_cast<__nv_bfloat16 *>(dy),
        reinterpret_cast<__nv_bfloat16 *>(da), reinterpret_cast<__nv_bfloat16 *>(db), act, accumulate_flags,
        positions, c, r.cgb, r.lanes, reinterpret_cast<const __nv_bfloat16 *>(xa), mean_a, partial_a,
        reinterpret_cast<const __nv_bfloat16 *>(xb), mean_b, partial_b);
    if (rows) *rows = r.gx;
    return launched();
}

extern "C" int bcnn_b200_actbwd_grad_bias_nhwc(float *g_bias, void *dy, const void *y, int act,
                                               size_t positions, int c, float *scratch, void *stream) {
    if (positions == 0 || c == 0) return 0;
    if (c % 8) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    const ReducePlan r = plan_reduce(positions, c);
    // without an activation the kernel only reads dy; y may then be NULL: pass dy as a stand-in
    const __nv_bfloat16 *yb = reinterpret_cast<const __nv_bfloat16 *>(act == ACT_NONE || !y ? dy : y);
    reduce_nhwc_kernel<1><<<dim3(r.gx, r.gy), 256, r.smem, st>>>(yb, reinterpret_cast<__nv_bfloat16 *>(dy),
                                                                 nullptr, nullptr, nullptr, nullptr, act,
                                                                 positions, c, r.cgb, r.lanes, scratch);
    int err = launched();
    if (err) return err;
    bias_bwd_finalize_nhwc_kernel<<<ceil_div(c, 32), dim3(32, 32), 0, st>>>(scratch, r.gx, c, g_bias);
    return launched();
}

extern "C" int bcnn_b200_bn_add_act_nhwc(const void *xa, const float *mean_a, const float *var_a,
                                         const float *gamma_a, const float *beta_a, const void *xb,
                                         const float *mean_b, const float *var_b, const float *gamma_b,
                                         const float *beta_b, void *y, size_t positions, int c, int act,
                                         void *stream) {
    if (positions == 0 || c == 0) return 0;
    if (c % 8) return (int)cudaErrorInvalidValue;
    BnOperand oa{reinterpret_cast<const __nv_bfloat16 *>(xa), mean_a, var_a, gamma_a, beta_a,
                 gamma_a ? (mean_a ? 1 : 2) : 0};
    BnOperand ob{reinterpret_cast<const __nv_bfloat16 *>(xb), mean_b, var_b, gamma_b, beta_b,
                 gamma_b ? (mean_b ? 1 : 2) : 0};
    const int cg = c / 8;
    const size_t vectors = positions * cg;
    bn_add_act_nhwc_kernel<<<fixed_group_grid(vectors, cg, 256, 4), 256, 0, as_stream(stream)>>>(
        oa, ob, reinterpret_cast<__nv_bfloat16 *>(y), vectors, cg, act);
    return launched();
}

extern "C" int bcnn_b200_eltwise_forward_bf16(const void *a, const void *b, void *y, size_t sz,
                                              size_t n_add, int act, void *stream) {
    if (sz == 0) return 0;
    if (sz % 8 || n_add % 8) return (int)cudaErrorInvalidValue;
    eltwise_fwd_bf16_kernel<<<stream_grid(sz / 32, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16 *>(a), reinterpret_cast<const __nv_bfloat16 *>(b),
        reinterpret_cast<__nv_bfloat16 *>(y), sz / 8, n_add / 8, act);
    return launched();
}
extern "C" int bcnn_b200_eltwise_backward_bf16(const void *y, void *dy, void *da, void *db, size_t sz,
                                               size_t n_add, int act, int accumulate_flags, void *stream) {
    if (sz == 0) return 0;
    if (sz % 8 || n_add % 8) return (int)cudaErrorInvalidValue;
    eltwise_bwd_bf16_kernel<<<stream_grid(sz / 8, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16 *>(y), reinterpret_cast<__nv_bfloat16 *>(dy),
        reinterpret_cast<__nv_bfloat16 *>(da), reinterpret_cast<__nv_bfloat16 *>(db), sz / 8, n_add / 8, act,
        accumulate_flags);
    return launched();
}

extern "C" int bcnn_b200_maxpool_forward_nhwc(const void *x, void *y, int *indexes, int n, int c, int h,
                                              int w, int ksize, int stride, int ho, int wo, void *stream) {
    const size_t vectors = (size_t)n * ho * wo * (c / 8);
    if (vectors == 0) return 0;
    if (c % 8 || (size_t)n * c * h * w >= (1ull << 31)) return (int)cudaErrorInvalidValue;
    const int grid = stream_grid(vectors, 256);
    const __nv_bfloat16 *xb = reinterpret_cast<const __nv_bfloat16 *>(x);
    __nv_bfloat16 *yb = reinterpret_cast<__nv_bfloat16 *>(y);
    cudaStream_t st = as_stream(stream);
    if (ksize == 2)
        maxpool_fwd_nhwc_kernel<2><<<grid, 256, 0, st>>>(xb, yb, indexes, n, c, h, w, ksize, stride, ho, wo, vectors, c / 8);
    else if (ksize == 3)
        maxpool_fwd_nhwc_kernel<3><<<grid, 256, 0, st>>>(xb, yb, indexes, n, c, h, w, ksize, stride, ho, wo, vectors, c / 8);
    else
        maxpool_fwd_nhwc_kernel<0><<<grid, 256, 0, st>>>(xb, yb, indexes, n, c, h, w, ksize, stride, ho, wo, vectors, c / 8);
    return launched();
}
extern "C" int bcnn_b200_maxpool_backward_nhwc(void *dx, const void *dy, const int *indexes, int n, int c,
                                               int h, int w, int ksize, int stride, int ho, int wo,
                                               int accumulate, void *stream) {
    const size_t vectors = (size_t)n * h * w * (c / 8);
    if (vectors == 0) return 0;
    if (c % 8) return (int)cudaErrorInvalidValue;
    if (stride == 2 && (ksize == 2 || ksize == 3)) {
        const int hb = (h + 1) / 2, wb = (w + 1) / 2;
        const size_t blocks = (size_t)n * hb * wb * (c / 8);
        __nv_bfloat16 *dxb = reinterpret_cast<__nv_bfloat16 *>(dx);
        const __nv_bfloat16 *dyb = reinterpret_cast<const __nv_bfloat16 *>(dy);
        if (ksize == 3)
            maxpool_bwd_s2_nhwc_kernel<3><<<stream_grid(blocks, 256), 256, 0, as_stream(stream)>>>(
                dxb, dyb, indexes, n, c, h, w, ho, wo, hb, wb, blocks, c / 8, accumulate);
        else
            maxpool_bwd_s2_nhwc_kernel<2><<<stream_grid(blocks, 256), 256, 0, as_stream(stream)>>>(
                dxb, dyb, indexes, n, c, h, w, ho, wo, hb, wb, blocks, c / 8, accumulate);
        return launched();
    }
    maxpool_bwd_nhwc_kernel<<<stream_grid(vectors, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<__nv_bfloat16 *>(dx), reinterpret_cast<const __nv_bfloat16 *>(dy), indexes, n, c, h, w,
        ksize, stride, ho, wo, vectors, c / 8, accumulate);
    return launched();
}

extern "C" int bcnn_b200_avgpool_forward_nhwc(const void *x, float *y, int n, int c, int hw, void *stream) {
    if ((size_t)n * c * hw == 0) return 0;
    if (c % 8) return (int)cudaErrorInvalidValue;
    avgpool_fwd_nhwc_kernel<<<dim3(ceil_div(c / 8, 32), n), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16 *>(x), y, c, hw);
    return launched();
}
extern "C" int bcnn_b200_avgpool_backward_nhwc(void *dx, const float *dy, int n, int c, int hw,
                                               int accumulate, void *stream) {
    const size_t vectors = (size_t)n * hw * (c / 8);
    if (vectors == 0) return 0;
    if (c % 8) return (int)cudaErrorInvalidValue;
    avgpool_bwd_nhwc_kernel<<<stream_grid(vectors, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<__nv_bfloat16 *>(dx), dy, c, hw, vectors, c / 8, accumulate);
    return launched();
}
