// depthwise.cu -- depthwise k x k convolution, forward and backward, for sm_100a.
//
// HBM-bound direct convolution (one multiply-add per tap, k*k taps per output):
// one thread per output (forward) or per input (data gradient) with coalesced rows;
// the weight gradient is a per-channel reduction over (batch, positions) done as
// grid = (channels, splits) CTAs + ticketed deterministic fold (the reference's
// kernel does a racy `*p += ...` from every thread,
// src/layers/bcnn_depthwise_conv_layer.cu:113).
//
// Semantics follow the CPU path, src/layers/bcnn_depthwise_conv_layer.c:165-547:
// zero padding, bias then activation fused in forward; gW and dX both accumulate.
#include "common.cuh"

using namespace b200;

namespace {

constexpr int MAX_SPLITS = 64;

__global__ void __launch_bounds__(256)
dw_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w,
              const float *__restrict__ bias, float *__restrict__ y, int c, int h, int wd, int ho,
              int wo, int k, int stride, int pad, int act, size_t total, FastDiv d_wo,
              FastDiv d_ho, FastDiv d_c) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gstride) {
        uint32_t t, ow, oh, plane, b, ch;
        d_wo.divmod((uint32_t)o, t, ow);
        d_ho.divmod(t, plane, oh);
        d_c.divmod(plane, b, ch);
        const float *img = x + (size_t)plane * h * wd;
        const float *wk = w + (size_t)ch * k * k;
        const int ih0 = (int)oh * stride - pad, iw0 = (int)ow * stride - pad;
        float v = 0.f;
        for (int kh = 0; kh < k; ++kh) {
            int ih = ih0 + kh;
            if ((unsigned)ih >= (unsigned)h) continue;
            for (int kw = 0; kw < k; ++kw) {
                int iw = iw0 + kw;
                if ((unsigned)iw >= (unsigned)wd) continue;
                v = fmaf(__ldg(wk + kh * k + kw), __ldg(img + ih * wd + iw), v);
            }
        }
        if (bias) v += __ldg(bias + ch);
        y[o] = act_fwd(v, act, 0.f);
    }
}

// dx[n,c,ih,iw] += sum_{kh,kw} w[c,kh,kw] * dy[n,c,(ih+p-kh)/s,(iw+p-kw)/s]
__global__ void __launch_bounds__(256)
dw_bwd_data_kernel(float *__restrict__ dx, const float *__restrict__ w,
                   const float *__restrict__ dy, int c, int h, int wd, int ho, int wo, int k,
                   int stride, int pad, size_t total, FastDiv d_w, FastDiv d_h, FastDiv d_c,
                   FastDiv d_s) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gstride) {
        uint32_t t, iw, ih, plane, b, ch;
        d_w.divmod((uint32_t)e, t, iw);
        d_h.divmod(t, plane, ih);
        d_c.divmod(plane, b, ch);
        const float *g = dy + (size_t)plane * ho * wo;
        const float *wk = w + (size_t)ch * k * k;
        float acc = 0.f;
        for (int kh = 0; kh < k; ++kh) {
            int th = (int)ih + pad - kh;
            if (th < 0) break;
            uint32_t oh, rh;
            d_s.divmod(th, oh, rh);
            if (rh != 0 || (int)oh >= ho) continue;
            for (int kw = 0; kw < k; ++kw) {
                int tw = (int)iw + pad - kw;
                if (tw < 0) break;
                uint32_t ow, rw;
                d_s.divmod(tw, ow, rw);
                if (rw != 0 || (int)ow >= wo) continue;
                acc = fmaf(__ldg(wk + kh * k + kw), __ldg(g + oh * wo + ow), acc);
            }
        }
        dx[e] += acc;
    }
}

// gw[c, tap] += sum_{n,oh,ow} x[n,c,oh*s-p+kh,ow*s-p+kw] * dy[n,c,oh,ow]
// KS > 0: compile-time kernel size with all taps kept in registers (single pass);
// KS == 0: generic, one pass over the plane per tap.
template <int KS>
__global__ void __launch_bounds__(256)
dw_bwd_weight_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                     float *__restrict__ gw, int n, int c, int h, int wd, int ho, int wo, int k,
                     int stride, int pad, float *__restrict__ partial,
                     unsigned int *__restrict__ tickets, FastDiv d_wo, FastDiv d_howo) {
    constexpr int NTAP = KS > 0 ? KS * KS : 1;
    __shared__ float red[NTAP * 8];
    __shared__ bool last;
    const int ch = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
    const int kk = k * k;
    const int b0 = (int)(((long long)n * split) / splits);
    const int b1 = (int)(((long long)n * (split + 1)) / splits);
    const int howo = ho * wo;
    const uint32_t total = (uint32_t)(b1 - b0) * howo;
    float *my_partial = partial + ((size_t)ch * MAX_SPLITS + split) * kk;

    if (KS > 0) {
        float acc[NTAP];
#pragma unroll
        for (int i = 0; i < NTAP; ++i) acc[i] = 0.f;
        for (uint32_t j = threadIdx.x; j < total; j += 256) {
            uint32_t b, pos, oh, ow;
            d_howo.divmod(j, b, pos);
            d_wo.divmod(pos, oh, ow);
            const size_t plane = (size_t)(b0 + b) * c + ch;
            const float g = __ldg(dy + plane * howo + pos);
            const float *img = x + plane * h * wd;
            const int ih0 = (int)oh * stride - pad, iw0 = (int)ow * stride - pad;
#pragma unroll
            for (int kh = 0; kh < KS; ++kh) {
                int ih = ih0 + kh;
                bool hv = (unsigned)ih < (unsigned)h;
#pragma unroll
                for (int kw = 0; kw < KS; ++kw) {
                    int iw = iw0 + kw;
                    float xv = (hv && (unsigned)iw < (unsigned)wd) ? __ldg(img + ih * wd + iw) : 0.f;
                    acc[kh * KS + kw] = fmaf(xv, g, acc[kh * KS + kw]);
                }
            }
        }
        block_sum<NTAP, 256>(acc, red);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < NTAP; ++i) my_partial[i] = acc[i];
        }
    } else {
        for (int tap = 0; tap < kk; ++tap) {
            const int kh = tap / k, kw = tap - kh * k;
            float acc[1] = {0.f};
            for (uint32_t j = threadIdx.x; j < total; j += 256) {
                uint32_t b, pos, oh, ow;
                d_howo.divmod(j, b, pos);
                d_wo.divmod(pos, oh, ow);
                const size_t plane = (size_t)(b0 + b) * c + ch;
                int ih = (int)oh * stride - pad + kh, iw = (int)ow * stride - pad + kw;
                if ((unsigned)ih < (unsigned)h && (unsigned)iw < (unsigned)wd)
                    acc[0] = fmaf(__ldg(x + plane * h * wd + ih * wd + iw),
                                  __ldg(dy + plane * howo + pos), acc[0]);
            }
            block_sum<1, 256>(acc, red);
            if (threadIdx.x == 0) my_partial[tap] = acc[0];
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(tickets + ch, 1u) == (unsigned)splits - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        for (int tap = threadIdx.x; tap < kk; tap += 256) {
            float s = 0.f;
            for (int i = 0; i < splits; ++i)
                s += __ldcg(partial + ((size_t)ch * MAX_SPLITS + i) * kk + tap);
            gw[(size_t)ch * kk + tap] += s;
        }
        if (threadIdx.x == 0) tickets[ch] = 0;
    }
}


// 3x3 fast path: one thread produces four horizontally adjacent outputs of one row, so every
// input row is read once per 4 outputs (a 16-byte load plus the halo columns when the row pitch
// allows, VEC) instead of three times per output, and the taps are unrolled: the generic kernel
// above spends ~200 instructions per output and is issue-bound at 10 % of HBM.
//   FLIP:  correlate with the filter rotated by 180 degrees and add to y -- the stride-1 data
//          gradient dx += w_rot (*) dy (pad' = 2 - pad).
//   ACT:   ACT_NONE / ACT_RELU / ACT_LRELU as a template parameter (the generic act_fwd switch drags
//          its double-precision branches into the loop); other activations use the generic kernel.
template <int S, bool VEC, bool FLIP, int ACT>
__global__ void __launch_bounds__(256)
dw3_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
               float *y, int c, int h, int wd, int ho, int wo, int pad, int act, uint32_t total,
               FastDiv d_g, FastDiv d_ho, FastDiv d_c) {
    constexpr int SPAN = 3 * S + 3;   // input columns feeding four outputs: 6 (stride 1) or 9 (stride 2)
    const uint32_t gstride = gridDim.x * 256u;
    for (uint32_t o = blockIdx.x * 256u + threadIdx.x; o < total; o += gstride) {
        uint32_t t, g, oh, plane, b, ch;
        d_g.divmod(o, t, g);
        d_ho.divmod(t, plane, oh);
        d_c.divmod(plane, b, ch);
        const float *img = x + (size_t)plane * h * wd;
        float wk[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) wk[i] = __ldg(w + (size_t)ch * 9 + (FLIP ? 8 - i : i));
        const int ow0 = (int)g * 4;
        const int iw0 = ow0 * S - pad;   // leftmost input column of the span
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = (int)oh * S - pad + kh;
            if ((unsigned)ih >= (unsigned)h) continue;
            const float *row = img + (size_t)ih * wd;
            float v[SPAN];
            if (VEC) {   // pad == 1, wd % 4 == 0: columns iw0 + 1 .. are 16-byte aligned
                v[0] = iw0 >= 0 ? __ldg(row + iw0) : 0.f;
                const float4 a = __ldg(reinterpret_cast<const float4 *>(row + iw0 + 1));
                v[1] = a.x; v[2] = a.y; v[3] = a.z; v[4] = a.w;
                if (S == 1) {
                    v[5] = iw0 + 5 < wd ? __ldg(row + iw0 + 5) : 0.f;
                } else {
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (iw0 + 5 < wd) q = __ldg(reinterpret_cast<const float4 *>(row + iw0 + 5));
                    v[5] = q.x; v[6] = q.y; v[7] = q.z; v[8] = q.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < SPAN; ++i)
                    v[i] = (unsigned)(iw0 + i) < (unsigned)wd ? __ldg(row + iw0 + i) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) acc[j] = fmaf(wk[kh * 3 + kw], v[j * S + kw], acc[j]);
        }
        float *out = y + ((size_t)plane * ho + oh) * wo + ow0;
        if (FLIP) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ow0 + j < wo) out[j] += acc[j];
        } else {
            const float bv = bias ? __ldg(bias + ch) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float r = acc[j] + bv;
                if (ACT == ACT_RELU) r = r * (float)(r > 0);
                else if (ACT == ACT_LRELU) r = r > 0 ? r : 0.1f * r;
                acc[j] = r;
            }
            if ((wo & 3) == 0) {
                *reinterpret_cast<float4 *>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (ow0 + j < wo) out[j] = acc[j];
            }
        }
    }
}

// Stride-2 3x3 data gradient, pad 1: dx[ih][iw] += sum over the taps with (ih + 1 - kh) and
// (iw + 1 - kw) even. One thread owns four adjacent input columns 4g .. 4g+3 of one row; they read
// dy columns 2g, 2g+1, 2g+2 of the one (ih even: kh = 1) or two (ih odd: kh = 0, 2) matching rows.
__global__ void __launch_bounds__(256)
dw3_s2_bwd_data_kernel(float *dx, const float *__restrict__ w, const float *__restrict__ dy, int c, int h,
                       int wd, int ho, int wo, uint32_t total, FastDiv d_g, FastDiv d_h, FastDiv d_c) {
    const uint32_t gstride = gridDim.x * 256u;
    for (uint32_t e = blockIdx.x * 256u + threadIdx.x; e < total; e += gstride) {
        uint32_t t, g, ih, plane, b, ch;
        d_g.divmod(e, t, g);
        d_h.divmod(t, plane, ih);
        d_c.divmod(plane, b, ch);
        const float *wk = w + (size_t)ch * 9;
        const float *gp = dy + (size_t)plane * ho * wo;
        const int ow0 = 2 * (int)g;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            // ih even: single row kh = 1 (oh = ih / 2); ih odd: kh = 0 (oh = (ih + 1) / 2), kh = 2 (oh = (ih - 1) / 2)
            const bool odd = (ih & 1) != 0;
            if (!odd && r == 1) break;
            const int kh = odd ? (r == 0 ? 0 : 2) : 1;
            const int oh = ((int)ih + 1 - kh) >> 1;
            if (oh < 0 || oh >= ho) continue;
            const float *row = gp + (size_t)oh * wo;
            const float g0 = ow0 < wo ? __ldg(row + ow0) : 0.f;
            const float g1 = ow0 + 1 < wo ? __ldg(row + ow0 + 1) : 0.f;
            const float g2 = ow0 + 2 < wo ? __ldg(row + ow0 + 2) : 0.f;
            const float w0 = __ldg(wk + kh * 3), w1 = __ldg(wk + kh * 3 + 1), w2 = __ldg(wk + kh * 3 + 2);
            acc[0] = fmaf(w1, g0, acc[0]);
            acc[1] = fmaf(w0, g1, fmaf(w2, g0, acc[1]));
            acc[2] = fmaf(w1, g1, acc[2]);
            acc[3] = fmaf(w0, g2, fmaf(w2, g1, acc[3]));
        }
        float *out = dx + ((size_t)plane * h + ih) * wd + 4 * g;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * (int)g + j < wd) out[j] += acc[j];
    }
}

// 3x3 weight gradient with the loads of dw3_fwd_kernel: a thread accumulates the nine taps over
// groups of four adjacent outputs, one dy float4 and one input span per row and group.
template <int S, bool VEC>
__global__ void __launch_bounds__(256)
dw3_bwd_weight_kernel(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ gw,
                      int n, int c, int h, int wd, int ho, int wo, int pad, float *__restrict__ partial,
                      unsigned int *__restrict__ tickets, FastDiv d_g, FastDiv d_hog) {
    constexpr int SPAN = 3 * S + 3;
    __shared__ float red[9 * 8];
    __shared__ bool last;
    const int ch = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
    const int b0 = (int)(((long long)n * split) / splits);
    const int b1 = (int)(((long long)n * (split + 1)) / splits);
    const int groups = (wo + 3) >> 2;
    const uint32_t total = (uint32_t)(b1 - b0) * ho * groups;
    float acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0.f;
    for (uint32_t j = threadIdx.x; j < total; j += 256) {
        uint32_t b, rest, oh, g;
        d_hog.divmod(j, b, rest);
        d_g.divmod(rest, oh, g);
        const size_t plane = (size_t)(b0 + b) * c + ch;
        const int ow0 = (int)g * 4;
        const float *grow = dy + (plane * ho + oh) * wo + ow0;
        float gv[4];
        if (VEC && (wo & 3) == 0) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(grow));
            gv[0] = q.x; gv[1] = q.y; gv[2] = q.z; gv[3] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) gv[i] = ow0 + i < wo ? __ldg(grow + i) : 0.f;
        }
        const float *img = x + plane * h * wd;
        const int iw0 = ow0 * S - pad;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = (int)oh * S - pad + kh;
            if ((unsigned)ih >= (unsigned)h) continue;
            const float *row = img + (size_t)ih * wd;
            float v[SPAN];
            if (VEC) {
                v[0] = iw0 >= 0 ? __ldg(row + iw0) : 0.f;
                const float4 a = __ldg(reinterpret_cast<const float4 *>(row + iw0 + 1));
                v[1] = a.x; v[2] = a.y; v[3] = a.z; v[4] = a.w;
                if (S == 1) {
                    v[5] = iw0 + 5 < wd ? __ldg(row + iw0 + 5) : 0.f;
                } else {
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (iw0 + 5 < wd) q = __ldg(reinterpret_cast<const float4 *>(row + iw0 + 5));
                    v[5] = q.x; v[6] = q.y; v[7] = q.z; v[8] = q.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < SPAN; ++i)
                    v[i] = (unsigned)(iw0 + i) < (unsigned)wd ? __ldg(row + iw0 + i) : 0.f;
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[kh * 3 + kw] = fmaf(v[i * S + kw], gv[i], acc[kh * 3 + kw]);
        }
    }
    block_sum<9, 256>(acc, red);
    float *my_partial = partial + ((size_t)ch * MAX_SPLITS + split) * 9;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) my_partial[i] = acc[i];
        __threadfence();
        last = (atomicAdd(tickets + ch, 1u) == (unsigned)splits - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        if (threadIdx.x < 9) {
            float s = 0.f;
            for (int i = 0; i < splits; ++i)
                s += __ldcg(partial + ((size_t)ch * MAX_SPLITS + i) * 9 + threadIdx.x);
            gw[(size_t)ch * 9 + threadIdx.x] += s;
        }
        if (threadIdx.x == 0) tickets[ch] = 0;
    }
}

template <bool FLIP>
int launch_dw3(const float *x, const float *w, const float *bias, float *y, int n, int c, int h, int wd,
               int ho, int wo, int stride, int pad, int act, cudaStream_t st) {
    const int groups = ceil_div(wo, 4);
    const size_t total = (size_t)n * c * ho * groups;
    if (total >= (1ull << 32)) return -1;
    const bool vec = pad == 1 && (wd & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (FLIP || (reinterpret_cast<uintptr_t>(y) & 15) == 0);
    const int grid = stream_grid(total, 256);
    const FastDiv dg(groups), dho(ho), dc(c);
    if (act != ACT_NONE && act != ACT_RELU && act != ACT_LRELU) return -1;
#define DW3(S, V, A) dw3_fwd_kernel<S, V, FLIP, A><<<grid, 256, 0, st>>>(x, w, bias, y, c, h, wd, ho, wo, pad, act, \
                                                                      (uint32_t)total, dg, dho, dc)
#define DW3A(S, V)                                                     \
    do {                                                               \
        if (FLIP || act == ACT_NONE) DW3(S, V, ACT_NONE);              \
        else if (act == ACT_RELU) DW3(S, V, ACT_RELU);                 \
        else DW3(S, V, ACT_LRELU);                                     \
    } while (0)
    if (stride == 1) { if (vec) DW3A(1, true); else DW3A(1, false); }
    else if (!FLIP) { if (vec) DW3A(2, true); else DW3A(2, false); }
    else return -1;
#undef DW3A
#undef DW3
    return launched();
}

}  // namespace

extern "C" int bcnn_b200_depthwise_forward(const float *x, const float *w, const float *bias,
                                           int act, float *y, int n, int c, int h, int wd,
                                           int ksize, int stride, int pad, void *stream) {
    int ho = (h + 2 * pad - ksize) / stride + 1, wo = (wd + 2 * pad - ksize) / stride + 1;
    size_t total = (size_t)n * c * ho * wo;
    if (total == 0) return 0;
    if (ksize == 3 && (stride == 1 || stride == 2) && pad <= 2) {
        int err = launch_dw3<false>(x, w, bias, y, n, c, h, wd, ho, wo, stride, pad, act, as_stream(stream));
        if (err >= 0) return err;
    }
    dw_fwd_kernel<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(
        x, w, bias, y, c, h, wd, ho, wo, ksize, stride, pad, act, total, FastDiv(wo), FastDiv(ho),
        FastDiv(c));
    return launched();
}

extern "C" size_t bcnn_b200_depthwise_scratch_floats(int n, int c, int ksize) {
    (void)n;
    return (size_t)c * MAX_SPLITS * ksize * ksize + (size_t)c + 64;
}

extern "C" int bcnn_b200_depthwise_backward(const float *x, const float *w, const float *dy,
                                            float *gw, float *dx, int n, int c, int h, int wd,
                                            int ksize, int stride, int pad, float *scratch,
                                            size_t scratch_floats, void *stream) {
    int ho = (h + 2 * pad - ksize) / stride + 1, wo = (wd + 2 * pad - ksize) / stride + 1;
    if ((size_t)n * c * ho * wo == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (scratch_floats < bcnn_b200_depthwise_scratch_floats(n, c, ksize))
        return (int)cudaErrorInvalidValue;
    if (gw) {
        int splits = ceil_div(2 * sm_count(), c);
        if (splits > n) splits = n;
        if (splits > MAX_SPLITS) splits = MAX_SPLITS;
        if (splits < 1) splits = 1;
        unsigned int *tickets = reinterpret_cast<unsigned int *>(
            scratch + (size_t)c * MAX_SPLITS * ksize * ksize);
        dim3 grid(c, splits);
        const bool vec = pad == 1 && (wd & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
        if (ksize == 3 && (stride == 1 || stride == 2) && pad <= 2) {
            const int groups = ceil_div(wo, 4);
            const FastDiv dg(groups), dhog(ho * groups);
#define DW3W(S, V) dw3_bwd_weight_kernel<S, V><<<grid, 256, 0, st>>>(x, dy, gw, n, c, h, wd, ho, wo, pad, scratch, \
                                                                  tickets, dg, dhog)
            if (stride == 1) { if (vec) DW3W(1, true); else DW3W(1, false); }
            else { if (vec) DW3W(2, true); else DW3W(2, false); }
#undef DW3W
        } else if (ksize == 3)
            dw_bwd_weight_kernel<3><<<grid, 256, 0, st>>>(x, dy, gw, n, c, h, wd, ho, wo, ksize,
                                                          stride, pad, scratch, tickets,
                                                          FastDiv(wo), FastDiv(ho * wo));
        else
            dw_bwd_weight_kernel<0><<<grid, 256, 0, st>>>(x, dy, gw, n, c, h, wd, ho, wo, ksize,
                                                          stride, pad, scratch, tickets,
                                                          FastDiv(wo), FastDiv(ho * wo));
        int err = launched();
        if (err) return err;
    }
    if (dx && ksize == 3 && stride == 1 && pad <= 2 && ho + 2 - 2 * pad == h) {
        // stride 1: dx += rot180(w) (*) dy with pad' = 2 - pad, the forward kernel on dy
        int err = launch_dw3<true>(dy, w, nullptr, dx, n, c, ho, wo, h, wd, 1, 2 - pad, 0, st);
        if (err >= 0) return err;
    }
    if (dx && ksize == 3 && stride == 2 && pad == 1 && (size_t)n * c * h * ceil_div(wd, 4) < (1ull << 32)) {
        const int groups = ceil_div(wd, 4);
        const size_t total = (size_t)n * c * h * groups;
        dw3_s2_bwd_data_kernel<<<stream_grid(total, 256), 256, 0, st>>>(
            dx, w, dy, c, h, wd, ho, wo, (uint32_t)total, FastDiv(groups), FastDiv(h), FastDiv(c));
        return launched();
    }
    if (dx) {
        size_t total = (size_t)n * c * h * wd;
        dw_bwd_data_kernel<<<stream_grid(total, 256), 256, 0, st>>>(
            dx, w, dy, c, h, wd, ho, wo, ksize, stride, pad, total, FastDiv(wd), FastDiv(h),
            FastDiv(c), FastDiv(stride));
        return launched();
    }
    return 0;
}
