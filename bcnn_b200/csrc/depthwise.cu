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

}  // namespace

extern "C" int bcnn_b200_depthwise_forward(const float *x, const float *w, const float *bias,
                                           int act, float *y, int n, int c, int h, int wd,
                                           int ksize, int stride, int pad, void *stream) {
    int ho = (h + 2 * pad - ksize) / stride + 1, wo = (wd + 2 * pad - ksize) / stride + 1;
    size_t total = (size_t)n * c * ho * wo;
    if (total == 0) return 0;
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
        if (ksize == 3)
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
    if (dx) {
        size_t total = (size_t)n * c * h * wd;
        dw_bwd_data_kernel<<<stream_grid(total, 256), 256, 0, st>>>(
            dx, w, dy, c, h, wd, ho, wo, ksize, stride, pad, total, FastDiv(wd), FastDiv(h),
            FastDiv(c), FastDiv(stride));
        return launched();
    }
    return 0;
}
