// pool.cu -- max pooling (bit-exact argmax) and global average pooling for sm_100a.
//
// HBM-bound kernels: the design goal is one coalesced, vectorised pass over the
// input and one over the outputs (fwd bytes = 4*E_in + 8*E_out).
//
// Semantics follow the reference CPU path, src/layers/bcnn_maxpool_layer.c:145-191:
// window origin (i*stride, j*stride), no leading pad, out-of-image taps read
// -FLT_MAX, strict '>' (first maximum in row-major window order wins), argmax is
// the flat NCHW offset including the batch, -1 for an empty window.
#include <float.h>

#include "common.cuh"

using namespace b200;

namespace {

// ---- forward, general shape: one thread per output element -----------------
__global__ void __launch_bounds__(256)
maxpool_fwd_generic(const float *__restrict__ x, float *__restrict__ y, int *__restrict__ idx,
                    int h, int w, int k, int stride, int ho, int wo, size_t total,
                    FastDiv div_wo, FastDiv div_ho) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gstride) {
        uint32_t t, j, i, plane;
        div_wo.divmod((uint32_t)o, t, j);
        div_ho.divmod(t, plane, i);
        const int base = (int)plane * h * w;  // plane = b*c + ch
        const float *px = x + base;
        float best = -FLT_MAX;
        int best_i = -1;
        const int h0 = i * stride, w0 = j * stride;
        for (int r = 0; r < k; ++r) {
            int ih = h0 + r;
            if (ih >= h) break;
            for (int s = 0; s < k; ++s) {
                int iw = w0 + s;
                if (iw >= w) break;
                float v = __ldg(px + ih * w + iw);
                if (v > best) {
                    best = v;
                    best_i = base + ih * w + iw;
                }
            }
        }
        y[o] = best;
        idx[o] = best_i;
    }
}

// ---- forward, 2x2 / stride 2 / even width: two outputs per thread, 128-bit loads
// One thread reads a float4 from each of the two input rows (4 columns = 2 windows)
// and writes float2 + int2. Requires w % 4 == 0 (so wo % 2 == 0 and rows 16B aligned).
__global__ void __launch_bounds__(256)
maxpool_fwd_k2s2(const float *__restrict__ x, float *__restrict__ y, int *__restrict__ idx,
                 int h, int w, int ho, int wo, size_t total_pairs, FastDiv div_wo2,
                 FastDiv div_ho) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    const int wo2 = wo >> 1;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total_pairs;
         p += gstride) {
        uint32_t t, jp, i, plane;
        div_wo2.divmod((uint32_t)p, t, jp);
        div_ho.divmod(t, plane, i);
        const int base = (int)plane * h * w;
        const int ih = 2 * i, iw = 4 * jp;
        const int off0 = base + ih * w + iw;
        float4 r0 = ld_stream4(x + off0);
        float4 r1;
        const bool has_r1 = (ih + 1) < h;  // odd h with SAME padding: bottom row missing
        if (has_r1) r1 = ld_stream4(x + off0 + w);
        else r1 = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
        // window 0: columns 0,1 ; window 1: columns 2,3. Scan order (r0.a, r0.b, r1.a, r1.b).
        float b0 = -FLT_MAX, b1 = -FLT_MAX;
        int i0 = -1, i1 = -1;
        if (r0.x > b0) { b0 = r0.x; i0 = off0; }
        if (r0.y > b0) { b0 = r0.y; i0 = off0 + 1; }
        if (r1.x > b0) { b0 = r1.x; i0 = off0 + w; }
        if (r1.y > b0) { b0 = r1.y; i0 = off0 + w + 1; }
        if (r0.z > b1) { b1 = r0.z; i1 = off0 + 2; }
        if (r0.w > b1) { b1 = r0.w; i1 = off0 + 3; }
        if (r1.z > b1) { b1 = r1.z; i1 = off0 + w + 2; }
        if (r1.w > b1) { b1 = r1.w; i1 = off0 + w + 3; }
        size_t o = ((size_t)plane * ho + i) * wo + 2 * jp;
        *reinterpret_cast<float2 *>(y + o) = make_float2(b0, b1);
        *reinterpret_cast<int2 *>(idx + o) = make_int2(i0, i1);
    }
}

// ---- forward, 3x3 / stride 2 / w % 8 == 0 (ResNet stem): four outputs per thread. The 3 x 9
// input patch is read as two float4 + one scalar per row; windows are scanned in the
// reference's (row, column) order with the strict '>' so ties resolve identically.
__global__ void __launch_bounds__(256)
maxpool_fwd_k3s2(const float *__restrict__ x, float *__restrict__ y, int *__restrict__ idx,
                 int h, int w, int ho, int wo, size_t total_quads, FastDiv div_wo4,
                 FastDiv div_ho) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += gstride) {
        uint32_t t, jq, i, plane;
        div_wo4.divmod((uint32_t)q, t, jq);
        div_ho.divmod(t, plane, i);
        const int base = (int)plane * h * w;
        const int ih0 = 2 * (int)i, iw0 = 8 * (int)jq;
        const bool has_c8 = iw0 + 8 < w;
        float v[3][9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (ih0 + r < h) {
                const float *row = x + base + (ih0 + r) * w + iw0;
                const float4 a = ld_stream4(row), b = ld_stream4(row + 4);
                v[r][0] = a.x; v[r][1] = a.y; v[r][2] = a.z; v[r][3] = a.w;
                v[r][4] = b.x; v[r][5] = b.y; v[r][6] = b.z; v[r][7] = b.w;
                v[r][8] = has_c8 ? __ldg(row + 8) : -FLT_MAX;
            } else {
#pragma unroll
                for (int s = 0; s < 9; ++s) v[r][s] = -FLT_MAX;
            }
        }
        float best[4];
        int bi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float b = -FLT_MAX;
            int at = -1;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const float val = v[r][2 * j + s];
                    if (val > b) { b = val; at = base + (ih0 + r) * w + iw0 + 2 * j + s; }
                }
            best[j] = b; bi[j] = at;
        }
        const size_t o = ((size_t)plane * ho + i) * wo + 4 * jq;
        *reinterpret_cast<float4 *>(y + o) = make_float4(best[0], best[1], best[2], best[3]);
        *reinterpret_cast<int4 *>(idx + o) = make_int4(bi[0], bi[1], bi[2], bi[3]);
    }
}

// ---- backward, 3x3 / stride 2 / w % 4 == 0: four input columns of one row per thread. Input
// column iw belongs to output columns ceil((iw - 2) / 2) .. iw / 2, so the quad 4jq .. 4jq+3
// sees output columns 2jq-1 .. 2jq+1 of at most two output rows; contributions are summed in
// increasing output index (the CPU scatter order) like the generic kernel.
__global__ void __launch_bounds__(256)
maxpool_bwd_k3s2(float *__restrict__ dx, const float *__restrict__ dy,
                 const int *__restrict__ idx, int h, int w, int ho, int wo, size_t total_quads,
                 FastDiv div_w4, FastDiv div_h) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += gstride) {
        uint32_t t, jq, ih, plane;
        div_w4.divmod((uint32_t)q, t, jq);
        div_h.divmod(t, plane, ih);
        const int e0 = ((int)plane * h + (int)ih) * w + 4 * (int)jq;
        float4 d = *reinterpret_cast<float4 *>(dx + e0);
        const int oh_hi = min((int)ih >> 1, ho - 1);
        const int oh_lo = max(0, ((int)ih - 1) >> 1);  // ceil((ih - 2) / 2) for ih >= 1, 0 for ih = 0
        const int ow0 = 2 * (int)jq;
        const size_t obase = (size_t)plane * ho * wo;
        for (int oh = oh_lo; oh <= oh_hi; ++oh) {
            const size_t o = obase + (size_t)oh * wo + ow0;
            int im = -1, i0, i1 = -1;
            float gm = 0.f, g0, g1 = 0.f;
            if (ow0 > 0) { im = __ldg(idx + o - 1); gm = __ldg(dy + o - 1); }
            i0 = __ldg(idx + o); g0 = __ldg(dy + o);
            if (ow0 + 1 < wo) { i1 = __ldg(idx + o + 1); g1 = __ldg(dy + o + 1); }
            if (im == e0) d.x += gm;             // column 4jq     <- windows 2jq-1, 2jq
            if (i0 == e0) d.x += g0;
            if (i0 == e0 + 1) d.y += g0;         // column 4jq + 1 <- window 2jq
            if (i0 == e0 + 2) d.z += g0;         // column 4jq + 2 <- windows 2jq, 2jq+1
            if (i1 == e0 + 2) d.z += g1;
            if (i1 == e0 + 3) d.w += g1;         // column 4jq + 3 <- window 2jq+1
        }
        *reinterpret_cast<float4 *>(dx + e0) = d;
    }
}

// ---- backward: gather form. One thread per INPUT element sums, in increasing
// output-index order (the CPU scatter order, bcnn_maxpool_layer.c:268-271), the dy of
// every window that selected it. No atomics, deterministic.
__global__ void __launch_bounds__(256)
maxpool_bwd_generic(float *__restrict__ dx, const float *__restrict__ dy,
                    const int *__restrict__ idx, int h, int w, int k, int stride, int ho,
                    int wo, size_t total_in, FastDiv div_w, FastDiv div_h) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total_in; e += gstride) {
        uint32_t t, iw, ih, plane;
        div_w.divmod((uint32_t)e, t, iw);
        div_h.divmod(t, plane, ih);
        // windows (oh, ow) with oh*stride <= ih < oh*stride + k
        int oh_hi = min((int)ih / stride, ho - 1);
        int oh_lo = max(0, ((int)ih - k + stride) / stride);  // ceil((ih-k+1)/stride)
        int ow_hi = min((int)iw / stride, wo - 1);
        int ow_lo = max(0, ((int)iw - k + stride) / stride);
        float acc = dx[e];
        const size_t obase = (size_t)plane * ho * wo;
        for (int oh = oh_lo; oh <= oh_hi; ++oh)
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                size_t o = obase + (size_t)oh * wo + ow;
                if (__ldg(idx + o) == (int)e) acc += __ldg(dy + o);
            }
        dx[e] = acc;
    }
}

// 2x2 / stride 2 / w % 4 == 0: each input belongs to exactly one window. One thread
// handles 4 input columns of one row: reads float2 dy + int2 idx, RMW float4 dx.
__global__ void __launch_bounds__(256)
maxpool_bwd_k2s2(float *__restrict__ dx, const float *__restrict__ dy,
                 const int *__restrict__ idx, int h, int w, int ho, int wo, size_t total_quads,
                 FastDiv div_w4, FastDiv div_h) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads;
         q += gstride) {
        uint32_t t, jq, ih, plane;
        div_w4.divmod((uint32_t)q, t, jq);
        div_h.divmod(t, plane, ih);
        const int e0 = ((int)plane * h + ih) * w + 4 * jq;
        const size_t o = ((size_t)plane * ho + (ih >> 1)) * wo + 2 * jq;
        float2 g = *reinterpret_cast<const float2 *>(dy + o);
        int2 id = *reinterpret_cast<const int2 *>(idx + o);
        float4 d = *reinterpret_cast<float4 *>(dx + e0);
        if (id.x == e0) d.x += g.x;
        if (id.x == e0 + 1) d.y += g.x;
        if (id.y == e0 + 2) d.z += g.y;
        if (id.y == e0 + 3) d.w += g.y;
        *reinterpret_cast<float4 *>(dx + e0) = d;
    }
}

// ---- global average pooling: one warp per (n,c) plane --------------------------
__global__ void __launch_bounds__(256)
avgpool_fwd(const float *__restrict__ x, float *__restrict__ y, int planes, int hw) {
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < planes; p += warps_per_grid) {
        const float *px = x + (size_t)p * hw;
        float s = 0.f;
        for (int i = lane; i < hw; i += 32) s += __ldg(px + i);
        s = warp_sum(s);
        if (lane == 0) y[p] = s / hw;
    }
}

__global__ void __launch_bounds__(256)
avgpool_bwd(float *__restrict__ dx, const float *__restrict__ dy, size_t total, FastDiv div_hw,
            int hw) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gstride) {
        uint32_t p = div_hw.div((uint32_t)e);
        dx[e] += __ldg(dy + p) / hw;
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int bcnn_b200_maxpool_forward(const float *x, float *y, int *indexes, int n, int c,
                                         int h, int w, int ksize, int stride, int ho, int wo,
                                         void *stream) {
    size_t total = (size_t)n * c * ho * wo;
    if (total == 0) return 0;
    cudaStream_t st = as_stream(stream);
    // fast path covers windows fully inside the image horizontally (w % 4 == 0 => wo == w/2)
    if (ksize == 2 && stride == 2 && (w % 4) == 0 && wo == w / 2 && ho == (h + 1) / 2 &&
        aligned16(x) && aligned16(y) && aligned16(indexes)) {
        size_t pairs = total / 2;
        maxpool_fwd_k2s2<<<stream_grid(pairs, 256), 256, 0, st>>>(
            x, y, indexes, h, w, ho, wo, pairs, FastDiv(wo / 2), FastDiv(ho));
    } else if (ksize == 3 && stride == 2 && (w % 8) == 0 && wo == w / 2 && ho == (h + 1) / 2 &&
               aligned16(x) && aligned16(y) && aligned16(indexes)) {
        size_t quads = total / 4;
        maxpool_fwd_k3s2<<<stream_grid(quads, 256), 256, 0, st>>>(
            x, y, indexes, h, w, ho, wo, quads, FastDiv(wo / 4), FastDiv(ho));
    } else {
        maxpool_fwd_generic<<<stream_grid(total, 256), 256, 0, st>>>(
            x, y, indexes, h, w, ksize, stride, ho, wo, total, FastDiv(wo), FastDiv(ho));
    }
    return launched();
}

extern "C" int bcnn_b200_maxpool_backward(float *dx, const float *dy, const int *indexes, int n,
                                          int c, int h, int w, int ksize, int stride, int ho,
                                          int wo, void *stream) {
    size_t total_in = (size_t)n * c * h * w;
    if (total_in == 0 || (size_t)n * c * ho * wo == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (ksize == 2 && stride == 2 && (w % 4) == 0 && wo == w / 2 && (h % 2) == 0 &&
        ho == h / 2 && aligned16(dx) && aligned16(dy) && aligned16(indexes)) {
        size_t quads = total_in / 4;
        maxpool_bwd_k2s2<<<stream_grid(quads, 256), 256, 0, st>>>(
            dx, dy, indexes, h, w, ho, wo, quads, FastDiv(w / 4), FastDiv(h));
    } else if (ksize == 3 && stride == 2 && (w % 4) == 0 && wo == (w + 1) / 2 && ho == (h + 1) / 2 &&
               aligned16(dx)) {
        size_t quads = total_in / 4;
        maxpool_bwd_k3s2<<<stream_grid(quads, 256), 256, 0, st>>>(
            dx, dy, indexes, h, w, ho, wo, quads, FastDiv(w / 4), FastDiv(h));
    } else {
        maxpool_bwd_generic<<<stream_grid(total_in, 256), 256, 0, st>>>(
            dx, dy, indexes, h, w, ksize, stride, ho, wo, total_in, FastDiv(w), FastDiv(h));
    }
    return launched();
}

extern "C" int bcnn_b200_avgpool_forward(const float *x, float *y, int planes, int hw,
                                         void *stream) {
    if (planes <= 0 || hw <= 0) return 0;
    avgpool_fwd<<<stream_grid((size_t)planes * 32, 256), 256, 0, as_stream(stream)>>>(x, y, planes,
                                                                                    hw);
    return launched();
}

extern "C" int bcnn_b200_avgpool_backward(float *dx, const float *dy, int planes, int hw,
                                          void *stream) {
    size_t total = (size_t)planes * hw;
    if (total == 0) return 0;
    avgpool_bwd<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(dx, dy, total, FastDiv(hw),
                                                                        hw);
    return launched();
}
