// activation.cu -- standalone activations, per-channel bias, fused
// (activation-backward + bias-gradient) reduction, residual add. All HBM-bound:
// 128-bit accesses, grid sized to the SM count, no shared-memory staging (no reuse).
//
// Arithmetic follows reference src/layers/bcnn_activation_layer.c:90-226 (all ten
// activations; the reference's own .cu covers six) and src/kernels/bcnn_mat.c:761-811.
#include <cstdlib>

#include "common.cuh"

using namespace b200;

namespace {

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- elementwise activation, in place ---------------------------------------
template <bool PER_CHANNEL>
__global__ void __launch_bounds__(256)
act_fwd_kernel(float *__restrict__ x, size_t n, int act, const float *__restrict__ slope,
               FastDiv div_hw, FastDiv div_c, bool vec) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {  // hw % 4 == 0 when PER_CHANNEL, so one float4 never straddles channels
        size_t n4 = n >> 2;
        for (size_t j = tid; j < n4; j += gstride) {
            float sl = 0.f;
            if (PER_CHANNEL) {
                uint32_t q, ch;
                div_c.divmod(div_hw.div((uint32_t)(j << 2)), q, ch);
                sl = __ldg(slope + ch);
            }
            float4 v = reinterpret_cast<float4 *>(x)[j];
            v.x = act_fwd(v.x, act, sl);
            v.y = act_fwd(v.y, act, sl);
            v.z = act_fwd(v.z, act, sl);
            v.w = act_fwd(v.w, act, sl);
            reinterpret_cast<float4 *>(x)[j] = v;
        }
        for (size_t j = (n4 << 2) + tid; j < n; j += gstride) x[j] = act_fwd(x[j], act, 0.f);
    } else {
        for (size_t j = tid; j < n; j += gstride) {
            float sl = 0.f;
            if (PER_CHANNEL) {
                uint32_t q, ch;
                div_c.divmod(div_hw.div((uint32_t)j), q, ch);
                sl = __ldg(slope + ch);
            }
            x[j] = act_fwd(x[j], act, sl);
        }
    }
}

__global__ void __launch_bounds__(256)
act_bwd_kernel(const float *__restrict__ y, float *__restrict__ dy, size_t n, int act, bool vec) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t done = 0;
    if (vec) {
        size_t n4 = n >> 2;
        for (size_t j = tid; j < n4; j += gstride) {
            float4 v = reinterpret_cast<const float4 *>(y)[j];
            float4 g = reinterpret_cast<float4 *>(dy)[j];
            g.x *= act_bwd_factor(v.x, act, 0.f);
            g.y *= act_bwd_factor(v.y, act, 0.f);
            g.z *= act_bwd_factor(v.z, act, 0.f);
            g.w *= act_bwd_factor(v.w, act, 0.f);
            reinterpret_cast<float4 *>(dy)[j] = g;
        }
        done = n4 << 2;
    }
    for (size_t j = done + tid; j < n; j += gstride) dy[j] *= act_bwd_factor(y[j], act, 0.f);
}

// PReLU backward: one CTA per channel does g_slope[c] += sum dy*y*(y<0) over the batch
// (bcnn_activation_layer.c:207-211), then dy *= (y>0 ? 1 : slope[c]).
__global__ void __launch_bounds__(256)
prelu_bwd_kernel(const float *__restrict__ y, float *__restrict__ dy,
                 const float *__restrict__ slope, float *__restrict__ g_slope, int n, int c,
                 int hw) {
    __shared__ float red[8];
    const int ch = blockIdx.x;
    const float sl = slope[ch];
    float acc[1] = {0.f};
    for (int b = 0; b < n; ++b) {
        size_t off = ((size_t)b * c + ch) * hw;
        for (int i = threadIdx.x; i < hw; i += 256) {
            float v = y[off + i], g = dy[off + i];
            acc[0] += g * v * (float)(v < 0);
            dy[off + i] = g * (v > 0 ? 1.0f : sl);
        }
    }
    block_sum<1, 256>(acc, red);
    if (threadIdx.x == 0 && g_slope != nullptr) g_slope[ch] += acc[0];
}

// ---- y[n,c,:] += b[c] ----------------------------------------------------------
__global__ void __launch_bounds__(256)
add_bias_kernel(float *__restrict__ y, const float *__restrict__ bias, size_t n, FastDiv div_hw,
                FastDiv div_c, bool vec) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        size_t n4 = n >> 2;
        for (size_t j = tid; j < n4; j += gstride) {
            uint32_t q, ch;
            div_c.divmod(div_hw.div((uint32_t)(j << 2)), q, ch);
            float b = __ldg(bias + ch);
            float4 v = reinterpret_cast<float4 *>(y)[j];
            v.x += b; v.y += b; v.z += b; v.w += b;
            reinterpret_cast<float4 *>(y)[j] = v;
        }
    } else {
        for (size_t j = tid; j < n; j += gstride) {
            uint32_t q, ch;
            div_c.divmod(div_hw.div((uint32_t)j), q, ch);
            y[j] += __ldg(bias + ch);
        }
    }
}

// ---- fused activation-backward + bias gradient ---------------------------------
// grid = (c, splits). Each CTA walks its share of the batch for one channel, applies
// dy *= act'(y) in place and accumulates sum(dy'); partial sums go to scratch and the
// last CTA of a channel (atomic ticket) folds them in split order -> deterministic.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
actbwd_gradbias_kernel(float *__restrict__ gb, float *__restrict__ dy, const float *__restrict__ y,
                       int act, int n, int c, int hw, float *__restrict__ partial,
                       unsigned int *__restrict__ tickets) {
    __shared__ float red[THREADS / 32];
    __shared__ bool last;
    const int ch = blockIdx.x, split = blockIdx.y, splits = gridDim.y;
    float acc[1] = {0.f};
    const bool vec = (hw & 3) == 0;
    if (hw < THREADS) {
        // small planes (fc layers: hw == 1): the threads walk (batch item, position) pairs of this split;
        // with one batch item per pass a batch-256 fc layer was 256 dependent loads per block (0.23 ms)
        const int nb = (n - split + splits - 1) / splits;
        for (int e = threadIdx.x; e < nb * hw; e += THREADS) {
            const int bi = e / hw, i = e - bi * hw;
            const size_t o = ((size_t)(split + bi * splits) * c + ch) * hw + i;
            float g = dy[o];
            if (act != ACT_NONE) {
                g *= act_bwd_factor(y[o], act, 0.f);
                dy[o] = g;
            }
            acc[0] += g;
        }
    } else
    for (int b = split; b < n; b += splits) {
        size_t off = ((size_t)b * c + ch) * hw;
        if (vec) {
            for (int i = threadIdx.x * 4; i < hw; i += THREADS * 4) {
                float4 g = *reinterpret_cast<float4 *>(dy + off + i);
                if (act != ACT_NONE) {
                    float4 v = *reinterpret_cast<const float4 *>(y + off + i);
                    g.x *= act_bwd_factor(v.x, act, 0.f);
                    g.y *= act_bwd_factor(v.y, act, 0.f);
                    g.z *= act_bwd_factor(v.z, act, 0.f);
                    g.w *= act_bwd_factor(v.w, act, 0.f);
                    *reinterpret_cast<float4 *>(dy + off + i) = g;
                }
                acc[0] += (g.x + g.y) + (g.z + g.w);
            }
        } else {
            for (int i = threadIdx.x; i < hw; i += THREADS) {
                float g = dy[off + i];
                if (act != ACT_NONE) {
                    g *= act_bwd_factor(y[off + i], act, 0.f);
                    dy[off + i] = g;
                }
                acc[0] += g;
            }
        }
    }
    block_sum<1, THREADS>(acc, red);
    if (threadIdx.x == 0) {
        partial[(size_t)ch * splits + split] = acc[0];
        __threadfence();
        unsigned int t = atomicAdd(tickets + ch, 1u);
        last = (t == (unsigned)splits - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float s = 0.f;
        for (int i = 0; i < splits; ++i) s += __ldcg(partial + (size_t)ch * splits + i);
        gb[ch] += s;
        tickets[ch] = 0;  // re-arm for the next launch
    }
}

// ---- residual add -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
eltwise_fwd_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ y,
                   size_t n, size_t n_add, int act, bool vec) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t done = 0;
    if (vec) {
        size_t n4 = n >> 2;
        for (size_t j = tid; j < n4; j += gstride) {
            float4 u = ld_stream4(a + (j << 2));
            float4 v = ((j << 2) < n_add) ? ld_stream4(b + (j << 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 r;
            r.x = act_fwd(u.x + v.x, act, 0.f);
            r.y = act_fwd(u.y + v.y, act, 0.f);
            r.z = act_fwd(u.z + v.z, act, 0.f);
            r.w = act_fwd(u.w + v.w, act, 0.f);
            reinterpret_cast<float4 *>(y)[j] = r;
        }
        done = n4 << 2;
    }
    for (size_t j = done + tid; j < n; j += gstride)
        y[j] = act_fwd(a[j] + (j < n_add ? b[j] : 0.f), act, 0.f);
}

__global__ void __launch_bounds__(256)
eltwise_bwd_kernel(const float *__restrict__ y, float *__restrict__ dy, float *__restrict__ da,
                   float *__restrict__ db, size_t n, size_t n_add, int act, bool vec, bool acc_a,
                   bool acc_b, bool reverse) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t done = 0;
    if (vec) {
        size_t n4 = n >> 2;
        for (size_t i = tid; i < n4; i += gstride) {
            // descending: dy was just written by an ascending convolution (its tail is in L2) and
            // the head written last here is what the next per-channel reduction reads first
            const size_t j = reverse ? n4 - 1 - i : i;
            float4 g = reinterpret_cast<float4 *>(dy)[j];
            if (act != ACT_NONE) {
                float4 v = reinterpret_cast<const float4 *>(y)[j];
                g.x *= act_bwd_factor(v.x, act, 0.f);
                g.y *= act_bwd_factor(v.y, act, 0.f);
                g.z *= act_bwd_factor(v.z, act, 0.f);
                g.w *= act_bwd_factor(v.w, act, 0.f);
                reinterpret_cast<float4 *>(dy)[j] = g;
            }
            if (da) {
                float4 u = g;
                if (acc_a) {
                    const float4 o = reinterpret_cast<float4 *>(da)[j];
                    u.x += o.x; u.y += o.y; u.z += o.z; u.w += o.w;
                }
                reinterpret_cast<float4 *>(da)[j] = u;
            }
            if (db) {
                const bool in = (j << 2) < n_add;
                if (in || !acc_b) {
                    float4 w = in ? g : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (acc_b) {
                        const float4 o = reinterpret_cast<float4 *>(db)[j];
                        w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w;
                    }
                    reinterpret_cast<float4 *>(db)[j] = w;
                }
            }
        }
        done = n4 << 2;
    }
    for (size_t j = done + tid; j < n; j += gstride) {
        float g = dy[j];
        if (act != ACT_NONE) {
            g *= act_bwd_factor(y[j], act, 0.f);
            dy[j] = g;
        }
        if (da) da[j] = acc_a ? da[j] + g : g;
        if (db) {
            if (acc_b) { if (j < n_add) db[j] += g; }
            else db[j] = j < n_add ? g : 0.f;
        }
    }
}


// ---- concat / upsample (YOLO second head glue) ---------------------------
// Channel concatenation is a strided block copy: image j of a source ([block] = C_src * H * W
// floats) lands at dst + j * dst_pitch + offset. The backward pass adds (or, for the first writer
// of the step, stores) the matching slice of the output gradient.
__global__ void __launch_bounds__(256)
block_copy_kernel(const float *__restrict__ in, float *__restrict__ out, uint32_t total4, uint32_t block4,
                  size_t in_pitch, size_t out_pitch, bool accumulate, FastDiv d_block4) {
    const uint32_t gstride = gridDim.x * 256u;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total4; i += gstride) {
        uint32_t j, r;
        d_block4.divmod(i, j, r);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(in + (size_t)j * in_pitch) + r);
        float4 *dst = reinterpret_cast<float4 *>(out + (size_t)j * out_pitch) + r;
        if (accumulate) {
            const float4 o = *dst;
            *dst = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
        } else {
            *dst = v;
        }
    }
}
__global__ void __launch_bounds__(256)
block_copy_scalar_kernel(const float *__restrict__ in, float *__restrict__ out, size_t total, uint32_t block,
                         size_t in_pitch, size_t out_pitch, bool accumulate, FastDiv d_block) {
    const size_t gstride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += gstride) {
        uint32_t j, r;
        d_block.divmod((uint32_t)i, j, r);
        const float v = __ldg(in + (size_t)j * in_pitch + r);
        float *dst = out + (size_t)j * out_pitch + r;
        *dst = accumulate ? *dst + v : v;
    }
}

int launch_block_copy(const float *in, float *out, int n, int block, size_t in_pitch, size_t out_pitch,
                      bool accumulate, cudaStream_t st) {
    if (n <= 0 || block <= 0) return 0;
    const size_t total = (size_t)n * block;
    if (total >= (1ull << 32)) return (int)cudaErrorInvalidValue;
    const bool vec = (block & 3) == 0 && (in_pitch & 3) == 0 && (out_pitch & 3) == 0 && aligned16(in) &&
                     aligned16(out);
    if (vec)
        block_copy_kernel<<<stream_grid(total / 4, 256), 256, 0, st>>>(
            in, out, (uint32_t)(total / 4), (uint32_t)(block / 4), in_pitch, out_pitch, accumulate,
            FastDiv((uint32_t)(block / 4)));
    else
        block_copy_scalar_kernel<<<stream_grid(total, 256), 256, 0, st>>>(
            in, out, total, (uint32_t)block, in_pitch, out_pitch, accumulate, FastDiv((uint32_t)block));
    return launched();
}

// Nearest-neighbour upsampling by an integer factor: one thread per OUTPUT element forward (rows
// of the output are contiguous), one thread per INPUT element backward, which sums its size x size
// block of the output gradient in the reference's raster order on top of the existing value.
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, size_t total, int h, int w,
                    FastDiv d_wo, FastDiv d_ho, FastDiv d_size) {
    const size_t gstride = (size_t)gridDim.x * 256;
    for (size_t o = (size_t)blockIdx.x * 256 + threadIdx.x; o < total; o += gstride) {
        uint32_t t, i, plane, j;
        d_wo.divmod((uint32_t)o, t, i);
        d_ho.divmod(t, plane, j);
        y[o] = __ldg(x + ((size_t)plane * h + d_size.div(j)) * w + d_size.div(i));
    }
}
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const float *__restrict__ dy, float *__restrict__ dx, size_t total, int h, int w,
                    int size, bool accumulate, FastDiv d_w, FastDiv d_h) {
    const size_t gstride = (size_t)gridDim.x * 256;
    const int wo = w * size;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += gstride) {
        uint32_t t, iw, plane, ih;
        d_w.divmod((uint32_t)e, t, iw);
        d_h.divmod(t, plane, ih);
        const float *g = dy + ((size_t)plane * h * size + (size_t)ih * size) * wo + (size_t)iw * size;
        float acc = accumulate ? dx[e] : 0.f;
        for (int a = 0; a < size; ++a)
            for (int b = 0; b < size; ++b) acc += __ldg(g + (size_t)a * wo + b);
        dx[e] = acc;
    }
}

}  // namespace

extern "C" int bcnn_b200_activation_forward(float *x, int sz, int act, const float *slope, int hw,
                                            int c, void *stream) {
    if (sz <= 0 || act == ACT_NONE) return 0;
    cudaStream_t st = as_stream(stream);
    if (act == ACT_PRELU) {
        bool vec = aligned16(x) && (hw % 4) == 0;
        act_fwd_kernel<true><<<stream_grid(vec ? sz / 4 + 1 : sz, 256), 256, 0, st>>>(
            x, (size_t)sz, act, slope, FastDiv(hw), FastDiv(c), vec);
    } else {
        bool vec = aligned16(x);
        act_fwd_kernel<false><<<stream_grid(vec ? sz / 4 + 1 : sz, 256), 256, 0, st>>>(
            x, (size_t)sz, act, nullptr, FastDiv(1), FastDiv(1), vec);
    }
    return launched();
}

extern "C" int bcnn_b200_activation_backward(const float *y, float *dy, int sz, int act,
                                             const float *slope, float *g_slope, int hw, int c,
                                             void *stream) {
    if (sz <= 0 || act == ACT_NONE) return 0;
    cudaStream_t st = as_stream(stream);
    if (act == ACT_PRELU) {
        int n = sz / (hw * c);
        prelu_bwd_kernel<<<c, 256, 0, st>>>(y, dy, slope, g_slope, n, c, hw);
    } else {
        bool vec = aligned16(y) && aligned16(dy);
        act_bwd_kernel<<<stream_grid(vec ? sz / 4 + 1 : sz, 256), 256, 0, st>>>(y, dy, (size_t)sz,
                                                                                act, vec);
    }
    return launched();
}

extern "C" int bcnn_b200_add_bias(float *y, const float *bias, int n, int c, int hw,
                                  void *stream) {
    size_t total = (size_t)n * c * hw;
    if (total == 0) return 0;
    bool vec = aligned16(y) && (hw % 4) == 0;
    add_bias_kernel<<<stream_grid(vec ? total / 4 : total, 256), 256, 0, as_stream(stream)>>>(
        y, bias, total, FastDiv(hw), FastDiv(c), vec);
    return launched();
}

// scratch layout (per layer, see bcnn_b200_bn_scratch_floats): [c * 64 * 4] floats of partial
// sums, then c uint tickets. The tickets must be zero before the first launch
// (bcnn_b200_malloc zero-fills) and are re-armed by the kernel; because the ticket offset
// depends on c, a scratch buffer belongs to ONE layer (one channel count).
static int reduce_splits(int n, int c) {
    int target = 2 * sm_count();
    int s = ceil_div(target, c);
    if (s > n) s = n;
    if (s > 64) s = 64;
    return s < 1 ? 1 : s;
}

extern "C" size_t bcnn_b200_bn_scratch_floats(int c) {
    // 4 partial sums per (channel, split) + one ticket per channel, splits <= 64
    return (size_t)c * 64 * 4 + (size_t)c + 64;
}

extern "C" int bcnn_b200_actbwd_grad_bias(float *gb, float *dy, const float *y, int act, int n,
                                          int c, int hw, float *scratch, void *stream) {
    if ((size_t)n * c * hw == 0) return 0;
    int splits = reduce_splits(n, c);
    unsigned int *tickets = reinterpret_cast<unsigned int *>(scratch + (size_t)c * 64 * 4);
    dim3 grid(c, splits);
    actbwd_gradbias_kernel<256><<<grid, 256, 0, as_stream(stream)>>>(gb, dy, y, act, n, c, hw,
                                                                     scratch, tickets);
    return launched();
}

extern "C" int bcnn_b200_eltwise_forward(const float *a, const float *b, float *y, int sz,
                                         int n_add, int act, void *stream) {
    if (sz <= 0) return 0;
    bool vec = aligned16(a) && aligned16(b) && aligned16(y) && (n_add % 4) == 0;
    eltwise_fwd_kernel<<<stream_grid(vec ? sz / 4 + 1 : sz, 256), 256, 0, as_stream(stream)>>>(
        a, b, y, (size_t)sz, (size_t)n_add, act, vec);
    return launched();
}

extern "C" int bcnn_b200_eltwise_backward(const float *y, float *dy, float *da, float *db, int sz,
                                          int n_add, int act, int accumulate_flags, void *stream) {
    if (sz <= 0) return 0;
    bool vec = aligned16(y) && aligned16(dy) && aligned16(da) && aligned16(db) && (n_add % 4) == 0;
    eltwise_bwd_kernel<<<stream_grid(vec ? sz / 4 + 1 : sz, 256), 256, 0, as_stream(stream)>>>(
        y, dy, da, db, (size_t)sz, (size_t)n_add, act, vec, (accumulate_flags & 1) != 0,
        (accumulate_flags & 2) != 0, getenv("BCNN_B200_NO_REVERSE") == nullptr);
    return launched();
}

extern "C" int bcnn_b200_concat_forward(const float *src, float *dst, int n, int src_sz, int dst_sz,
                                        int dst_offset, void *stream) {
    return launch_block_copy(src, dst + dst_offset, n, src_sz, (size_t)src_sz, (size_t)dst_sz, false,
                             as_stream(stream));
}

extern "C" int bcnn_b200_concat_backward(const float *dst_grad, float *src_grad, int n, int src_sz,
                                         int dst_sz, int dst_offset, int accumulate, void *stream) {
    return launch_block_copy(dst_grad + dst_offset, src_grad, n, src_sz, (size_t)dst_sz, (size_t)src_sz,
                             accumulate != 0, as_stream(stream));
}

// YOLOv3 head, inference part of bcnn_forward_yolo_layer_cpu (src/layers/bcnn_yolo.c:226-250):
// y = x with the logistic function on the box-centre offsets (entries 0, 1) and on objectness +
// class scores (entries coords ..) of every anchor group; the size entries pass through. The
// reference does this on the host after a D2H copy of the head (:418-431); here it is one pass
// over the tensor on the device (8 B / element).
__global__ void __launch_bounds__(256)
yolo_activate_kernel(const float *__restrict__ x, float *__restrict__ y, size_t total, FastDiv d_hw,
                     FastDiv d_group, int coords) {
    const size_t gstride = (size_t)gridDim.x * 256;
    for (size_t o = (size_t)blockIdx.x * 256 + threadIdx.x; o < total; o += gstride) {
        uint32_t plane, pos, grp, entry;
        d_hw.divmod((uint32_t)o, plane, pos);     // plane = batch * channels + channel
        d_group.divmod(plane, grp, entry);        // channels is a multiple of the group size
        float v = __ldg(x + o);
        if (entry < 2u || entry >= (uint32_t)coords) v = act_fwd(v, ACT_LOGISTIC, 0.f);
        y[o] = v;
    }
}

extern "C" int bcnn_b200_yolo_activate(const float *x, float *y, int n, int boxes_per_cell,
                                       int classes, int coords, int hw, void *stream) {
    const size_t group = (size_t)coords + classes + 1;
    const size_t total = (size_t)n * boxes_per_cell * group * hw;
    if (total == 0) return 0;
    if (coords < 2 || total >= (1ull << 32)) return (int)cudaErrorInvalidValue;
    yolo_activate_kernel<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(
        x, y, total, FastDiv((uint32_t)hw), FastDiv((uint32_t)group), coords);
    return launched();
}

extern "C" int bcnn_b200_upsample_forward(const float *x, float *y, int n, int c, int h, int w, int size,
                                          void *stream) {
    const size_t total = (size_t)n * c * h * w * size * size;
    if (total == 0) return 0;
    if (size < 1 || total >= (1ull << 32)) return (int)cudaErrorInvalidValue;
    upsample_fwd_kernel<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(
        x, y, total, h, w, FastDiv((uint32_t)(w * size)), FastDiv((uint32_t)(h * size)),
        FastDiv((uint32_t)size));
    return launched();
}

extern "C" int bcnn_b200_upsample_backward(const float *dy, float *dx, int n, int c, int h, int w,
                                           int size, int accumulate, void *stream) {
    const size_t total = (size_t)n * c * h * w;
    if (total == 0) return 0;
    if (size < 1 || total * size * size >= (1ull << 32)) return (int)cudaErrorInvalidValue;
    upsample_bwd_kernel<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(
        dy, dx, total, h, w, size, accumulate != 0, FastDiv((uint32_t)w), FastDiv((uint32_t)h));
    return launched();
}
