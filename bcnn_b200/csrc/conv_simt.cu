// conv_simt.cu -- FP32 SIMT implicit-GEMM convolution (fprop / dgrad / wgrad).
//
// This is the 1e-5 verification path and the fallback for shapes the tcgen05
// kernel does not take (K < 32 first layers, grouped convs). No im2col buffer is
// ever materialised: the B operand is gathered from the NCHW tensor while the tile
// is staged into shared memory, the whole batch is folded into the GEMM N (fprop,
// dgrad) or K (wgrad) dimension, and bias + activation are fused into the fprop
// epilogue.
//
//   fprop : Y[co, (b,oh,ow)]  = sum_{(c,kh,kw)} W[co,(c,kh,kw)] * X[b,c,oh*s-p+kh,ow*s-p+kw]
//   dgrad : dX[c, (b,ih,iw)]  = sum_{(co,kh,kw)} W[co,c,kh,kw]  * dY[b,co,(ih+p-kh)/s,(iw+p-kw)/s]
//   wgrad : dW[co,(c,kh,kw)] += sum_{(b,oh,ow)} dY[b,co,oh,ow]  * X[b,c,oh*s-p+kh,ow*s-p+kw]
//
// Replaces the reference's per-image im2col + GEMM (+ col2im) loops,
// src/layers/bcnn_conv_layer.c:438-462 (fwd) and :534-583 (bwd), with the semantics
// kept: wgrad accumulates into gW (beta = 1), dgrad overwrites dX (beta = 0).
//
// Tile: BM x 128 x 16, 256 threads, (BM/16) x 8 outputs per thread, double-buffered
// shared memory with register prefetch. wgrad uses split-K over the (batch, position)
// axis with a deterministic second-stage reduction.
#include "common.cuh"

using namespace b200;

namespace {

constexpr int BN = 128;
constexpr int BK = 16;
constexpr int NT = 256;
constexpr int PAD = 4;

enum Mode { FPROP = 0, DGRAD = 1, WGRAD = 2 };

struct ConvP {
    const float *x;     // input activations (fprop, wgrad)
    const float *w;     // weights (fprop, dgrad)
    const float *dy;    // output gradient (dgrad, wgrad)
    float *out;         // y / dx / gw (or split-K partials)
    const float *bias;  // fprop only, may be null
    int act, accumulate;
    int batch, cin, h, wd, cout, ho, wo, ks, stride, pad, groups, cin_g, cout_g;
    int M, N, K;        // GEMM extents per group
    int kchunk, splits; // split-K (wgrad)
    int kk, hw, howo;
    size_t out_split_stride;  // elements between split-K partial slabs
    FastDiv d_howo, d_wo, d_kk, d_ks, d_hw, d_w, d_stride;
};

template <int BM, int MODE>
__global__ void __launch_bounds__(NT)
igemm_kernel(const ConvP p) {
    constexpr int TM = BM / 16;
    constexpr int A_PER_T = BM * BK / NT;  // A elements each thread stages per k-tile
    constexpr int B_PER_T = BN * BK / NT;  // 8
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int g = blockIdx.z / p.splits;
    const int split = blockIdx.z - g * p.splits;
    const int k_begin = split * p.kchunk;
    const int k_end = min(p.K, k_begin + p.kchunk);

    // ---- per-thread, loop-invariant staging coordinates ----
    // A is always staged k-fast: k_l = t % 16, rows m_l = t / 16 + 16 * i.
    const int a_kl = t & 15;
    const int a_ml = t >> 4;
    // B is staged n-fast for fprop/dgrad (n_l = t % 128, k_l = t / 128 + 2 i) and k-fast
    // for wgrad (k_l = t % 16, n_l = t / 16 + 16 i).
    int bn_valid = 0, b_base = 0, b_c0 = 0, b_c1 = 0;  // fprop/dgrad: decoded column n
    int wn_off[B_PER_T], wn_khkw[B_PER_T];             // wgrad: decoded columns n_i
    if (MODE == FPROP) {
        int n = n0 + (t & 127);
        bn_valid = n < p.N;
        uint32_t b, pos, oh, ow;
        p.d_howo.divmod(bn_valid ? n : 0, b, pos);
        p.d_wo.divmod(pos, oh, ow);
        b_base = (b * p.cin + g * p.cin_g) * p.hw;
        b_c0 = (int)oh * p.stride - p.pad;
        b_c1 = (int)ow * p.stride - p.pad;
    } else if (MODE == DGRAD) {
        int n = n0 + (t & 127);
        bn_valid = n < p.N;
        uint32_t b, pos, ih, iw;
        p.d_hw.divmod(bn_valid ? n : 0, b, pos);
        p.d_w.divmod(pos, ih, iw);
        b_base = (b * p.cout + g * p.cout_g) * p.howo;
        b_c0 = (int)ih + p.pad;
        b_c1 = (int)iw + p.pad;
    } else {
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            int n = n0 + (t >> 4) + 16 * i;
            if (n < p.N) {
                uint32_t c, r, kh, kw;
                p.d_kk.divmod(n, c, r);
                p.d_ks.divmod(r, kh, kw);
                wn_off[i] = c * p.hw;
                wn_khkw[i] = (int)(kh << 16 | kw);
            } else {
                wn_off[i] = -1;
                wn_khkw[i] = 0;
            }
        }
    }

    float a_reg[A_PER_T], b_reg[B_PER_T];

    auto fetch = [&](int k0) {
        // ---------------- A ----------------
        {
            const int k = k0 + a_kl;
            const bool kv = k < k_end;
            if (MODE == FPROP) {
#pragma unroll
                for (int i = 0; i < A_PER_T; ++i) {
                    int m = m0 + a_ml + 16 * i;
                    a_reg[i] = (kv && m < p.M)
                                   ? __ldg(p.w + ((size_t)(g * p.cout_g + m)) * p.K + k)
                                   : 0.f;
                }
            } else if (MODE == DGRAD) {
                uint32_t co, r;
                p.d_kk.divmod(kv ? k : 0, co, r);
                const size_t wbase = ((size_t)(g * p.cout_g + co) * p.cin_g) * p.kk + r;
#pragma unroll
                for (int i = 0; i < A_PER_T; ++i) {
                    int m = m0 + a_ml + 16 * i;
                    a_reg[i] = (kv && m < p.M) ? __ldg(p.w + wbase + (size_t)m * p.kk) : 0.f;
                }
            } else {
                uint32_t b, pos;
                p.d_howo.divmod(kv ? k : 0, b, pos);
                const size_t dbase = ((size_t)b * p.cout + g * p.cout_g) * p.howo + pos;
#pragma unroll
                for (int i = 0; i < A_PER_T; ++i) {
                    int m = m0 + a_ml + 16 * i;
                    a_reg[i] = (kv && m < p.M) ? __ldg(p.dy + dbase + (size_t)m * p.howo) : 0.f;
                }
            }
        }
        // ---------------- B ----------------
        if (MODE == FPROP) {
#pragma unroll
            for (int i = 0; i < B_PER_T; ++i) {
                const int k = k0 + (t >> 7) + 2 * i;
                float v = 0.f;
                if (bn_valid && k < k_end) {
                    uint32_t c, r, kh, kw;
                    p.d_kk.divmod(k, c, r);
                    p.d_ks.divmod(r, kh, kw);
                    int ih = b_c0 + (int)kh, iw = b_c1 + (int)kw;
                    if ((unsigned)ih < (unsigned)p.h && (unsigned)iw < (unsigned)p.wd)
                        v = __ldg(p.x + b_base + ((int)c * p.h + ih) * p.wd + iw);
                }
                b_reg[i] = v;
            }
        } else if (MODE == DGRAD) {
#pragma unroll
            for (int i = 0; i < B_PER_T; ++i) {
                const int k = k0 + (t >> 7) + 2 * i;
                float v = 0.f;
                if (bn_valid && k < k_end) {
                    uint32_t co, r, kh, kw;
                    p.d_kk.divmod(k, co, r);
                    p.d_ks.divmod(r, kh, kw);
                    int th = b_c0 - (int)kh, tw = b_c1 - (int)kw;
                    if (th >= 0 && tw >= 0) {
                        uint32_t oh, rh, ow, rw;
                        p.d_stride.divmod(th, oh, rh);
                        p.d_stride.divmod(tw, ow, rw);
                        if (rh == 0 && rw == 0 && (int)oh < p.ho && (int)ow < p.wo)
                            v = __ldg(p.dy + b_base + ((int)co * p.ho + (int)oh) * p.wo + (int)ow);
                    }
                }
                b_reg[i] = v;
            }
        } else {
            const int k = k0 + (t & 15);
            const bool kv = k < k_end;
            uint32_t b, pos, oh, ow;
            p.d_howo.divmod(kv ? k : 0, b, pos);
            p.d_wo.divmod(pos, oh, ow);
            const int xb = (b * p.cin + g * p.cin_g) * p.hw;
            const int ih0 = (int)oh * p.stride - p.pad, iw0 = (int)ow * p.stride - p.pad;
#pragma unroll
            for (int i = 0; i < B_PER_T; ++i) {
                float v = 0.f;
                if (kv && wn_off[i] >= 0) {
                    int ih = ih0 + (wn_khkw[i] >> 16), iw = iw0 + (wn_khkw[i] & 0xffff);
                    if ((unsigned)ih < (unsigned)p.h && (unsigned)iw < (unsigned)p.wd)
                        v = __ldg(p.x + xb + wn_off[i] + ih * p.wd + iw);
                }
                b_reg[i] = v;
            }
        }
    };

    auto stage = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER_T; ++i) As[buf][a_kl][a_ml + 16 * i] = a_reg[i];
        if (MODE == WGRAD) {
#pragma unroll
            for (int i = 0; i < B_PER_T; ++i) Bs[buf][t & 15][(t >> 4) + 16 * i] = b_reg[i];
        } else {
#pragma unroll
            for (int i = 0; i < B_PER_T; ++i) Bs[buf][(t >> 7) + 2 * i][t & 127] = b_reg[i];
        }
    };

    float acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    int buf = 0;
    if (k_begin < k_end) {
        fetch(k_begin);
        stage(0);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = (k0 + BK) < k_end;
        if (more) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[8];
            if constexpr (TM == 8) {
                float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 8]);
                float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 8 + 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            } else if constexpr (TM == 4) {
                float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][ty * TM + i];
            }
            float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][64 + tx * 4]);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) stage(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ---------------- epilogue ----------------
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        if (MODE == WGRAD) {
            float *dst = p.out + (size_t)split * p.out_split_stride +
                         ((size_t)(g * p.cout_g + m)) * p.N;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (n < p.N) {
                    if (p.splits == 1) dst[n] += acc[i][j];
                    else dst[n] = acc[i][j];
                }
            }
        } else {
            const int plane = (MODE == FPROP) ? p.howo : p.hw;
            const int chans = (MODE == FPROP) ? p.cout : p.cin;
            const int ch = (MODE == FPROP) ? g * p.cout_g + m : g * p.cin_g + m;
            const float bv = (MODE == FPROP && p.bias) ? __ldg(p.bias + ch) : 0.f;
            if (p.splits > 1) {
                // split-K: raw partial sums into slab `split`; splitk_finish_kernel folds the slabs
                // in split order and applies bias / activation / accumulation
                float *slab = p.out + (size_t)split * p.out_split_stride;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                    if (n >= p.N) continue;
                    uint32_t bj, pj;
                    if (MODE == FPROP) p.d_howo.divmod(n, bj, pj);
                    else p.d_hw.divmod(n, bj, pj);
                    slab[((size_t)bj * chans + ch) * plane + pj] = acc[i][j];
                }
                continue;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int n = n0 + half * 64 + tx * 4;
                if (n >= p.N) continue;
                uint32_t b, pos;
                if (MODE == FPROP) p.d_howo.divmod(n, b, pos);
                else p.d_hw.divmod(n, b, pos);
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = acc[i][half * 4 + j];
                    if (MODE == FPROP) v[j] = act_fwd(v[j] + bv, p.act, 0.f);
                }
                float *dst = p.out + ((size_t)b * chans + ch) * plane + pos;
                if ((plane & 3) == 0 && n + 3 < p.N) {  // 4 columns in one image, 16B aligned
                    float4 o = make_float4(v[0], v[1], v[2], v[3]);
                    if (MODE == DGRAD && p.accumulate) {
                        float4 old = *reinterpret_cast<float4 *>(dst);
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    *reinterpret_cast<float4 *>(dst) = o;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (n + j >= p.N) break;
                        uint32_t bj, pj;
                        if (MODE == FPROP) p.d_howo.divmod(n + j, bj, pj);
                        else p.d_hw.divmod(n + j, bj, pj);
                        float *d = p.out + ((size_t)bj * chans + ch) * plane + pj;
                        if (MODE == DGRAD && p.accumulate) *d += v[j];
                        else *d = v[j];
                    }
                }
            }
        }
    }
}

// gw[i] += sum_s partial[s][i], in split order.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(float *__restrict__ gw, const float *__restrict__ partial, size_t n,
                     int splits) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += __ldg(partial + (size_t)k * n + i);
        gw[i] += s;
    }
}

// out[i] = act(sum_s partial[s][i] + bias[channel])  (fprop)  or  out[i] (+)= sum_s partial[s][i]
// (dgrad), slabs folded in split order.
__global__ void __launch_bounds__(256)
splitk_finish_kernel(float *__restrict__ out, const float *__restrict__ partial, size_t n, int splits,
                     const float *__restrict__ bias, int act, int accumulate, FastDiv d_plane,
                     FastDiv d_chans) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += __ldg(partial + (size_t)k * n + i);
        if (bias) {
            uint32_t q, ch;
            d_chans.divmod(d_plane.div((uint32_t)i), q, ch);
            s += __ldg(bias + ch);
        }
        s = act_fwd(s, act, 0.f);
        out[i] = accumulate ? out[i] + s : s;
    }
}

// Split-K for fprop / dgrad launches whose tile grid would leave most SMs idle (fully-connected
// layers: a few tiles with a K of thousands): enough splits for about two CTAs per SM, at least
// 128 reduction steps each.
int fwd_splits(int M, int N, int K, int groups) {
    const int bm = M > 64 ? 128 : (M > 32 ? 64 : 32);
    const long long tiles = (long long)ceil_div(M, bm) * ceil_div(N, BN) * groups;
    if (tiles * 2 > sm_count() || K < 512) return 1;
    long long want = (2LL * sm_count() + tiles - 1) / tiles;
    if (want > K / 128) want = K / 128;
    if (want > 32) want = 32;
    return want < 1 ? 1 : (int)want;
}

template <int MODE>
int launch(ConvP &p, cudaStream_t st);

// Runs an fprop / dgrad launch, through split-K partial slabs in `workspace` when that pays.
template <int MODE>
int launch_fwd_like(ConvP &p, float *out, size_t out_elems, int plane, int chans, void *workspace,
                    size_t workspace_bytes, cudaStream_t st) {
    const int s = out_elems < (1ull << 31) ? fwd_splits(p.M, p.N, p.K, p.groups) : 1;
    if (s <= 1 || workspace == nullptr || workspace_bytes < (size_t)s * out_elems * sizeof(float)) {
        p.out = out;
        return launch<MODE>(p, st);
    }
    const float *bias = p.bias;
    const int act = p.act, accumulate = p.accumulate;
    p.out = reinterpret_cast<float *>(workspace);
    p.splits = s;
    p.out_split_stride = out_elems;
    p.kchunk = ceil_div(ceil_div(p.K, s), BK) * BK;
    int err = launch<MODE>(p, st);
    if (err) return err;
    splitk_finish_kernel<<<stream_grid(out_elems, 256), 256, 0, st>>>(
        out, reinterpret_cast<const float *>(workspace), out_elems, s, bias, act, accumulate,
        FastDiv((uint32_t)plane), FastDiv((uint32_t)chans));
    return launched();
}

void fill_geom(ConvP &p, const bcnn_b200_conv_desc *d) {
    p.batch = d->batch; p.cin = d->cin; p.h = d->h; p.wd = d->w;
    p.cout = d->cout; p.ho = d->ho; p.wo = d->wo;
    p.ks = d->ksize; p.stride = d->stride; p.pad = d->pad; p.groups = d->groups;
    p.cin_g = d->cin / d->groups; p.cout_g = d->cout / d->groups;
    p.kk = d->ksize * d->ksize; p.hw = d->h * d->w; p.howo = d->ho * d->wo;
    p.d_howo = FastDiv(p.howo); p.d_wo = FastDiv(d->wo); p.d_kk = FastDiv(p.kk);
    p.d_ks = FastDiv(d->ksize); p.d_hw = FastDiv(p.hw); p.d_w = FastDiv(d->w);
    p.d_stride = FastDiv(d->stride);
    p.splits = 1; p.kchunk = 0; p.out_split_stride = 0;
    p.act = 0; p.accumulate = 0; p.bias = nullptr;
    p.x = p.w = p.dy = nullptr; p.out = nullptr;
}

template <int MODE>
int launch(ConvP &p, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
    if (p.kchunk == 0) p.kchunk = p.K;
    dim3 grid(ceil_div(p.N, BN), 1, p.groups * p.splits);
    if (p.M > 64) {
        grid.y = ceil_div(p.M, 128);
        igemm_kernel<128, MODE><<<grid, NT, 0, st>>>(p);
    } else if (p.M > 32) {
        grid.y = 1;
        igemm_kernel<64, MODE><<<grid, NT, 0, st>>>(p);
    } else {
        grid.y = 1;
        igemm_kernel<32, MODE><<<grid, NT, 0, st>>>(p);
    }
    return launched();
}

int wgrad_splits(const bcnn_b200_conv_desc *d) {
    int cout_g = d->cout / d->groups;
    int ncol = (d->cin / d->groups) * d->ksize * d->ksize;
    long long K = (long long)d->batch * d->ho * d->wo;
    int bm = cout_g > 64 ? 128 : (cout_g > 32 ? 64 : 32);
    long long tiles = (long long)ceil_div(cout_g, bm) * ceil_div(ncol, BN) * d->groups;
    long long want = (2LL * sm_count() + tiles - 1) / tiles;
    long long max_by_k = K / 256;  // keep at least 256 reduction steps per split
    if (want > max_by_k) want = max_by_k;
    if (want > 64) want = 64;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace

// ---------------------------------------------------------------------------
// Entry points used by conv.cu (the dispatcher behind the C ABI).
// ---------------------------------------------------------------------------
namespace b200 {

size_t conv_simt_workspace_bytes(const bcnn_b200_conv_desc *d) {
    size_t need = 0;
    int s = wgrad_splits(d);
    if (s > 1) need = (size_t)s * d->cout * (d->cin / d->groups) * d->ksize * d->ksize * sizeof(float);
    const int kk = d->ksize * d->ksize, cin_g = d->cin / d->groups, cout_g = d->cout / d->groups;
    s = fwd_splits(cout_g, d->batch * d->ho * d->wo, cin_g * kk, d->groups);
    size_t b = s > 1 ? (size_t)s * d->batch * d->cout * d->ho * d->wo * sizeof(float) : 0;
    if (b > need) need = b;
    s = fwd_splits(cin_g, d->batch * d->h * d->w, cout_g * kk, d->groups);
    b = s > 1 ? (size_t)s * d->batch * d->cin * d->h * d->w * sizeof(float) : 0;
    return b > need ? b : need;
}

int conv_simt_forward(const bcnn_b200_conv_desc *d, const float *x, const float *w,
                      const float *bias, int act, float *y, void *workspace, size_t workspace_bytes,
                      cudaStream_t st) {
    ConvP p;
    fill_geom(p, d);
    p.x = x; p.w = w; p.bias = bias; p.act = act;
    p.M = p.cout_g; p.N = d->batch * p.howo; p.K = p.cin_g * p.kk;
    return launch_fwd_like<FPROP>(p, y, (size_t)d->batch * d->cout * p.howo, p.howo, d->cout, workspace,
                                  workspace_bytes, st);
}

int conv_simt_backward_data(const bcnn_b200_conv_desc *d, const float *w, const float *dy,
                            float *dx, int accumulate, void *workspace, size_t workspace_bytes,
                            cudaStream_t st) {
    ConvP p;
    fill_geom(p, d);
    p.w = w; p.dy = dy; p.accumulate = accumulate;
    p.M = p.cin_g; p.N = d->batch * p.hw; p.K = p.cout_g * p.kk;
    return launch_fwd_like<DGRAD>(p, dx, (size_t)d->batch * d->cin * p.hw, p.hw, d->cin, workspace,
                                  workspace_bytes, st);
}

int conv_simt_backward_weights(const bcnn_b200_conv_desc *d, const float *x, const float *dy,
                               float *gw, void *workspace, size_t workspace_bytes,
                               cudaStream_t st) {
    ConvP p;
    fill_geom(p, d);
    p.x = x; p.dy = dy;
    p.M = p.cout_g; p.N = p.cin_g * p.kk; p.K = d->batch * p.howo;
    size_t wsize = (size_t)d->cout * p.N;
    int s = wgrad_splits(d);
    if (s > 1 && (workspace == nullptr || workspace_bytes < (size_t)s * wsize * sizeof(float)))
        s = 1;  // not enough scratch: fall back to a single pass over K
    p.splits = s;
    if (s == 1) {
        p.out = gw;
        p.kchunk = p.K;
        return launch<WGRAD>(p, st);
    }
    p.out = reinterpret_cast<float *>(workspace);
    p.out_split_stride = wsize;
    p.kchunk = ceil_div(ceil_div(p.K, s), BK) * BK;
    int err = launch<WGRAD>(p, st);
    if (err) return err;
    splitk_reduce_kernel<<<stream_grid(wsize, 256), 256, 0, st>>>(
        gw, reinterpret_cast<const float *>(workspace), wsize, s);
    return launched();
}

}  // namespace b200
