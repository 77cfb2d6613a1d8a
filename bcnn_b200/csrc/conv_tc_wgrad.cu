// conv_tc_wgrad.cu -- weight gradient of the convolution on tcgen05 tensor cores.
//
//   dW[co, ci, tap] += sum_{(b,oh,ow)} dY[b,co,oh,ow] * X[b,ci,oh*s-p+kh,ow*s-p+kw]
//
// One CTA owns one filter tap, 128 output channels (UMMA M) and up to 256 input channels
// (UMMA N) and reduces over a slice of the (batch, position) axis (UMMA K, split-K across
// CTAs): D[128 co x N ci] += A[128 x 64 pos] * B[N x 64 pos]^T per k-block.
// In NCHW a channel's positions are contiguous, so BOTH operands are K-major rows that
// are gathered straight from the FP32 tensors (dY rows are plain contiguous runs; X rows
// are the tap-shifted, zero-padded view), converted to BF16 in registers and written in
// the 128-byte-swizzled layout the UMMA descriptors expect. The global loads of k-block
// i+1 are issued into registers before k-block i is stored, synchronised and multiplied,
// so their latency overlaps that work.
// Accumulation is FP32 in TMEM; split-K partials go to the workspace and are folded into
// dW (+=, the reference's beta = 1) in split order by a second kernel: deterministic.
#include "common.cuh"
#include "conv_impl.cuh"
#include "tc_ptx.cuh"

using namespace b200;
using namespace b200::tc;

namespace {

constexpr int TILE_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 256;
constexpr int A_BYTES = TILE_M * BLOCK_K * 2;  // 16 KiB
constexpr int STAGES = 3;
constexpr int A_CHUNKS = TILE_M * 8 / NTHREADS;  // 16-byte chunks per thread for A = 4
constexpr int B_CHUNKS_MAX = 256 * 8 / NTHREADS; // up to 8 for N = 256

struct WgParams {
    const float *x, *dy;
    float *out;  // split-K partial slabs [split][cout][cin][kk], or gw itself when splits == 1
    int batch, cin, h, w, cout, ho, wo, ks, stride, pad;
    int kk, hw, howo, total_pos;
    int n_tile;        // input channels per CTA (multiple of 16, <= 256)
    int ci_tiles;
    int kb_total, kb_per_split, splits;
    size_t split_stride;
    FastDiv d_howo, d_wo, d_ks;
};

// Gather 8 consecutive reduction positions g0..g0+7 of dY channel `co` (zero beyond the
// end of the batch or for padded rows).
__device__ __forceinline__ void load_dy_chunk(const WgParams &p, int co, int g0, bool row_valid,
                                              float (&v)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (!row_valid || g0 >= p.total_pos) return;
    uint32_t b, pos;
    p.d_howo.divmod(g0, b, pos);
    const float *base = p.dy + ((size_t)b * p.cout + co) * p.howo + pos;
    if ((int)pos + 8 <= p.howo) {  // same image: one contiguous run
        if (((reinterpret_cast<uintptr_t>(base)) & 15) == 0) {
            float4 lo = __ldg(reinterpret_cast<const float4 *>(base));
            float4 hi = __ldg(reinterpret_cast<const float4 *>(base) + 1);
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
            v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __ldg(base + e);
        }
    } else {  // the run crosses into the next image
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int g = g0 + e;
            if (g < p.total_pos) {
                uint32_t be, pe;
                p.d_howo.divmod(g, be, pe);
                v[e] = __ldg(p.dy + ((size_t)be * p.cout + co) * p.howo + pe);
            }
        }
    }
}

// Gather the tap-shifted X values of channel `ci` for reduction positions g0..g0+7.
__device__ __forceinline__ void load_x_chunk(const WgParams &p, int ci, int g0, bool row_valid,
                                             int dh, int dw, float (&v)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (!row_valid || g0 >= p.total_pos) return;
    uint32_t b, pos, oh, ow;
    p.d_howo.divmod(g0, b, pos);
    p.d_wo.divmod(pos, oh, ow);
    int ib = (int)b, ioh = (int)oh, iow = (int)ow;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (ib < p.batch) {
            int ih = ioh * p.stride + dh, iw = iow * p.stride + dw;
            if ((unsigned)ih < (unsigned)p.h && (unsigned)iw < (unsigned)p.w)
                v[e] = __ldg(p.x + ((size_t)ib * p.cin + ci) * p.hw + ih * p.w + iw);
        }
        if (++iow == p.wo) {
            iow = 0;
            if (++ioh == p.ho) { ioh = 0; ++ib; }
        }
    }
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 r;
    r.x = pack_bf16x2(v[0], v[1]);
    r.y = pack_bf16x2(v[2], v[3]);
    r.z = pack_bf16x2(v[4], v[5]);
    r.w = pack_bf16x2(v[6], v[7]);
    return r;
}

template <int B_CHUNKS>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tc_kernel(const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int n_tile = p.n_tile;
    const int b_bytes = n_tile * 128;
    const int stage_bytes = A_BYTES + b_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * stage_bytes);
    // bars[0..S): MMAs of the stage retired ; bars[S]: accumulator ready
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + STAGES + 1);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t tmem_cols = n_tile <= 32 ? 32 : (n_tile <= 64 ? 64 : (n_tile <= 128 ? 128 : 256));
    if (t == 0) {
        for (int i = 0; i < STAGES + 1; ++i) mbar_init(smem_u32(bars + i), 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int split = blockIdx.x;
    const int tap = blockIdx.y % p.kk;
    const int ci0 = (blockIdx.y / p.kk) * n_tile;
    const int co0 = blockIdx.z * TILE_M;
    uint32_t kh, kw;
    p.d_ks.divmod(tap, kh, kw);
    const int dh = (int)kh - p.pad, dw = (int)kw - p.pad;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
    const uint32_t idesc = make_idesc(TILE_M, n_tile);

    // chunk q = t + 256*i : row = q >> 3, 16-byte chunk (8 positions) = q & 7
    const int kchunk = t & 7;
    const int row_base = t >> 3;  // + 32*i

    uint4 a_pk[A_CHUNKS], b_pk[B_CHUNKS];          // k-block being stored (BF16, packed)
    float a_raw[A_CHUNKS][8], b_raw[B_CHUNKS][8];  // next k-block, loads in flight (FP32)

    auto issue_loads = [&](int kb) {
        const int g0 = kb * BLOCK_K + kchunk * 8;
#pragma unroll
        for (int i = 0; i < A_CHUNKS; ++i) {
            const int r = row_base + 32 * i;
            load_dy_chunk(p, co0 + r, g0, co0 + r < p.cout, a_raw[i]);
        }
#pragma unroll
        for (int i = 0; i < B_CHUNKS; ++i) {
            const int r = row_base + 32 * i;
            load_x_chunk(p, ci0 + r, g0, r < n_tile && ci0 + r < p.cin, dh, dw, b_raw[i]);
        }
    };
    auto convert = [&]() {
#pragma unroll
        for (int i = 0; i < A_CHUNKS; ++i) a_pk[i] = pack8(a_raw[i]);
#pragma unroll
        for (int i = 0; i < B_CHUNKS; ++i) b_pk[i] = pack8(b_raw[i]);
    };

    if (kb_begin < kb_end) {
        issue_loads(kb_begin);
        convert();
    }
    int it = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int s = it % STAGES;
        const int round = it / STAGES;
        uint8_t *a_stage = smem + (size_t)s * stage_bytes;
        uint8_t *b_stage = a_stage + A_BYTES;
        // the next block's global loads go out first: they fly while this block is stored,
        // the CTA synchronises and the MMAs are issued
        const bool more = kb + 1 < kb_end;
        if (more) issue_loads(kb + 1);
        if (it >= STAGES) mbar_wait(smem_u32(bars + s), (round - 1) & 1);
#pragma unroll
        for (int i = 0; i < A_CHUNKS; ++i) {
            const int r = row_base + 32 * i;
            *reinterpret_cast<uint4 *>(a_stage + (r >> 3) * 1024 + (r & 7) * 128 +
                                       ((kchunk ^ (r & 7)) << 4)) = a_pk[i];
        }
#pragma unroll
        for (int i = 0; i < B_CHUNKS; ++i) {
            const int r = row_base + 32 * i;
            if (r < n_tile)
                *reinterpret_cast<uint4 *>(b_stage + (r >> 3) * 1024 + (r & 7) * 128 +
                                           ((kchunk ^ (r & 7)) << 4)) = b_pk[i];
        }
        fence_proxy_async();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            const uint64_t da = make_desc_sw128(smem_u32(a_stage));
            const uint64_t db = make_desc_sw128(smem_u32(b_stage));
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                          (it > 0 || k > 0) ? 1u : 0u);
            umma_commit(smem_u32(bars + s));
            if (kb == kb_end - 1) umma_commit(smem_u32(bars + STAGES));
        }
        if (more) convert();
    }

    // ---------------- epilogue: D[co, ci] -> out[split][co][ci][tap]
    float *out = p.out + (size_t)split * p.split_stride;
    const bool have_acc = kb_begin < kb_end;
    if (have_acc) {
        mbar_wait(smem_u32(bars + STAGES), 0);
        tc_fence_after();
    }
    {
        const int quarter = warp & 3, col_half = warp >> 2;
        const int co = co0 + quarter * 32 + lane;
        const int chunks32 = (n_tile + 31) / 32;
        for (int ck = col_half; ck < chunks32; ck += 2) {
            uint32_t r[32];
            if (have_acc) {
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ck * 32), r);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            if (co >= p.cout) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int ci = ci0 + ck * 32 + j;
                if (ck * 32 + j < n_tile && ci < p.cin) {
                    float *d = out + ((size_t)co * p.cin + ci) * p.kk + tap;
                    const float val = __uint_as_float(r[j]);
                    if (p.splits == 1) *d += val;
                    else *d = val;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(float *__restrict__ gw, const float *__restrict__ partial, size_t n,
                    int splits) {
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += __ldg(partial + (size_t)k * n + i);
        gw[i] += s;
    }
}

struct WgPlan {
    int n_tile, ci_tiles, co_tiles, kb_total, splits, kb_per_split;
    size_t smem_bytes, partial_bytes;
};

WgPlan make_wg_plan(const bcnn_b200_conv_desc *d) {
    WgPlan pl;
    int n = d->cin;
    if (n > 256) {
        int tiles = ceil_div(n, 256);
        n = ceil_div(ceil_div(n, tiles), 16) * 16;
    } else {
        n = ceil_div(n, 16) * 16;
    }
    pl.n_tile = n;
    pl.ci_tiles = ceil_div(d->cin, n);
    pl.co_tiles = ceil_div(d->cout, TILE_M);
    const int kk = d->ksize * d->ksize;
    const long long total_pos = (long long)d->batch * d->ho * d->wo;
    pl.kb_total = (int)((total_pos + BLOCK_K - 1) / BLOCK_K);
    const long long tiles = (long long)pl.ci_tiles * pl.co_tiles * kk;
    long long want = (2LL * sm_count() + tiles - 1) / tiles;   // ~2 CTAs' worth per SM
    long long max_by_k = pl.kb_total / 8;                      // >= 8 k-blocks per split
    if (want > max_by_k) want = max_by_k;
    if (want > 128) want = 128;
    if (want < 1) want = 1;
    pl.kb_per_split = ceil_div(pl.kb_total, (int)want);
    pl.splits = ceil_div(pl.kb_total, pl.kb_per_split);
    pl.smem_bytes = (size_t)STAGES * (A_BYTES + n * 128) + 1024 + 256;
    const size_t wsize = (size_t)d->cout * d->cin * kk;
    pl.partial_bytes = pl.splits > 1 ? (size_t)pl.splits * wsize * sizeof(float) : 0;
    return pl;
}

bool wg_shape_ok(const bcnn_b200_conv_desc *d) {
    if (d->groups != 1) return false;
    if (d->cin < 16 || d->cout < 32) return false;  // thin first layers stay on the SIMT kernel
    if ((long long)d->batch * d->ho * d->wo >= (1LL << 31) - 64) return false;
    if ((long long)d->batch * d->ho * d->wo < 512) return false;  // tiny reductions (fc-shaped)
    return true;
}

template <int B_CHUNKS>
int launch_wg(const WgParams &p, dim3 grid, size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<B_CHUNKS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    wgrad_tc_kernel<B_CHUNKS><<<grid, NTHREADS, smem, st>>>(p);
    return launched();
}

}  // namespace

namespace b200 {

bool conv_tc_supports_wgrad(const bcnn_b200_conv_desc *d) { return wg_shape_ok(d); }

size_t conv_tc_wgrad_workspace_bytes(const bcnn_b200_conv_desc *d) {
    return wg_shape_ok(d) ? make_wg_plan(d).partial_bytes : 0;
}

int conv_tc_backward_weights(const bcnn_b200_conv_desc *d, const float *x, const float *dy,
                             float *gw, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    WgPlan pl = make_wg_plan(d);
    if (pl.splits > 1 && (workspace == nullptr || workspace_bytes < pl.partial_bytes))
        return (int)cudaErrorInvalidValue;
    WgParams p;
    p.x = x; p.dy = dy;
    p.batch = d->batch; p.cin = d->cin; p.h = d->h; p.w = d->w;
    p.cout = d->cout; p.ho = d->ho; p.wo = d->wo;
    p.ks = d->ksize; p.stride = d->stride; p.pad = d->pad;
    p.kk = d->ksize * d->ksize; p.hw = d->h * d->w; p.howo = d->ho * d->wo;
    p.total_pos = d->batch * p.howo;
    p.n_tile = pl.n_tile; p.ci_tiles = pl.ci_tiles;
    p.kb_total = pl.kb_total; p.kb_per_split = pl.kb_per_split; p.splits = pl.splits;
    const size_t wsize = (size_t)d->cout * d->cin * p.kk;
    p.split_stride = pl.splits > 1 ? wsize : 0;
    p.out = pl.splits > 1 ? reinterpret_cast<float *>(workspace) : gw;
    p.d_howo = FastDiv(p.howo); p.d_wo = FastDiv(d->wo); p.d_ks = FastDiv(d->ksize);
    dim3 grid(pl.splits, pl.ci_tiles * p.kk, pl.co_tiles);
    int err;
    const int b_chunks = ceil_div(pl.n_tile * 8, NTHREADS);
    if (b_chunks <= 2) err = launch_wg<2>(p, grid, pl.smem_bytes, st);
    else if (b_chunks <= 4) err = launch_wg<4>(p, grid, pl.smem_bytes, st);
    else err = launch_wg<8>(p, grid, pl.smem_bytes, st);
    if (err) return err;
    if (pl.splits > 1) {
        wgrad_reduce_kernel<<<stream_grid(wsize, 256), 256, 0, st>>>(
            gw, reinterpret_cast<const float *>(workspace), wsize, pl.splits);
        return launched();
    }
    return 0;
}

}  // namespace b200
