// conv_tc.cu -- BF16 implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   fprop : Y[(b,oh,ow), co]  = sum_{tap,(ci)} X[b,ci,oh*s-p+kh,ow*s-p+kw] * W[co,ci,kh,kw]
//   dgrad : dX[(b,ih,iw), ci] = sum_{tap,(co)} dY[b,co,(ih+p-kh)/s,(iw+p-kw)/s] * W[co,ci,kh,kw]
//
// GEMM view per CTA: D[128 positions x N channels] += A[128 x 64] * B[N x 64]^T per k-block,
// K ordered tap-major / channel-minor so a k-block is 64 consecutive channels of one filter
// tap. UMMA shape M = 128 (cta_group::1), N = 16..256, K = 16 (kind::f16, BF16 inputs,
// FP32 accumulation in TMEM).
//
//   B (weights)     : repacked once per call into BF16 tiles that are byte-for-byte the
//                     128B-swizzled K-major shared-memory image; one elected thread stages a
//                     whole tile per k-block with a single bulk async copy (TMA engine,
//                     cp.async.bulk ... mbarrier::complete_tx).
//   A (activations) : gathered straight from the NCHW FP32 tensor by all 256 threads
//                     (coalesced along the position axis), converted to BF16 in registers
//                     and stored into the same swizzled K-major layout -- the im2col matrix
//                     never exists in memory, and no NHWC/BF16 shadow of the activations is
//                     kept.
//   pipeline        : 3-4 smem stages; tcgen05.mma is issued by one thread and tracked with
//                     tcgen05.commit -> mbarrier; the global loads of k-block k+1 are issued
//                     into registers before block k is stored / synchronised / multiplied.
//   epilogue        : tcgen05.ld (32x32b.x32) TMEM -> registers, fused bias + activation,
//                     coalesced NCHW FP32 stores (a warp's 32 lanes are 32 consecutive
//                     positions of one channel plane).
//
// wgrad lives in conv_tc_wgrad.cu.
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_impl.cuh"
#include "tc_ptx.cuh"

using namespace b200;
using namespace b200::tc;

namespace {

constexpr int TILE_M = 128;       // positions per CTA (UMMA M)
constexpr int BLOCK_K = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 256;
constexpr int A_STAGE_BYTES = TILE_M * BLOCK_K * 2;  // 16 KiB

enum TcMode { TC_FPROP = 0, TC_DGRAD = 1 };

struct TcParams {
    const float *src;       // fprop: x ; dgrad: dy
    float *dst;             // fprop: y ; dgrad: dx
    const float *bias;      // fprop only
    const __nv_bfloat16 *wpack;  // [n_tiles][k_blocks][n_tile rows][64] swizzled
    int act, accumulate;
    // geometry of the *gathered* tensor (src) and the *written* tensor (dst)
    int batch;
    int src_c, src_h, src_w;   // channels / extent of src
    int dst_c, dst_h, dst_w;   // channels / extent of dst (positions enumerate dst)
    int ks, stride, pad;
    int kc_blocks;             // ceil(src_c / 64): k-blocks per tap
    int k_blocks;              // ks*ks*kc_blocks
    int n_tile;                // channels of dst per CTA (multiple of 16, <= 256)
    int stages;
    int total_pos;             // batch * dst_h * dst_w
    FastDiv d_plane, d_w, d_ks, d_kc, d_stride;
};

// ---------------------------------------------------------------- weight repack
// wpack[tile][kb][row][64] with the 16-byte chunk index XOR-swizzled by (row & 7):
// exactly the shared-memory image the UMMA descriptor above expects, so a k-block tile is
// one contiguous bulk copy. rows >= dst_c and channels >= src_c are zero.
//   fprop: row = co, k-channel = ci   -> W[co][ci][tap]
//   dgrad: row = ci, k-channel = co   -> W[co][ci][tap]   (transposed roles)
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ wpack, int mode,
                    int cout, int cin, int kk, int n_tile, int n_tiles, int kc_blocks) {
    const int rows_total = n_tiles * n_tile;
    const int k_blocks = kk * kc_blocks;
    const size_t chunks = (size_t)rows_total * k_blocks * 8;  // 16-byte chunks
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks;
         i += (size_t)gridDim.x * blockDim.x) {
        int chunk = (int)(i & 7);
        size_t r = i >> 3;
        int row_in_tile = (int)(r % n_tile);
        size_t r2 = r / n_tile;
        int kb = (int)(r2 % k_blocks);
        int tile = (int)(r2 / k_blocks);
        int row = tile * n_tile + row_in_tile;
        int tap = kb / kc_blocks, cb = kb - tap * kc_blocks;
        const int row_c = (mode == TC_FPROP) ? cout : cin;  // extent of the row channel
        const int k_c = (mode == TC_FPROP) ? cin : cout;    // extent of the k channel
        uint32_t packed[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int kc = cb * 64 + chunk * 8 + j * 2 + e;
                float val = 0.f;
                if (row < row_c && kc < k_c) {
                    int co = (mode == TC_FPROP) ? row : kc;
                    int ci = (mode == TC_FPROP) ? kc : row;
                    val = __ldg(w + ((size_t)co * cin + ci) * kk + tap);
                }
                v[e] = val;
            }
            packed[j] = pack_bf16x2(v[0], v[1]);
        }
        size_t tile_base = ((size_t)tile * k_blocks + kb) * (size_t)n_tile * 64;  // elements
        size_t off = tile_base + (size_t)row_in_tile * 64 + (size_t)((chunk ^ (row_in_tile & 7)) * 8);
        *reinterpret_cast<uint4 *>(wpack + off) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

// ---------------------------------------------------------------- main kernel
template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][A 16 KiB][B n_tile*128 B] (1024-aligned), then barriers + tmem slot
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_stage_bytes = p.n_tile * 128;
    const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    // bars[0..S): weights landed (tx) ; bars[S..2S): MMAs of the stage retired ; bars[2S]: accumulator ready
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 1);

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int S = p.stages;
    const int n_tile = p.n_tile;
    const uint32_t tmem_cols = n_tile <= 32 ? 32 : (n_tile <= 64 ? 64 : (n_tile <= 128 ? 128 : 256));

    if (t == 0) {
        for (int i = 0; i < 2 * S + 1; ++i) mbar_init(smem_u32(bars + i), 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- loop-invariant gather coordinates: this thread owns position row (t & 127) and the
    // 32 channels [32*(t>>7), +32) of every k-block
    const int row = t & 127;
    const int chalf = t >> 7;
    const int pos = blockIdx.x * TILE_M + row;
    const bool pos_valid = pos < p.total_pos;
    uint32_t b_img, rem, ph, pw;
    p.d_plane.divmod(pos_valid ? pos : 0, b_img, rem);
    p.d_w.divmod(rem, ph, pw);
    const int src_plane = p.src_h * p.src_w;
    const float *src_img = p.src + (size_t)b_img * p.src_c * src_plane;
    const int tile_n = blockIdx.y;
    const __nv_bfloat16 *wtile = p.wpack + (size_t)tile_n * p.k_blocks * n_tile * 64;
    const uint32_t idesc = make_idesc(TILE_M, n_tile);
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);

    float v_raw[32];  // next k-block: global loads in flight (FP32)
    uint4 pk[4];      // current k-block: packed BF16, ready to store

    // gather A for k-block kb: 32 channels x 1 position per thread
    auto issue_loads = [&](int kb) {
        uint32_t tap, cb, kh, kw;
        p.d_kc.divmod(kb, tap, cb);
        p.d_ks.divmod(tap, kh, kw);
        int sh, sw;
        bool valid = pos_valid;
        if (MODE == TC_FPROP) {
            sh = (int)ph * p.stride - p.pad + (int)kh;
            sw = (int)pw * p.stride - p.pad + (int)kw;
        } else {
            int th = (int)ph + p.pad - (int)kh, tw = (int)pw + p.pad - (int)kw;
            valid = valid && th >= 0 && tw >= 0;
            uint32_t qh, rh, qw, rw;
            p.d_stride.divmod(th < 0 ? 0 : th, qh, rh);
            p.d_stride.divmod(tw < 0 ? 0 : tw, qw, rw);
            valid = valid && rh == 0 && rw == 0;
            sh = (int)qh;
            sw = (int)qw;
        }
        valid = valid && (unsigned)sh < (unsigned)p.src_h && (unsigned)sw < (unsigned)p.src_w;
        const int c0 = (int)cb * 64 + chalf * 32;
        const float *gp = src_img + (size_t)c0 * src_plane + (valid ? sh * p.src_w + sw : 0);
        const int c_left = p.src_c - c0;  // channels still inside the tensor
#pragma unroll
        for (int j = 0; j < 32; ++j)
            v_raw[j] = (valid && j < c_left) ? __ldg(gp + (size_t)j * src_plane) : 0.f;
    };
    auto convert = [&]() {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pk[q].x = pack_bf16x2(v_raw[q * 8 + 0], v_raw[q * 8 + 1]);
            pk[q].y = pack_bf16x2(v_raw[q * 8 + 2], v_raw[q * 8 + 3]);
            pk[q].z = pack_bf16x2(v_raw[q * 8 + 4], v_raw[q * 8 + 5]);
            pk[q].w = pack_bf16x2(v_raw[q * 8 + 6], v_raw[q * 8 + 7]);
        }
    };

    issue_loads(0);
    convert();
    for (int kb = 0; kb < p.k_blocks; ++kb) {
        const int s = kb % S;
        const int round = kb / S;
        uint8_t *a_stage = smem + (size_t)s * stage_bytes;
        uint8_t *b_stage = a_stage + A_STAGE_BYTES;
        // the next block's loads go out first: their latency overlaps the stores, the CTA
        // barrier and the MMA issue of this block
        const bool more = kb + 1 < p.k_blocks;
        if (more) issue_loads(kb + 1);
        // stage s is free once the MMAs issued S k-blocks ago have retired
        if (kb >= S) mbar_wait(smem_u32(bars + S + s), (round - 1) & 1);

        if (t == 0) {  // weights: one bulk async copy of the pre-swizzled tile
            mbar_expect_tx(smem_u32(bars + s), (uint32_t)b_stage_bytes);
            bulk_copy_g2s(smem_u32(b_stage), wtile + (size_t)kb * n_tile * 64,
                          (uint32_t)b_stage_bytes, smem_u32(bars + s));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t chunk = (uint32_t)(chalf * 4 + q);
            *reinterpret_cast<uint4 *>(a_stage + row_off + ((chunk ^ (uint32_t)(row & 7)) << 4)) = pk[q];
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();

        if (t == 0) {
            mbar_wait(smem_u32(bars + s), round & 1);  // weight tile landed
            tc_fence_after();
            const uint64_t da = make_desc_sw128(smem_u32(a_stage));
            const uint64_t db = make_desc_sw128(smem_u32(b_stage));
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(smem_u32(bars + S + s));       // frees this stage when the MMAs retire
            if (kb == p.k_blocks - 1) umma_commit(smem_u32(bars + 2 * S));  // accumulator ready
        }
        if (more) convert();
    }

    // ---------------- epilogue: TMEM -> registers -> bias/activation -> NCHW global
    mbar_wait(smem_u32(bars + 2 * S), 0);
    tc_fence_after();
    {
        const int quarter = warp & 3;          // TMEM lane quarter this warp may read
        const int col_half = warp >> 2;        // warps 0-3: low half of the columns, 4-7: high
        const int m = quarter * 32 + lane;     // position row of this thread
        const int epos = blockIdx.x * TILE_M + m;
        const bool e_valid = epos < p.total_pos;
        uint32_t eb, erem;
        p.d_plane.divmod(e_valid ? epos : 0, eb, erem);
        const int dst_plane = p.dst_h * p.dst_w;
        float *dst_img = p.dst + (size_t)eb * p.dst_c * dst_plane + erem;
        const int chunks32 = (n_tile + 31) / 32;
        for (int ck = col_half; ck < chunks32; ck += 2) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ck * 32), r);
            if (!e_valid) continue;
            float old[32];
            if (MODE == TC_DGRAD && p.accumulate) {  // issue every RMW load before the stores
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int ch = tile_n * n_tile + ck * 32 + j;
                    old[j] = (ck * 32 + j < n_tile && ch < p.dst_c)
                                 ? __ldcs(dst_img + (size_t)ch * dst_plane) : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int ch = tile_n * n_tile + ck * 32 + j;
                if (ck * 32 + j < n_tile && ch < p.dst_c) {
                    float val = __uint_as_float(r[j]);
                    float *d = dst_img + (size_t)ch * dst_plane;
                    if (MODE == TC_FPROP) {
                        if (p.bias) val += __ldg(p.bias + ch);
                        val = act_fwd(val, p.act, 0.f);
                        *d = val;
                    } else {
                        *d = p.accumulate ? (old[j] + val) : val;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------- host side
struct TcPlan {
    int n_tile, n_tiles, kc_blocks, k_blocks, stages;
    size_t wpack_bytes, smem_bytes;
};

TcPlan make_plan(int dst_c, int src_c, int ks) {
    TcPlan pl;
    int n = dst_c;
    if (n > 256) {
        // balanced tiles <= 256, multiple of 16
        int tiles = ceil_div(n, 256);
        n = ceil_div(ceil_div(n, tiles), 16) * 16;
    } else {
        n = ceil_div(n, 16) * 16;
    }
    pl.n_tile = n;
    pl.n_tiles = ceil_div(dst_c, n);
    pl.kc_blocks = ceil_div(src_c, 64);
    pl.k_blocks = ks * ks * pl.kc_blocks;
    pl.stages = n > 128 ? 4 : 3;
    pl.wpack_bytes = (size_t)pl.n_tiles * pl.k_blocks * n * 64 * sizeof(__nv_bfloat16);
    pl.smem_bytes = (size_t)pl.stages * (A_STAGE_BYTES + n * 128) + 1024 /*align*/ + 256 /*barriers*/;
    return pl;
}

bool shape_ok(const bcnn_b200_conv_desc *d, int src_c) {
    if (d->groups != 1) return false;
    if (src_c < 32) return false;               // K too thin for 64-wide k-blocks (first layers)
    if ((long long)d->batch * d->h * d->w >= (1LL << 31)) return false;
    return true;
}

template <int MODE>
int launch_tc(const bcnn_b200_conv_desc *d, const float *src, const float *w, const float *bias,
              int act, float *dst, int accumulate, void *workspace, size_t workspace_bytes,
              cudaStream_t st) {
    const int src_c = MODE == TC_FPROP ? d->cin : d->cout;
    const int dst_c = MODE == TC_FPROP ? d->cout : d->cin;
    TcPlan pl = make_plan(dst_c, src_c, d->ksize);
    if (workspace == nullptr || workspace_bytes < pl.wpack_bytes) return (int)cudaErrorInvalidValue;
    __nv_bfloat16 *wpack = reinterpret_cast<__nv_bfloat16 *>(workspace);
    const int kk = d->ksize * d->ksize;
    size_t chunks = (size_t)pl.n_tiles * pl.n_tile * pl.k_blocks * 8;
    pack_weights_kernel<<<stream_grid(chunks, 256), 256, 0, st>>>(w, wpack, MODE, d->cout, d->cin,
                                                                kk, pl.n_tile, pl.n_tiles,
                                                                pl.kc_blocks);
    int err = launched();
    if (err) return err;

    TcParams p;
    p.src = src; p.dst = dst; p.bias = bias; p.wpack = wpack;
    p.act = act; p.accumulate = accumulate;
    p.batch = d->batch;
    if (MODE == TC_FPROP) {
        p.src_c = d->cin; p.src_h = d->h; p.src_w = d->w;
        p.dst_c = d->cout; p.dst_h = d->ho; p.dst_w = d->wo;
    } else {
        p.src_c = d->cout; p.src_h = d->ho; p.src_w = d->wo;
        p.dst_c = d->cin; p.dst_h = d->h; p.dst_w = d->w;
    }
    p.ks = d->ksize; p.stride = d->stride; p.pad = d->pad;
    p.kc_blocks = pl.kc_blocks; p.k_blocks = pl.k_blocks;
    p.n_tile = pl.n_tile; p.stages = pl.stages;
    p.total_pos = d->batch * p.dst_h * p.dst_w;
    p.d_plane = FastDiv(p.dst_h * p.dst_w);
    p.d_w = FastDiv(p.dst_w);
    p.d_ks = FastDiv(d->ksize);
    p.d_kc = FastDiv(pl.kc_blocks);
    p.d_stride = FastDiv(d->stride);

    static bool attr_set[2] = {false, false};
    if (!attr_set[MODE]) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<MODE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set[MODE] = true;
    }
    dim3 grid(ceil_div(p.total_pos, TILE_M), pl.n_tiles);
    conv_tc_kernel<MODE><<<grid, NTHREADS, pl.smem_bytes, st>>>(p);
    return launched();
}

}  // namespace

namespace b200 {

bool conv_tc_supports_fprop(const bcnn_b200_conv_desc *d) { return shape_ok(d, d->cin); }
bool conv_tc_supports_dgrad(const bcnn_b200_conv_desc *d) { return shape_ok(d, d->cout); }

size_t conv_tc_workspace_bytes(const bcnn_b200_conv_desc *d) {
    size_t need = conv_tc_wgrad_workspace_bytes(d);
    if (conv_tc_supports_fprop(d)) {
        size_t a = make_plan(d->cout, d->cin, d->ksize).wpack_bytes;
        if (a > need) need = a;
    }
    if (conv_tc_supports_dgrad(d)) {
        size_t b = make_plan(d->cin, d->cout, d->ksize).wpack_bytes;
        if (b > need) need = b;
    }
    return need;
}

int conv_tc_forward(const bcnn_b200_conv_desc *d, const float *x, const float *w,
                    const float *bias, int act, float *y, void *workspace,
                    size_t workspace_bytes, cudaStream_t st) {
    return launch_tc<TC_FPROP>(d, x, w, bias, act, y, 0, workspace, workspace_bytes, st);
}

int conv_tc_backward_data(const bcnn_b200_conv_desc *d, const float *w, const float *dy, float *dx,
                          int accumulate, void *workspace, size_t workspace_bytes,
                          cudaStream_t st) {
    return launch_tc<TC_DGRAD>(d, dy, w, nullptr, 0, dx, accumulate, workspace, workspace_bytes, st);
}

}  // namespace b200
