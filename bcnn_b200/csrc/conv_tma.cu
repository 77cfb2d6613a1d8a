// conv_tma.cu -- TF32 implicit-GEMM convolution on tcgen05 tensor cores, operands staged by
// the TMA straight from the FP32 NCHW tensors (sm_100a).
//
// No im2col matrix, no layout change and no precision-conversion pass exists anywhere: a
// 4-D tensor map over the NCHW activation (dims w, h, c, n) lets one TMA box fetch
// "32 consecutive output columns x 32 channels" of one filter tap -- shifted by (kh - pad,
// kw - pad), zero-filled outside the image by the TMA's out-of-bounds rule -- into
// 128-byte-swizzled shared memory, which is exactly a tcgen05 *MN-major* TF32 operand atom
// (positions contiguous, channels strided). kind::tf32 consumes the FP32 bits as they are.
//
//   fprop / stride-1 dgrad  (conv_tma_fwd_kernel):
//       D[128 positions x N channels] += A[128 x 32ch] (MN-major, TMA) * B[N x 32ch]^T
//       (K-major packed weights, one bulk copy per k-block); K order: tap-major, channel-minor.
//       A tile = 4 atoms of 32 positions: 4 column chunks of one output row, 2 x 2, or 1 x 4 rows
//       (1x1 convolutions see the image plane as one long row, so tiles are dense).
//       dgrad with stride 1 is the same kernel on dY with flipped taps and pad' = k-1-pad.
//   wgrad  (conv_tma_wgrad_kernel):
//       D[128 co x N ci] += dY[128 co x 32 pos] * X_shift[N ci x 32 pos]^T, both operands
//       K-major (positions contiguous in NCHW) and both fetched by TMA; split-K over
//       (image, row, column chunk) across CTAs, deterministic second-stage reduction.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA
// issuer (one lane), warps 2-5 = epilogue (tcgen05.ld -> bias/activation -> coalesced NCHW
// stores). smem ring of 2-4 stages with full/empty mbarriers; two CTAs fit per SM so one
// CTA's epilogue overlaps the other's main loop.
//
// Shapes the TMA cannot address (row pitch not a multiple of 16 bytes, stride > 1, groups)
// stay on the register-gather tcgen05 kernels of conv_tc.cu / conv_tc_wgrad.cu.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_impl.cuh"
#include "tc_ptx.cuh"

using namespace b200;
using namespace b200::tc;

namespace {

constexpr int TILE_M = 128;
constexpr int BLOCK_K = 32;                 // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;                   // kind::tf32
constexpr int ATOM_BYTES = 32 * BLOCK_K * 4;   // 4 KiB: 32 positions x 32 channels
constexpr int A_STAGE_BYTES = TILE_M * BLOCK_K * 4;  // 16 KiB
constexpr int NTHREADS = 192;

// ------------------------------------------------------------------ PTX helpers (TF32 / TMA)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1,
                                            int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];"
        :: "r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}
// MN-major 32-bit operand: the only layout tcgen05 accepts is SWIZZLE_128B_BASE32B (layout type
// 1; cute::UMMA Layout_MN_SW128_32B_Atom, Swizzle<2,5,2>): 32 fp32 along MN are contiguous
// (128 B), 32-byte chunks are XOR-swizzled by (K-row & 3), 4 K-rows of 128 B form a 512 B atom.
// LBO = byte stride between MN atoms, SBO = byte stride between groups of 4 K-rows. The TMA
// produces it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
// D = F32, A = B = TF32; a_mn != 0 marks A as MN-major (bit 15); B is K-major.
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n, int a_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------ tensor maps (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// Map over an NCHW fp32 tensor seen as (w, h, c, n); box = 32 columns x 1 row x box_c channels.
// Out-of-bounds elements (negative or past-the-end coordinates) read as zero.
bool make_map(CUtensorMap *map, const float *base, int w, int h, int c, int n, int box_c,
              bool mn_major) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)c, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4, (cuuint64_t)w * h * c * 4};
    cuuint32_t box[4] = {32, 1, (cuuint32_t)box_c, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ------------------------------------------------------------------ weight repack (TF32 = fp32 bits)
// wpack[tile][kb][row][32 floats], 16-byte chunks XOR-swizzled by (row & 7): byte-for-byte the
// K-major SWIZZLE_128B shared-memory image, so a k-block tile is one bulk copy.
//   fprop: row = co, k = ci, tap as is            -> W[co][ci][tap]
//   dgrad: row = ci, k = co, tap flipped (kk-1-t) -> W[co][ci][kk-1-tap]
__global__ void __launch_bounds__(256)
pack_weights_tf32_kernel(const float *__restrict__ w, float *__restrict__ wpack, int dgrad, int cout,
                         int cin, int kk, int n_tile, int n_tiles, int kc_blocks) {
    const int k_blocks = kk * kc_blocks;
    const size_t chunks = (size_t)n_tiles * n_tile * k_blocks * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks;
         i += (size_t)gridDim.x * blockDim.x) {
        const int chunk = (int)(i & 7);
        size_t r = i >> 3;
        const int row_in_tile = (int)(r % n_tile);
        size_t r2 = r / n_tile;
        const int kb = (int)(r2 % k_blocks);
        const int tile = (int)(r2 / k_blocks);
        const int row = tile * n_tile + row_in_tile;
        const int tap = kb / kc_blocks, cb = kb - tap * kc_blocks;
        const int row_c = dgrad ? cin : cout;
        const int k_c = dgrad ? cout : cin;
        const int wtap = dgrad ? kk - 1 - tap : tap;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int kc = cb * BLOCK_K + chunk * 4 + e;
            float val = 0.f;
            if (row < row_c && kc < k_c) {
                const int co = dgrad ? kc : row, ci = dgrad ? row : kc;
                val = __ldg(w + ((size_t)co * cin + ci) * kk + wtap);
            }
            v[e] = val;
        }
        const size_t tile_base = ((size_t)tile * k_blocks + kb) * (size_t)n_tile * BLOCK_K;
        const size_t off = tile_base + (size_t)row_in_tile * BLOCK_K + (size_t)((chunk ^ (row_in_tile & 7)) * 4);
        *reinterpret_cast<float4 *>(wpack + off) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// ------------------------------------------------------------------ fprop / stride-1 dgrad
struct FwdParams {
    float *dst;
    const float *bias;
    const float *wpack;
    int act, accumulate;
    int src_c, dst_c;
    int out_w, out_h;     // output plane as the kernel sees it (1x1: (H*W, 1))
    int ks, pad;
    int kc_blocks, k_blocks, n_tile, n_tiles, stages;
    int wc, rows;         // tile = wc column chunks of 32 x rows output rows, wc * rows == 4
    int tiles_w;          // tiles per output row band
    int tiles_img;        // position tiles per image
    int total_tiles;      // n_tiles * tiles_img * batch
    FastDiv d_ntiles, d_tiles_img, d_tiles_w;
};

constexpr int FWD_EPI_WARPS = 8;
constexpr int FWD_THREADS = 64 + 32 * FWD_EPI_WARPS;

struct TileCoord { int tile_n, img, w0, h0; };

__device__ __forceinline__ TileCoord decode_tile(const FwdParams &p, int tile) {
    uint32_t rest, tn, img, pt, th, tw;
    p.d_ntiles.divmod((uint32_t)tile, rest, tn);
    p.d_tiles_img.divmod(rest, img, pt);
    p.d_tiles_w.divmod(pt, th, tw);
    TileCoord c;
    c.tile_n = (int)tn; c.img = (int)img;
    c.w0 = (int)tw * p.wc * 32; c.h0 = (int)th * p.rows;
    return c;
}

// Epilogue store of one 32-column chunk held in registers (lane = position, j = channel).
// FULL: all 32 channels exist, so the loop carries no per-element predicate.
template <int ACT, bool HAS_BIAS, bool ACCUM, bool FULL>
__device__ __forceinline__ void store_chunk(const uint32_t (&v)[32], float *d, uint32_t plane,
                                            const float *bias, int nvalid) {
    float old[ACCUM ? 32 : 1];
    if (ACCUM) {  // all 32 read-modify-write loads go out before the first dependent store
#pragma unroll
        for (int j = 0; j < 32; ++j)
            old[j] = (FULL || j < nvalid) ? __ldcs(d + (size_t)((uint32_t)j * plane)) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (FULL || j < nvalid) {
            float val = __uint_as_float(v[j]);
            if (HAS_BIAS) val += __ldg(bias + j);
            if (ACT == ACT_RELU) val = fmaxf(val, 0.f);
            else if (ACT == ACT_LRELU) val = val > 0 ? val : 0.1f * val;
            if (ACCUM) val += old[j];
            d[(size_t)((uint32_t)j * plane)] = val;
        }
    }
}
template <int ACT, bool HAS_BIAS, bool ACCUM>
__device__ __forceinline__ void store_chunk_any(const uint32_t (&v)[32], float *d, uint32_t plane,
                                                const float *bias, int nvalid) {
    if (nvalid == 32) store_chunk<ACT, HAS_BIAS, ACCUM, true>(v, d, plane, bias, 32);
    else if (nvalid > 0) store_chunk<ACT, HAS_BIAS, ACCUM, false>(v, d, plane, bias, nvalid);
}

// Persistent kernel: every CTA walks tiles blockIdx.x, +gridDim.x, ... (channel tile fastest, so
// CTAs working at the same time share an activation tile through L2). The TMA producer runs
// ahead across tile boundaries; two TMEM accumulators let the epilogue of tile i overlap the
// MMAs of tile i+1.
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv_tma_fwd_kernel(const __grid_constant__ CUtensorMap tm_src, const FwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int S = p.stages;
    const int n_tile = p.n_tile;
    const int b_stage_bytes = n_tile * BLOCK_K * 4;
    const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_bytes);
    uint64_t *full = bars, *empty = bars + S, *acc_full = bars + 2 * S, *acc_empty = bars + 2 * S + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t acc_cols = n_tile <= 32 ? 32 : (n_tile <= 64 ? 64 : (n_tile <= 128 ? 128 : 256));
    const uint32_t tmem_cols = 2 * acc_cols;

    if (t == 0) {
        for (int i = 0; i < 2 * S; ++i) mbar_init(smem_u32(bars + i), 1);
        mbar_init(smem_u32(acc_full), 1);
        mbar_init(smem_u32(acc_full + 1), 1);
        mbar_init(smem_u32(acc_empty), FWD_EPI_WARPS);
        mbar_init(smem_u32(acc_empty + 1), FWD_EPI_WARPS);
        fence_barrier_init();
        prefetch_tensormap(&tm_src);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord c = decode_tile(p, tile);
                const float *wtile = p.wpack + (size_t)c.tile_n * p.k_blocks * n_tile * BLOCK_K;
                int kb = 0;
                for (int kh = 0; kh < p.ks; ++kh) {
                    for (int kw = 0; kw < p.ks; ++kw) {
                        for (int cb = 0; cb < p.kc_blocks; ++cb, ++kb, ++it) {
                            const uint32_t s = it % (uint32_t)S;
                            mbar_wait(smem_u32(empty + s), ((it / (uint32_t)S) & 1) ^ 1);
                            const uint32_t fb = smem_u32(full + s);
                            uint8_t *a_stage = smem + (size_t)s * stage_bytes;
                            mbar_expect_tx(fb, (uint32_t)stage_bytes);
                            for (int a = 0; a < 4; ++a) {  // atom a = (column chunk, row) of the tile
                                const int wci = a / p.rows, r = a - wci * p.rows;
                                tma_load_4d(smem_u32(a_stage + a * ATOM_BYTES), &tm_src,
                                            c.w0 + wci * 32 + kw - p.pad, c.h0 + r + kh - p.pad,
                                            cb * BLOCK_K, c.img, fb);
                            }
                            bulk_copy_g2s(smem_u32(a_stage + A_STAGE_BYTES),
                                          wtile + (size_t)kb * n_tile * BLOCK_K, (uint32_t)b_stage_bytes, fb);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer
            const uint32_t idesc = make_idesc_tf32(TILE_M, n_tile, 1);
            uint32_t it = 0, local = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
                const uint32_t buf = local & 1, use = local >> 1;
                mbar_wait(smem_u32(acc_empty + buf), (use & 1) ^ 1);  // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * acc_cols;
                for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                    const uint32_t s = it % (uint32_t)S;
                    mbar_wait(smem_u32(full + s), (it / (uint32_t)S) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
                    for (int g = 0; g < BLOCK_K / UMMA_K; ++g) {
                        // A: K-group g = 8 channel rows = 1 KiB further inside every 4 KiB atom
                        const uint64_t da = make_desc_mn_sw128(a_addr + g * 1024, ATOM_BYTES, 512);
                        // B: 8 fp32 = 32 bytes further along the 128-byte K row
                        const uint64_t db = make_desc_sw128(b_addr) + (uint64_t)(2 * g);
                        umma_tf32(d_tmem, da, db, idesc, (kb > 0 || g > 0) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(empty + s));
                }
                umma_commit(smem_u32(acc_full + buf));
            }
        }
    } else {
        // ---------------- epilogue warps: TMEM lane quarter = warp & 3 = one atom of 32 positions;
        // the two warps sharing a quarter split the 32-column chunks between them
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const int wci = q / p.rows, r = q - wci * p.rows;
        const uint32_t plane = (uint32_t)(p.out_w * p.out_h);
        const int chunks32 = (n_tile + 31) / 32;
        uint32_t local = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
            const TileCoord c = decode_tile(p, tile);
            const uint32_t buf = local & 1, use = local >> 1;
            const int ow = c.w0 + wci * 32 + lane, oh = c.h0 + r;
            const bool valid = ow < p.out_w && oh < p.out_h;
            float *dst = p.dst + (size_t)c.img * p.dst_c * plane + (size_t)oh * p.out_w + ow;
            if (lane == 0) mbar_wait(smem_u32(acc_full + buf), use & 1);
            __syncwarp();
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * acc_cols + ((uint32_t)(q * 32) << 16);
            for (int ck = half; ck < chunks32; ck += 2) {
                uint32_t v[32];
                tmem_ld32(d_tmem + (uint32_t)(ck * 32), v);
                const int ch0 = c.tile_n * n_tile + ck * 32;
                int nvalid = min(32, min(n_tile - ck * 32, p.dst_c - ch0));
                if (!valid) nvalid = 0;
                float *d = dst + (size_t)ch0 * plane;
                const float *b = p.bias ? p.bias + ch0 : nullptr;
                if (p.accumulate) store_chunk_any<ACT_NONE, false, true>(v, d, plane, b, nvalid);
                else if (p.act == ACT_RELU) {
                    if (b) store_chunk_any<ACT_RELU, true, false>(v, d, plane, b, nvalid);
                    else store_chunk_any<ACT_RELU, false, false>(v, d, plane, b, nvalid);
                } else if (p.act == ACT_LRELU) {
                    if (b) store_chunk_any<ACT_LRELU, true, false>(v, d, plane, b, nvalid);
                    else store_chunk_any<ACT_LRELU, false, false>(v, d, plane, b, nvalid);
                } else if (p.act == ACT_NONE) {
                    if (b) store_chunk_any<ACT_NONE, true, false>(v, d, plane, b, nvalid);
                    else store_chunk_any<ACT_NONE, false, false>(v, d, plane, b, nvalid);
                } else {  // rare activations: generic arithmetic of the reference
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < nvalid) {
                            float val = __uint_as_float(v[j]);
                            if (b) val += __ldg(b + j);
                            d[(size_t)((uint32_t)j * plane)] = act_fwd(val, p.act, 0.f);
                        }
                    }
                }
            }
            // all of this warp's tcgen05.ld have completed (wait::ld inside tmem_ld32)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

struct FwdPlan {
    int n_tile, n_tiles, kc_blocks, k_blocks, stages;
    int view_w, view_h;      // source plane as the TMA sees it
    int out_w, out_h;
    int wc, rows, tiles_w, tiles_h;
    size_t wpack_bytes, smem_bytes;
};

bool tma_disabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("BCNN_B200_NO_TMA");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}

// Geometry shared by fprop (src = x) and stride-1 dgrad (src = dy): src plane (sh, sw) with
// src_c channels -> dst plane (dh, dw) with dst_c channels.
bool plan_fwd(int src_c, int sh, int sw, int dst_c, int dh, int dw, int ks, FwdPlan *pl) {
    const bool flat = (ks == 1);  // pad 0, stride 1: positions are one long contiguous row
    pl->view_w = flat ? sh * sw : sw;
    pl->view_h = flat ? 1 : sh;
    pl->out_w = flat ? dh * dw : dw;
    pl->out_h = flat ? 1 : dh;
    if (pl->view_w % 4 != 0) return false;  // TMA: row pitch must be a multiple of 16 bytes
    static int nmax = 0;
    if (!nmax) {
        const char *e = getenv("BCNN_B200_FWD_NMAX");
        nmax = e ? atoi(e) : 256;
        if (nmax < 16 || nmax > 256) nmax = 256;
    }
    int n = dst_c;
    if (n > nmax) {
        int tiles = ceil_div(n, nmax);
        n = ceil_div(ceil_div(n, tiles), 16) * 16;
    } else {
        n = ceil_div(n, 16) * 16;
    }
    pl->n_tile = n;
    pl->n_tiles = ceil_div(dst_c, n);
    pl->kc_blocks = ceil_div(src_c, BLOCK_K);
    pl->k_blocks = ks * ks * pl->kc_blocks;
    const int chunks = ceil_div(pl->out_w, 32);
    pl->wc = chunks >= 4 ? 4 : (chunks >= 2 ? 2 : 1);
    pl->rows = 4 / pl->wc;
    pl->tiles_w = ceil_div(chunks, pl->wc);
    pl->tiles_h = ceil_div(pl->out_h, pl->rows);
    const int stage = A_STAGE_BYTES + n * BLOCK_K * 4;
    int stages = (200 * 1024) / stage;  // persistent: one CTA per SM owns the shared memory
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    pl->stages = stages;
    pl->wpack_bytes = (size_t)pl->n_tiles * pl->k_blocks * n * BLOCK_K * sizeof(float);
    pl->smem_bytes = (size_t)stages * stage + 1024 + 256;
    return true;
}

bool fwd_shape_ok(const bcnn_b200_conv_desc *d, bool dgrad) {
    if (tma_disabled() || !encode_fn()) return false;
    if (d->groups != 1 || d->stride != 1) return false;
    const int src_c = dgrad ? d->cout : d->cin;
    if (src_c < 16) return false;  // K too thin (first layers): SIMT kernel
    if (d->ksize != 1 || d->pad != 0) return false;  // taps shift the inner TMA coordinate by 4 B
    FwdPlan pl;
    if (dgrad) return plan_fwd(d->cout, d->ho, d->wo, d->cin, d->h, d->w, d->ksize, &pl);
    return plan_fwd(d->cin, d->h, d->w, d->cout, d->ho, d->wo, d->ksize, &pl);
}

int launch_fwd(const bcnn_b200_conv_desc *d, bool dgrad, const float *src, const float *w,
               const float *bias, int act, float *dst, int accumulate, void *workspace,
               size_t workspace_bytes, cudaStream_t st) {
    FwdPlan pl;
    const int src_c = dgrad ? d->cout : d->cin, dst_c = dgrad ? d->cin : d->cout;
    const int sh = dgrad ? d->ho : d->h, sw = dgrad ? d->wo : d->w;
    const int dh = dgrad ? d->h : d->ho, dw = dgrad ? d->w : d->wo;
    if (!plan_fwd(src_c, sh, sw, dst_c, dh, dw, d->ksize, &pl)) return (int)cudaErrorInvalidValue;
    if (workspace == nullptr || workspace_bytes < pl.wpack_bytes) return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(src) & 15) != 0) return (int)cudaErrorMisalignedAddress;
    float *wpack = reinterpret_cast<float *>(workspace);
    const int kk = d->ksize * d->ksize;
    const size_t chunks = (size_t)pl.n_tiles * pl.n_tile * pl.k_blocks * 8;
    pack_weights_tf32_kernel<<<stream_grid(chunks, 256), 256, 0, st>>>(
        w, wpack, dgrad ? 1 : 0, d->cout, d->cin, kk, pl.n_tile, pl.n_tiles, pl.kc_blocks);
    int err = launched();
    if (err) return err;

    CUtensorMap tm;
    if (!make_map(&tm, src, pl.view_w, pl.view_h, src_c, d->batch, BLOCK_K, true))
        return (int)cudaErrorInvalidValue;
    FwdParams p;
    p.dst = dst; p.bias = bias; p.wpack = wpack; p.act = act; p.accumulate = accumulate;
    p.src_c = src_c; p.dst_c = dst_c;
    p.out_w = pl.out_w; p.out_h = pl.out_h;
    p.ks = d->ksize; p.pad = dgrad ? d->ksize - 1 - d->pad : d->pad;
    p.kc_blocks = pl.kc_blocks; p.k_blocks = pl.k_blocks; p.n_tile = pl.n_tile; p.stages = pl.stages;
    p.n_tiles = pl.n_tiles;
    p.wc = pl.wc; p.rows = pl.rows; p.tiles_w = pl.tiles_w;
    p.tiles_img = pl.tiles_w * pl.tiles_h;
    const long long total = (long long)pl.n_tiles * p.tiles_img * d->batch;
    if (total >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    p.total_tiles = (int)total;
    p.d_ntiles = FastDiv((uint32_t)pl.n_tiles);
    p.d_tiles_img = FastDiv((uint32_t)p.tiles_img);
    p.d_tiles_w = FastDiv((uint32_t)pl.tiles_w);

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tma_fwd_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    conv_tma_fwd_kernel<<<grid, FWD_THREADS, pl.smem_bytes, st>>>(tm, p);
    return launched();
}

// ------------------------------------------------------------------ wgrad
struct WgParams {
    float *out;  // split-K partial slabs [split][cout][cin][kk], or gw itself when splits == 1
    int cin, cout, kk, ks, pad;
    int n_tile;
    int chunks_w, out_h;   // k-blocks per image = out_h * chunks_w (32 output columns each)
    int kb_total, kb_per_split, splits;
    size_t split_stride;
    FastDiv d_img, d_cw;
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_tma_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                      const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int S = 4;
    const int n_tile = p.n_tile;
    const int b_stage_bytes = n_tile * BLOCK_K * 4;
    const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 1);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t tmem_cols = n_tile <= 32 ? 32 : (n_tile <= 64 ? 64 : (n_tile <= 128 ? 128 : 256));
    if (t == 0) {
        for (int i = 0; i < 2 * S + 1; ++i) mbar_init(smem_u32(bars + i), 1);
        fence_barrier_init();
        prefetch_tensormap(&tm_dy);
        prefetch_tensormap(&tm_x);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int split = blockIdx.x;
    const int tap = blockIdx.y % p.kk;
    const int ci0 = (blockIdx.y / p.kk) * n_tile;
    const int co0 = blockIdx.z * TILE_M;
    const int kh = tap / p.ks, kw = tap - kh * p.ks;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
    const int iters = kb_end - kb_begin;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                if (it >= S) mbar_wait(smem_u32(bars + S + s), ((it / S) - 1) & 1);
                uint32_t img, rem, oh, cw;
                p.d_img.divmod((uint32_t)(kb_begin + it), img, rem);
                p.d_cw.divmod(rem, oh, cw);
                const uint32_t full = smem_u32(bars + s);
                uint8_t *a_stage = smem + (size_t)s * stage_bytes;
                mbar_expect_tx(full, (uint32_t)stage_bytes);
                tma_load_4d(smem_u32(a_stage), &tm_dy, (int)cw * 32, (int)oh, co0, (int)img, full);
                tma_load_4d(smem_u32(a_stage + A_STAGE_BYTES), &tm_x, (int)cw * 32 + kw - p.pad,
                            (int)oh + kh - p.pad, ci0, (int)img, full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TILE_M, n_tile, 0);
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                mbar_wait(smem_u32(bars + s), (it / S) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t da = make_desc_sw128(a_addr);
                const uint64_t db = make_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
                for (int g = 0; g < BLOCK_K / UMMA_K; ++g)
                    umma_tf32(tmem_base, da + (uint64_t)(2 * g), db + (uint64_t)(2 * g), idesc,
                              (it > 0 || g > 0) ? 1u : 0u);
                umma_commit(smem_u32(bars + S + s));
                if (it == iters - 1) umma_commit(smem_u32(bars + 2 * S));
            }
        }
    } else {
        // epilogue: D[co lane, ci column] -> out[split][co][ci][tap]
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        float *out = p.out + (size_t)split * p.split_stride;
        if (iters > 0) {
            mbar_wait(smem_u32(bars + 2 * S), 0);
            tc_fence_after();
        }
        const int chunks32 = (n_tile + 31) / 32;
        for (int ck = 0; ck < chunks32; ++ck) {
            uint32_t v[32];
            if (iters > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ck * 32), v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (co >= p.cout) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int ci = ci0 + ck * 32 + j;
                if (ck * 32 + j < n_tile && ci < p.cin) {
                    float *d = out + ((size_t)co * p.cin + ci) * p.kk + tap;
                    const float val = __uint_as_float(v[j]);
                    if (p.splits == 1) *d += val;
                    else *d = val;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

__global__ void __launch_bounds__(256)
wgrad_reduce_tma_kernel(float *__restrict__ gw, const float *__restrict__ partial, size_t n, int splits) {
    const size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += __ldg(partial + (size_t)k * n + i);
        gw[i] += s;
    }
}

struct WgPlan {
    int n_tile, ci_tiles, co_tiles;
    int view_w, view_h, out_w, out_h, chunks_w;
    int kb_total, splits, kb_per_split;
    size_t smem_bytes, partial_bytes;
};

bool plan_wgrad(const bcnn_b200_conv_desc *d, WgPlan *pl) {
    const bool flat = d->ksize == 1;
    pl->view_w = flat ? d->h * d->w : d->w;
    pl->view_h = flat ? 1 : d->h;
    pl->out_w = flat ? d->ho * d->wo : d->wo;
    pl->out_h = flat ? 1 : d->ho;
    if (pl->view_w % 4 != 0 || pl->out_w % 4 != 0) return false;
    int n = d->cin;
    if (n > 128) {
        int tiles = ceil_div(n, 128);
        n = ceil_div(ceil_div(n, tiles), 16) * 16;
    } else {
        n = ceil_div(n, 16) * 16;
    }
    pl->n_tile = n;
    pl->ci_tiles = ceil_div(d->cin, n);
    pl->co_tiles = ceil_div(d->cout, TILE_M);
    pl->chunks_w = ceil_div(pl->out_w, 32);
    const long long kb_total = (long long)d->batch * pl->out_h * pl->chunks_w;
    if (kb_total >= (1LL << 31)) return false;
    pl->kb_total = (int)kb_total;
    const int kk = d->ksize * d->ksize;
    const long long tiles = (long long)pl->ci_tiles * pl->co_tiles * kk;
    long long want = (2LL * sm_count() + tiles - 1) / tiles;
    long long max_by_k = pl->kb_total / 16;  // >= 16 k-blocks per split
    if (want > max_by_k) want = max_by_k;
    if (want > 256) want = 256;
    if (want < 1) want = 1;
    pl->kb_per_split = ceil_div(pl->kb_total, (int)want);
    pl->splits = ceil_div(pl->kb_total, pl->kb_per_split);
    pl->smem_bytes = (size_t)4 * (A_STAGE_BYTES + n * BLOCK_K * 4) + 1024 + 256;
    const size_t wsize = (size_t)d->cout * d->cin * kk;
    pl->partial_bytes = pl->splits > 1 ? (size_t)pl->splits * wsize * sizeof(float) : 0;
    return true;
}

bool wg_shape_ok(const bcnn_b200_conv_desc *d) {
    if (tma_disabled() || !encode_fn()) return false;
    if (d->groups != 1 || d->stride != 1) return false;
    if (d->cin < 16 || d->cout < 32) return false;
    if (d->ksize != 1 || d->pad != 0) return false;
    if ((long long)d->batch * d->ho * d->wo < 512) return false;
    WgPlan pl;
    return plan_wgrad(d, &pl);
}

}  // namespace

namespace b200 {

bool conv_tma_supports_fprop(const bcnn_b200_conv_desc *d) { return fwd_shape_ok(d, false); }
bool conv_tma_supports_dgrad(const bcnn_b200_conv_desc *d) { return fwd_shape_ok(d, true); }
bool conv_tma_supports_wgrad(const bcnn_b200_conv_desc *d) { return wg_shape_ok(d); }

size_t conv_tma_workspace_bytes(const bcnn_b200_conv_desc *d) {
    size_t need = 0;
    FwdPlan pl;
    if (fwd_shape_ok(d, false) && plan_fwd(d->cin, d->h, d->w, d->cout, d->ho, d->wo, d->ksize, &pl))
        need = pl.wpack_bytes;
    if (fwd_shape_ok(d, true) && plan_fwd(d->cout, d->ho, d->wo, d->cin, d->h, d->w, d->ksize, &pl) &&
        pl.wpack_bytes > need)
        need = pl.wpack_bytes;
    WgPlan wp;
    if (wg_shape_ok(d) && plan_wgrad(d, &wp) && wp.partial_bytes > need) need = wp.partial_bytes;
    return need;
}

int conv_tma_forward(const bcnn_b200_conv_desc *d, const float *x, const float *w, const float *bias,
                     int act, float *y, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    return launch_fwd(d, false, x, w, bias, act, y, 0, workspace, workspace_bytes, st);
}

int conv_tma_backward_data(const bcnn_b200_conv_desc *d, const float *w, const float *dy, float *dx,
                           int accumulate, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    return launch_fwd(d, true, dy, w, nullptr, 0, dx, accumulate, workspace, workspace_bytes, st);
}

int conv_tma_backward_weights(const bcnn_b200_conv_desc *d, const float *x, const float *dy, float *gw,
                              void *workspace, size_t workspace_bytes, cudaStream_t st) {
    WgPlan pl;
    if (!plan_wgrad(d, &pl)) return (int)cudaErrorInvalidValue;
    if (pl.splits > 1 && (workspace == nullptr || workspace_bytes < pl.partial_bytes))
        return (int)cudaErrorInvalidValue;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) != 0)
        return (int)cudaErrorMisalignedAddress;
    CUtensorMap tm_dy, tm_x;
    if (!make_map(&tm_dy, dy, pl.out_w, pl.out_h, d->cout, d->batch, TILE_M, false) ||
        !make_map(&tm_x, x, pl.view_w, pl.view_h, d->cin, d->batch, pl.n_tile, false))
        return (int)cudaErrorInvalidValue;
    WgParams p;
    p.cin = d->cin; p.cout = d->cout; p.ks = d->ksize; p.kk = d->ksize * d->ksize; p.pad = d->pad;
    p.n_tile = pl.n_tile; p.chunks_w = pl.chunks_w; p.out_h = pl.out_h;
    p.kb_total = pl.kb_total; p.kb_per_split = pl.kb_per_split; p.splits = pl.splits;
    const size_t wsize = (size_t)d->cout * d->cin * p.kk;
    p.split_stride = pl.splits > 1 ? wsize : 0;
    p.out = pl.splits > 1 ? reinterpret_cast<float *>(workspace) : gw;
    p.d_img = FastDiv((uint32_t)(pl.out_h * pl.chunks_w));
    p.d_cw = FastDiv((uint32_t)pl.chunks_w);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tma_wgrad_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid(pl.splits, pl.ci_tiles * p.kk, pl.co_tiles);
    conv_tma_wgrad_kernel<<<grid, NTHREADS, pl.smem_bytes, st>>>(tm_dy, tm_x, p);
    int err = launched();
    if (err) return err;
    if (pl.splits > 1) {
        wgrad_reduce_tma_kernel<<<stream_grid(wsize, 256), 256, 0, st>>>(
            gw, reinterpret_cast<const float *>(workspace), wsize, pl.splits);
        return launched();
    }
    return 0;
}

}  // namespace b200
