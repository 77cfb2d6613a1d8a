// conv_tma.cu -- TF32 implicit-GEMM convolution on tcgen05 tensor cores with every operand
// staged by the TMA (sm_100a). Accumulation is FP32 in TMEM; kind::tf32 consumes FP32 bits as
// they are, so no precision-conversion pass exists anywhere.
//
// Two ways to let the TMA see an activation tensor (bcnn tensors are FP32 NCHW):
//
//  (1) DIRECT (1x1, stride 1, pad 0, H*W % 4 == 0): a 4-D map (w, h, c, n) over the NCHW tensor
//      itself, the image plane seen as one long row. A box "32 positions x 32 channels" lands
//      in shared memory as a tcgen05 *MN-major* TF32 atom (positions contiguous, channels
//      strided; SWIZZLE_128B_ATOM_32B, the only MN-major layout 32-bit operands accept).
//  (2) NHWC SHADOW (k > 1, stride 2, odd widths): the inner TMA coordinate must be a multiple of
//      16 bytes, so a filter tap cannot shift an NCHW row by one pixel. The tensor is therefore
//      transposed once into an NHWC copy in the workspace (one HBM-bound pass); over it the
//      4-D map (c, w, h, n) lets a single box fetch "32 channels x TW x TH x TN positions" of
//      one tap at any (kh - pad, kw - pad) shift and any stride (TMA element strides), zero-
//      filled outside the image by the TMA's out-of-bounds rule: a K-major SWIZZLE_128B tile.
//
//   conv_tma_fwd_kernel   fprop and stride-1 dgrad (same kernel on dY with flipped taps and
//       pad' = k-1-pad): D[128 positions x N channels] += A[128 x 32ch] * B[N x 32ch]^T,
//       B = K-major packed weights (one bulk copy per k-block); K order tap-major, channel-
//       minor. Persistent CTAs (one per SM) walk the tile list; the TMA producer runs ahead
//       across tiles; two TMEM accumulators overlap the epilogue of tile i with the MMAs of
//       tile i+1. Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps
//       2-9 epilogue (tcgen05.ld -> bias/activation -> coalesced NCHW stores).
//   conv_tma_wgrad_kernel  dW[co, ci, tap] = sum_pos dY[co, pos] * X[ci, pos (+) tap]:
//       D[128 co x N ci], reduction over positions. DIRECT: both operands K-major rows of 32
//       positions. NHWC: both operands MN-major (channels contiguous), K-block = bw x bh
//       positions of one tap-shifted window. Split-K across CTAs, deterministic second-stage
//       reduction in split order.
//
// Shapes outside both (groups, C % 4 != 0, thin first layers, stride-2 dgrad) stay on the
// register-gather tcgen05 kernels of conv_tc.cu / conv_tc_wgrad.cu or the FP32 SIMT kernels.
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
constexpr int WG_THREADS = 192;

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
// Shared -> global tensor stores of the bulk async-group kind (3-D map: position, channel, image).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src_smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, uint32_t src_smem, int c0, int c1,
                                                  int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// 4-D flavours: BF16 NHWC results seen as (channel, w, h, image)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src_smem, int c0, int c1, int c2,
                                             int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap *map, uint32_t src_smem, int c0, int c1,
                                                  int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every earlier bulk group of this thread has finished READING its shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// at most one bulk group (the most recent) of this thread may still be reading shared memory
__device__ __forceinline__ void bulk_wait_read_le1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
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
// MN-major 16-bit operand, plain SWIZZLE_128B (layout type 2): 64 bf16 along MN are contiguous
// (128 B), 8 K-rows form a 1 KiB atom with 16-byte chunks XOR-swizzled by (K-row & 7).
// LBO = byte stride between MN atoms, SBO = byte stride between groups of 8 K-rows.
__device__ __forceinline__ uint64_t make_desc_mn_sw128_b16(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D = F32, A = B = BF16 (kind::f16); a_mn / b_mn mark an MN-major operand.
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) |
           ((uint32_t)(b_mn ? 1 : 0) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D = F32, A = B = TF32; a_mn / b_mn mark an MN-major operand (bits 15 / 16).
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) |
           ((uint32_t)(b_mn ? 1 : 0) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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
bool make_map_nchw(CUtensorMap *map, const float *base, int w, int h, int c, int n, int box_c,
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

// Map over an NCHW fp32 result seen as (position, channel, image), the plane as one row (plane % 4
// == 0): the store box is box_pos consecutive positions x 16 channels x box_n images, no swizzle.
// Elements of a box that fall outside the tensor are not written.
bool make_map_out(CUtensorMap *map, float *base, int plane, int c, int n, int box_pos, int box_n) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)plane, (cuuint64_t)c, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)plane * 4, (cuuint64_t)plane * c * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_pos, 16, (cuuint32_t)box_n};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Map over a BF16 NHWC result seen as (c, w, h, n): the store box is 64 channels x tw x th x tn
// positions, SWIZZLE_128B (row r of the staged tile is 128 bytes whose 16-byte chunks are XOR-ed
// with r & 7). Elements of a box that fall outside the tensor are not written.
bool make_map_out16(CUtensorMap *map, void *base, int c, int w, int h, int n, int tw, int th, int tn) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)w * h * c * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Map over an NHWC fp32 shadow seen as (c, w, h, n); box = 32 channels x bw x bh x bn positions,
// walking w and h with the convolution stride.
// bf16 shadows: 64 channels per 128-byte row, plain SWIZZLE_128B for both operand majors.
bool make_map_nhwc(CUtensorMap *map, const void *base, int c, int w, int h, int n, int bw, int bh,
                   int bn, int stride, bool mn_major, bool bf16) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t es = bf16 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * es, (cuuint64_t)w * c * es, (cuuint64_t)w * h * c * es};
    cuuint32_t box[4] = {bf16 ? 64u : 32u, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride),
                         (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    (mn_major && !bf16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ------------------------------------------------------------------ NCHW -> NHWC shadow
// out[n][p][c] = in[n][c][p]; 32 x 32 tiles through padded shared memory, 128-byte coalesced on
// both sides.
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float *__restrict__ in, float *__restrict__ out, int C, int P) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *src = in + (size_t)n * C * P;
    float *dst = out + (size_t)n * C * P;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, p = p0 + tx;
        tile[ty + 8 * i][tx] = (c < C && p < P) ? __ldg(src + (size_t)c * P + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + ty + 8 * i, c = c0 + tx;
        if (p < P && c < C) dst[(size_t)p * C + c] = tile[tx][ty + 8 * i];
    }
}

// bf16 flavour: 64 channels x 32 positions per CTA, each thread stores a bf16 pair so a warp
// still writes 128 contiguous bytes. C % 8 == 0 (16-byte rows for the TMA).
__global__ void __launch_bounds__(256)
nchw_to_nhwc_bf16_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, int C, int P) {
    __shared__ float tile[64][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
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
                pack_bf16x2(tile[2 * tx][ty + 8 * i], tile[2 * tx + 1][ty + 8 * i]);
    }
}

int launch_transpose(const float *in, void *out, int n, int c, int p, bool bf16, cudaStream_t st) {
    if (bf16) {
        dim3 grid(ceil_div(p, 32), ceil_div(c, 64), n);
        nchw_to_nhwc_bf16_kernel<<<grid, 256, 0, st>>>(in, reinterpret_cast<__nv_bfloat16 *>(out), c, p);
    } else {
        dim3 grid(ceil_div(p, 32), ceil_div(c, 32), n);
        nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>(in, reinterpret_cast<float *>(out), c, p);
    }
    return launched();
}

// Shadows and packed weights are bf16 unless BCNN_B200_SHADOW=fp32 (TF32 on the shadow routes too).
bool shadow_bf16() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("BCNN_B200_SHADOW");
        v = (e && (e[0] == 'f' || e[0] == 'F')) ? 0 : 1;
    }
    return v == 1;
}

// ------------------------------------------------------------------ weight repack (TF32 = fp32 bits)
// wpack[tile][kb][row][32 floats], 16-byte chunks XOR-swizzled by (row & 7): byte-for-byte the
// K-major SWIZZLE_128B shared-memory image, so a k-block tile is one bulk copy.
//   fprop: row = co, k = ci, tap as is            -> W[co][ci][tap]
//   dgrad: row = ci, k = co, tap flipped (kk-1-t) -> W[co][ci][kk-1-tap]
//   sub-sampled dgrad classes (stride > 1): taps.idx lists the filter taps of the class
struct TapMap {
    int n;            // taps of this (sub-)convolution, in the kernel's (kh, kw) walk order
    short idx[64];    // tap t reads W[..][..][idx[t]]
};

// BF16: rows of 64 bf16 (same 128 bytes, same 16-byte chunk swizzle), 8 elements per chunk.
// One 16-byte chunk i of a packed image.
template <bool BF16>
__device__ __forceinline__ void pack_chunk(const float *__restrict__ w, uint8_t *__restrict__ wpack, int dgrad,
                                           int cout, int cin, int kk, int n_tile, int kc_blocks,
                                           const TapMap &taps, size_t i) {
    constexpr int EPC = BF16 ? 8 : 4;   // elements per 16-byte chunk
    const int k_blocks = taps.n * kc_blocks;
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
    const int wtap = taps.idx[tap];
    float v[EPC];
#pragma unroll
    for (int e = 0; e < EPC; ++e) {
        const int kc = (cb * 8 + chunk) * EPC + e;
        float val = 0.f;
        if (row < row_c && kc < k_c) {
            const int co = dgrad ? kc : row, ci = dgrad ? row : kc;
            val = __ldg(w + ((size_t)co * cin + ci) * kk + wtap);
        }
        v[e] = val;
    }
    const size_t tile_base = ((size_t)tile * k_blocks + kb) * (size_t)n_tile * 128;
    const size_t off = tile_base + (size_t)row_in_tile * 128 + (size_t)((chunk ^ (row_in_tile & 7)) * 16);
    if constexpr (BF16) {
        *reinterpret_cast<uint4 *>(wpack + off) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                       pack_bf16x2(v[EPC - 2], v[EPC - 1]));
    } else {
        *reinterpret_cast<float4 *>(wpack + off) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

template <bool BF16>
__global__ void __launch_bounds__(256)
pack_weights_tf32_kernel(const float *__restrict__ w, uint8_t *__restrict__ wpack, int dgrad, int cout,
                         int cin, int kk, int n_tile, int n_tiles, int kc_blocks, const TapMap taps) {
    const size_t chunks = (size_t)n_tiles * n_tile * taps.n * kc_blocks * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks;
         i += (size_t)gridDim.x * blockDim.x)
        pack_chunk<BF16>(w, wpack, dgrad, cout, cin, kk, n_tile, kc_blocks, taps, i);
}

// Every packed BF16 image of a net in ONE launch (the resident step packed 114 images -- fprop, dgrad
// and the classes of strided dgrads of 53 layers -- in 114 launches of ~4 us): a table of jobs in
// device memory, CTA -> job by bisection of first_block, 2048 chunks per CTA.
struct PackJob {
    const float *w;
    uint8_t *dst;
    int dgrad, cout, cin, kk, n_tile, n_tiles, kc_blocks;
    unsigned int first_block;   // first CTA of this job; the table ends with a sentinel job holding the grid size
    TapMap taps;
};
constexpr int PACK_JOB_CHUNKS = 2048;

__global__ void __launch_bounds__(256)
pack_jobs_kernel(const PackJob *__restrict__ jobs, int count) {
    int lo = 0, hi = count;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (blockIdx.x >= jobs[mid].first_block) lo = mid;
        else hi = mid;
    }
    const PackJob &j = jobs[lo];
    const size_t chunks = (size_t)j.n_tiles * j.n_tile * j.taps.n * j.kc_blocks * 8;
    const size_t base = (size_t)(blockIdx.x - j.first_block) * PACK_JOB_CHUNKS;
    for (int k = threadIdx.x; k < PACK_JOB_CHUNKS; k += 256) {
        const size_t i = base + k;
        if (i < chunks) pack_chunk<true>(j.w, j.dst, j.dgrad, j.cout, j.cin, j.kk, j.n_tile, j.kc_blocks, j.taps, i);
    }
}

// ------------------------------------------------------------------ fprop / stride-1 dgrad
struct FwdParams {
    float *dst;
    // batch-norm statistics fused into the epilogue (nullptr = off): every epilogue warp sums the
    // accumulator rows (and their squares) of all the tiles its CTA walks; the lane quarters are folded
    // at the end and the CTA leaves one row: stat_partial[(row * 2 + {0, 1}) * dst_c + channel],
    // row = blockIdx.x / n_tiles; the grid is a multiple of n_tiles, so a CTA stays on one channel
    // tile. Every entry is written exactly once; bn_stats_from_partials folds the rows in a fixed order.
    float *stat_partial;
    const float *bias;
    const uint8_t *wpack;
    int act, accumulate;
    // tstore: the epilogue stages each 32-channel chunk in shared memory and a bulk tensor store
    // (reduce-add when accumulating) writes it through tm_dst; the tile is tile_pos consecutive
    // positions of the plane (x tn images). 0: per-thread stores (scattered dgrad classes, planes
    // whose size is not a multiple of 4 elements).
    int tstore, tile_pos;
    // out16: the result is a BF16 NHWC tensor (resident activations). 1: each half of the epilogue
    // warps stages 128 positions x 64 channels (128-byte rows, SWIZZLE_128B) and one bulk tensor
    // store (reduce-add when accumulating) writes the box through the 4-D map tm_dst; 2: per-thread
    // 16-byte stores (scattered strided-dgrad classes, channel tiles that are not 64-aligned).
    int out16;
    // halo mode (resident, stride 1, k > 1): the input patch of a tile -- (th + ksh - 1) rows of
    // vw = tw + ksw - 1 columns -- is loaded ONCE per 64-channel block and the taps are MMAs whose A
    // descriptor starts kh * vw + kw rows into it (tile row m = vh * vw + vcol; columns >= tw are
    // junk the epilogue skips). Activation tiles (A ring, sa slots) and weight tap tiles (B ring, sb
    // slots) travel through separate rings; b_resident: every weight tile of the layer fits the B ring
    // and is loaded once per CTA.
    int halo, vw, sa, sb, b_resident;
    uint32_t a_slot_bytes, ring_bytes;   // ring_bytes: everything in front of the staging buffers
    FastDiv d_vw;
    int narrow;      // n_tile <= 64: both epilogue halves share the single 64-channel group
    int dbg;         // timing decomposition (BCNN_B200_DBG_EPI bit mask; results are garbage)
    // stats_frag: fused statistics of BF16 results through fragment-shaped TMEM loads and shuffles
    // (frag_col_sums) instead of the shared-memory transpose; then the scratch of the transpose holds
    // a second output staging buffer per epilogue half (stage_bufs == 2: a store drains while the
    // next group is staged)
    int stats_frag, stage_bufs;
    int src_c, dst_c, batch;
    int out_w, out_h;     // output plane as the kernel sees it (DIRECT: (H*W, 1))
    int ksh, ksw, pad_h, pad_w, stride;   // tap window (rows x columns) and its leading pads
    // output scatter: logical position (oh, ow) lands at (oh * o_s + o_oy, ow * o_s + o_ox) of a
    // plane dst_w wide holding dst_plane elements (strided-dgrad classes; 1 / 0 / 0 otherwise)
    int o_s, o_oy, o_ox, dst_w, dst_plane;
    int kc_blocks, k_blocks, n_tile, n_tiles, stages;
    // DIRECT: tile = wc column chunks of 32 x rows output rows (wc * rows == 4), one image
    int wc, rows;
    // NHWC: tile = tw x th x tn output positions (<= 128), row m = w + tw * (h + th * n)
    int tw, th, tn;
    uint32_t a_bytes;     // TMA bytes of the activation operand per stage
    int tiles_w, tiles_h, tiles_b;   // tiles along w, h and the batch
    int total_tiles;      // n_tiles * tiles_w * tiles_h * tiles_b
    FastDiv d_ntiles, d_tiles_w, d_tiles_h, d_tw, d_th;
};

constexpr int FWD_EPI_WARPS = 8;
constexpr int FWD_THREADS = 64 + 32 * FWD_EPI_WARPS;
// output staging of the bulk-store epilogue: per half of the epilogue warps one chunk of 128
// positions x 32 channels, laid out [16-channel sub-box][image][channel][position]
constexpr int OUT_STAGE_BYTES = TILE_M * 32 * 4;
constexpr int STAT_SCRATCH_BYTES = 8 * 32 * 36 * 4;   // chunk_col_sums scratch of the 8 epilogue warps
// alignment slack, barriers, output staging, statistics scratch
constexpr int FWD_BAR_BYTES = 512;
constexpr int FWD_EXTRA_SMEM = 1024 + FWD_BAR_BYTES + 2 * OUT_STAGE_BYTES + STAT_SCRATCH_BYTES;

struct TileCoord { int tile_n, m_tile, img, w0, h0; };

template <bool NHWC>
__device__ __forceinline__ TileCoord decode_tile(const FwdParams &p, int tile) {
    uint32_t rest, tn, rest2, tw, tb, th;
    p.d_ntiles.divmod((uint32_t)tile, rest, tn);
    p.d_tiles_w.divmod(rest, rest2, tw);
    p.d_tiles_h.divmod(rest2, tb, th);
    TileCoord c;
    c.tile_n = (int)tn;
    c.m_tile = (int)rest;
    if (NHWC) {
        c.img = (int)tb * p.tn; c.w0 = (int)tw * p.tw; c.h0 = (int)th * p.th;
    } else {
        c.img = (int)tb; c.w0 = (int)tw * p.wc * 32; c.h0 = (int)th * p.rows;
    }
    return c;
}

// Rare activations (everything but none / ReLU / leaky ReLU) in the epilogues: one out-of-line copy of the
// reference's arithmetic instead of 32 inlined copies per call site (the kernel was 2 MB of SASS).
__device__ __noinline__ float act_fwd_rare(float x, int act) { return act_fwd(x, act, 0.f); }

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

// Per-channel sum / sum of squares of one 32-column chunk of a warp's 32 accumulator rows
// (lane = position row, v[j] = channel j): the warp writes its 32 x 32 block to its own scratch
// (row pitch 36 floats: 16-byte aligned and conflict-free for both accesses) and every lane sums
// one channel column. 8 vector stores + 32 loads per chunk instead of the 62 shuffles of a
// transposing butterfly, which bound the epilogue (and with it thin-K layers) on the shuffle pipe.
constexpr int STAT_PITCH = 36;
constexpr int STAT_WARP_FLOATS = 32 * STAT_PITCH;
__device__ __forceinline__ void chunk_col_sums(const uint32_t (&v)[32], bool valid, float *wscr, int lane,
                                               float &a1, float &a2) {
    float4 *row = reinterpret_cast<float4 *>(wscr + lane * STAT_PITCH);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        row[j] = valid ? make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                     __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const float x = wscr[r * STAT_PITCH + lane];
        s1[r & 3] += x;
        s2[r & 3] = fmaf(x, x, s2[r & 3]);
    }
    __syncwarp();
    a1 += (s1[0] + s1[1]) + (s1[2] + s1[3]);
    a2 += (s2[0] + s2[1]) + (s2[2] + s2[3]);
}

// Register-only flavour of the same sums (resident epilogues): the warp's 32 accumulator rows x 8*NB
// columns are read a second time in the m16n8 fragment shape (tmem_ld_frag*), where a thread holds 4
// rows of 2 columns per 8-column block: 3 adds per column in the thread, then a halving butterfly
// over the 8 threads that share a column pair (lane bits 4, 3, 2): 7 * NB / 4 shuffles per quantity
// instead of a 32 x 32 trip through shared memory per chunk, which bound thin-K layers on the
// shared-memory pipe (profiles/r2_epilogue_decomposition.txt: 90 of 210 us on 1x1 64->256 @56).
// vmask: ballot of the rows that exist. Results: NB == 8: s[c] / q[c] = column 2 * lane + c (c = 0, 1);
// NB == 4: s[0] / q[0] = column frag32_col(lane).
template <int N>
__device__ __forceinline__ void halve_across(float (&v)[N], bool upper, int lane_mask) {
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
        const float send = upper ? v[j] : v[j + N / 2];
        const float keep = upper ? v[j + N / 2] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, lane_mask);
    }
}
__device__ __forceinline__ int frag32_col(int lane) {
    return ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + (lane & 3) * 2 + ((lane >> 2) & 1);
}
template <int NB>
__device__ __forceinline__ void frag_col_sums(const uint32_t (&lo)[4 * NB], const uint32_t (&hi)[4 * NB],
                                              uint32_t vmask, int lane, float *s_out, float *q_out) {
    float s[2 * NB], q[2 * NB];
    const int r = lane >> 2;
    if (vmask == 0xffffffffu) {
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) {
            const int b = i >> 1, c = i & 1;
            const float x0 = __uint_as_float(lo[4 * b + c]), x1 = __uint_as_float(lo[4 * b + 2 + c]);
            const float x2 = __uint_as_float(hi[4 * b + c]), x3 = __uint_as_float(hi[4 * b + 2 + c]);
            s[i] = (x0 + x1) + (x2 + x3);
            q[i] = fmaf(x0, x0, x1 * x1) + fmaf(x2, x2, x3 * x3);
        }
    } else {
        const bool m0 = (vmask >> r) & 1u, m1 = (vmask >> (r + 8)) & 1u;
        const bool m2 = (vmask >> (r + 16)) & 1u, m3 = (vmask >> (r + 24)) & 1u;
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) {
            const int b = i >> 1, c = i & 1;
            const float x0 = m0 ? __uint_as_float(lo[4 * b + c]) : 0.f;
            const float x1 = m1 ? __uint_as_float(lo[4 * b + 2 + c]) : 0.f;
            const float x2 = m2 ? __uint_as_float(hi[4 * b + c]) : 0.f;
            const float x3 = m3 ? __uint_as_float(hi[4 * b + 2 + c]) : 0.f;
            s[i] = (x0 + x1) + (x2 + x3);
            q[i] = fmaf(x0, x0, x1 * x1) + fmaf(x2, x2, x3 * x3);
        }
    }
    halve_across(s, (lane & 16) != 0, 16);
    halve_across(q, (lane & 16) != 0, 16);
    float s2[NB], q2[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { s2[i] = s[i]; q2[i] = q[i]; }
    halve_across(s2, (lane & 8) != 0, 8);
    halve_across(q2, (lane & 8) != 0, 8);
    float s3[NB / 2], q3[NB / 2];
#pragma unroll
    for (int i = 0; i < NB / 2; ++i) { s3[i] = s2[i]; q3[i] = q2[i]; }
    halve_across(s3, (lane & 4) != 0, 4);
    halve_across(q3, (lane & 4) != 0, 4);
#pragma unroll
    for (int i = 0; i < NB / 4; ++i) { s_out[i] += s3[i]; q_out[i] += q3[i]; }
}

// Persistent kernel: every CTA walks tiles blockIdx.x, +gridDim.x, ... (channel tile fastest, so
// CTAs working at the same time share an activation tile through L2). The TMA producer runs
// ahead across tile boundaries; two TMEM accumulators let the epilogue of tile i overlap the
// MMAs of tile i+1.
template <bool NHWC, bool BF16>
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv_tma_fwd_kernel(const __grid_constant__ CUtensorMap tm_src, const __grid_constant__ CUtensorMap tm_dst,
                    const FwdParams p) {
    constexpr int KC = BF16 ? 64 : 32;   // channels per 128-byte k-block row
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int S = p.stages;
    const int n_tile = p.n_tile;
    const int b_stage_bytes = n_tile * BLOCK_K * 4;
    const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    // [operand ring(s)][2 output staging buffers (1 KiB aligned)][statistics scratch, or staging
    // buffers 2 and 3][barriers]
    const size_t ring = p.ring_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + ring + 2 * OUT_STAGE_BYTES + STAT_SCRATCH_BYTES);
    // plain: full[S], empty[S]; halo: fullA[sa], emptyA[sa], fullB[sb], emptyB[sb]
    const int nring = p.halo ? 2 * (p.sa + p.sb) : 2 * S;
    uint64_t *full = bars, *empty = bars + S, *acc_full = bars + nring, *acc_empty = bars + nring + 2;
    uint64_t *full_a = bars, *empty_a = bars + p.sa, *full_b = bars + 2 * p.sa, *empty_b = bars + 2 * p.sa + p.sb;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + nring + 4);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t acc_cols = n_tile <= 32 ? 32 : (n_tile <= 64 ? 64 : (n_tile <= 128 ? 128 : 256));
    const uint32_t tmem_cols = 2 * acc_cols;

    if (t == 0) {
        for (int i = 0; i < nring; ++i) mbar_init(smem_u32(bars + i), 1);
        mbar_init(smem_u32(acc_full), 1);
        mbar_init(smem_u32(acc_full + 1), 1);
        mbar_init(smem_u32(acc_empty), FWD_EPI_WARPS);
        mbar_init(smem_u32(acc_empty + 1), FWD_EPI_WARPS);
        fence_barrier_init();
        prefetch_tensormap(&tm_src);
        if (p.tstore || p.out16 == 1) prefetch_tensormap(&tm_dst);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // The two single-thread roles below keep their loop state in running counters (ring slot, phase
    // bit, byte offsets): with `it % S`, `it / S` and per-tap divisions the issue loop cost ~700 cycles
    // per k-block (SASS: two MUFU.RCP division chains), more than the MMAs of a k-block take, and
    // bounded every layer with more than a few k-blocks per tile.
    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer
            const uint32_t tx_bytes = p.a_bytes + (uint32_t)b_stage_bytes;
            const uint32_t smem0 = smem_u32(smem);
            if (NHWC && p.halo) {
                const uint32_t ring_b = smem0 + (uint32_t)p.sa * p.a_slot_bytes;
                const uint32_t fa0 = smem_u32(full_a), ea0 = smem_u32(empty_a);
                const uint32_t fb0 = smem_u32(full_b), eb0 = smem_u32(empty_b);
                uint32_t sa = 0, pa = 1, sb = 0, pb = 1;   // ring slots and the parity an empty slot shows
                const int taps = p.ksh * p.ksw;
                bool b_loaded = false;
                for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                    const TileCoord c = decode_tile<NHWC>(p, tile);
                    const uint8_t *wtile = p.wpack + (size_t)c.tile_n * p.k_blocks * b_stage_bytes;
                    for (int cb = 0; cb < p.kc_blocks; ++cb) {
                        mbar_wait(ea0 + 8 * sa, pa);
                        mbar_expect_tx(fa0 + 8 * sa, p.a_bytes);
                        tma_load_4d(smem0 + sa * p.a_slot_bytes, &tm_src, cb * KC, c.w0 - p.pad_w,
                                    c.h0 - p.pad_h, c.img, fa0 + 8 * sa);
                        if (++sa == (uint32_t)p.sa) { sa = 0; pa ^= 1; }
                        if (p.b_resident && b_loaded) continue;
                        // tap tiles of this channel block: k-block index tap * kc_blocks + cb
                        const uint8_t *wk = wtile + (size_t)cb * b_stage_bytes;
                        for (int tap = 0; tap < taps; ++tap, wk += (size_t)p.kc_blocks * b_stage_bytes) {
                            mbar_wait(eb0 + 8 * sb, pb);
                            mbar_expect_tx(fb0 + 8 * sb, (uint32_t)b_stage_bytes);
                            bulk_copy_g2s(ring_b + sb * (uint32_t)b_stage_bytes, wk, (uint32_t)b_stage_bytes,
                                          fb0 + 8 * sb);
                            if (++sb == (uint32_t)p.sb) { sb = 0; pb ^= 1; }
                        }
                    }
                    b_loaded = true;
                }
            } else {
                const uint32_t f0 = smem_u32(full), e0 = smem_u32(empty);
                uint32_t s = 0, ph = 1;
                for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                    const TileCoord c = decode_tile<NHWC>(p, tile);
                    const uint8_t *wk = p.wpack + (size_t)c.tile_n * p.k_blocks * b_stage_bytes;
                    const int x0 = c.w0 * (NHWC ? p.stride : 1) - p.pad_w;
                    const int y0 = c.h0 * (NHWC ? p.stride : 1) - p.pad_h;
                    for (int kh = 0; kh < p.ksh; ++kh) {
                        for (int kw = 0; kw < p.ksw; ++kw) {
                            for (int cb = 0; cb < p.kc_blocks; ++cb, wk += b_stage_bytes) {
                                mbar_wait(e0 + 8 * s, ph);
                                const uint32_t fb = f0 + 8 * s;
                                const uint32_t a_stage = smem0 + s * (uint32_t)stage_bytes;
                                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                                if (p.dbg & 32) { mbar_arrive(fb); continue; }
                                mbar_expect_tx(fb, tx_bytes);
                                if (NHWC) {
                                    tma_load_4d(a_stage, &tm_src, cb * KC, x0 + kw, y0 + kh, c.img, fb);
                                } else {
                                    for (int a = 0; a < 4; ++a) {  // atom a = (column chunk, row) of the tile
                                        const int wci = a / p.rows, r = a - wci * p.rows;
                                        tma_load_4d(a_stage + a * ATOM_BYTES, &tm_src, x0 + wci * 32 + kw,
                                                    y0 + r + kh, cb * BLOCK_K, c.img, fb);
                                    }
                                }
                                bulk_copy_g2s(a_stage + A_STAGE_BYTES, wk, (uint32_t)b_stage_bytes, fb);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: the whole warp walks the loop (uniform state), one elected lane issues
        const uint32_t idesc = BF16 ? make_idesc_bf16(TILE_M, n_tile, 0, 0)
                                    : make_idesc_tf32(TILE_M, n_tile, NHWC ? 0 : 1, 0);
        const uint64_t desc_k = make_desc_sw128(0);   // K-major SWIZZLE_128B descriptor without its address
        const uint32_t smem0 = smem_u32(smem);
        const uint32_t accf0 = smem_u32(acc_full), acce0 = smem_u32(acc_empty);
        uint32_t local = 0;
        if (NHWC && p.halo) {
            const uint32_t ring_b = smem0 + (uint32_t)p.sa * p.a_slot_bytes;
            const uint32_t fa0 = smem_u32(full_a), ea0 = smem_u32(empty_a);
            const uint32_t fb0 = smem_u32(full_b), eb0 = smem_u32(empty_b);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;   // ring slots and the parity a filled slot shows
            const uint32_t row_step = (uint32_t)(p.vw - p.ksw) * 128u;   // from the last tap of a row to the next row
            if (p.b_resident) {
                // Resident weights: after the tap tiles have landed (once per CTA) a tile is one wait per
                // channel block and a straight run of MMAs whose descriptors advance by constants, all
                // inside one elected region (~16 instructions per tap instead of ~75).
                const int slots = p.kc_blocks * p.ksh * p.ksw;
                for (int i = 0; i < slots; ++i) mbar_wait(fb0 + 8 * (uint32_t)i, 0u);
                tc_fence_after();
                const uint64_t b_step = (uint64_t)((uint32_t)b_stage_bytes >> 4);
                for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
                    const uint32_t buf = local & 1, use = local >> 1;
                    mbar_wait(acce0 + 8 * buf, (use & 1) ^ 1);  // epilogue drained this buffer
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * acc_cols;
                    const uint64_t db0 = desc_k | (uint64_t)((ring_b & 0x3FFFF) >> 4);
                    uint32_t acc = 0;
                    for (int cb = 0; cb < p.kc_blocks; ++cb) {
                        mbar_wait(fa0 + 8 * sa, pa);
                        tc_fence_after();
                        if (elect_one()) {
                            uint64_t da = desc_k | (uint64_t)(((smem0 + sa * p.a_slot_bytes) & 0x3FFFF) >> 4);
                            uint64_t db = db0 + b_step * (uint64_t)(cb * p.ksh * p.ksw);   // slot of (cb, tap 0)
                            for (int kh = 0; kh < p.ksh; ++kh, da += (uint64_t)(row_step >> 4)) {
                                for (int kw = 0; kw < p.ksw; ++kw, da += 8u, db += b_step) {
#pragma unroll
                                    for (int g = 0; g < BLOCK_K / UMMA_K; ++g)
                                        umma_bf16(d_tmem, da + (uint64_t)(2 * g), db + (uint64_t)(2 * g), idesc,
                                                  acc | (uint32_t)g);
                                    acc = 1;
                                }
                            }
                            umma_commit(ea0 + 8 * sa);
                        }
                        __syncwarp();
                        acc = 1;
                        if (++sa == (uint32_t)p.sa) { sa = 0; pa ^= 1; }
                    }
                    if (elect_one()) umma_commit(accf0 + 8 * buf);
                    __syncwarp();
                }
            } else
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
                const uint32_t buf = local & 1, use = local >> 1;
                mbar_wait(acce0 + 8 * buf, (use & 1) ^ 1);  // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * acc_cols;
                uint32_t first = 0;   // 0 until the first MMA of the tile went out
                for (int cb = 0; cb < p.kc_blocks; ++cb) {
                    mbar_wait(fa0 + 8 * sa, pa);
                    tc_fence_after();
                    uint32_t a_addr = smem0 + sa * p.a_slot_bytes;   // row kh * vw + kw of the patch
                    uint32_t rb = (uint32_t)cb * (uint32_t)(p.ksh * p.ksw);   // resident weights: slot of (cb, tap 0)
                    for (int kh = 0; kh < p.ksh; ++kh, a_addr += row_step) {
                        for (int kw = 0; kw < p.ksw; ++kw, a_addr += 128u, ++rb) {
                            const uint32_t slot = p.b_resident ? rb : sb;
                            if (!p.b_resident || local == 0) {   // resident tiles arrive during the first tile
                                mbar_wait(fb0 + 8 * slot, p.b_resident ? 0u : pb);
                                tc_fence_after();
                            }
                            const uint64_t da = desc_k | (uint64_t)((a_addr & 0x3FFFF) >> 4);
                            const uint64_t db = desc_k | (uint64_t)(((ring_b + slot * (uint32_t)b_stage_bytes) & 0x3FFFF) >> 4);
                            if (elect_one()) {
#pragma unroll
                                for (int g = 0; g < BLOCK_K / UMMA_K; ++g)
                                    umma_bf16(d_tmem, da + (uint64_t)(2 * g), db + (uint64_t)(2 * g), idesc,
                                              first | (uint32_t)g);
                                if (!p.b_resident) umma_commit(eb0 + 8 * sb);
                            }
                            __syncwarp();
                            first = 1;
                            if (!p.b_resident && ++sb == (uint32_t)p.sb) { sb = 0; pb ^= 1; }
                        }
                    }
                    if (elect_one()) umma_commit(ea0 + 8 * sa);
                    __syncwarp();
                    if (++sa == (uint32_t)p.sa) { sa = 0; pa ^= 1; }
                }
                if (elect_one()) umma_commit(accf0 + 8 * buf);
                __syncwarp();
            }
        } else {
            const uint32_t f0 = smem_u32(full), e0 = smem_u32(empty);
            uint32_t s = 0, ph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
                const uint32_t buf = local & 1, use = local >> 1;
                mbar_wait(acce0 + 8 * buf, (use & 1) ^ 1);  // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * acc_cols;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(f0 + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem0 + s * (uint32_t)stage_bytes;
                    const uint32_t b_addr = a_addr + A_STAGE_BYTES;
                    const uint64_t db = desc_k | (uint64_t)((b_addr & 0x3FFFF) >> 4);
                    if (!(p.dbg & 16) && elect_one()) {
#pragma unroll
                        for (int g = 0; g < BLOCK_K / UMMA_K; ++g) {
                            // DIRECT A: K-group g = 8 channel rows = 1 KiB further inside every 4 KiB atom
                            // NHWC A / B: 8 fp32 (16 bf16) = 32 bytes further along the 128-byte K row
                            const uint64_t da = NHWC ? (desc_k | (uint64_t)((a_addr & 0x3FFFF) >> 4)) + (uint64_t)(2 * g)
                                                     : make_desc_mn_sw128(a_addr + g * 1024, ATOM_BYTES, 512);
                            if (BF16) umma_bf16(d_tmem, da, db + (uint64_t)(2 * g), idesc, (kb > 0 || g > 0) ? 1u : 0u);
                            else umma_tf32(d_tmem, da, db + (uint64_t)(2 * g), idesc, (kb > 0 || g > 0) ? 1u : 0u);
                        }
                    }
                    if (p.dbg & 64) {   // (timing experiment, with bit 16: plain arrive instead of a commit)
                        if (elect_one()) mbar_arrive(e0 + 8 * s);
                    } else if (elect_one()) umma_commit(e0 + 8 * s);
                    __syncwarp();
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit(accf0 + 8 * buf);
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue warps: TMEM lane quarter = warp & 3 (32 tile rows); the two
        // warps sharing a quarter split the 32-column chunks between them
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        // fused statistics: running sums of this warp's <= 4 column chunks (lane = channel)
        float acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
        // bulk-store epilogue: staging buffer of this half, position of this thread's row in it
        const int hw = ew & 3;   // warp within the half
        float *stage = reinterpret_cast<float *>(smem + ring + (size_t)half * OUT_STAGE_BYTES);
        float *wscr = reinterpret_cast<float *>(smem + ring + 2 * OUT_STAGE_BYTES) + (size_t)ew * STAT_WARP_FLOATS;
        const int sub_stride = p.tn * 16 * p.tile_pos;   // floats of one 16-channel sub-box
        uint32_t stores = 0;                              // bulk stores issued by this half so far
        const uint32_t plane = (uint32_t)p.dst_plane;
        const int chunks32 = (n_tile + 31) / 32;
        // position of this thread's tile row relative to the tile origin
        int rw, rh, rn;
        if (NHWC && p.halo) {   // virtual rows of vw columns; columns >= tw and rows >= th are junk
            uint32_t m = (uint32_t)(q * 32 + lane), h_, w_;
            p.d_vw.divmod(m, h_, w_);
            rw = (int)w_; rh = (int)h_; rn = 0;
        } else if (NHWC) {
            uint32_t m = (uint32_t)(q * 32 + lane), t2, w_, n_, h_;
            p.d_tw.divmod(m, t2, w_);
            p.d_th.divmod(t2, n_, h_);
            rw = (int)w_; rh = (int)h_; rn = (int)n_;
        } else {
            const int wci = q / p.rows;
            rw = wci * 32 + lane; rh = q - wci * p.rows; rn = 0;
        }
        const bool in_box = (NHWC && p.halo) ? (rw < p.tw && rh < p.th) : (NHWC ? (rn < p.tn) : (rh == 0));
        const int row_off = NHWC ? rn * 16 * p.tile_pos + rh * p.tw + rw : rw;
        uint32_t local = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
            const TileCoord c = decode_tile<NHWC>(p, tile);
            const uint32_t buf = local & 1, use = local >> 1;
            const int ow = c.w0 + rw, oh = c.h0 + rh, img = c.img + rn;
            bool valid = ow < p.out_w && oh < p.out_h && img < p.batch;
            if (NHWC) valid = valid && in_box;
            // (address arithmetic of the per-thread store paths is done inside them: the bulk-store paths,
            // which run 99 % of the tiles, need none of it -- 155 instructions per warp and tile here
            // were 10 % of the samples of a thin-K layer, ncu source page, gpurun r2zy)
            if (lane == 0) mbar_wait(smem_u32(acc_full + buf), use & 1);
            __syncwarp();
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * acc_cols + ((uint32_t)(q * 32) << 16);
            if (p.out16) {
                // BF16 NHWC result: this half owns the 64-channel groups gi = half, half + 2
                auto dst16_of = [&]() {   // this thread's position in the result (per-thread store paths only)
                    return reinterpret_cast<__nv_bfloat16 *>(p.dst) +
                           ((size_t)img * plane + (size_t)(oh * p.o_s + p.o_oy) * p.dst_w +
                            (size_t)(ow * p.o_s + p.o_ox)) * (size_t)p.dst_c;
                };
                // row of the staged box: halo tiles drop their junk columns (box rows are tw wide)
                const int r = p.halo ? rh * p.tw + rw : q * 32 + lane;
                const bool stats = p.stat_partial != nullptr && !(p.dbg & 1);
                const uint32_t vmask = (stats && p.stats_frag) ? __ballot_sync(0xffffffffu, valid) : 0u;
                if (p.narrow) {
                    // Narrow tiles (one 64-channel group): both halves work on it, half h on the 32-
                    // channel chunk h, through ONE staging buffer and one store -- with the group left
                    // to a single half the epilogue chain of thin layers (the stem, 64-channel 1x1 /
                    // 3x3, every dgrad into 64 channels) took 2 us per tile with four warps idle.
                    const int ck = half;
                    const int ch0 = c.tile_n * n_tile + ck * 32;
                    uint32_t pkn[16];
                    if (ck < chunks32) {
                        uint32_t v[32];
                        tmem_ld32(d_tmem + (uint32_t)(ck * 32), v);
                        if (p.bias != nullptr || p.act != ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float val = __uint_as_float(v[j]);
                                if (p.bias != nullptr && ch0 + j < p.dst_c) val += __ldg(p.bias + ch0 + j);
                                if (p.act == ACT_RELU) val = fmaxf(val, 0.f);
                                else if (p.act == ACT_LRELU) val = val > 0 ? val : 0.1f * val;
                                else val = act_fwd_rare(val, p.act);
                                v[j] = __float_as_uint(val);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            pkn[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pkn[j] = 0u;
                    }
                    if (p.out16 == 1 && (p.dbg & 8)) {
                    } else if (p.out16 == 1) {
                        // staging buffers 0 and 2 in turn (stage_bufs == 2), else buffer 0
                        uint8_t *sbn = smem + ring + (size_t)((stores & (uint32_t)(p.stage_bufs - 1)) * 2) * OUT_STAGE_BYTES;
                        ++stores;
                        if (ew == 0 && lane == 0) {   // the store issued from this buffer has read it
                            if (p.stage_bufs == 2) bulk_wait_read_le1();
                            else bulk_wait_read_all();
                        }
                        named_bar_sync(3, 256);
                        if (in_box) {
                            const uint32_t row = smem_u32(sbn) + (uint32_t)r * 128u;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                st_shared_v4(row + (uint32_t)(((half * 4 + j) ^ (r & 7)) << 4), pkn[4 * j],
                                             pkn[4 * j + 1], pkn[4 * j + 2], pkn[4 * j + 3]);
                        }
                        fence_proxy_async();
                        named_bar_sync(3, 256);
                        if (ew == 0 && lane == 0) {
                            const int g0 = c.tile_n * n_tile;
                            if (p.accumulate) tma_reduce_add_4d(&tm_dst, smem_u32(sbn), g0, c.w0, c.h0, c.img);
                            else tma_store_4d(&tm_dst, smem_u32(sbn), g0, c.w0, c.h0, c.img);
                            bulk_commit_group();
                        }
                    } else if (valid && ck < chunks32) {
                        __nv_bfloat16 *dst16 = dst16_of();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (ch0 + j * 8 < p.dst_c && ck * 32 + j * 8 < n_tile) {
                                uint4 o = make_uint4(pkn[4 * j], pkn[4 * j + 1], pkn[4 * j + 2], pkn[4 * j + 3]);
                                uint4 *gp = reinterpret_cast<uint4 *>(dst16 + ch0 + j * 8);
                                if (p.accumulate) {
                                    const uint4 old = *gp;
                                    const uint32_t ov[4] = {old.x, old.y, old.z, old.w};
                                    uint32_t nv[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float lo = __uint_as_float(ov[e] << 16) + __uint_as_float(nv[e] << 16);
                                        const float hi2 = __uint_as_float(ov[e] & 0xffff0000u) +
                                                          __uint_as_float(nv[e] & 0xffff0000u);
                                        nv[e] = pack_bf16x2(lo, hi2);
                                    }
                                    o = make_uint4(nv[0], nv[1], nv[2], nv[3]);
                                }
                                *gp = o;
                            }
                        }
                    }
                    // statistics from a second read of the accumulators (cheaper than keeping the
                    // registers alive across the store), while the bulk store drains
                    if (stats && ck < chunks32) {
                        if (p.stats_frag) {
                            uint32_t lo[16], hi[16];
                            tmem_ld_frag32(d_tmem + (uint32_t)(ck * 32), lo, hi);
                            frag_col_sums<4>(lo, hi, vmask, lane, &acc1[0], &acc2[0]);
                        } else {
                            uint32_t v[32];
                            tmem_ld32(d_tmem + (uint32_t)(ck * 32), v);
                            chunk_col_sums(v, valid, wscr, lane, acc1[0], acc2[0]);
                        }
                    }
                } else
#pragma unroll
                for (int gj = 0; gj < 2; ++gj) {
                    const int gi = half + 2 * gj;
                    if (gi * 64 >= n_tile) break;
                    uint32_t pk[2][16];
#pragma unroll
                    for (int sc = 0; sc < 2; ++sc) {
                        const int ck = 2 * gi + sc;
                        const int ch0 = c.tile_n * n_tile + ck * 32;
                        if (ck < chunks32) {
                            uint32_t v[32];
                            tmem_ld32(d_tmem + (uint32_t)(ck * 32), v);
                            if (p.bias != nullptr || p.act != ACT_NONE) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    float val = __uint_as_float(v[j]);
                                    if (p.bias != nullptr && ch0 + j < p.dst_c) val += __ldg(p.bias + ch0 + j);
                                    if (p.act == ACT_RELU) val = fmaxf(val, 0.f);
                                    else if (p.act == ACT_LRELU) val = val > 0 ? val : 0.1f * val;
                                    else val = act_fwd_rare(val, p.act);
                                    v[j] = __float_as_uint(val);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                pk[sc][j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) pk[sc][j] = 0u;
                        }
                    }
                    const int ch0 = c.tile_n * n_tile + gi * 64;
                    if (p.out16 == 1 && (p.dbg & 8)) {
                    } else if (p.out16 == 1) {
                        // this half's staging buffers half and 2 + half in turn (stage_bufs == 2): the
                        // bulk store of one group drains while the next group is staged
                        uint8_t *sb = smem + ring +
                                      (size_t)((stores & (uint32_t)(p.stage_bufs - 1)) * 2 + half) * OUT_STAGE_BYTES;
                        ++stores;
                        if (hw == 0 && lane == 0) {   // the store issued from this buffer has read it
                            if (p.stage_bufs == 2) bulk_wait_read_le1();
                            else bulk_wait_read_all();
                        }
                        named_bar_sync(1 + half, 128);
                        if (in_box) {
                            const uint32_t row = smem_u32(sb) + (uint32_t)r * 128u;
#pragma unroll
                            for (int k8 = 0; k8 < 8; ++k8) {
                                const uint32_t *w4 = &pk[k8 >> 2][(k8 & 3) * 4];
                                st_shared_v4(row + (uint32_t)((k8 ^ (r & 7)) << 4), w4[0], w4[1], w4[2], w4[3]);
                            }
                        }
                        fence_proxy_async();
                        named_bar_sync(1 + half, 128);
                        if (hw == 0 && lane == 0) {
                            if (p.accumulate) tma_reduce_add_4d(&tm_dst, smem_u32(sb), ch0, c.w0, c.h0, c.img);
                            else tma_store_4d(&tm_dst, smem_u32(sb), ch0, c.w0, c.h0, c.img);
                            bulk_commit_group();
                        }
                    } else if (valid) {
                        __nv_bfloat16 *dst16 = dst16_of();
#pragma unroll
                        for (int k8 = 0; k8 < 8; ++k8) {
                            if (ch0 + k8 * 8 < p.dst_c && gi * 64 + k8 * 8 < n_tile) {
                                const uint32_t *w4 = &pk[k8 >> 2][(k8 & 3) * 4];
                                uint4 o = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                                uint4 *gp = reinterpret_cast<uint4 *>(dst16 + ch0 + k8 * 8);
                                if (p.accumulate) {
                                    const uint4 old = *gp;
                                    const uint32_t ov[4] = {old.x, old.y, old.z, old.w};
                                    uint32_t nv[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float lo = __uint_as_float(ov[e] << 16) + __uint_as_float(nv[e] << 16);
                                        const float hi2 = __uint_as_float(ov[e] & 0xffff0000u) +
                                                          __uint_as_float(nv[e] & 0xffff0000u);
                                        nv[e] = pack_bf16x2(lo, hi2);
                                    }
                                    o = make_uint4(nv[0], nv[1], nv[2], nv[3]);
                                }
                                *gp = o;
                            }
                        }
                    }
                    // batch-norm statistics of this group while its bulk store drains: the accumulators
                    // are read again, cheaper than keeping 64 registers alive across the store. (Summing
                    // the staged BF16 tile instead -- each warp reading back its own 32 rows, lane = channel
                    // pair -- was measured slower twice: 0.237 vs 0.191 ms on 1x1 64->256 @56 in gpurun r2j,
                    // 0.163 vs 0.145 ms after the fragment-shaped second read, gpurun r2zj.)
                    if (stats) {
                        if (p.stats_frag) {
                            uint32_t lo[32], hi[32];
                            tmem_ld_frag64(d_tmem + (uint32_t)(gi * 64), lo, hi);
                            frag_col_sums<8>(lo, hi, vmask, lane, &acc1[gj * 2], &acc2[gj * 2]);
                        } else {
#pragma unroll
                            for (int sc = 0; sc < 2; ++sc)
                                if (2 * gi + sc < chunks32) {
                                    uint32_t v[32];
                                    tmem_ld32(d_tmem + (uint32_t)((2 * gi + sc) * 32), v);
                                    chunk_col_sums(v, valid, wscr, lane, acc1[gj * 2 + sc], acc2[gj * 2 + sc]);
                                }
                        }
                    }
                }
            } else if (p.tstore) {
                const int pos0 = NHWC ? c.h0 * p.dst_w + c.w0 : c.w0;
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int ck = half + 2 * ci;
                    if (ck < chunks32) {
                        uint32_t v[32];
                        tmem_ld32(d_tmem + (uint32_t)(ck * 32), v);
                        const int ch0 = c.tile_n * n_tile + ck * 32;
                        if (p.bias != nullptr || p.act != ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float val = __uint_as_float(v[j]);
                                if (p.bias != nullptr && ch0 + j < p.dst_c) val += __ldg(p.bias + ch0 + j);
                                if (p.act == ACT_RELU) val = fmaxf(val, 0.f);
                                else if (p.act == ACT_LRELU) val = val > 0 ? val : 0.1f * val;
                                else val = act_fwd_rare(val, p.act);
                                v[j] = __float_as_uint(val);
                            }
                        }
                        // two 16-channel sub-boxes per chunk, ping-ponging between the two halves of
                        // the staging buffer: a sub-box is rewritten once the bulk store issued two
                        // stores ago has finished reading it, so the store of one sub-box drains
                        // while the warps fill the other
                        const int nsub = (n_tile - ck * 32) > 16 ? 2 : 1;
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            if (sub < nsub) {
                                float *sb = stage + (stores & 1) * sub_stride;
                                if (hw == 0 && lane == 0) bulk_wait_read_le1();
                                named_bar_sync(1 + half, 128);
                                if (in_box) {
                                    float *sp = sb + row_off;
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        sp[j * p.tile_pos] = __uint_as_float(v[sub * 16 + j]);
                                }
                                fence_proxy_async();
                                named_bar_sync(1 + half, 128);
                                if (hw == 0 && lane == 0) {
                                    if (p.accumulate) tma_reduce_add_3d(&tm_dst, smem_u32(sb), pos0, ch0 + sub * 16, c.img);
                                    else tma_store_3d(&tm_dst, smem_u32(sb), pos0, ch0 + sub * 16, c.img);
                                    bulk_commit_group();
                                }
                                ++stores;
                            }
                        }
                        if (p.stat_partial != nullptr) chunk_col_sums(v, valid, wscr, lane, acc1[ci], acc2[ci]);
                    }
                }
            } else
            for (int ck = half; ck < chunks32; ck += 2) {
                uint32_t v[32];
                tmem_ld32(d_tmem + (uint32_t)(ck * 32), v);
                const int ch0 = c.tile_n * n_tile + ck * 32;
                int nvalid = min(32, min(n_tile - ck * 32, p.dst_c - ch0));
                if (!valid) nvalid = 0;
                float *dst = p.dst + (size_t)img * p.dst_c * plane +
                             (size_t)(oh * p.o_s + p.o_oy) * p.dst_w + (ow * p.o_s + p.o_ox);
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
                            d[(size_t)((uint32_t)j * plane)] = act_fwd_rare(val, p.act);
                        }
                    }
                }
            }
            // all of this warp's tcgen05.ld have completed (wait::ld inside tmem_ld32)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + buf));
        }
        if (p.stat_partial != nullptr) {
            // The four lane quarters of each half are folded here, in quarter order, through the (now
            // idle) staging region, so the CTA leaves ONE row: the finalize kernel reads a quarter of
            // the rows it used to (148 instead of 592 for a full grid).
            if ((p.tstore || p.out16 == 1) && hw == 0 && lane == 0) bulk_wait_read_all();
            named_bar_sync(3, 32 * FWD_EPI_WARPS);
            float *fold = reinterpret_cast<float *>(smem + ring) + (size_t)half * 3 * 8 * 32;
            if (q != 0) {
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    fold[((q - 1) * 8 + ci) * 32 + lane] = acc1[ci];
                    fold[((q - 1) * 8 + 4 + ci) * 32 + lane] = acc2[ci];
                }
            }
            named_bar_sync(3, 32 * FWD_EPI_WARPS);
            if (q == 0) {
#pragma unroll
                for (int qq = 0; qq < 3; ++qq)
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        acc1[ci] += fold[(qq * 8 + ci) * 32 + lane];
                        acc2[ci] += fold[(qq * 8 + 4 + ci) * 32 + lane];
                    }
            }
            // row = CTA index among those of this channel tile
            const uint32_t tile_n = blockIdx.x % (uint32_t)p.n_tiles;
            const size_t row = (size_t)(blockIdx.x / (uint32_t)p.n_tiles);
            if (q == 0)
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                // out16: this half owns the 64-channel groups half, half + 2 (two chunks each); narrow
                // tiles (n_tile <= 64): chunk `half`, kept in slot 0
                int ck = p.out16 ? 2 * (half + 2 * (ci >> 1)) + (ci & 1) : half + 2 * ci;
                if (p.out16 && p.narrow) ck = ci == 0 ? half : chunks32;
                int col = ck * 32 + lane;   // column of the channel tile this lane's sums belong to
                if (p.out16 && p.stats_frag) {
                    // frag_col_sums: slots 2 gj + c hold column 2 * lane + c of group half + 2 gj; narrow
                    // tiles: slot 0 holds column frag32_col(lane) of chunk `half`
                    col = p.narrow ? ck * 32 + frag32_col(lane) : (half + 2 * (ci >> 1)) * 64 + 2 * lane + (ci & 1);
                }
                const int ch = (int)tile_n * n_tile + col;
                if (ck < chunks32 && col < n_tile && ch < p.dst_c) {
                    p.stat_partial[(row * 2 + 0) * p.dst_c + ch] = acc1[ci];
                    p.stat_partial[(row * 2 + 1) * p.dst_c + ch] = acc2[ci];
                }
            }
        }
        // shared memory must outlive the bulk stores that read it
        if ((p.tstore || p.out16 == 1) && hw == 0 && lane == 0) bulk_wait_read_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// One launch of conv_tma_fwd_kernel: a stride-`stride` correlation of a ksh x ksw tap window
// over the source tensor (sh x sw, src_c channels), producing dh x dw logical positions of
// dst_c channels that are scattered into the destination plane.
struct FwdGeom {
    int batch, src_c, sh, sw;
    int dst_c, dh, dw;
    int ksh, ksw, pad_h, pad_w, stride;
    int o_s, o_oy, o_ox, dst_w, dst_plane;
    bool direct;   // source may be read in place through the NCHW map
    // resident activations: the source already is a BF16 NHWC tensor (no shadow to make) and the
    // result is written as BF16 NHWC (dst_h = dst_plane / dst_w rows of dst_w positions)
    bool resident = false;
};

struct FwdPlan {
    bool nhwc, bf16;
    int n_tile, n_tiles, kc_blocks, k_blocks, stages;
    int view_w, view_h;      // DIRECT: source plane as the TMA sees it
    int out_w, out_h;
    int wc, rows, tw, th, tn;
    int tiles_w, tiles_h, tiles_b;
    uint32_t a_bytes;
    size_t shadow_bytes, wpack_bytes, smem_bytes;
    bool tstore;             // bulk tensor-store epilogue (tile = tile_pos consecutive positions)
    int out16;               // BF16 NHWC result: 1 bulk tensor store, 2 per-thread stores (FwdParams)
    bool halo;               // halo tiles (FwdParams.halo)
    int vw, ph, sa, sb, b_resident;
    uint32_t a_slot_bytes, ring_bytes;
    int tile_pos;
    int stat_rows;           // rows of the fused batch-norm partials: one per CTA of a channel tile
    size_t stat_bytes;
};

bool tma_disabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("BCNN_B200_NO_TMA");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}
bool nhwc_disabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("BCNN_B200_NO_NHWC");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}
bool env_off(const char *name) {
    const char *e = getenv(name);
    return e && e[0] && e[0] != '0';
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

bool plan_fwd(const FwdGeom &g, FwdPlan *pl) {
    if (g.src_c < 16) return false;  // K too thin: im2col route or SIMT kernel
    const bool direct = g.direct && g.ksh == 1 && g.ksw == 1 && g.stride == 1 && g.pad_h == 0 &&
                        g.pad_w == 0 && g.o_s == 1 && (g.sh * g.sw) % 4 == 0;
    pl->nhwc = !direct;
    if (!direct && (nhwc_disabled() || g.src_c % 4 != 0 || g.stride > 4)) return false;
    if (g.ksh * g.ksw > 64) return false;   // TapMap capacity
    if (g.resident && (g.src_c % 8 != 0 || g.dst_c % 8 != 0)) return false;   // 16-byte channel vectors
    pl->bf16 = !direct && (g.resident || shadow_bf16()) && g.src_c % 8 == 0;
    const int kc = pl->bf16 ? 64 : 32;       // channels per 128-byte k-block row
    static int nmax = 0;
    if (!nmax) {
        const char *e = getenv("BCNN_B200_FWD_NMAX");
        nmax = e ? atoi(e) : 256;
        if (nmax < 16 || nmax > 256) nmax = 256;
    }
    int n = g.dst_c;
    if (n > nmax) {
        int tiles = ceil_div(n, nmax);
        n = ceil_div(ceil_div(n, tiles), 16) * 16;
    } else {
        n = ceil_div(n, 16) * 16;
    }
    // resident results are stored in boxes of 64 channels: several channel tiles must not overlap
    if (g.resident && ceil_div(g.dst_c, n) > 1 && n % 64 != 0) n = ceil_div(n, 64) * 64;
    pl->n_tile = n;
    pl->n_tiles = ceil_div(g.dst_c, n);
    pl->kc_blocks = ceil_div(g.src_c, kc);
    pl->k_blocks = g.ksh * g.ksw * pl->kc_blocks;
    pl->wc = pl->rows = 1; pl->tw = pl->th = pl->tn = 1;
    if (direct) {
        pl->view_w = g.sh * g.sw; pl->view_h = 1;
        pl->out_w = g.dh * g.dw; pl->out_h = 1;
        const int chunks = ceil_div(pl->out_w, 32);
        pl->wc = chunks >= 4 ? 4 : (chunks >= 2 ? 2 : 1);
        pl->rows = 4 / pl->wc;
        pl->tiles_w = ceil_div(chunks, pl->wc);
        pl->tiles_h = ceil_div(pl->out_h, pl->rows);
        pl->tiles_b = g.batch;
        pl->a_bytes = A_STAGE_BYTES;
        pl->shadow_bytes = 0;
        pl->tile_pos = pl->wc * 32;
        pl->tstore = g.o_s == 1 && g.dst_plane % 4 == 0;
    } else {
        pl->view_w = g.sw; pl->view_h = g.sh;
        pl->out_w = g.dw; pl->out_h = g.dh;
        // tile = tw x th x tn output positions, <= 128 rows, TMA box dims <= 256
        int tw = g.dw < 128 ? g.dw : 128;
        while (tw * g.stride > 256) --tw;
        const int tiles_w = ceil_div(g.dw, tw);
        tw = ceil_div(g.dw, tiles_w);  // balanced
        int th_max = 128 / tw;
        if (th_max < 1) th_max = 1;
        while (th_max * g.stride > 256) --th_max;
        int th = g.dh < th_max ? g.dh : th_max;
        int tiles_h = ceil_div(g.dh, th);
        th = ceil_div(g.dh, tiles_h);
        // bulk-store epilogue: a tile must be a run of consecutive plane positions whose length is a
        // multiple of 4 elements (16 bytes): full-width rows, th a multiple of 4 / gcd(tw, 4)
        if (tiles_w == 1 && tiles_h > 1 && (tw * th) % 4 != 0) {
            const int m = (tw % 2 == 0) ? 2 : 4;
            int up = ceil_div(th, m) * m;
            if (up > th_max) up = (th / m) * m;
            if (up >= m) {
                th = up;
                tiles_h = ceil_div(g.dh, th);
            }
        }
        int tn = 1;
        if (tiles_w == 1 && tiles_h == 1) {
            tn = 128 / (tw * th);
            if (tn > g.batch) tn = g.batch;
            if (tn < 1) tn = 1;
        }
        pl->tw = tw; pl->th = th; pl->tn = tn;
        pl->tiles_w = tiles_w; pl->tiles_h = tiles_h; pl->tiles_b = ceil_div(g.batch, tn);
        pl->a_bytes = (uint32_t)(tw * th * tn * 128);
        pl->shadow_bytes = align256((size_t)g.batch * g.src_c * g.sh * g.sw * (pl->bf16 ? 2 : 4));
        pl->tile_pos = tw * th;
        const bool contiguous = tiles_w == 1 || (th == 1 && g.dw % tw == 0);
        pl->tstore = g.o_s == 1 && g.dst_plane % 4 == 0 && g.dst_w == g.dw && contiguous &&
                     pl->tile_pos % 4 == 0 && !env_off("BCNN_B200_NO_TSTORE");
    }
    if (env_off("BCNN_B200_NO_TSTORE")) pl->tstore = false;
    pl->out16 = 0;
    if (g.resident) {
        if (direct) return false;
        pl->shadow_bytes = 0;
        pl->out16 = (g.o_s == 1 && g.dst_w == g.dw && g.dst_c % 64 == 0 && !env_off("BCNN_B200_NO_TSTORE")) ? 1 : 2;
        pl->tstore = false;
    }
    const int stage = A_STAGE_BYTES + n * BLOCK_K * 4;
    int stages = (227 * 1024 - FWD_EXTRA_SMEM) / stage;  // persistent: one CTA per SM owns the shared memory
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    {
        static const int smax = [] { const char *e = getenv("BCNN_B200_FWD_STAGES"); return e ? atoi(e) : 0; }();
        if (smax >= 2 && stages > smax) stages = smax;
    }
    pl->stages = stages;
    pl->wpack_bytes = align256((size_t)pl->n_tiles * pl->k_blocks * n * BLOCK_K * sizeof(float));
    pl->smem_bytes = (size_t)stages * stage + FWD_EXTRA_SMEM;
    pl->ring_bytes = (uint32_t)((size_t)stages * stage);
    pl->halo = false;
    pl->vw = pl->ph = pl->sa = pl->sb = pl->b_resident = 0;
    pl->a_slot_bytes = 0;
    // Halo tiles: resident 3x3-class layers (any k > 1 at stride 1, BF16), see FwdParams.halo.
    // EXPERIMENTAL, opt-in (BCNN_B200_HALO=1): parity-green (tests/test_nhwc_bf16_gpu.py passes with
    // it), L2 -> SM traffic halves (ncu: 1.45 GB -> 0.74 GB on 3x3 64->64 @56, 0.23 GB with resident
    // weights), but the kernel is 25-30 % SLOWER than the per-tap route on every ResNet-50 3x3 shape
    // (0.20 vs 0.19 ms, 0.134 vs 0.103, 0.090 vs 0.072; gpurun r2w) whatever the ring depths: the
    // per-tap route was not bound by that traffic after all. Kept for the next round's profiling.
    // BCNN_B200_HALO: 1 every eligible layer, 0 none; default: layers whose weights stay resident in
    // shared memory (64 -> 64 channels), where a tile is one patch load and a straight run of MMAs
    static const int halo_mode = [] { const char *e = getenv("BCNN_B200_HALO"); return e ? atoi(e) : 2; }();
    if (halo_mode && g.resident && pl->bf16 && g.stride == 1 && g.o_s == 1 && g.ksh * g.ksw > 1 &&
        g.ksw - 1 + 8 <= 128 && !env_off("BCNN_B200_NO_HALO")) {
        int tw = g.dw;
        if (tw + g.ksw - 1 > 128) {   // wide planes: balanced column segments
            const int segs = ceil_div(g.dw, 128 - (g.ksw - 1));
            tw = ceil_div(g.dw, segs);
        }
        const int vw = tw + g.ksw - 1;
        int th = 128 / vw;
        if (th > g.dh) th = g.dh;
        if (th >= 1) {
            const int tiles_h = ceil_div(g.dh, th);
            th = ceil_div(g.dh, tiles_h);
        }
        // one image per tile: small planes (7 x 7) fill the 128 MMA rows better when the plain route
        // stacks several images into a tile
        const char *e_nmax = getenv("BCNN_B200_HALO_NMAX");
        const int halo_nmax = e_nmax ? atoi(e_nmax) : 256;
        if (th >= 1 && th * vw >= 96 && n <= halo_nmax) {
            const int ph = th + g.ksh - 1;
            const int reach = 128 + (g.ksh - 1) * vw + g.ksw - 1;   // rows the descriptors may touch
            const int rows = vw * ph > reach ? vw * ph : reach;
            const uint32_t a_slot = (uint32_t)((rows * 128 + 1023) & ~1023);
            const uint32_t b_slot = (uint32_t)n * 128u;
            // ring split: the activation patches need depth (a patch is one tile's worth of MMAs, its
            // load latency must hide behind the tiles in front of it), the weight tap tiles need at
            // least four slots; weights stay resident when everything still fits
            const long long avail = 227 * 1024 - FWD_EXTRA_SMEM;
            const int taps_all = g.ksh * g.ksw * pl->kc_blocks;
            const char *e_sa = getenv("BCNN_B200_HALO_SA"), *e_sb = getenv("BCNN_B200_HALO_SB");
            int sa = e_sa ? atoi(e_sa) : 3;
            bool resident_b = false;
            if (pl->n_tiles == 1 && !env_off("BCNN_B200_HALO_NO_RESIDENT_B")) {   // weights resident: 3, else 2 patches in flight
                if ((long long)sa * a_slot + (long long)taps_all * b_slot <= avail) resident_b = true;
                else if (!e_sa && 2LL * a_slot + (long long)taps_all * b_slot <= avail) { sa = 2; resident_b = true; }
            }
            int sb = resident_b ? taps_all : (int)((avail - (long long)sa * a_slot) / b_slot);
            if (!resident_b && sb > 8) sb = 8;
            if (e_sb && !resident_b) sb = atoi(e_sb);
            while (!resident_b && sb < 4 && sa > 2) {   // trade activation depth for weight depth
                --sa;
                sb = (int)((avail - (long long)sa * a_slot) / b_slot);
            }
            if ((long long)sa * a_slot + (long long)sb * b_slot > avail) sb = 0;
            // the box must fit the TMA limits and the rings must hold a few tap tiles
            if (sb >= 3 && vw <= 256 && ph <= 256 && 2 * (sa + sb) + 5 <= FWD_BAR_BYTES / 8 &&
                (halo_mode == 1 || resident_b)) {
                pl->halo = true;
                pl->vw = vw; pl->ph = ph; pl->sa = sa; pl->sb = sb; pl->b_resident = resident_b ? 1 : 0;
                pl->a_slot_bytes = a_slot;
                pl->tw = tw; pl->th = th; pl->tn = 1;
                pl->tiles_w = ceil_div(g.dw, tw); pl->tiles_h = ceil_div(g.dh, th); pl->tiles_b = g.batch;
                pl->a_bytes = (uint32_t)(vw * ph * 128);
                pl->tile_pos = tw * th;
                pl->ring_bytes = (uint32_t)(sa * a_slot + sb * b_slot);
                pl->smem_bytes = (size_t)pl->ring_bytes + FWD_EXTRA_SMEM;
                pl->out16 = (g.dst_w == g.dw && g.dst_c % 64 == 0 && !env_off("BCNN_B200_NO_TSTORE")) ? 1 : 2;
            }
        }
    }
    const long long total = (long long)pl->n_tiles * pl->tiles_w * pl->tiles_h * pl->tiles_b;
    if (total >= (1LL << 31)) return false;
    // fused statistics: grid = a multiple of n_tiles (0 rows: more channel tiles than SMs, no fusion)
    const long long per_tile = sm_count() / pl->n_tiles;
    const long long m_tiles = total / pl->n_tiles;
    pl->stat_rows = (pl->tstore || pl->out16) ? (int)(per_tile < m_tiles ? per_tile : m_tiles) : 0;
    pl->stat_bytes = align256((size_t)pl->stat_rows * 2 * g.dst_c * sizeof(float));
    return true;
}

int stat_grid(const FwdPlan &pl) { return pl.stat_rows * pl.n_tiles; }

// ---- how a (descriptor, pass) maps onto launches of the kernel
enum FwdRoute { ROUTE_NONE = 0, ROUTE_PLAIN, ROUTE_IM2COL, ROUTE_STRIDED_DGRAD };

// Thin first layers (Cin = 3): an explicit im2col buffer [n, ho, wo, Kp] (Kp = Cin * k * k rounded
// up to 8 elements) is an NHWC tensor with Kp channels; over it the convolution is 1x1.
int im2col_kp(const bcnn_b200_conv_desc *d) { return ceil_div(d->cin * d->ksize * d->ksize, 8) * 8; }
bool im2col_shape(const bcnn_b200_conv_desc *d) {
    static int off = -1;
    if (off < 0) off = env_off("BCNN_B200_NO_IM2COL") ? 1 : 0;
    return !off && d->groups == 1 && d->cin < 16 && d->cin * d->ksize * d->ksize >= 24 &&
           d->cout >= 16 && (long long)d->batch * d->ho * d->wo >= 4096;
}

FwdGeom geom_plain(const bcnn_b200_conv_desc *d, bool dgrad) {
    FwdGeom g;
    g.batch = d->batch;
    if (dgrad) {  // stride 1: correlation of dy with the flipped filter, pad' = k - 1 - pad
        g.src_c = d->cout; g.sh = d->ho; g.sw = d->wo;
        g.dst_c = d->cin; g.dh = d->h; g.dw = d->w;
        g.pad_h = g.pad_w = d->ksize - 1 - d->pad;
        g.stride = 1;
    } else {
        g.src_c = d->cin; g.sh = d->h; g.sw = d->w;
        g.dst_c = d->cout; g.dh = d->ho; g.dw = d->wo;
        g.pad_h = g.pad_w = d->pad;
        g.stride = d->stride;
    }
    g.ksh = g.ksw = d->ksize;
    g.o_s = 1; g.o_oy = g.o_ox = 0; g.dst_w = g.dw; g.dst_plane = g.dh * g.dw;
    g.direct = true;
    return g;
}

FwdGeom geom_im2col(const bcnn_b200_conv_desc *d) {
    FwdGeom g;
    g.batch = d->batch;
    g.src_c = im2col_kp(d); g.sh = d->ho; g.sw = d->wo;
    g.dst_c = d->cout; g.dh = d->ho; g.dw = d->wo;
    g.ksh = g.ksw = 1; g.pad_h = g.pad_w = 0; g.stride = 1;
    g.o_s = 1; g.o_oy = g.o_ox = 0; g.dst_w = g.dw; g.dst_plane = g.dh * g.dw;
    g.direct = false;
    return g;
}

// Strided dgrad, class (ph, pw) of input positions (h, w) = (s * i + ph, s * j + pw):
//   dx[h] = sum over kh with (ph + pad - kh) % s == 0 of W[kh] * dy[i + (ph + pad - kh) / s]
// i.e. a stride-1 correlation of dy with the sub-sampled filter. Along one axis the class has
// `n` taps; sub-tap t reads dy[i - lead + t] and uses filter tap kfirst - s * t.
struct ClassAxis { int n, lead, kfirst, extent; };
ClassAxis class_axis(int ph, int s, int pad, int k, int in_extent) {
    ClassAxis a;
    const int k0 = (ph + pad) % s;                 // smallest matching filter tap
    a.n = k0 < k ? (k - 1 - k0) / s + 1 : 0;
    const int k_last = k0 + s * (a.n - 1);         // largest matching filter tap: smallest offset
    const int q_min = (ph + pad - k_last) / s;     // exact division by construction
    a.lead = -q_min;
    a.kfirst = k_last;
    a.extent = ph < in_extent ? (in_extent - ph + s - 1) / s : 0;
    return a;
}

FwdGeom geom_dgrad_class(const bcnn_b200_conv_desc *d, const ClassAxis &ah, const ClassAxis &aw, int ph,
                         int pw) {
    FwdGeom g;
    g.batch = d->batch;
    g.src_c = d->cout; g.sh = d->ho; g.sw = d->wo;
    g.dst_c = d->cin; g.dh = ah.extent; g.dw = aw.extent;
    g.ksh = ah.n; g.ksw = aw.n; g.pad_h = ah.lead; g.pad_w = aw.lead; g.stride = 1;
    g.o_s = d->stride; g.o_oy = ph; g.o_ox = pw; g.dst_w = d->w; g.dst_plane = d->h * d->w;
    g.direct = false;
    return g;
}

bool strided_dgrad_shape(const bcnn_b200_conv_desc *d) {
    static int off = -1;
    if (off < 0) off = env_off("BCNN_B200_NO_STRIDED_DGRAD") ? 1 : 0;
    return !off && d->groups == 1 && d->stride > 1 && d->stride <= 4 && d->ksize <= 8 &&
           d->pad < d->ksize && d->cout % 4 == 0 && d->cout >= 16;
}

// Workspace bytes of the strided-dgrad route (0 = not applicable): dy shadow + all class packs.
size_t strided_dgrad_bytes(const bcnn_b200_conv_desc *d) {
    if (!strided_dgrad_shape(d)) return 0;
    size_t shadow = 0, packs = 0;
    for (int ph = 0; ph < d->stride; ++ph)
        for (int pw = 0; pw < d->stride; ++pw) {
            const ClassAxis ah = class_axis(ph, d->stride, d->pad, d->ksize, d->h);
            const ClassAxis aw = class_axis(pw, d->stride, d->pad, d->ksize, d->w);
            if (ah.n == 0 || aw.n == 0 || ah.extent == 0 || aw.extent == 0) continue;
            FwdPlan pl;
            if (!plan_fwd(geom_dgrad_class(d, ah, aw, ph, pw), &pl)) return 0;
            shadow = pl.shadow_bytes;
            packs += pl.wpack_bytes;
        }
    return shadow + packs;
}

FwdRoute route_fwd(const bcnn_b200_conv_desc *d, bool dgrad, FwdPlan *pl) {
    if (tma_disabled() || !encode_fn() || d->groups != 1) return ROUTE_NONE;
    if (dgrad) {
        if (d->stride != 1) return strided_dgrad_bytes(d) ? ROUTE_STRIDED_DGRAD : ROUTE_NONE;
        if (d->pad > d->ksize - 1) return ROUTE_NONE;
        return plan_fwd(geom_plain(d, true), pl) ? ROUTE_PLAIN : ROUTE_NONE;
    }
    if (im2col_shape(d)) return plan_fwd(geom_im2col(d), pl) ? ROUTE_IM2COL : ROUTE_NONE;
    return plan_fwd(geom_plain(d, false), pl) ? ROUTE_PLAIN : ROUTE_NONE;
}

template <bool NHWC, bool BF16>
int launch_fwd_kernel(const CUtensorMap &tm, const CUtensorMap &tm_dst, const FwdParams &p, size_t smem,
                      cudaStream_t st, int grid_override) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tma_fwd_kernel<NHWC, BF16>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    if (grid_override > 0) grid = grid_override;
    conv_tma_fwd_kernel<NHWC, BF16><<<grid, FWD_THREADS, smem, st>>>(tm, tm_dst, p);
    return launched();
}

// src: the NCHW tensor (DIRECT plans) or the NHWC-shaped shadow (all others).
int run_fwd(const FwdGeom &g, const FwdPlan &pl, const void *src, const uint8_t *wpack, const float *bias,
            int act, void *dst_any, int accumulate, cudaStream_t st, float *stat_partial = nullptr) {
    float *dst = reinterpret_cast<float *>(dst_any);
    CUtensorMap tm;
    if (pl.nhwc) {
        const int bw = pl.halo ? pl.vw : pl.tw, bh = pl.halo ? pl.ph : pl.th;
        if (!make_map_nhwc(&tm, src, g.src_c, g.sw, g.sh, g.batch, bw, bh, pl.tn, g.stride, false, pl.bf16))
            return (int)cudaErrorInvalidValue;
    } else if (!make_map_nchw(&tm, reinterpret_cast<const float *>(src), pl.view_w, pl.view_h, g.src_c,
                              g.batch, BLOCK_K, true)) {
        return (int)cudaErrorInvalidValue;
    }
    FwdParams p;
    p.dst = dst; p.bias = bias; p.wpack = wpack; p.act = act; p.accumulate = accumulate;
    p.stat_partial = stat_partial;
    p.src_c = g.src_c; p.dst_c = g.dst_c; p.batch = g.batch;
    p.out_w = pl.out_w; p.out_h = pl.out_h;
    p.ksh = g.ksh; p.ksw = g.ksw; p.pad_h = g.pad_h; p.pad_w = g.pad_w; p.stride = g.stride;
    p.o_s = g.o_s; p.o_oy = g.o_oy; p.o_ox = g.o_ox;
    p.dst_w = pl.nhwc ? g.dst_w : pl.out_w; p.dst_plane = g.dst_plane;
    p.kc_blocks = pl.kc_blocks; p.k_blocks = pl.k_blocks; p.n_tile = pl.n_tile; p.stages = pl.stages;
    p.n_tiles = pl.n_tiles;
    p.wc = pl.wc; p.rows = pl.rows; p.tw = pl.tw; p.th = pl.th; p.tn = pl.tn;
    p.a_bytes = pl.a_bytes;
    p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.tiles_b = pl.tiles_b;
    p.total_tiles = pl.n_tiles * pl.tiles_w * pl.tiles_h * pl.tiles_b;
    p.d_ntiles = FastDiv((uint32_t)pl.n_tiles);
    p.d_tiles_w = FastDiv((uint32_t)pl.tiles_w);
    p.d_tiles_h = FastDiv((uint32_t)pl.tiles_h);
    p.d_tw = FastDiv((uint32_t)pl.tw);
    p.d_th = FastDiv((uint32_t)pl.th);
    p.tstore = pl.tstore ? 1 : 0;
    p.out16 = pl.out16;
    p.narrow = (pl.out16 && pl.n_tile <= 64 && !env_off("BCNN_B200_NO_NARROW")) ? 1 : 0;
    p.halo = pl.halo ? 1 : 0;
    p.vw = pl.vw; p.sa = pl.sa; p.sb = pl.sb; p.b_resident = pl.b_resident;
    p.a_slot_bytes = pl.a_slot_bytes; p.ring_bytes = pl.ring_bytes;
    p.d_vw = FastDiv((uint32_t)(pl.vw > 0 ? pl.vw : 1));
    // fused statistics of BF16 results in registers (BCNN_B200_STATS_SMEM=1: the shared-memory transpose
    // of the FP32-tensor path); the transpose scratch then serves as second staging buffers
    static const bool stats_smem = env_off("BCNN_B200_STATS_SMEM"), one_buf = env_off("BCNN_B200_ONE_STAGE_BUF");
    p.stats_frag = (pl.out16 && !stats_smem) ? 1 : 0;
    p.stage_bufs = (pl.out16 == 1 && (p.stats_frag || stat_partial == nullptr) && !one_buf) ? 2 : 1;
    {
        const char *e2 = getenv("BCNN_B200_DBG_EPI");
        p.dbg = e2 ? atoi(e2) : 0;
    }
    p.tile_pos = pl.tile_pos;
    CUtensorMap tm_dst = tm;   // placeholder when the epilogue stores from registers
    if (pl.tstore &&
        !make_map_out(&tm_dst, dst, g.dst_plane, g.dst_c, g.batch, pl.tile_pos, pl.nhwc ? pl.tn : 1))
        return (int)cudaErrorInvalidValue;
    if (pl.out16 == 1 && !make_map_out16(&tm_dst, dst_any, g.dst_c, g.dst_w, g.dst_plane / g.dst_w, g.batch,
                                         pl.tw, pl.th, pl.tn))
        return (int)cudaErrorInvalidValue;
    const size_t smem = pl.smem_bytes;
    const int grid = stat_partial ? stat_grid(pl) : 0;
    if (!pl.nhwc) return launch_fwd_kernel<false, false>(tm, tm_dst, p, smem, st, grid);
    return pl.bf16 ? launch_fwd_kernel<true, true>(tm, tm_dst, p, smem, st, grid)
                   : launch_fwd_kernel<true, false>(tm, tm_dst, p, smem, st, grid);
}

// Packed images somebody else keeps up to date (bcnn_b200_conv_prepacked_*: the net runtime packs every
// layer's images in one launch at the start of a training step): (weights, fprop / dgrad) -> image.
struct Prepacked { const float *w; int dgrad; const uint8_t *image; };
constexpr int PREPACKED_MAX = 512;
Prepacked g_prepacked[PREPACKED_MAX];
int g_prepacked_n = 0;
bool g_prepacked_on = false;
const uint8_t *prepacked_lookup(const float *w, int dgrad) {
    if (!g_prepacked_on) return nullptr;
    for (int i = 0; i < g_prepacked_n; ++i)
        if (g_prepacked[i].w == w && g_prepacked[i].dgrad == dgrad) return g_prepacked[i].image;
    return nullptr;
}

int launch_pack(const float *w, uint8_t *wpack, bool dgrad, int cout, int cin, int kk, const FwdPlan &pl,
                const TapMap &taps, cudaStream_t st) {
    const size_t chunks = (size_t)pl.n_tiles * pl.n_tile * pl.k_blocks * 8;
    if (pl.bf16)
        pack_weights_tf32_kernel<true><<<stream_grid(chunks, 256), 256, 0, st>>>(
            w, wpack, dgrad ? 1 : 0, cout, cin, kk, pl.n_tile, pl.n_tiles, pl.kc_blocks, taps);
    else
        pack_weights_tf32_kernel<false><<<stream_grid(chunks, 256), 256, 0, st>>>(
            w, wpack, dgrad ? 1 : 0, cout, cin, kk, pl.n_tile, pl.n_tiles, pl.kc_blocks, taps);
    return launched();
}

int launch_im2col(const bcnn_b200_conv_desc *d, const float *x, void *col, bool bf16, cudaStream_t st);

// The dY shadow of a layer's backward pass: taken from `sh` when an earlier call of the same pass
// left one of the right format there, else transposed into sh->dy (kept for the next call) or, with no
// shadow storage, into `fallback` (the workspace).
int dy_shadow(const float *dy, int batch, int c, int plane, bool bf16, size_t bytes,
              bcnn_b200_conv_shadows *sh, void *fallback, const void **out, cudaStream_t st) {
    const int fmt = bf16 ? BCNN_B200_SHADOW_NHWC_BF16 : BCNN_B200_SHADOW_NHWC_F32;
    if (sh && sh->dy && sh->dy_fmt == fmt && sh->dy_bytes >= bytes) {
        *out = sh->dy;
        return 0;
    }
    void *dst = fallback;
    if (sh && sh->dy && sh->dy_bytes >= bytes && (reinterpret_cast<uintptr_t>(sh->dy) & 255) == 0) {
        dst = sh->dy;
        sh->dy_fmt = fmt;
    }
    *out = dst;
    return launch_transpose(dy, dst, batch, c, plane, bf16, st);
}

int launch_strided_dgrad(const bcnn_b200_conv_desc *d, const float *dy, const float *w, float *dx,
                         int accumulate, void *workspace, size_t workspace_bytes,
                         bcnn_b200_conv_shadows *sh, cudaStream_t st) {
    const size_t need = strided_dgrad_bytes(d);
    if (!need || workspace == nullptr || workspace_bytes < need) return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return (int)cudaErrorMisalignedAddress;
    const bool bf16 = shadow_bf16() && d->cout % 8 == 0;   // what plan_fwd decides for every class
    const void *shadow = nullptr;
    int err = dy_shadow(dy, d->batch, d->cout, d->ho * d->wo, bf16,
                        align256((size_t)d->batch * d->cout * d->ho * d->wo * (bf16 ? 2 : 4)), sh,
                        workspace, &shadow, st);
    if (err) return err;
    const int s = d->stride, kk = d->ksize * d->ksize;
    // classes without taps (e.g. 1x1 stride 2: three of four) receive no gradient
    bool holes = false;
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw)
            if (class_axis(ph, s, d->pad, d->ksize, d->h).n == 0 ||
                class_axis(pw, s, d->pad, d->ksize, d->w).n == 0)
                holes = true;
    if (holes && !accumulate) {
        cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)d->batch * d->cin * d->h * d->w * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    size_t off = 0;
    bool have_shadow_bytes = false;
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw) {
            const ClassAxis ah = class_axis(ph, s, d->pad, d->ksize, d->h);
            const ClassAxis aw = class_axis(pw, s, d->pad, d->ksize, d->w);
            if (ah.n == 0 || aw.n == 0 || ah.extent == 0 || aw.extent == 0) continue;
            const FwdGeom g = geom_dgrad_class(d, ah, aw, ph, pw);
            FwdPlan pl;
            if (!plan_fwd(g, &pl)) return (int)cudaErrorInvalidValue;
            if (!have_shadow_bytes) { off = pl.shadow_bytes; have_shadow_bytes = true; }
            uint8_t *wpack = reinterpret_cast<uint8_t *>(workspace) + off;
            off += pl.wpack_bytes;
            TapMap taps;
            taps.n = ah.n * aw.n;
            for (int th = 0; th < ah.n; ++th)
                for (int tw = 0; tw < aw.n; ++tw)
                    taps.idx[th * aw.n + tw] = (short)((ah.kfirst - s * th) * d->ksize + (aw.kfirst - s * tw));
            err = launch_pack(w, wpack, true, d->cout, d->cin, kk, pl, taps, st);
            if (err) return err;
            err = run_fwd(g, pl, shadow, wpack, nullptr, 0, dx, accumulate, st);
            if (err) return err;
        }
    return 0;
}

int launch_fwd(const bcnn_b200_conv_desc *d, bool dgrad, const float *src, const float *w,
               const float *bias, int act, float *dst, int accumulate, void *workspace,
               size_t workspace_bytes, bcnn_b200_conv_shadows *sh, cudaStream_t st,
               const float **stat_partial = nullptr, int *stat_rows = nullptr) {
    FwdPlan pl;
    const FwdRoute route = route_fwd(d, dgrad, &pl);
    if (route == ROUTE_NONE) return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(src) & 15) != 0) return (int)cudaErrorMisalignedAddress;
    if (route == ROUTE_STRIDED_DGRAD)
        return launch_strided_dgrad(d, src, w, dst, accumulate, workspace, workspace_bytes, sh, st);
    if (workspace == nullptr || workspace_bytes < pl.shadow_bytes + pl.wpack_bytes)
        return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return (int)cudaErrorMisalignedAddress;
    // fused batch-norm partials live behind the packed weights when the workspace has the room
    float *stats = nullptr;
    if (stat_partial && !dgrad && !accumulate && pl.stat_rows > 0 &&
        workspace_bytes >= pl.shadow_bytes + pl.wpack_bytes + pl.stat_bytes) {
        stats = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(workspace) + pl.shadow_bytes +
                                          pl.wpack_bytes);
        *stat_partial = stats;
        *stat_rows = pl.stat_rows;
    }
    void *shadow = workspace;
    // fprop: the x shadow goes to the layer's own storage when there is one (wgrad reuses it)
    const bool keep_x = !dgrad && sh && sh->x && sh->x_bytes >= pl.shadow_bytes &&
                        (reinterpret_cast<uintptr_t>(sh->x) & 255) == 0;
    if (keep_x) shadow = sh->x;
    uint8_t *wpack = reinterpret_cast<uint8_t *>(workspace) + pl.shadow_bytes;
    const int kk = d->ksize * d->ksize;
    TapMap taps;
    int err;
    if (route == ROUTE_IM2COL) {
        // W [cout][cin * kk] is already the K-major matrix of the 1x1 problem
        taps.n = 1; taps.idx[0] = 0;
        err = launch_pack(w, wpack, false, d->cout, d->cin * kk, 1, pl, taps, st);
        if (err) return err;
        err = launch_im2col(d, src, shadow, pl.bf16, st);
        if (err) return err;
        if (keep_x) sh->x_fmt = pl.bf16 ? BCNN_B200_SHADOW_IM2COL_BF16 : BCNN_B200_SHADOW_IM2COL_F32;
        return run_fwd(geom_im2col(d), pl, shadow, wpack, bias, act, dst, accumulate, st, stats);
    }
    taps.n = kk;
    for (int t = 0; t < kk; ++t) taps.idx[t] = (short)(dgrad ? kk - 1 - t : t);
    err = launch_pack(w, wpack, dgrad, d->cout, d->cin, kk, pl, taps, st);
    if (err) return err;
    const FwdGeom g = geom_plain(d, dgrad);
    const void *operand = src;
    if (pl.nhwc && dgrad) {
        err = dy_shadow(src, g.batch, g.src_c, g.sh * g.sw, pl.bf16, pl.shadow_bytes, sh, workspace,
                        &operand, st);
        if (err) return err;
    } else if (pl.nhwc) {
        err = launch_transpose(src, shadow, g.batch, g.src_c, g.sh * g.sw, pl.bf16, st);
        if (err) return err;
        if (keep_x) sh->x_fmt = pl.bf16 ? BCNN_B200_SHADOW_NHWC_BF16 : BCNN_B200_SHADOW_NHWC_F32;
        operand = shadow;
    }
    return run_fwd(g, pl, operand, wpack, bias, act, dst, accumulate, st, stats);
}


// ------------------------------------------------------------------ resident (BF16 NHWC) fprop / dgrad
// Source and result are BF16 NHWC tensors: no shadow is made, the TMA reads the activation tensor
// itself and the epilogue writes the result once in the layout the next layer's loads want.
FwdGeom geom_resident(const bcnn_b200_conv_desc *d, bool dgrad) {
    FwdGeom g = geom_plain(d, dgrad);
    g.direct = false;
    g.resident = true;
    const long long positions = (long long)d->batch * d->h * d->w;
    if (d->ksize == 1 && d->stride == 1 && d->pad == 0 && positions < (1LL << 31)) {
        // 1x1: the batch is one row of N*H*W positions, so every tile has 128 full rows whatever
        // the plane size (7x7 planes would fill 98 of 128 otherwise)
        g.batch = 1;
        g.sh = g.dh = 1;
        g.sw = g.dw = g.dst_w = g.dst_plane = (int)positions;
    }
    return g;
}

bool resident_desc_ok(const bcnn_b200_conv_desc *d) {
    return !tma_disabled() && encode_fn() && d->groups == 1 && d->cout % 8 == 0 && d->ksize <= 8 &&
           (d->cin % 8 == 0 || im2col_shape(d));
}

FwdRoute route_fwd_resident(const bcnn_b200_conv_desc *d, bool dgrad, FwdPlan *pl) {
    if (!resident_desc_ok(d)) return ROUTE_NONE;
    if (dgrad) {
        if (d->cin % 8 != 0) return ROUTE_NONE;
        if (d->stride != 1) {
            if (!strided_dgrad_shape(d)) return ROUTE_NONE;
            for (int ph = 0; ph < d->stride; ++ph)
                for (int pw = 0; pw < d->stride; ++pw) {
                    const ClassAxis ah = class_axis(ph, d->stride, d->pad, d->ksize, d->h);
                    const ClassAxis aw = class_axis(pw, d->stride, d->pad, d->ksize, d->w);
                    if (ah.n == 0 || aw.n == 0 || ah.extent == 0 || aw.extent == 0) continue;
                    FwdGeom g = geom_dgrad_class(d, ah, aw, ph, pw);
                    g.resident = true;
                    if (!plan_fwd(g, pl)) return ROUTE_NONE;
                }
            return ROUTE_STRIDED_DGRAD;
        }
        if (d->pad > d->ksize - 1) return ROUTE_NONE;
        return plan_fwd(geom_resident(d, true), pl) ? ROUTE_PLAIN : ROUTE_NONE;
    }
    if (im2col_shape(d)) {
        FwdGeom g = geom_im2col(d);
        g.resident = true;
        return plan_fwd(g, pl) ? ROUTE_IM2COL : ROUTE_NONE;
    }
    return plan_fwd(geom_resident(d, false), pl) ? ROUTE_PLAIN : ROUTE_NONE;
}

size_t im2col_bytes(const bcnn_b200_conv_desc *d) {
    return align256((size_t)d->batch * d->ho * d->wo * im2col_kp(d) * 2);
}

size_t strided_dgrad_pack_bytes_resident(const bcnn_b200_conv_desc *d) {
    size_t packs = 0;
    for (int ph = 0; ph < d->stride; ++ph)
        for (int pw = 0; pw < d->stride; ++pw) {
            const ClassAxis ah = class_axis(ph, d->stride, d->pad, d->ksize, d->h);
            const ClassAxis aw = class_axis(pw, d->stride, d->pad, d->ksize, d->w);
            if (ah.n == 0 || aw.n == 0 || ah.extent == 0 || aw.extent == 0) continue;
            FwdGeom g = geom_dgrad_class(d, ah, aw, ph, pw);
            g.resident = true;
            FwdPlan pl;
            if (!plan_fwd(g, &pl)) return 0;
            packs += pl.wpack_bytes;
        }
    return packs;
}

// filter taps of one class of a strided dgrad, in the kernel's walk order
void class_taps(const bcnn_b200_conv_desc *d, const ClassAxis &ah, const ClassAxis &aw, TapMap *taps) {
    const int s = d->stride;
    taps->n = ah.n * aw.n;
    for (int th = 0; th < ah.n; ++th)
        for (int tw = 0; tw < aw.n; ++tw)
            taps->idx[th * aw.n + tw] = (short)((ah.kfirst - s * th) * d->ksize + (aw.kfirst - s * tw));
}

// The packed images the resident fprop (dgrad == 0) / dgrad (1) of `d` consume, as jobs for
// pack_jobs_kernel, in the order and at the offsets launch_fwd_resident / launch_dgrad_resident read
// them from a prepacked base. Returns the number of jobs (-1: the route does not exist, -2: more than
// max_jobs); *bytes = size of the images.
int plan_pack_jobs(const bcnn_b200_conv_desc *d, int dgrad, const float *w, uint8_t *dst, PackJob *jobs,
                   int max_jobs, size_t *bytes) {
    FwdPlan pl;
    const FwdRoute route = route_fwd_resident(d, dgrad != 0, &pl);
    if (route == ROUTE_NONE) return -1;
    const int kk = d->ksize * d->ksize;
    int n = 0;
    size_t off = 0;
    auto add = [&](const FwdPlan &p, int cin, int kkk, const TapMap &taps) {
        if (n >= max_jobs) return false;
        if (!p.bf16) return false;
        PackJob &j = jobs[n++];
        j.w = w; j.dst = dst ? dst + off : nullptr; j.dgrad = dgrad ? 1 : 0; j.cout = d->cout; j.cin = cin; j.kk = kkk;
        j.n_tile = p.n_tile; j.n_tiles = p.n_tiles; j.kc_blocks = p.kc_blocks; j.first_block = 0; j.taps = taps;
        off += p.wpack_bytes;
        return true;
    };
    TapMap taps;
    if (route == ROUTE_IM2COL) {
        taps.n = 1; taps.idx[0] = 0;
        if (!add(pl, d->cin * kk, 1, taps)) return -2;
    } else if (route == ROUTE_PLAIN) {
        taps.n = kk;
        for (int t = 0; t < kk; ++t) taps.idx[t] = (short)(dgrad ? kk - 1 - t : t);
        if (!add(pl, d->cin, kk, taps)) return -2;
    } else {   // strided dgrad: one image per class of input positions
        const int s = d->stride;
        for (int ph = 0; ph < s; ++ph)
            for (int pw = 0; pw < s; ++pw) {
                const ClassAxis ah = class_axis(ph, s, d->pad, d->ksize, d->h);
                const ClassAxis aw = class_axis(pw, s, d->pad, d->ksize, d->w);
                if (ah.n == 0 || aw.n == 0 || ah.extent == 0 || aw.extent == 0) continue;
                FwdGeom g = geom_dgrad_class(d, ah, aw, ph, pw);
                g.resident = true;
                if (!plan_fwd(g, &pl)) return -1;
                class_taps(d, ah, aw, &taps);
                if (!add(pl, d->cin, kk, taps)) return -2;
            }
    }
    *bytes = off;
    return n;
}

// fprop: x is the BF16 NHWC activation, or the FP32 NCHW input of a thin first layer (im2col route)
int launch_fwd_resident(const bcnn_b200_conv_desc *d, const void *x, const float *w, const float *bias,
                        int act, void *y16, void *workspace, size_t workspace_bytes,
                        bcnn_b200_conv_shadows *sh, cudaStream_t st, const float **stat_partial,
                        int *stat_rows) {
    FwdPlan pl;
    const FwdRoute route = route_fwd_resident(d, false, &pl);
    if (route == ROUTE_NONE) return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(y16) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
        return (int)cudaErrorMisalignedAddress;
    uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
    size_t off = 0;
    const void *operand = x;
    const int kk = d->ksize * d->ksize;
    TapMap taps;
    if (route == ROUTE_IM2COL) {
        const size_t cb = im2col_bytes(d);
        void *col = ws;
        const bool keep = sh && sh->x && sh->x_bytes >= cb && (reinterpret_cast<uintptr_t>(sh->x) & 255) == 0;
        if (keep) col = sh->x;
        else off += cb;
        if (workspace == nullptr || workspace_bytes < off + pl.wpack_bytes) return (int)cudaErrorInvalidValue;
        int err = launch_im2col(d, reinterpret_cast<const float *>(x), col, true, st);
        if (err) return err;
        if (keep) sh->x_fmt = BCNN_B200_SHADOW_IM2COL_BF16;
        operand = col;
        taps.n = 1; taps.idx[0] = 0;
    } else {
        taps.n = kk;
        for (int t = 0; t < kk; ++t) taps.idx[t] = (short)t;
    }
    if (workspace == nullptr || workspace_bytes < off + pl.wpack_bytes) return (int)cudaErrorInvalidValue;
    uint8_t *wpack = ws + off;
    off += pl.wpack_bytes;
    float *stats = nullptr;
    if (stat_partial && pl.stat_rows > 0 && workspace_bytes >= off + pl.stat_bytes) {
        stats = reinterpret_cast<float *>(ws + off);
        *stat_partial = stats;
        *stat_rows = pl.stat_rows;
    }
    const uint8_t *image = prepacked_lookup(w, 0);
    if (!image) {
        int err = route == ROUTE_IM2COL ? launch_pack(w, wpack, false, d->cout, d->cin * kk, 1, pl, taps, st)
                                        : launch_pack(w, wpack, false, d->cout, d->cin, kk, pl, taps, st);
        if (err) return err;
        image = wpack;
    }
    FwdGeom g = route == ROUTE_IM2COL ? geom_im2col(d) : geom_resident(d, false);
    g.resident = true;
    return run_fwd(g, pl, operand, image, bias, act, y16, 0, st, stats);
}

int launch_dgrad_resident(const bcnn_b200_conv_desc *d, const float *w, const void *dy16, void *dx16,
                          int accumulate, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    FwdPlan pl;
    const FwdRoute route = route_fwd_resident(d, true, &pl);
    if (route == ROUTE_NONE) return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(dy16) & 15) != 0 || (reinterpret_cast<uintptr_t>(dx16) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
        return (int)cudaErrorMisalignedAddress;
    uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
    const int kk = d->ksize * d->ksize;
    if (route == ROUTE_PLAIN) {
        if (workspace == nullptr || workspace_bytes < pl.wpack_bytes) return (int)cudaErrorInvalidValue;
        TapMap taps;
        taps.n = kk;
        for (int t = 0; t < kk; ++t) taps.idx[t] = (short)(kk - 1 - t);
        const uint8_t *image = prepacked_lookup(w, 1);
        if (!image) {
            int err = launch_pack(w, ws, true, d->cout, d->cin, kk, pl, taps, st);
            if (err) return err;
            image = ws;
        }
        return run_fwd(geom_resident(d, true), pl, dy16, image, nullptr, 0, dx16, accumulate, st);
    }
    // strided: one stride-1 launch per class of input positions, scattered into dx
    if (workspace == nullptr || workspace_bytes < strided_dgrad_pack_bytes_resident(d))
        return (int)cudaErrorInvalidValue;
    const int s = d->stride;
    bool holes = false;
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw)
            if (class_axis(ph, s, d->pad, d->ksize, d->h).n == 0 || class_axis(pw, s, d->pad, d->ksize, d->w).n == 0)
                holes = true;
    if (holes && !accumulate) {
        cudaError_t e = cudaMemsetAsync(dx16, 0, (size_t)d->batch * d->cin * d->h * d->w * 2, st);
        if (e != cudaSuccess) return (int)e;
    }
    size_t off = 0;
    const uint8_t *pre = prepacked_lookup(w, 1);
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw) {
            const ClassAxis ah = class_axis(ph, s, d->pad, d->ksize, d->h);
            const ClassAxis aw = class_axis(pw, s, d->pad, d->ksize, d->w);
            if (ah.n == 0 || aw.n == 0 || ah.extent == 0 || aw.extent == 0) continue;
            FwdGeom g = geom_dgrad_class(d, ah, aw, ph, pw);
            g.resident = true;
            if (!plan_fwd(g, &pl)) return (int)cudaErrorInvalidValue;
            uint8_t *wpack = ws + off;
            const uint8_t *image = pre ? pre + off : nullptr;   // the classes' images in launch order
            off += pl.wpack_bytes;
            int err;
            if (!image) {
                TapMap taps;
                class_taps(d, ah, aw, &taps);
                err = launch_pack(w, wpack, true, d->cout, d->cin, kk, pl, taps, st);
                if (err) return err;
                image = wpack;
            }
            err = run_fwd(g, pl, dy16, image, nullptr, 0, dx16, accumulate, st);
            if (err) return err;
        }
    return 0;
}

// ------------------------------------------------------------------ wgrad
struct WgParams {
    float *out;  // split-K partial slabs [split][cout][cin][kk], or gw itself when splits == 1
    int cin, cout, kk, ks, pad, stride;
    int n_tile;
    int bw, bh;            // k-block = bw x bh output positions (DIRECT: 32 x 1)
    int kpos;              // bw * bh
    int nb;                // NHWC: 32-channel atoms of the B operand
    uint32_t stage_bytes, a_bytes, atom_bytes;
    int blocks_w, blocks_h;   // k-blocks per image = blocks_h * blocks_w
    int kb_total, kb_per_split, splits, stages;
    size_t split_stride;
    FastDiv d_img, d_bw;
};

template <bool NHWC, bool BF16>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_tma_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                      const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int S = p.stages;
    const int n_tile = p.n_tile;
    const uint32_t stage_bytes = p.stage_bytes;
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

    // running ring slot / phase / block coordinates instead of divisions per k-block (see the fprop kernel)
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t smem0 = smem_u32(smem), bar0 = smem_u32(bars);
            uint32_t s = 0, ph = 1;
            uint32_t img, rem, hb, wb;
            p.d_img.divmod((uint32_t)kb_begin, img, rem);
            p.d_bw.divmod(rem, hb, wb);
            for (int it = 0; it < iters; ++it) {
                mbar_wait(bar0 + 8 * ((uint32_t)S + s), ph);
                const uint32_t full = bar0 + 8 * s;
                const uint32_t a_stage = smem0 + s * stage_bytes;
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                mbar_expect_tx(full, stage_bytes);
                if (NHWC) {
                    const int ow0 = (int)wb * p.bw, oh0 = (int)hb * p.bh;
                    constexpr int CA = BF16 ? 64 : 32;   // channels per atom (128-byte rows)
                    for (int a = 0; a < TILE_M / CA; ++a)
                        tma_load_4d(a_stage + a * p.atom_bytes, &tm_dy, co0 + CA * a, ow0, oh0, (int)img, full);
                    for (int b = 0; b < p.nb; ++b)
                        tma_load_4d(a_stage + p.a_bytes + b * p.atom_bytes, &tm_x, ci0 + CA * b,
                                    ow0 * p.stride + kw - p.pad, oh0 * p.stride + kh - p.pad, (int)img, full);
                } else {
                    tma_load_4d(a_stage, &tm_dy, (int)wb * 32, (int)hb, co0, (int)img, full);
                    tma_load_4d(a_stage + p.a_bytes, &tm_x, (int)wb * 32 + kw - p.pad, (int)hb + kh - p.pad, ci0,
                                (int)img, full);
                }
                // next k-block: column block fastest, then row block, then image
                if (++wb == (uint32_t)p.blocks_w) {
                    wb = 0;
                    if (++hb == (uint32_t)p.blocks_h) { hb = 0; ++img; }
                }
            }
        }
    } else if (warp == 1) {
        // the whole warp walks the loop (uniform state), one elected lane issues
        const uint32_t idesc = BF16 ? make_idesc_bf16(TILE_M, n_tile, 1, 1)
                                    : make_idesc_tf32(TILE_M, n_tile, NHWC ? 1 : 0, NHWC ? 1 : 0);
        const int mmas = p.kpos / (BF16 ? 16 : UMMA_K);
        const uint32_t smem0 = smem_u32(smem), bar0 = smem_u32(bars);
        // descriptors without their address: MN-major atoms (BF16 / TF32) or K-major rows (DIRECT)
        const uint64_t desc0 = BF16 ? make_desc_mn_sw128_b16(0, p.atom_bytes, 1024)
                                    : (NHWC ? make_desc_mn_sw128(0, p.atom_bytes, 512) : make_desc_sw128(0));
        const uint32_t g_step = BF16 ? (2048u >> 4) : (NHWC ? (1024u >> 4) : 2u);   // per MMA, in 16-byte units
        uint32_t s = 0, ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(bar0 + 8 * s, ph);
            tc_fence_after();
            const uint32_t a_addr = smem0 + s * stage_bytes;
            uint64_t da = desc0 | (uint64_t)((a_addr & 0x3FFFF) >> 4);
            uint64_t db = desc0 | (uint64_t)(((a_addr + p.a_bytes) & 0x3FFFF) >> 4);
            if (elect_one()) {
                for (int g = 0; g < mmas; ++g, da += g_step, db += g_step) {
                    if (BF16) umma_bf16(tmem_base, da, db, idesc, (uint32_t)(it | g));
                    else umma_tf32(tmem_base, da, db, idesc, (uint32_t)(it | g));
                }
                umma_commit(bar0 + 8 * ((uint32_t)S + s));
                if (it == iters - 1) umma_commit(bar0 + 8 * (2u * (uint32_t)S));
            }
            __syncwarp();
            if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
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
        // Every 32 x 32 block (lane = co row, register = ci) goes through this warp's shared-memory
        // scratch (the operand ring is idle by now) and leaves transposed: a store instruction then
        // covers 32 consecutive ci of ONE co row (one 128-byte line for 1x1 layers, kk lines otherwise)
        // instead of 32 rows -- the per-thread row stores spent 32 LSU line cycles per instruction,
        // as long as the main loop of a small-K layer.
        float *wscr = reinterpret_cast<float *>(smem) + (size_t)q * STAT_WARP_FLOATS;
        (void)co;
        for (int ck = 0; ck < chunks32; ++ck) {
            uint32_t v[32];
            if (iters > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ck * 32), v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            float4 *row = reinterpret_cast<float4 *>(wscr + lane * STAT_PITCH);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                row[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                     __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            __syncwarp();
            const int ci = ci0 + ck * 32 + lane;
            if (ck * 32 + lane < n_tile && ci < p.cin) {
                float *d = out + ((size_t)(co0 + q * 32) * p.cin + ci) * p.kk + tap;
                const size_t row_stride = (size_t)p.cin * p.kk;
                const int nrows = min(32, p.cout - (co0 + q * 32));
                if (p.splits == 1) {
#pragma unroll 8
                    for (int r = 0; r < nrows; ++r) d[(size_t)r * row_stride] += wscr[r * STAT_PITCH + lane];
                } else {
#pragma unroll 8
                    for (int r = 0; r < nrows; ++r) d[(size_t)r * row_stride] = wscr[r * STAT_PITCH + lane];
                }
            }
            __syncwarp();
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
    bool nhwc, bf16;
    int n_tile, ci_tiles, co_tiles, nb;
    int view_w, view_h, out_w, out_h;   // DIRECT views
    int bw, bh, blocks_w, blocks_h;
    uint32_t atom_bytes, a_bytes, stage_bytes;
    int kb_total, splits, kb_per_split, stages;
    size_t smem_bytes, partial_bytes, shadow_x_bytes, shadow_dy_bytes;
};

// The wgrad problem the kernel sees. Thin first layers go through the im2col buffer: a 1x1
// problem whose "input channels" are the Cin * k * k patch elements (cin = logical count and
// leading dimension of gw, cin_phys = channels of the buffer, a multiple of 4).
struct WgEff {
    int batch, cin, cin_phys, h, w, cout, ho, wo, ksize, stride, pad;
    bool im2col;
};
// resident: x and dy are BF16 NHWC tensors (or the kept im2col buffer); 1x1 problems are flattened
// to one row of N*H*W positions, so a k-block is 64 consecutive positions whatever the plane size
WgEff wg_effective(const bcnn_b200_conv_desc *d, bool resident = false) {
    WgEff e;
    e.batch = d->batch; e.cout = d->cout; e.ho = d->ho; e.wo = d->wo;
    e.im2col = im2col_shape(d);
    if (e.im2col) {
        e.cin = d->cin * d->ksize * d->ksize; e.cin_phys = im2col_kp(d);
        e.h = d->ho; e.w = d->wo; e.ksize = 1; e.stride = 1; e.pad = 0;
    } else {
        e.cin = e.cin_phys = d->cin; e.h = d->h; e.w = d->w;
        e.ksize = d->ksize; e.stride = d->stride; e.pad = d->pad;
    }
    const long long positions = (long long)e.batch * e.ho * e.wo;
    if (resident && e.ksize == 1 && e.stride == 1 && e.pad == 0 && positions < (1LL << 31)) {
        e.batch = 1;
        e.h = e.ho = 1;
        e.w = e.wo = (int)positions;
    }
    return e;
}

bool plan_wgrad(const bcnn_b200_conv_desc *desc, WgPlan *pl, bool resident = false) {
    if (tma_disabled() || !encode_fn()) return false;
    if (desc->groups != 1) return false;
    const WgEff e = wg_effective(desc, resident);
    const WgEff *d = &e;
    if (d->cin < 16 || d->cout < 32) return false;
    if ((long long)d->batch * d->ho * d->wo < 512) return false;  // tiny reductions (fc-shaped)
    if (resident && (d->cin_phys % 8 != 0 || d->cout % 8 != 0)) return false;
    const bool direct = !resident && !d->im2col && d->ksize == 1 && d->stride == 1 && d->pad == 0 &&
                        (d->h * d->w) % 4 == 0;
    pl->nhwc = !direct;
    if (!direct && (nhwc_disabled() || d->cin_phys % 4 != 0 || d->cout % 4 != 0 || d->stride > 4))
        return false;
    // N tile (input channels): up to 128 columns leave room for two co-resident CTAs per SM; layers
    // on small planes with many channels are bound by L2 -> SM operand traffic instead (every dY
    // tile is re-read per N tile, every X tile per M tile), so they take 256-column tiles
    static int wg_nmax = 0;
    if (!wg_nmax) {
        const char *e = getenv("BCNN_B200_WG_NMAX");
        wg_nmax = e ? atoi(e) : 256;
        if (wg_nmax != 128 && wg_nmax != 256) wg_nmax = 256;
    }
    const int nmax = d->cin >= 256 ? wg_nmax : 128;
    int n = d->cin;
    if (n > nmax) {
        int tiles = ceil_div(n, nmax);
        n = ceil_div(ceil_div(n, tiles), 16) * 16;
    } else {
        n = ceil_div(n, 16) * 16;
    }
    pl->n_tile = n;
    pl->ci_tiles = ceil_div(d->cin, n);
    pl->co_tiles = ceil_div(d->cout, TILE_M);
    pl->bf16 = !direct && (resident || shadow_bf16()) && d->cin_phys % 8 == 0 && d->cout % 8 == 0;
    pl->nb = ceil_div(n, pl->bf16 ? 64 : 32);
    if (direct) {
        pl->view_w = d->h * d->w; pl->view_h = 1;
        pl->out_w = d->ho * d->wo; pl->out_h = 1;
        pl->bw = 32; pl->bh = 1;
        pl->blocks_w = ceil_div(pl->out_w, 32); pl->blocks_h = 1;
        pl->atom_bytes = 0;
        pl->a_bytes = A_STAGE_BYTES;
        pl->stage_bytes = A_STAGE_BYTES + (uint32_t)n * BLOCK_K * 4;
        pl->shadow_x_bytes = pl->shadow_dy_bytes = 0;
    } else {
        // k-block = bw x bh output positions: multiple of 8, at most 64; columns past the row end
        // are out of bounds in dY and read as zero
        // (bf16: K = 16 positions per MMA, so bw * bh is kept a multiple of 16; rows past the
        // image are out of bounds too and read as zero)
        int bw = d->wo >= 64 ? 64 : ceil_div(d->wo, 8) * 8;
        int bh = 64 / bw;
        if (bh < 1) bh = 1;
        if (bh > d->ho) bh = d->ho;
        while (bw * d->stride > 256) bw -= 8;
        if (pl->bf16 && (bw * bh) % 16 != 0) ++bh;
        pl->bw = bw; pl->bh = bh;
        pl->blocks_w = ceil_div(d->wo, bw);
        pl->blocks_h = ceil_div(d->ho, bh);
        pl->atom_bytes = (uint32_t)(bw * bh * 128);
        const int a_atoms = pl->bf16 ? 2 : 4;
        const size_t es = pl->bf16 ? 2 : 4;
        pl->a_bytes = a_atoms * pl->atom_bytes;
        pl->stage_bytes = (uint32_t)(a_atoms + pl->nb) * pl->atom_bytes;
        pl->shadow_x_bytes = align256((size_t)d->batch * d->cin_phys * d->h * d->w * es);
        pl->shadow_dy_bytes = align256((size_t)d->batch * d->cout * d->ho * d->wo * es);
        if (resident) {   // the tensors themselves are the operands; only a thin first layer's
            pl->shadow_dy_bytes = 0;   // im2col buffer may have to be rebuilt
            if (!d->im2col) pl->shadow_x_bytes = 0;
        }
    }
    const long long kb_total = (long long)d->batch * pl->blocks_h * pl->blocks_w;
    if (kb_total >= (1LL << 31)) return false;
    pl->kb_total = (int)kb_total;
    // Shared memory: two co-resident CTAs per SM (one's epilogue under the other's main loop)
    // when three stages fit in half the SM, else one CTA per SM.
    int ctas_per_sm = 1;
    int stages = (int)((100u * 1024u) / pl->stage_bytes);
    if (stages >= 3) {
        ctas_per_sm = 2;
        if (stages > 4) stages = 4;
    } else {
        stages = (int)((200u * 1024u) / pl->stage_bytes);
        if (stages > 6) stages = 6;
        if (stages < 2) return false;
    }
    pl->stages = stages;
    pl->smem_bytes = (size_t)stages * pl->stage_bytes + 1024 + 256;
    // Split-K so that the whole grid is ONE resident wave (a partial extra wave costs a full
    // CTA lifetime): floor, never above ctas_per_sm * SMs unless the tile count alone exceeds it.
    const int kk = d->ksize * d->ksize;
    const long long tiles = (long long)pl->ci_tiles * pl->co_tiles * kk;
    long long want = ((long long)ctas_per_sm * sm_count()) / tiles;
    long long max_by_k = pl->kb_total / 16;  // >= 16 k-blocks per split
    if (want > max_by_k) want = max_by_k;
    if (want > 256) want = 256;
    if (want < 1) want = 1;
    pl->kb_per_split = ceil_div(pl->kb_total, (int)want);
    pl->splits = ceil_div(pl->kb_total, pl->kb_per_split);
    const size_t wsize = (size_t)d->cout * d->cin * kk;
    pl->partial_bytes = pl->splits > 1 ? align256((size_t)pl->splits * wsize * sizeof(float)) : 0;
    return true;
}

template <bool NHWC, bool BF16>
int launch_wgrad_kernel(const CUtensorMap &tm_dy, const CUtensorMap &tm_x, const WgParams &p, dim3 grid,
                        size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tma_wgrad_kernel<NHWC, BF16>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    conv_tma_wgrad_kernel<NHWC, BF16><<<grid, WG_THREADS, smem, st>>>(tm_dy, tm_x, p);
    return launched();
}

// ------------------------------------------------------------------ im2col for thin first layers
// col[n][oh][ow][k], k = (ci * ks + kh) * ks + kw fastest (the order of W[co][ci][kh][kw]), padded
// with zeros to kp = 4 * kp4 columns. One thread writes one float4; a warp covers 512 contiguous
// bytes of the buffer and gathers from a few image rows that stay in L1.
template <bool BF16>
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const float *__restrict__ x, void *__restrict__ col, int cin, int h, int w, int ks,
                   int stride, int pad, int kdim, uint32_t total4, FastDiv d_kp4, FastDiv d_wo,
                   FastDiv d_ho, FastDiv d_ks, FastDiv d_kk) {
    constexpr int EPT = BF16 ? 8 : 4;   // elements per thread = one 16-byte store
    const uint32_t gstride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gstride) {
        uint32_t pos, q4, t, ow, n, oh;
        d_kp4.divmod(i, pos, q4);
        d_wo.divmod(pos, t, ow);
        d_ho.divmod(t, n, oh);
        const float *xn = x + (size_t)n * cin * h * w;
        const int ih0 = (int)oh * stride - pad, iw0 = (int)ow * stride - pad;
        float v[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const uint32_t k = q4 * EPT + e;
            uint32_t ci, r, kh, kw;
            d_kk.divmod(k, ci, r);
            d_ks.divmod(r, kh, kw);
            const int ih = ih0 + (int)kh, iw = iw0 + (int)kw;
            const bool ok = k < (uint32_t)kdim && ih >= 0 && ih < h && iw >= 0 && iw < w;
            v[e] = ok ? __ldg(xn + ((size_t)ci * h + ih) * w + iw) : 0.f;
        }
        uint8_t *out = reinterpret_cast<uint8_t *>(col) + (size_t)i * 16;
        if constexpr (BF16)
            *reinterpret_cast<uint4 *>(out) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                           pack_bf16x2(v[EPT - 2], v[EPT - 1]));
        else
            *reinterpret_cast<float4 *>(out) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// BF16 flavour through shared memory: one CTA per (image, output row). The ks input rows of every
// channel are staged once (coalesced, zero-filled outside the image, with the left / right padding
// columns materialised), then every thread assembles 16-byte chunks of the row's col entries from
// them: no divisions and no scattered global loads in the inner loop (the gather kernel above spends
// 0.76 ms on ResNet-50's stem at batch 256, 4x the time its bytes need).
// smem: rows[cin * ks][wp] floats (wp = w + 2 * pad), then the k -> row / column offset table.
__global__ void __launch_bounds__(256)
im2col_rows_bf16_kernel(const float *__restrict__ x, uint4 *__restrict__ col, int cin, int h, int w, int ks,
                        int stride, int pad, int kdim, int kp8, int ho, int wo) {
    extern __shared__ float sm[];
    const int wp = w + 2 * pad;
    const int nrows = cin * ks;
    float *rows = sm;
    // k = 8 q + e -> (ci * ks + kh) * wp + kw, stored [e][q] so that a warp's lanes (consecutive q) hit
    // consecutive banks
    int *koff = reinterpret_cast<int *>(sm + (size_t)nrows * wp);
    const int oh = blockIdx.x % ho, n = blockIdx.x / ho;
    const int t = threadIdx.x;
    for (int k = t; k < kp8 * 8; k += (int)blockDim.x) {
        int off = -1;
        if (k < kdim) {
            const int ci = k / (ks * ks), r = k - ci * ks * ks, kh = r / ks, kw = r - kh * ks;
            off = (ci * ks + kh) * wp + kw;
        }
        koff[(k & 7) * kp8 + (k >> 3)] = off;
    }
    const int ih0 = oh * stride - pad;
    if ((w & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        // a warp per staged row, lanes on its 16-byte vectors; the loads of up to four rows go out before
        // the first shared-memory store (one row per pass left a global-memory round trip per pass
        // exposed: 0.56 ms for ResNet-50's stem at batch 256, 27 % of what its bytes need)
        const int lane = t & 31, warp = t >> 5, nwarps = ((int)blockDim.x + 31) >> 5, w4 = w >> 2;
        const int active = (int)blockDim.x - warp * 32 < 32 ? (int)blockDim.x - warp * 32 : 32;   // lanes of a partial last warp
        for (int r0 = warp; r0 < nrows; r0 += 4 * nwarps) {
            for (int c0 = lane; c0 < w4; c0 += 2 * active) {
                float4 v[4][2];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + u * nwarps;
                    const int ci = r / ks, kh = r - ci * ks, ih = ih0 + kh;
                    const bool row_ok = r < nrows && ih >= 0 && ih < h;
                    const float4 *src = reinterpret_cast<const float4 *>(x + (((size_t)n * cin + ci) * h + ih) * w);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int c4 = c0 + j * active;
                        v[u][j] = (row_ok && c4 < w4) ? __ldg(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + u * nwarps;
                    if (r >= nrows) continue;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int c4 = c0 + j * active;
                        if (c4 >= w4) continue;
                        float *d = rows + (size_t)r * wp + pad + 4 * c4;
                        d[0] = v[u][j].x; d[1] = v[u][j].y; d[2] = v[u][j].z; d[3] = v[u][j].w;
                    }
                }
            }
        }
        for (int i = t; i < nrows * 2 * pad; i += (int)blockDim.x) {   // left / right padding columns
            const int r = i / (2 * pad), c = i - r * 2 * pad;
            rows[(size_t)r * wp + (c < pad ? c : w + c)] = 0.f;
        }
    } else
    for (int i = t; i < nrows * wp; i += (int)blockDim.x) {
        const int r = i / wp, cpos = i - r * wp;
        const int ci = r / ks, kh = r - ci * ks;
        const int ih = ih0 + kh, iw = cpos - pad;
        float v = 0.f;
        if (ih >= 0 && ih < h && iw >= 0 && iw < w) v = __ldg(x + (((size_t)n * cin + ci) * h + ih) * w + iw);
        rows[i] = v;
    }
    __syncthreads();
    uint4 *out = col + ((size_t)n * ho + oh) * wo * kp8;
    // thread t owns chunk q = t % kp8 of every col row it writes (the block is a multiple of kp8 wide):
    // its 8 patch offsets live in registers, the loop walks output columns
    const int lanes = (int)blockDim.x / kp8;       // output columns per pass
    const int q = t % kp8, ow0 = t / kp8;
    if (ow0 < lanes) {
        int off[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) off[e] = koff[e * kp8 + q];
        for (int ow = ow0; ow < wo; ow += lanes) {
            const int base = ow * stride;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? rows[off[e] + base] : 0.f;
            out[(size_t)ow * kp8 + q] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                   pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
    }
}

int launch_im2col(const bcnn_b200_conv_desc *d, const float *x, void *col, bool bf16, cudaStream_t st) {
    const int kk = d->ksize * d->ksize, kp = im2col_kp(d);
    if (bf16 && !env_off("BCNN_B200_NO_FAST_IM2COL")) {
        const size_t smem = ((size_t)d->cin * d->ksize * (d->w + 2 * d->pad) + kp) * sizeof(float);
        if (smem <= 96 * 1024 && (long long)d->batch * d->ho < (1LL << 31)) {
            static bool attr_set = false;
            if (!attr_set) {
                cudaError_t e = cudaFuncSetAttribute(im2col_rows_bf16_kernel,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
                if (e != cudaSuccess) return (int)e;
                attr_set = true;
            }
            const int kp8 = kp / 8;
            const int threads = kp8 <= 256 ? (256 / kp8) * kp8 : 256;
            if (kp8 > 256) return (int)cudaErrorInvalidValue;
            im2col_rows_bf16_kernel<<<d->batch * d->ho, threads, smem, st>>>(
                x, reinterpret_cast<uint4 *>(col), d->cin, d->h, d->w, d->ksize, d->stride, d->pad, d->cin * kk,
                kp / 8, d->ho, d->wo);
            return launched();
        }
    }
    const int ept = bf16 ? 8 : 4;
    const size_t total4 = (size_t)d->batch * d->ho * d->wo * (kp / ept);
    if (total4 >= (1ull << 31)) return (int)cudaErrorInvalidValue;
    const int grid = stream_grid(total4, 256, 16);
    if (bf16)
        im2col_nhwc_kernel<true><<<grid, 256, 0, st>>>(
            x, col, d->cin, d->h, d->w, d->ksize, d->stride, d->pad, d->cin * kk, (uint32_t)total4,
            FastDiv((uint32_t)(kp / ept)), FastDiv((uint32_t)d->wo), FastDiv((uint32_t)d->ho),
            FastDiv((uint32_t)d->ksize), FastDiv((uint32_t)kk));
    else
        im2col_nhwc_kernel<false><<<grid, 256, 0, st>>>(
            x, col, d->cin, d->h, d->w, d->ksize, d->stride, d->pad, d->cin * kk, (uint32_t)total4,
            FastDiv((uint32_t)(kp / ept)), FastDiv((uint32_t)d->wo), FastDiv((uint32_t)d->ho),
            FastDiv((uint32_t)d->ksize), FastDiv((uint32_t)kk));
    return launched();
}

}  // namespace

namespace b200 {

size_t conv_pack_job_bytes() { return sizeof(PackJob); }
int conv_nhwc_pack_jobs(const bcnn_b200_conv_desc *d, int dgrad, const float *w, void *dst, void *jobs,
                        int max_jobs, size_t *bytes) {
    if (!resident_desc_ok(d)) return -1;
    return plan_pack_jobs(d, dgrad, w, reinterpret_cast<uint8_t *>(dst), reinterpret_cast<PackJob *>(jobs),
                          max_jobs, bytes);
}
unsigned int conv_pack_table_finish(void *jobs, int count) {
    PackJob *j = reinterpret_cast<PackJob *>(jobs);
    unsigned int block = 0;
    for (int i = 0; i < count; ++i) {
        j[i].first_block = block;
        const size_t chunks = (size_t)j[i].n_tiles * j[i].n_tile * j[i].taps.n * j[i].kc_blocks * 8;
        block += (unsigned int)((chunks + PACK_JOB_CHUNKS - 1) / PACK_JOB_CHUNKS);
    }
    return block;
}
int conv_pack_run(const void *jobs_dev, int count, unsigned int grid, cudaStream_t st) {
    if (count <= 0 || grid == 0) return 0;
    pack_jobs_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const PackJob *>(jobs_dev), count);
    return launched();
}
void conv_prepacked_set(const float *w, int dgrad, const void *image) {
    int k = -1;
    for (int i = 0; i < g_prepacked_n; ++i)
        if (g_prepacked[i].w == w && g_prepacked[i].dgrad == dgrad) k = i;
    if (!image) {
        if (k >= 0) g_prepacked[k] = g_prepacked[--g_prepacked_n];
        return;
    }
    if (k < 0) {
        if (g_prepacked_n == PREPACKED_MAX) return;
        k = g_prepacked_n++;
    }
    g_prepacked[k].w = w;
    g_prepacked[k].dgrad = dgrad;
    g_prepacked[k].image = reinterpret_cast<const uint8_t *>(image);
}
void conv_prepacked_enable(int on) { g_prepacked_on = on != 0; }

bool conv_tma_supports_fprop(const bcnn_b200_conv_desc *d) {
    FwdPlan pl;
    return route_fwd(d, false, &pl) != ROUTE_NONE;
}
bool conv_tma_supports_dgrad(const bcnn_b200_conv_desc *d) {
    FwdPlan pl;
    return route_fwd(d, true, &pl) != ROUTE_NONE;
}
bool conv_tma_supports_wgrad(const bcnn_b200_conv_desc *d) {
    WgPlan pl;
    return plan_wgrad(d, &pl);
}

size_t conv_tma_workspace_bytes(const bcnn_b200_conv_desc *d) {
    size_t need = 0;
    FwdPlan pl;
    for (int dgrad = 0; dgrad < 2; ++dgrad) {
        const FwdRoute r = route_fwd(d, dgrad != 0, &pl);
        size_t b = 0;
        if (r == ROUTE_STRIDED_DGRAD) b = strided_dgrad_bytes(d);
        else if (r != ROUTE_NONE) b = pl.shadow_bytes + pl.wpack_bytes + (dgrad ? 0 : pl.stat_bytes);
        if (b > need) need = b;
    }
    WgPlan wp;
    if (plan_wgrad(d, &wp)) {
        const size_t w = wp.shadow_x_bytes + wp.shadow_dy_bytes + wp.partial_bytes;
        if (w > need) need = w;
    }
    return need;
}

// Shadow storage a layer may keep (0 = the TMA route of `d` uses no such shadow).
size_t conv_tma_x_shadow_bytes(const bcnn_b200_conv_desc *d) {
    FwdPlan pl;
    const FwdRoute r = route_fwd(d, false, &pl);
    return (r == ROUTE_PLAIN || r == ROUTE_IM2COL) ? pl.shadow_bytes : 0;
}
size_t conv_tma_dy_shadow_bytes(const bcnn_b200_conv_desc *d) {
    size_t need = 0;
    FwdPlan pl;
    const FwdRoute r = route_fwd(d, true, &pl);
    if (r == ROUTE_STRIDED_DGRAD) {
        const bool bf16 = shadow_bf16() && d->cout % 8 == 0;
        need = align256((size_t)d->batch * d->cout * d->ho * d->wo * (bf16 ? 2 : 4));
    } else if (r == ROUTE_PLAIN) {
        need = pl.shadow_bytes;
    }
    WgPlan wp;
    if (plan_wgrad(d, &wp) && wp.shadow_dy_bytes > need) need = wp.shadow_dy_bytes;
    return need;
}

int conv_tma_forward(const bcnn_b200_conv_desc *d, const float *x, const float *w, const float *bias,
                     int act, float *y, void *workspace, size_t workspace_bytes,
                     bcnn_b200_conv_shadows *sh, cudaStream_t st) {
    return launch_fwd(d, false, x, w, bias, act, y, 0, workspace, workspace_bytes, sh, st);
}

// fprop without bias / activation that also leaves the batch-norm partial sums of y in the workspace:
// *stat_partial stays nullptr when the workspace is too small for them (the caller then runs the
// stand-alone statistics kernel over y).
int conv_tma_forward_stats(const bcnn_b200_conv_desc *d, const float *x, const float *w, float *y,
                           void *workspace, size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                           const float **stat_partial, int *stat_rows, cudaStream_t st) {
    *stat_partial = nullptr;
    *stat_rows = 0;
    return launch_fwd(d, false, x, w, nullptr, 0, y, 0, workspace, workspace_bytes, sh, st, stat_partial,
                      stat_rows);
}

int conv_tma_backward_data(const bcnn_b200_conv_desc *d, const float *w, const float *dy, float *dx,
                           int accumulate, void *workspace, size_t workspace_bytes,
                           bcnn_b200_conv_shadows *sh, cudaStream_t st) {
    return launch_fwd(d, true, dy, w, nullptr, 0, dx, accumulate, workspace, workspace_bytes, sh, st);
}

static int backward_weights_impl(const bcnn_b200_conv_desc *desc, const void *x_any, const void *dy_any,
                                 float *gw, void *workspace, size_t workspace_bytes,
                                 bcnn_b200_conv_shadows *sh, bool resident, cudaStream_t st) {
    const float *x = reinterpret_cast<const float *>(x_any);
    const float *dy = reinterpret_cast<const float *>(dy_any);
    WgPlan pl;
    if (!plan_wgrad(desc, &pl, resident)) return (int)cudaErrorInvalidValue;
    const WgEff e = wg_effective(desc, resident);
    const WgEff *d = &e;
    const size_t need = pl.shadow_x_bytes + pl.shadow_dy_bytes + pl.partial_bytes;
    if (need > 0 && (workspace == nullptr || workspace_bytes < need)) return (int)cudaErrorInvalidValue;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
        return (int)cudaErrorMisalignedAddress;
    uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
    const void *x_shadow = ws;
    const void *dy_sh = ws + pl.shadow_x_bytes;
    float *partial = reinterpret_cast<float *>(ws + pl.shadow_x_bytes + pl.shadow_dy_bytes);
    CUtensorMap tm_dy, tm_x;
    int err;
    if (pl.nhwc) {
        // x: the shadow (or im2col buffer) the forward pass of this layer kept, when its format fits
        const int x_fmt = d->im2col ? (pl.bf16 ? BCNN_B200_SHADOW_IM2COL_BF16 : BCNN_B200_SHADOW_IM2COL_F32)
                                    : (pl.bf16 ? BCNN_B200_SHADOW_NHWC_BF16 : BCNN_B200_SHADOW_NHWC_F32);
        if (resident && !d->im2col) {
            x_shadow = x_any;
        } else if (sh && sh->x && sh->x_fmt == x_fmt && sh->x_bytes >= pl.shadow_x_bytes) {
            x_shadow = sh->x;
        } else {
            err = d->im2col ? launch_im2col(desc, x, ws, pl.bf16, st)
                            : launch_transpose(x, ws, d->batch, d->cin, d->h * d->w, pl.bf16, st);
            if (err) return err;
        }
        if (resident) {
            dy_sh = dy_any;
        } else {
            err = dy_shadow(dy, d->batch, d->cout, d->ho * d->wo, pl.bf16, pl.shadow_dy_bytes, sh,
                            ws + pl.shadow_x_bytes, &dy_sh, st);
            if (err) return err;
        }
        if (!make_map_nhwc(&tm_dy, dy_sh, d->cout, d->wo, d->ho, d->batch, pl.bw, pl.bh, 1, 1, true,
                           pl.bf16) ||
            !make_map_nhwc(&tm_x, x_shadow, d->cin_phys, d->w, d->h, d->batch, pl.bw, pl.bh, 1, d->stride,
                           true, pl.bf16))
            return (int)cudaErrorInvalidValue;
    } else if (!make_map_nchw(&tm_dy, dy, pl.out_w, pl.out_h, d->cout, d->batch, TILE_M, false) ||
               !make_map_nchw(&tm_x, x, pl.view_w, pl.view_h, d->cin, d->batch, pl.n_tile, false)) {
        return (int)cudaErrorInvalidValue;
    }
    WgParams p;
    p.cin = d->cin; p.cout = d->cout; p.ks = d->ksize; p.kk = d->ksize * d->ksize; p.pad = d->pad;
    p.stride = d->stride;
    p.n_tile = pl.n_tile; p.bw = pl.bw; p.bh = pl.bh; p.kpos = pl.bw * pl.bh; p.nb = pl.nb;
    p.stage_bytes = pl.stage_bytes; p.a_bytes = pl.a_bytes; p.atom_bytes = pl.atom_bytes;
    p.blocks_w = pl.blocks_w; p.blocks_h = pl.blocks_h;
    p.kb_total = pl.kb_total; p.kb_per_split = pl.kb_per_split; p.splits = pl.splits;
    p.stages = pl.stages;
    const size_t wsize = (size_t)d->cout * d->cin * p.kk;
    p.split_stride = pl.splits > 1 ? wsize : 0;
    p.out = pl.splits > 1 ? partial : gw;
    p.d_img = FastDiv((uint32_t)(pl.blocks_h * pl.blocks_w));
    p.d_bw = FastDiv((uint32_t)pl.blocks_w);
    dim3 grid(pl.splits, pl.ci_tiles * p.kk, pl.co_tiles);
    err = !pl.nhwc ? launch_wgrad_kernel<false, false>(tm_dy, tm_x, p, grid, pl.smem_bytes, st)
          : pl.bf16 ? launch_wgrad_kernel<true, true>(tm_dy, tm_x, p, grid, pl.smem_bytes, st)
                    : launch_wgrad_kernel<true, false>(tm_dy, tm_x, p, grid, pl.smem_bytes, st);
    if (err) return err;
    if (pl.splits > 1) {
        wgrad_reduce_tma_kernel<<<stream_grid(wsize, 256), 256, 0, st>>>(gw, partial, wsize, pl.splits);
        return launched();
    }
    return 0;
}

int conv_tma_backward_weights(const bcnn_b200_conv_desc *desc, const float *x, const float *dy, float *gw,
                              void *workspace, size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                              cudaStream_t st) {
    return backward_weights_impl(desc, x, dy, gw, workspace, workspace_bytes, sh, false, st);
}

// ---- resident (BF16 NHWC) entry points
int conv_nhwc_supported(const bcnn_b200_conv_desc *d) {
    FwdPlan pl;
    WgPlan wp;
    int mask = 0;
    if (route_fwd_resident(d, false, &pl) != ROUTE_NONE) mask |= 1;
    if (route_fwd_resident(d, true, &pl) != ROUTE_NONE) mask |= 2;
    if (resident_desc_ok(d) && plan_wgrad(d, &wp, true)) mask |= 4;
    return mask;
}

size_t conv_nhwc_workspace_bytes(const bcnn_b200_conv_desc *d) {
    size_t need = 0;
    FwdPlan pl;
    FwdRoute r = route_fwd_resident(d, false, &pl);
    if (r != ROUTE_NONE) need = (r == ROUTE_IM2COL ? im2col_bytes(d) : 0) + pl.wpack_bytes + pl.stat_bytes;
    r = route_fwd_resident(d, true, &pl);
    size_t b = r == ROUTE_STRIDED_DGRAD ? strided_dgrad_pack_bytes_resident(d) : (r == ROUTE_PLAIN ? pl.wpack_bytes : 0);
    if (b > need) need = b;
    WgPlan wp;
    if (resident_desc_ok(d) && plan_wgrad(d, &wp, true)) {
        b = wp.shadow_x_bytes + wp.partial_bytes;
        if (b > need) need = b;
    }
    return need;
}

size_t conv_nhwc_x_keep_bytes(const bcnn_b200_conv_desc *d) {
    FwdPlan pl;
    return route_fwd_resident(d, false, &pl) == ROUTE_IM2COL ? im2col_bytes(d) : 0;
}

int conv_nhwc_forward(const bcnn_b200_conv_desc *d, const void *x, const float *w, const float *bias, int act,
                      void *y16, void *workspace, size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                      const float **stat_partial, int *stat_rows, cudaStream_t st) {
    if (stat_partial) { *stat_partial = nullptr; *stat_rows = 0; }
    return launch_fwd_resident(d, x, w, bias, act, y16, workspace, workspace_bytes, sh, st, stat_partial,
                               stat_rows);
}

int conv_nhwc_backward_data(const bcnn_b200_conv_desc *d, const float *w, const void *dy16, void *dx16,
                            int accumulate, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    return launch_dgrad_resident(d, w, dy16, dx16, accumulate, workspace, workspace_bytes, st);
}

int conv_nhwc_backward_weights(const bcnn_b200_conv_desc *d, const void *x, const void *dy16, float *gw,
                               void *workspace, size_t workspace_bytes, bcnn_b200_conv_shadows *sh,
                               cudaStream_t st) {
    if (!resident_desc_ok(d)) return (int)cudaErrorInvalidValue;
    return backward_weights_impl(d, x, dy16, gw, workspace, workspace_bytes, sh, true, st);
}

}  // namespace b200
