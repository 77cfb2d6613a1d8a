// common.cuh -- shared device/host helpers for the bcnn_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "bcnn_b200.h"

namespace b200 {

// bcnn_activation values (include/bcnn/bcnn.h).
enum Act : int {
    ACT_NONE = 0, ACT_TANH, ACT_RELU, ACT_RAMP, ACT_SOFTPLUS, ACT_LRELU, ACT_ABS,
    ACT_CLAMP, ACT_PRELU, ACT_LOGISTIC
};

extern unsigned long long g_launch_count;  // host-side counter (device.cu)
int sm_count();                            // cached SM count of the current device

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Launch epilogue: count the launch and surface launch-time errors.
static inline int launched() {
    ++g_launch_count;
    return (int)cudaGetLastError();
}

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ static inline size_t ceil_div_sz(size_t a, size_t b) { return (a + b - 1) / b; }

// Grid for a bandwidth-bound grid-stride kernel: enough CTAs to fill the SMs a few
// times over, never more than the work needs.
static inline int stream_grid(size_t work_items, int threads, int ctas_per_sm = 8) {
    size_t need = ceil_div_sz(work_items, (size_t)threads);
    size_t cap = (size_t)sm_count() * ctas_per_sm;
    size_t g = need < cap ? need : cap;
    return (int)(g ? g : 1);
}

// Division by a runtime-constant divisor (n < 2^31): q = (umulhi(n, mul) + n) >> shr.
struct FastDiv {
    uint32_t d, mul, shr;
    FastDiv() : d(1), mul(0), shr(0) {}
    explicit FastDiv(uint32_t div) : d(div) {
        if (div <= 1) { d = 1; mul = 0; shr = 0; return; }
        uint32_t s = 0;
        while ((1ull << s) < div) ++s;
        shr = s;
        mul = (uint32_t)((((1ull << 32) * ((1ull << s) - div)) / div) + 1);
    }
    __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
        return (__umulhi(n, mul) + n) >> shr;
#else
        return (uint32_t)(((((uint64_t)n * mul) >> 32) + n) >> shr);
#endif
    }
    __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &r) const {
        q = div(n);
        r = n - q * d;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of up to 4 values at once; result valid in thread 0.
template <int NV, int THREADS>
__device__ __forceinline__ void block_sum(float (&v)[NV], float *smem /* NV * THREADS/32 */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[i * (THREADS / 32) + wid] = v[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float t = lane < THREADS / 32 ? smem[i * (THREADS / 32) + lane] : 0.f;
            v[i] = warp_sum(t);
        }
    }
}

// Streaming (read-once / write-once) 128-bit accesses that bypass L1 allocation.
__device__ __forceinline__ float4 ld_stream4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// The same with L2 eviction priorities: read-once operands leave L2 first, results that the next
// kernel consumes stay last (a stream then keeps its freshly written tail / head in the 126 MB L2
// instead of the operands it will never touch again).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_stream4_hint(const float *p, uint64_t policy) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ float4 ld4_hint(const float *p, uint64_t policy) {   // coherent flavour
    float4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(policy) : "memory");
    return r;
}
__device__ __forceinline__ void st_stream4_hint(float *p, float4 v, uint64_t policy) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}

// Activation forward, arithmetic of bcnn_forward_activation_cpu
// (reference src/layers/bcnn_activation_layer.c:90-146): transcendentals go through
// double exp/log and are rounded to float where the reference casts.
__device__ __forceinline__ float act_fwd(float x, int act, float slope) {
    switch (act) {
        case ACT_TANH: {
            double e = exp((double)(2 * x));
            return (float)(e - 1) / ((float)e + 1);
        }
        case ACT_RELU: return x * (float)(x > 0);
        case ACT_LRELU: return x > 0 ? x : 0.1f * x;
        case ACT_RAMP: return x * (float)(x > 0) + 0.1f * x;
        case ACT_SOFTPLUS: return (float)log((double)(1.0f + (float)exp((double)x)));
        case ACT_ABS: return fabsf(x);
        case ACT_CLAMP: return x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
        case ACT_LOGISTIC: return 1.0f / (1.0f + (float)exp((double)(-x)));
        case ACT_PRELU: return x > 0 ? x : slope * x;
        default: return x;
    }
}

// Activation derivative factor evaluated on the POST-activation value y
// (bcnn_backward_activation_cpu, bcnn_activation_layer.c:165-226).
__device__ __forceinline__ float act_bwd_factor(float y, int act, float slope) {
    switch (act) {
        case ACT_TANH: return 1 - y * y;
        case ACT_RELU: return (float)(y > 0);
        case ACT_LRELU: return y > 0 ? 1.0f : 0.1f;
        case ACT_RAMP: return (float)(y > 0) + 0.1f;
        case ACT_SOFTPLUS: return 1.0f / (1.0f + (float)exp((double)(-y)));
        case ACT_ABS: return y >= 0 ? 1.0f : -1.0f;
        case ACT_CLAMP: return (float)(y > 0.0f && y < 1.0f);
        case ACT_LOGISTIC: return (1 - y) * y;
        case ACT_PRELU: return y > 0 ? 1.0f : slope;
        default: return 1.0f;
    }
}

}  // namespace b200
