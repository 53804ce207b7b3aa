// Shared device helpers for the DualAR decode kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace smol {

constexpr int kHeadDim = 64;  // both shipped configs; head_dim = dim / n_head (reference rq_transformer.py:65)
constexpr int kWarp = 32;

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

// 16-byte read-only streaming load (weights are read once per step: do not pollute L1).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) { return *reinterpret_cast<const uint4*>(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- Philox4x32-10 (counter-based RNG; same constants as oracle/sampler_oracle.py) ----
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// exp(x) * 2^30 truncated to uint32 for x <= 0, IEEE mul/add only (no FMA contraction) so the
// numpy restatement is bit-identical (DESIGN.md "Sampling").
__device__ __forceinline__ uint32_t exp_det_weight(float x) {
    if (!(x >= -21.5f)) return 0u;
    float t = __fmul_rn(x, 1.4426950408889634f);
    float n = rintf(t);
    float f = __fsub_rn(t, n);
    float p = 1.5252733804059841e-05f;
    p = __fadd_rn(__fmul_rn(p, f), 0.00015403530393381608f);
    p = __fadd_rn(__fmul_rn(p, f), 0.0013333558146428443f);
    p = __fadd_rn(__fmul_rn(p, f), 0.009618129107628477f);
    p = __fadd_rn(__fmul_rn(p, f), 0.05550410866482158f);
    p = __fadd_rn(__fmul_rn(p, f), 0.2402265069591007f);
    p = __fadd_rn(__fmul_rn(p, f), 0.6931471805599453f);
    p = __fadd_rn(__fmul_rn(p, f), 1.0f);
    float scale = __uint_as_float((uint32_t)((int)n + 127) << 23);
    float val = __fmul_rn(p, scale);
    return __float2uint_rz(__fmul_rn(val, 1073741824.0f));
}

}  // namespace smol
