// decompress.cuh -- per-token residual decoding shared by decompress.cu and the fused MaxSim kernel.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace plaid {

// Explicit shared-space accesses (32-bit shared addresses): pointers carved out of dynamic shared
// memory lose their address space in the compiler and would otherwise become generic LD/ST.
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)  { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_v2u32(uint32_t a) {
    uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_v2f32(uint32_t a) {
    float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ float4 lds_v4f32(uint32_t a) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts_v4u32(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Bucket weights of the 8 dimensions 8h..8h+7 (h = 0..15) of one token whose packed row sits at shared
// address `row`; sW = shared address of the weight table W[256][8/NBITS] (fp32).
template <int NBITS>
__device__ __forceinline__ void token_weights8(uint32_t row, uint32_t sW, int h, float (&w)[8]) {
    if constexpr (NBITS == 2) {           // 4 dims per byte: bytes 2h, 2h+1; W row = 16 B
        const uint32_t x = lds_u16(row + 2 * h);
        const float4 a = lds_v4f32(sW + ((x & 0xff) << 4)), b = lds_v4f32(sW + ((x >> 8) << 4));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else if constexpr (NBITS == 4) {    // 2 dims per byte: bytes 4h..4h+3; W row = 8 B
        const uint32_t x = lds_u32(row + 4 * h);
        const float2 a = lds_v2f32(sW + ((x & 0xff) << 3)), b = lds_v2f32(sW + (((x >> 8) & 0xff) << 3));
        const float2 c = lds_v2f32(sW + (((x >> 16) & 0xff) << 3)), d = lds_v2f32(sW + ((x >> 24) << 3));
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y; w[6] = d.x; w[7] = d.y;
    } else if constexpr (NBITS == 1) {    // 8 dims per byte: byte h; W row = 32 B
        const uint32_t x = lds_u8(row + h);
        const float4 a = lds_v4f32(sW + (x << 5)), b = lds_v4f32(sW + (x << 5) + 16);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {                              // 1 dim per byte: bytes 8h..8h+7; W row = 4 B
        const uint2 x = lds_v2u32(row + 8 * h);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            w[i] = lds_f32(sW + (((x.x >> (8 * i)) & 0xff) << 2));
            w[4 + i] = lds_f32(sW + (((x.y >> (8 * i)) & 0xff) << 2));
        }
    }
}

// centroid elements 8h..8h+7 widened to fp32
__device__ __forceinline__ void load_centroid8(const float* c, int h, float (&e)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(c) + 2 * h), b = __ldg(reinterpret_cast<const float4*>(c) + 2 * h + 1);
    e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
}
__device__ __forceinline__ void load_centroid8(const __half* c, int h, float (&e)[8]) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(c) + h);   // one 128-bit load = 8 fp16
    const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[i]));
        e[2 * i] = f.x;
        e[2 * i + 1] = f.y;
    }
}

}  // namespace plaid
