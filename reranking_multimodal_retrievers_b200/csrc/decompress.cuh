// decompress.cuh -- per-token residual decoding shared by decompress.cu and the fused MaxSim kernel.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace plaid {

// Bucket weights of the 8 dimensions 8h..8h+7 (h = 0..15) of one token whose packed row sits at `row` (smem).
template <int NBITS>
__device__ __forceinline__ void token_weights8(const uint8_t* row, const float* sW, int h, float (&w)[8]) {
    if constexpr (NBITS == 2) {           // 4 dims per byte: bytes 2h, 2h+1
        const uchar2 x = reinterpret_cast<const uchar2*>(row)[h];
        const float4 a = reinterpret_cast<const float4*>(sW)[x.x], b = reinterpret_cast<const float4*>(sW)[x.y];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else if constexpr (NBITS == 4) {    // 2 dims per byte: bytes 4h..4h+3
        const uchar4 x = reinterpret_cast<const uchar4*>(row)[h];
        const float2 a = reinterpret_cast<const float2*>(sW)[x.x], b = reinterpret_cast<const float2*>(sW)[x.y];
        const float2 c = reinterpret_cast<const float2*>(sW)[x.z], d = reinterpret_cast<const float2*>(sW)[x.w];
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y; w[6] = d.x; w[7] = d.y;
    } else if constexpr (NBITS == 1) {    // 8 dims per byte: byte h
        const int x = row[h];
        const float4 a = reinterpret_cast<const float4*>(sW)[x * 2], b = reinterpret_cast<const float4*>(sW)[x * 2 + 1];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {                              // 1 dim per byte: bytes 8h..8h+7
        const uint2 x = reinterpret_cast<const uint2*>(row)[h];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            w[i] = sW[(x.x >> (8 * i)) & 0xff];
            w[4 + i] = sW[(x.y >> (8 * i)) & 0xff];
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
