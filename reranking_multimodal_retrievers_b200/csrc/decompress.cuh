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
__device__ __forceinline__ uint4 lds_v4u32(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts_v4u32(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// Same store without the compiler-level memory barrier: ordered against the other volatile shared
// accesses and the fences/barriers that publish it, but ordinary loads (centroid rows) may move across it.
__device__ __forceinline__ void sts_v4u32_relaxed(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(x), "r"(y), "r"(z), "r"(w));
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

// ======================= fp16 decoding (search pipeline) =======================
// The arithmetic of the reference's GPU branch: half centroid + half bucket weight -> half
// (CB/indexing/codecs/decompress_residuals.cu:35-37), L2-normalised and kept in half
// (CB/indexing/codecs/residual.py:273).  Used by both the fused MaxSim kernel and
// plaid_decompress_normalize_f16, so the two produce the same bits.
//
// Weight table in shared memory: entry x (a packed residual byte) holds the 8/NBITS fp16 weights it
// expands to.  Entries are 128 B apart and each is replicated across the 128-byte line, one copy per
// lane slot: lane l reads copy l % (128 / ES), so the lanes of a wavefront never meet in a bank and the
// random byte values cost no bank conflicts (the fp32 table of the first version spent ~60% of the
// shared-memory load wavefronts on them).
static constexpr int kLutBytes = 256 * 128;
template <int NBITS> __host__ __device__ constexpr int lut_entry_bytes() { return NBITS == 8 ? 4 : 16 / NBITS; }

template <int NBITS>
__device__ __forceinline__ void lut_fill_f16(const float* __restrict__ W, uint8_t* lut) {
    constexpr int ES = lut_entry_bytes<NBITS>(), KEYS = 8 / NBITS, SLOTS = 128 / ES;
    for (int i = threadIdx.x; i < 256 * SLOTS; i += blockDim.x) {
        const int e = i / SLOTS, sl = i - e * SLOTS;
        __half* dst = reinterpret_cast<__half*>(lut + e * 128 + sl * ES);
#pragma unroll
        for (int k = 0; k < KEYS; k++) dst[k] = __float2half_rn(W[e * KEYS + k]);
        if (NBITS == 8) dst[1] = __float2half_rn(0.0f);
    }
}
// shared address of this lane's copy of entry 0
template <int NBITS>
__device__ __forceinline__ uint32_t lut_lane_base(uint32_t lut_sa, int lane) {
    constexpr int ES = lut_entry_bytes<NBITS>();
    return lut_sa + (uint32_t)(lane & (128 / ES - 1)) * ES;
}
// fp16 weights of dimensions 8h..8h+7 of the token whose packed row sits at shared address `row`, as 4 half2 words
// (byte extraction with PRMT, entry address with one multiply-add)
template <int NBITS>
__device__ __forceinline__ void token_weights_h8(uint32_t row, uint32_t lut, int h, uint32_t (&w)[4]) {
    if constexpr (NBITS == 2) {
        const uint32_t x = lds_u16(row + 2 * h);
        const uint2 a = lds_v2u32(__byte_perm(x, 0, 0x4440) * 128u + lut), b = lds_v2u32(__byte_perm(x, 0, 0x4441) * 128u + lut);
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
    } else if constexpr (NBITS == 4) {
        const uint32_t x = lds_u32(row + 4 * h);
        w[0] = lds_u32(__byte_perm(x, 0, 0x4440) * 128u + lut);
        w[1] = lds_u32(__byte_perm(x, 0, 0x4441) * 128u + lut);
        w[2] = lds_u32(__byte_perm(x, 0, 0x4442) * 128u + lut);
        w[3] = lds_u32(__byte_perm(x, 0, 0x4443) * 128u + lut);
    } else if constexpr (NBITS == 1) {
        const uint32_t x = lds_u8(row + h);
        const uint4 a = lds_v4u32(x * 128u + lut);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else {
        const uint2 x = lds_v2u32(row + 8 * h);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t xx = i < 2 ? x.x : x.y;
            const uint32_t lo = lds_u16(__byte_perm(xx, 0, 0x4440 + 2 * (i & 1)) * 128u + lut);
            const uint32_t hi = lds_u16(__byte_perm(xx, 0, 0x4441 + 2 * (i & 1)) * 128u + lut);
            w[i] = lo | (hi << 16);
        }
    }
}

__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// A token is decoded by 8 lanes (a quarter warp); lane q of the quarter owns dimensions 8q..8q+7 (`lo`) and
// 64+8q..64+8q+7 (`hi`), i.e. 16-byte chunk q of both 64-dim k-halves of the operand tile.
// centroid row address: base + code * 256 bytes in one 64-bit multiply-add
__device__ __forceinline__ const uint4* centroid_row(const void* base_q, uint32_t code) {
    uint64_t a;
    asm("mad.wide.u32 %0, %1, 256, %2;" : "=l"(a) : "r"(code), "l"(reinterpret_cast<uint64_t>(base_q)));
    return reinterpret_cast<const uint4*>(a);
}
// v = centroid + weight (half2); returns the lane's partial sum of squares (two fp16 accumulators of 4 products
// each per half2 component, combined in fp32)
__device__ __forceinline__ float token_sum_h16(const uint4& clo, const uint4& chi, const uint32_t (&wlo)[4],
                                               const uint32_t (&whi)[4], __half2 (&v)[8]) {
    v[0] = __hadd2(u32_as_h2(clo.x), u32_as_h2(wlo[0]));
    v[1] = __hadd2(u32_as_h2(clo.y), u32_as_h2(wlo[1]));
    v[2] = __hadd2(u32_as_h2(clo.z), u32_as_h2(wlo[2]));
    v[3] = __hadd2(u32_as_h2(clo.w), u32_as_h2(wlo[3]));
    v[4] = __hadd2(u32_as_h2(chi.x), u32_as_h2(whi[0]));
    v[5] = __hadd2(u32_as_h2(chi.y), u32_as_h2(whi[1]));
    v[6] = __hadd2(u32_as_h2(chi.z), u32_as_h2(whi[2]));
    v[7] = __hadd2(u32_as_h2(chi.w), u32_as_h2(whi[3]));
    __half2 a = __hmul2(v[0], v[0]), b = __hmul2(v[4], v[4]);
#pragma unroll
    for (int i = 1; i < 4; i++) {
        a = __hfma2(v[i], v[i], a);
        b = __hfma2(v[4 + i], v[4 + i], b);
    }
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return (fa.x + fa.y) + (fb.x + fb.y);
}
// sum over the 8 lanes of the quarter warp
__device__ __forceinline__ float quarter_sum(float ss) {
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    return ss;
}
// v = centroid + weight only (the scale comes from the precomputed per-token table)
__device__ __forceinline__ void token_add_h16(const uint4& clo, const uint4& chi, const uint32_t (&wlo)[4],
                                              const uint32_t (&whi)[4], __half2 (&v)[8]) {
    v[0] = __hadd2(u32_as_h2(clo.x), u32_as_h2(wlo[0]));
    v[1] = __hadd2(u32_as_h2(clo.y), u32_as_h2(wlo[1]));
    v[2] = __hadd2(u32_as_h2(clo.z), u32_as_h2(wlo[2]));
    v[3] = __hadd2(u32_as_h2(clo.w), u32_as_h2(wlo[3]));
    v[4] = __hadd2(u32_as_h2(chi.x), u32_as_h2(whi[0]));
    v[5] = __hadd2(u32_as_h2(chi.y), u32_as_h2(whi[1]));
    v[6] = __hadd2(u32_as_h2(chi.z), u32_as_h2(whi[2]));
    v[7] = __hadd2(u32_as_h2(chi.w), u32_as_h2(whi[3]));
}
// 1 / max(||x||, 1e-12) rounded to half: the scale factor of a token, exactly as token_scale_h16 applies it
__device__ __forceinline__ __half token_inv_h16(float ss) { return __float2half_rn(rsqrt_ftz(fmaxf(ss, 1e-24f))); }
// scale by a precomputed half factor (0 for pad rows)
__device__ __forceinline__ void token_scale_pre_h16(const __half2 (&v)[8], uint32_t inv_bits, uint4& lo, uint4& hi) {
    const __half2 i2 = u32_as_h2(__byte_perm(inv_bits, 0, 0x1010));
    lo = make_uint4(h2_as_u32(__hmul2(v[0], i2)), h2_as_u32(__hmul2(v[1], i2)), h2_as_u32(__hmul2(v[2], i2)),
                    h2_as_u32(__hmul2(v[3], i2)));
    hi = make_uint4(h2_as_u32(__hmul2(v[4], i2)), h2_as_u32(__hmul2(v[5], i2)), h2_as_u32(__hmul2(v[6], i2)),
                    h2_as_u32(__hmul2(v[7], i2)));
}
// x / max(||x||, 1e-12) = x * rsqrt(max(||x||^2, 1e-24)) (index_storage.py:175); pad rows become zeros
__device__ __forceinline__ void token_scale_h16(const __half2 (&v)[8], float ss, bool real, uint4& lo, uint4& hi) {
    const float inv = real ? rsqrt_ftz(fmaxf(ss, 1e-24f)) : 0.0f;
    const __half2 i2 = __float2half2_rn(inv);
    lo = make_uint4(h2_as_u32(__hmul2(v[0], i2)), h2_as_u32(__hmul2(v[1], i2)), h2_as_u32(__hmul2(v[2], i2)),
                    h2_as_u32(__hmul2(v[3], i2)));
    hi = make_uint4(h2_as_u32(__hmul2(v[4], i2)), h2_as_u32(__hmul2(v[5], i2)), h2_as_u32(__hmul2(v[6], i2)),
                    h2_as_u32(__hmul2(v[7], i2)));
}

}  // namespace plaid
