// queries.cu -- query preparation: zero-row removal (CB/searcher.py:124-130), bf16 conversion,
// padding to the tile shapes of the tensor-core kernels.
#include "common.cuh"
#include <cuda_fp16.h>

namespace plaid {

// One CTA per query slot b in [0, B_pad).  Warp w tests rows w, w+nw, ...: sum |q| over 128 dims
// (> 0 keeps the row; a NaN sum drops it, as `torch.abs(Q).sum(-1) > 0` does).  Kept rows are
// compacted in order, rounded to bf16 (RNE); everything else in the [Lq_pad, 128] slab is zeroed.
__global__ void __launch_bounds__(256) prepare_queries_kernel(const float* __restrict__ Q, int B, int Lq,
                                                              int remove_zero, int Lq_pad,
                                                              __nv_bfloat16* __restrict__ Qb,
                                                              __half* __restrict__ Qh, int32_t* __restrict__ qlens) {
    extern __shared__ int s_flag[];  // [Lq] keep flags, then [Lq] destination rows
    int* s_dst = s_flag + Lq;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = warp_index(), nw = blockDim.x >> 5;
    __nv_bfloat16* out = Qb + (size_t)b * Lq_pad * kDim;
    __half* outh = Qh ? Qh + (size_t)b * Lq_pad * kDim : nullptr;      // optional fp16 twin (MaxSim operand)
    if (b >= B) {
        for (int i = threadIdx.x; i < Lq_pad * kDim / 8; i += blockDim.x) {
            reinterpret_cast<int4*>(out)[i] = make_int4(0, 0, 0, 0);
            if (outh) reinterpret_cast<int4*>(outh)[i] = make_int4(0, 0, 0, 0);
        }
        if (threadIdx.x == 0) qlens[b] = 0;
        return;
    }
    const float* q = Q + (size_t)b * Lq * kDim;
    for (int r = warp; r < Lq; r += nw) {
        float4 v = reinterpret_cast<const float4*>(q + (size_t)r * kDim)[lane];
        float s = fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_flag[r] = remove_zero ? (s > 0.0f ? 1 : 0) : 1;
    }
    __syncthreads();
    if (warp == 0) {  // ordered compaction: warp-wide ballot scan over the rows
        int base = 0;
        for (int r0 = 0; r0 < Lq; r0 += 32) {
            const int r = r0 + lane;
            const int f = (r < Lq) ? s_flag[r] : 0;
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (r < Lq) s_dst[r] = f ? base + __popc(m & ((1u << lane) - 1)) : -1;
            base += __popc(m);
        }
        if (lane == 0) {
            qlens[b] = base;
            s_flag[0] = base;  // reuse as broadcast slot (flags are no longer needed)
        }
    }
    __syncthreads();
    const int kept = s_flag[0];
    for (int r = warp; r < Lq; r += nw) {
        const int d = s_dst[r];
        if (d < 0) continue;
        float4 v = reinterpret_cast<const float4*>(q + (size_t)r * kDim)[lane];
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        reinterpret_cast<uint2*>(out + (size_t)d * kDim)[lane] = pk;
        if (outh) {
            __half2 hl = __floats2half2_rn(v.x, v.y), hh = __floats2half2_rn(v.z, v.w);
            reinterpret_cast<uint2*>(outh + (size_t)d * kDim)[lane] =
                make_uint2(*reinterpret_cast<uint32_t*>(&hl), *reinterpret_cast<uint32_t*>(&hh));
        }
    }
    for (int r = kept + warp; r < Lq_pad; r += nw) {
        reinterpret_cast<uint2*>(out + (size_t)r * kDim)[lane] = make_uint2(0, 0);
        if (outh) reinterpret_cast<uint2*>(outh + (size_t)r * kDim)[lane] = make_uint2(0, 0);
    }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n) {
            float4 v = *reinterpret_cast<const float4*>(src + i);
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            *reinterpret_cast<uint2*>(dst + i) =
                make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        } else {
            for (int64_t j = i; j < n; j++) dst[j] = __float2bfloat16_rn(src[j]);
        }
    }
}

}  // namespace plaid

extern "C" int plaid_prepare_queries(const float* Q, int B, int Lq, int remove_zero_rows, int B_pad, int Lq_pad,
                                     void* Qb_bf16, void* Qh_f16, int32_t* qlens, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(Q && Qb_bf16 && qlens, PLAID_ERR_ARG, "plaid_prepare_queries: null pointer");
    PLAID_CHECK_ARG(B >= 0 && B_pad >= B && (B_pad % 4) == 0, PLAID_ERR_ARG,
                    "plaid_prepare_queries: B_pad=%d must be a multiple of 4 and >= B=%d", B_pad, B);
    PLAID_CHECK_ARG(Lq >= 1 && Lq_pad >= Lq && (Lq_pad % 32) == 0 && Lq <= 4096, PLAID_ERR_ARG,
                    "plaid_prepare_queries: Lq_pad=%d must be a multiple of 32 and >= Lq=%d (<= 4096)", Lq_pad, Lq);
    if (B_pad == 0) return PLAID_OK;
    prepare_queries_kernel<<<B_pad, 256, 2 * Lq * sizeof(int), (cudaStream_t)stream>>>(
        Q, B, Lq, remove_zero_rows, Lq_pad, reinterpret_cast<__nv_bfloat16*>(Qb_bf16), reinterpret_cast<__half*>(Qh_f16), qlens);
    PLAID_LAUNCH_OK("prepare_queries_kernel");
    return PLAID_OK;
}

extern "C" int plaid_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(src && dst_bf16 && n >= 0, PLAID_ERR_ARG, "plaid_f32_to_bf16: bad argument");
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_bf16) & 7) == 0,
                    PLAID_ERR_ARG, "plaid_f32_to_bf16: pointers must be 16/8-byte aligned");
    if (n == 0) return PLAID_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    f32_to_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), n);
    PLAID_LAUNCH_OK("f32_to_bf16_kernel");
    return PLAID_OK;
}
