// maxsim_backward.cu -- gradient of the padded MaxSim, for the training-time scoring of FLMR:
// FLMRModelForRetrieval.score / compute_ib_loss_new (src/models/flmr/models/flmr/modeling_flmr.py:932-947,1089-1125)
// run colbert_score (flmr_utils.py:22-48) under autograd; here the forward is the tcgen05 kernel of maxsim.cu and
// the backward is made of three small kernels over the same bf16-rounded operands:
//
//   score[p] = sum_{k < qlen} max_t sim[p, t, k],  sim = <D[d, t], Q[q, k]>, masked positions count as -9999
//   A  argmax  : idx[p, k] = the (first) t that attains the maximum
//   B  dQ      : dQ[q, k, :] = sum over the pairs p of q of g[p] * D[d_p, idx[p, k], :]     (no atomics: a CTA owns q)
//   C  dD      : dD[d, t, :] = sum over the pairs p of d, k with idx[p, k] == t of g[p] * Q[q_p, k, :]
//                (a CTA owns passage d and accumulates in shared memory; global atomics only when Ld x 128 floats
//                 do not fit)
// Pairing: all_pairs != 0 -> p = q * n + d for every query and passage (in-batch negatives, the [B, B*n_docs] score
// matrix of compute_ib_loss_new); else passage d belongs to query d / docs_per_query and p = d (colbert_score).
// The argmax is recomputed on the CUDA cores (bf16 products are exact in fp32; only the order of the 128 additions
// differs from the tensor cores', so two passage tokens closer than that can swap -- a sub-gradient either way).
#include "common.cuh"
#include <cuda_bf16.h>

namespace plaid {

static constexpr int kBwWarps = 8;
static constexpr int kBwRow = kDim + 2;     // bf16 row stride in shared memory: 65 words -> lanes reading different rows never share a bank

__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

struct BwPairing {
    long long n;          // passages
    int nQ, docs_per_query, all_pairs;
    __device__ __forceinline__ long long pairs() const { return all_pairs ? (long long)nQ * n : n; }
    __device__ __forceinline__ void split(long long p, int& q, long long& d) const {
        if (all_pairs) { q = (int)(p / n); d = p - (long long)q * n; }
        else { d = p; q = (int)(p / docs_per_query); }
    }
};

// ---- A: one CTA per pair; the passage's rows are staged in shared memory, a warp takes query tokens k = warp, warp + 8, ...,
// lane l scores passage tokens l, l + 32, ... against the query row (read as broadcasts), then a warp arg-max.
__global__ void __launch_bounds__(kBwWarps * 32)
maxsim_argmax_kernel(const __nv_bfloat16* __restrict__ Qb, const int32_t* __restrict__ qlens, int Lq_pad,
                     const __nv_bfloat16* __restrict__ D, const uint8_t* __restrict__ mask, int Ld, BwPairing pr,
                     int32_t* __restrict__ idx) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __nv_bfloat16* sD = reinterpret_cast<__nv_bfloat16*>(s_raw);                 // [Ld][kBwRow]
    __nv_bfloat16* sQ = sD + (((size_t)Ld * kBwRow + 7) & ~size_t(7));           // [kBwWarps][kDim]: the warp's query row (16-byte aligned)
    const int lane = threadIdx.x & 31, warp = warp_index();
    for (long long p = blockIdx.x; p < pr.pairs(); p += gridDim.x) {
        int q; long long d;
        pr.split(p, q, d);
        __syncthreads();
        const uint32_t* src = reinterpret_cast<const uint32_t*>(D + (size_t)d * Ld * kDim);
        for (int i = threadIdx.x; i < Ld * (kDim / 2); i += blockDim.x) {
            const int t = i / (kDim / 2), c = i - t * (kDim / 2);
            reinterpret_cast<uint32_t*>(sD + (size_t)t * kBwRow)[c] = __ldg(src + i);
        }
        __syncthreads();
        const int lq = min(qlens[q], Lq_pad);
        const uint8_t* mrow = mask + (size_t)d * Ld;
        for (int k = warp; k < lq; k += kBwWarps) {
            reinterpret_cast<uint2*>(sQ + warp * kDim)[lane] =
                __ldg(reinterpret_cast<const uint2*>(Qb + ((size_t)q * Lq_pad + k) * kDim) + lane);
            __syncwarp();
            float best = -INFINITY;
            int bt = 0x7fffffff;
            for (int t = lane; t < Ld; t += 32) {
                const uint32_t* drow = reinterpret_cast<const uint32_t*>(sD + (size_t)t * kBwRow);
                const uint32_t* qrow = reinterpret_cast<const uint32_t*>(sQ + warp * kDim);
                float acc = 0.0f;
#pragma unroll 16
                for (int c = 0; c < kDim / 2; c++) {
                    const float2 a = bf2_to_f2(drow[c]), b = bf2_to_f2(qrow[c]);
                    acc = fmaf(a.x, b.x, acc);
                    acc = fmaf(a.y, b.y, acc);
                }
                if (!mrow[t]) acc = -9999.0f;                 // flmr_utils.py:26: scores_padded[D_padding] = -9999
                if (acc > best) { best = acc; bt = t; }       // ascending t per lane: the first maximum stays
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
                if (ob > best || (ob == best && ot < bt)) { best = ob; bt = ot; }
            }
            if (lane == 0) idx[p * Lq_pad + k] = bt;
            __syncwarp();
        }
    }
}

// ---- B: one CTA per query q, a warp per query token k (strided); sums g[p] * D[d_p, idx[p,k], :] over q's pairs.
__global__ void __launch_bounds__(kBwWarps * 32)
maxsim_dq_kernel(const int32_t* __restrict__ qlens, int Lq_pad, const __nv_bfloat16* __restrict__ D, int Ld, BwPairing pr,
                 const int32_t* __restrict__ idx, const float* __restrict__ grad, float* __restrict__ dQ) {
    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = warp_index();
    const int lq = min(qlens[q], Lq_pad);
    long long p0, p1, dstep;          // the pairs of q: p = p0 + j, passage d = d0 + j
    long long d0;
    if (pr.all_pairs) { p0 = (long long)q * pr.n; p1 = p0 + pr.n; d0 = 0; }
    else { p0 = (long long)q * pr.docs_per_query; p1 = min(p0 + pr.docs_per_query, pr.n); d0 = p0; }
    (void)dstep;
    for (int k = warp; k < Lq_pad; k += kBwWarps) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < lq) {
            for (long long p = p0; p < p1; p++) {
                const float g = grad[p];
                const int t = idx[p * Lq_pad + k];
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(D + ((size_t)(d0 + (p - p0)) * Ld + t) * kDim) + lane);
                const float2 a = bf2_to_f2(v.x), b = bf2_to_f2(v.y);
                acc.x = fmaf(g, a.x, acc.x); acc.y = fmaf(g, a.y, acc.y);
                acc.z = fmaf(g, b.x, acc.z); acc.w = fmaf(g, b.y, acc.w);
            }
        }
        reinterpret_cast<float4*>(dQ + ((size_t)q * Lq_pad + k) * kDim)[lane] = acc;
    }
}

// ---- C: one CTA per passage d; dD[d] accumulated in shared memory (USE_SMEM) or straight into the zeroed output.
template <bool USE_SMEM>
__global__ void __launch_bounds__(kBwWarps * 32)
maxsim_dd_kernel(const __nv_bfloat16* __restrict__ Qb, const int32_t* __restrict__ qlens, int Lq_pad, int Ld, BwPairing pr,
                 const int32_t* __restrict__ idx, const float* __restrict__ grad, float* __restrict__ dD) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    float* sAcc = reinterpret_cast<float*>(s_raw);            // [Ld][kDim]
    const long long d = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = warp_index();
    float* out = dD + (size_t)d * Ld * kDim;
    if (USE_SMEM) {
        for (int i = threadIdx.x; i < Ld * kDim; i += blockDim.x) sAcc[i] = 0.0f;
        __syncthreads();
    }
    const int q_lo = pr.all_pairs ? 0 : (int)(d / pr.docs_per_query), q_hi = pr.all_pairs ? pr.nQ : q_lo + 1;
    for (int q = q_lo; q < q_hi; q++) {
        const long long p = pr.all_pairs ? (long long)q * pr.n + d : d;
        const float g = grad[p];
        const int lq = min(qlens[q], Lq_pad);
        for (int k = warp; k < lq; k += kBwWarps) {
            const int t = idx[p * Lq_pad + k];
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(Qb + ((size_t)q * Lq_pad + k) * kDim) + lane);
            const float2 a = bf2_to_f2(v.x), b = bf2_to_f2(v.y);
            float* dst = (USE_SMEM ? sAcc : out) + (size_t)t * kDim + lane * 4;
            atomicAdd(dst, g * a.x); atomicAdd(dst + 1, g * a.y); atomicAdd(dst + 2, g * b.x); atomicAdd(dst + 3, g * b.y);
        }
    }
    if (USE_SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < Ld * kDim / 4; i += blockDim.x)
            reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(sAcc)[i];
    }
}

}  // namespace plaid

extern "C" int plaid_colbert_score_backward(const void* Qb_bf16, const int32_t* qlens, int nQ, int Lq_pad,
                                            const void* D_bf16, const uint8_t* D_mask, int64_t n, int Ld, int docs_per_query,
                                            int all_pairs, const float* grad, int32_t* ws_argmax, float* dQ, float* dD,
                                            void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(Qb_bf16 && qlens && D_bf16 && D_mask && grad && ws_argmax, PLAID_ERR_ARG, "plaid_colbert_score_backward: null pointer");
    PLAID_CHECK_ARG(nQ >= 1 && n >= 0 && Ld >= 1 && Lq_pad >= 1 && (all_pairs || docs_per_query >= 1), PLAID_ERR_ARG,
                    "plaid_colbert_score_backward: bad sizes");
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(Qb_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(D_bf16) & 15) == 0,
                    PLAID_ERR_ARG, "plaid_colbert_score_backward: Q / D must be 16-byte aligned");
    if (n == 0) return PLAID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const BwPairing pr{(long long)n, nQ, docs_per_query > 0 ? docs_per_query : 1, all_pairs};
    const long long pairs = all_pairs ? (long long)nQ * n : (long long)n;
    const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(Qb_bf16);
    const __nv_bfloat16* D = reinterpret_cast<const __nv_bfloat16*>(D_bf16);
    const size_t smem_a = ((((size_t)Ld * kBwRow + 7) & ~size_t(7)) + kBwWarps * kDim) * sizeof(__nv_bfloat16);
    PLAID_CHECK_ARG(smem_a <= 200 * 1024, PLAID_ERR_UNSUPPORTED, "plaid_colbert_score_backward: Ld=%d passage tokens exceed shared memory", Ld);
    static int conf_a[kMaxDevices] = {0}, conf_c[kMaxDevices] = {0};
    if (int rc = ensure_dynamic_smem((const void*)maxsim_argmax_kernel, (int)smem_a, conf_a)) return rc;
    const long long grid_a = pairs < 148ll * 64 ? pairs : 148ll * 64;
    maxsim_argmax_kernel<<<(int)grid_a, kBwWarps * 32, smem_a, st>>>(Q, qlens, Lq_pad, D, D_mask, Ld, pr, ws_argmax);
    PLAID_LAUNCH_OK("maxsim_argmax_kernel");
    if (dQ) {
        maxsim_dq_kernel<<<nQ, kBwWarps * 32, 0, st>>>(qlens, Lq_pad, D, Ld, pr, ws_argmax, grad, dQ);
        PLAID_LAUNCH_OK("maxsim_dq_kernel");
    }
    if (dD) {
        PLAID_CHECK_ARG(n <= 0x7fffffffll, PLAID_ERR_UNSUPPORTED, "plaid_colbert_score_backward: too many passages");
        const size_t smem_c = (size_t)Ld * kDim * sizeof(float);
        if (smem_c <= 200 * 1024) {
            if (int rc = ensure_dynamic_smem((const void*)maxsim_dd_kernel<true>, (int)smem_c, conf_c)) return rc;
            maxsim_dd_kernel<true><<<(int)n, kBwWarps * 32, smem_c, st>>>(Q, qlens, Lq_pad, Ld, pr, ws_argmax, grad, dD);
        } else {
            PLAID_CUDA_OK(cudaMemsetAsync(dD, 0, (size_t)n * Ld * kDim * sizeof(float), st));
            maxsim_dd_kernel<false><<<(int)n, kBwWarps * 32, 0, st>>>(Q, qlens, Lq_pad, Ld, pr, ws_argmax, grad, dD);
        }
        PLAID_LAUNCH_OK("maxsim_dd_kernel");
    }
    return PLAID_OK;
}
