// segmented.cu -- the operator-level reductions and gathers the reference binds as class attributes:
//   ColBERT.segmented_maxsim      (CB/modeling/segmented_maxsim.cpp:22-93)
//   colbert_score_reduce          (CB/modeling/colbert.py:237-263, 'colbert' interaction)
//   StridedTensor.segmented_lookup (CB/search/segmented_lookup.cpp:36-125)
// These take an already materialised similarity matrix / tensor, as the reference operators do; the
// search pipeline itself never materialises it (maxsim.cu).
#include "common.cuh"

namespace plaid {

// One warp per passage.  Lane l keeps the running max of columns l, l+32, ... (nq <= 128), rows are
// read as coalesced nq-float requests.  Zero-initialised max (segmented_maxsim.cpp:58-59), then a
// left-to-right fp32 sum over the columns.
__global__ void __launch_bounds__(256)
segmented_maxsim_kernel(const float* __restrict__ scores, int nq, const int64_t* __restrict__ lengths,
                        const int64_t* __restrict__ row_offsets, int ndocs, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= ndocs) return;
    const int64_t r0 = row_offsets[d], len = lengths[d];
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t r = 0; r < len; r++) {
        const float* row = scores + (r0 + r) * nq;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = lane + 32 * u;
            if (k < nq) m[u] = fmaxf(m[u], __ldg(row + k));
        }
    }
    float s = 0.f;
    for (int k = 0; k < nq; k++) {
        const int u = k >> 5;
        const float mv = (u == 0) ? m[0] : (u == 1) ? m[1] : (u == 2) ? m[2] : m[3];
        s += __shfl_sync(0xffffffffu, mv, k & 31);
    }
    if (lane == 0) out[d] = s;
}

// scores_padded fp32 [n, Ld, Lq], mask u8 [n, Ld]: one warp per passage, lanes over query tokens.
__global__ void __launch_bounds__(256)
score_reduce_kernel(const float* __restrict__ sp, const uint8_t* __restrict__ mask, int64_t n, int Ld, int Lq,
                    float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t d = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= n) return;
    float s = 0.f;
    for (int k0 = 0; k0 < Lq; k0 += 32) {
        const int k = k0 + lane;
        float m = -INFINITY;
        for (int t = 0; t < Ld; t++) {
            float v = (k < Lq) ? __ldg(sp + (d * Ld + t) * Lq + k) : 0.f;
            if (!mask[d * Ld + t]) v = -9999.0f;  // colbert.py:240-241
            m = fmaxf(m, v);
        }
        const int kn = min(32, Lq - k0);
        for (int j = 0; j < kn; j++) s += __shfl_sync(0xffffffffu, m, j);
    }
    if (lane == 0) out[d] = s;
}

// ragged row gather, 16-byte vectorised when row_bytes allows
__global__ void __launch_bounds__(256)
segmented_lookup_kernel(const uint8_t* __restrict__ input, int64_t row_bytes, const int64_t* __restrict__ lengths,
                        const int64_t* __restrict__ offsets, const int64_t* __restrict__ out_offsets, int n,
                        uint8_t* __restrict__ out, int vec16) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int64_t bytes = lengths[i] * row_bytes;
        const uint8_t* src = input + offsets[i] * row_bytes;
        uint8_t* dst = out + out_offsets[i] * row_bytes;
        if (vec16) {
            for (int64_t j = threadIdx.x; j < (bytes >> 4); j += blockDim.x)
                reinterpret_cast<int4*>(dst)[j] = ld_stream_v4(src + (j << 4));
        } else {
            for (int64_t j = threadIdx.x; j < bytes; j += blockDim.x) dst[j] = src[j];
        }
    }
}

}  // namespace plaid

extern "C" int plaid_segmented_maxsim(const float* scores, int nq, const int64_t* lengths, const int64_t* row_offsets,
                                      int ndocs, float* out, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(scores && lengths && row_offsets && out, PLAID_ERR_ARG, "plaid_segmented_maxsim: null pointer");
    PLAID_CHECK_ARG(nq >= 1 && nq <= 128, PLAID_ERR_UNSUPPORTED, "plaid_segmented_maxsim: nq=%d outside [1,128]", nq);
    if (ndocs <= 0) return PLAID_OK;
    segmented_maxsim_kernel<<<(ndocs + 7) / 8, 256, 0, (cudaStream_t)stream>>>(scores, nq, lengths, row_offsets, ndocs, out);
    PLAID_LAUNCH_OK("segmented_maxsim_kernel");
    return PLAID_OK;
}

extern "C" int plaid_colbert_score_reduce(const float* scores_padded, const uint8_t* D_mask, int64_t n, int Ld, int Lq,
                                          float* scores, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(scores_padded && D_mask && scores, PLAID_ERR_ARG, "plaid_colbert_score_reduce: null pointer");
    PLAID_CHECK_ARG(n >= 0 && Ld >= 1 && Lq >= 1, PLAID_ERR_ARG, "plaid_colbert_score_reduce: bad sizes");
    if (n == 0) return PLAID_OK;
    score_reduce_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(scores_padded, D_mask, n, Ld, Lq, scores);
    PLAID_LAUNCH_OK("score_reduce_kernel");
    return PLAID_OK;
}

extern "C" int plaid_segmented_lookup(const uint8_t* input, int64_t row_bytes, const int64_t* lengths,
                                      const int64_t* offsets, const int64_t* out_offsets, int n, uint8_t* out,
                                      void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(input && lengths && offsets && out_offsets && out, PLAID_ERR_ARG, "plaid_segmented_lookup: null pointer");
    PLAID_CHECK_ARG(row_bytes >= 1 && n >= 0, PLAID_ERR_ARG, "plaid_segmented_lookup: bad sizes");
    if (n == 0) return PLAID_OK;
    const int vec16 = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(input) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int grid = n < 148 * 8 ? n : 148 * 8;
    segmented_lookup_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(input, row_bytes, lengths, offsets, out_offsets, n, out,
                                                                   vec16);
    PLAID_LAUNCH_OK("segmented_lookup_kernel");
    return PLAID_OK;
}
