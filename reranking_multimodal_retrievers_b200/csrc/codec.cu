// codec.cu -- index-build side of the residual codec: residual = embedding - centroid[code],
// bucketize against the bucket cutoffs, write the bucket index bit by bit and pack 8 bits per byte.
//
// Replaces ResidualCodec.compress / binarize (CB/indexing/codecs/residual.py:169-203: fp32 subtraction,
// torch.bucketize(right=False), `>> arange_bits & 1`, np.packbits / codecs/packbits.cu:10-57).  Bit layout of the
// reference: dimension d contributes nbits consecutive bits, least significant bit of its bucket FIRST, and the
// flat bit string is packed MSB-first -- which is why decompression goes through reversed_bit_map.
// Half a warp owns a token (lane h: dimensions 8h..8h+7 = nbits whole bytes), so the embedding row is read with two
// 128-bit loads per lane, the fp16 centroid row with one, and the 16*nbits output bytes of a token leave as one
// contiguous run.  HBM-bound: 516 B in (+ 256 B of centroid row from L2), 16*nbits B out per token.
#include "common.cuh"
#include <cuda_fp16.h>

namespace plaid {

template <int NBITS>
__global__ void __launch_bounds__(256)
compress_residuals_kernel(const float* __restrict__ embs, const int32_t* __restrict__ codes, const __half* __restrict__ centroids,
                          const float* __restrict__ cutoffs, int64_t n, int C, uint8_t* __restrict__ out, int* __restrict__ bad_code) {
    __shared__ float s_cut[256];
    constexpr int NCUT = (1 << NBITS) - 1;
    for (int i = threadIdx.x; i < NCUT; i += blockDim.x) s_cut[i] = cutoffs[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, h = lane & 15, half = lane >> 4;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2) + half; t < n; t += 2 * warps) {
        const int code = codes[t];
        if (code < 0 || code >= C) {              // the reference would index out of bounds
            if (h == 0) atomicExch(bad_code, 1);
            continue;
        }
        const float4 e0 = __ldcs(reinterpret_cast<const float4*>(embs + t * kDim) + 2 * h);
        const float4 e1 = __ldcs(reinterpret_cast<const float4*>(embs + t * kDim) + 2 * h + 1);
        const uint4 craw = __ldg(reinterpret_cast<const uint4*>(centroids + (size_t)code * kDim) + h);
        const uint32_t cu[4] = {craw.x, craw.y, craw.z, craw.w};
        const float e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
        uint64_t bits = 0;                        // this lane's 8 * NBITS bits, first bit in the MSB of the first byte
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float2 c2 = __half22float2(*reinterpret_cast<const __half2*>(&cu[i >> 1]));
            const float r = e[i] - ((i & 1) ? c2.y : c2.x);          // fp32, as residual.py:179
            int bucket = 0;                       // torch.bucketize(right=False): number of cutoffs strictly below r
            if constexpr (NBITS <= 4) {
#pragma unroll
                for (int k = 0; k < NCUT; k++) bucket += (s_cut[k] < r) ? 1 : 0;
            } else {
                int lo = 0, hi = NCUT;            // cutoffs are sorted
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_cut[mid] < r) lo = mid + 1; else hi = mid;
                }
                bucket = lo;
            }
#pragma unroll
            for (int b = 0; b < NBITS; b++)       // bit b of the bucket goes to flat position i*NBITS + b
                bits |= (uint64_t)((bucket >> b) & 1) << (8 * NBITS - 1 - (i * NBITS + b));
        }
        uint8_t* dst = out + t * (16 * NBITS) + h * NBITS;
#pragma unroll
        for (int k = 0; k < NBITS; k++) dst[k] = (uint8_t)(bits >> (8 * (NBITS - 1 - k)));
    }
}

// ResidualCodec.packbits (CB/indexing/codecs/residual.py:130, codecs/packbits.cu:10-57): a flat u8 array of 0/1 flags
// (any non-zero byte counts as 1, like the reference's __ballot_sync on the byte) -> one byte per 8 flags, first flag
// in the most significant bit (np.packbits order).  A thread turns 8 flags (one 64-bit load) into one byte.
__global__ void __launch_bounds__(256)
packbits_kernel(const uint8_t* __restrict__ bits, int64_t nbytes_out, uint8_t* __restrict__ packed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes_out; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 w = __ldcs(reinterpret_cast<const uint2*>(bits) + i);
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            o |= (((w.x >> (8 * j)) & 0xffu) ? 1u : 0u) << (7 - j);
            o |= (((w.y >> (8 * j)) & 0xffu) ? 1u : 0u) << (3 - j);
        }
        packed[i] = (uint8_t)o;
    }
}

}  // namespace plaid

extern "C" int plaid_compress_residuals(const float* embs, const int32_t* codes, const void* centroids_f16,
                                        const float* bucket_cutoffs, int64_t n, int C, int nbits, uint8_t* residuals,
                                        int* bad_code_flag, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(embs && codes && centroids_f16 && bucket_cutoffs && residuals && bad_code_flag, PLAID_ERR_ARG,
                    "plaid_compress_residuals: null pointer");
    PLAID_CHECK_ARG(n >= 0 && C > 0, PLAID_ERR_ARG, "plaid_compress_residuals: bad sizes");
    PLAID_CHECK_ARG(nbits == 1 || nbits == 2 || nbits == 4 || nbits == 8, PLAID_ERR_UNSUPPORTED,
                    "plaid_compress_residuals: nbits=%d not in {1,2,4,8}", nbits);
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(embs) & 15) == 0 && (reinterpret_cast<uintptr_t>(centroids_f16) & 15) == 0,
                    PLAID_ERR_ARG, "plaid_compress_residuals: embs/centroids must be 16-byte aligned");
    if (n == 0) return PLAID_OK;
    int64_t blocks = (n + 15) / 16;               // 8 warps x 2 tokens per CTA and pass
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    const __half* cent = reinterpret_cast<const __half*>(centroids_f16);
    switch (nbits) {
        case 1: compress_residuals_kernel<1><<<(int)blocks, 256, 0, st>>>(embs, codes, cent, bucket_cutoffs, n, C, residuals, bad_code_flag); break;
        case 2: compress_residuals_kernel<2><<<(int)blocks, 256, 0, st>>>(embs, codes, cent, bucket_cutoffs, n, C, residuals, bad_code_flag); break;
        case 4: compress_residuals_kernel<4><<<(int)blocks, 256, 0, st>>>(embs, codes, cent, bucket_cutoffs, n, C, residuals, bad_code_flag); break;
        default: compress_residuals_kernel<8><<<(int)blocks, 256, 0, st>>>(embs, codes, cent, bucket_cutoffs, n, C, residuals, bad_code_flag); break;
    }
    PLAID_LAUNCH_OK("compress_residuals_kernel");
    return PLAID_OK;
}

extern "C" int plaid_packbits(const uint8_t* bits, int64_t nflags, uint8_t* packed, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(bits && packed && nflags >= 0, PLAID_ERR_ARG, "plaid_packbits: bad argument");
    PLAID_CHECK_ARG((nflags % 8) == 0, PLAID_ERR_ARG, "plaid_packbits: %lld flags is not a multiple of 8", (long long)nflags);
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(bits) & 7) == 0, PLAID_ERR_ARG, "plaid_packbits: bits must be 8-byte aligned");
    if (nflags == 0) return PLAID_OK;
    const int64_t nout = nflags / 8;
    int64_t blocks = (nout + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    packbits_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(bits, nout, packed);
    PLAID_LAUNCH_OK("packbits_kernel");
    return PLAID_OK;
}
