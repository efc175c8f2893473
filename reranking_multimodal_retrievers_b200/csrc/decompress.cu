// decompress.cu -- residual decompression (2/4-bit bucket unpack + centroid add).
//
// Replaces CB/search/decompress_residuals.cpp:27-155 (CPU, byte-at-a-time table walks) and
// CB/indexing/codecs/decompress_residuals.cu (one thread per packed byte, two global RMWs per
// element).  Here a warp pulls 512 B of packed residuals per step with one 128-bit load per lane
// (= 32/16/8/4 whole tokens at 1/2/4/8 bits), parks them in shared memory, and then emits two
// tokens per iteration: each half-warp owns a token, lane h of the half its dimensions 8h..8h+7, so
// the fp16 centroid row is read with one 128-bit load per lane and the output row is written as one
// fully coalesced request.  The reference's two byte tables
// (reversed_bit_map, decompression_lookup_table; CB/indexing/codecs/residual.py:54-89) and
// bucket_weights are folded into one 256-row weight table W[x][l] kept in shared memory.
#include "common.cuh"
#include "decompress.cuh"

namespace plaid {

static constexpr int kDecWarps = 4;

__global__ void build_weight_table_kernel(const float* __restrict__ bucket_weights, const uint8_t* __restrict__ rbm,
                                          const uint8_t* __restrict__ lookup, int nbits, float* __restrict__ W) {
    const int keys = 8 / nbits, x = threadIdx.x;
    if (x >= 256) return;
    const int y = rbm[x];
    for (int l = 0; l < keys; l++) W[x * keys + l] = bucket_weights[lookup[y * keys + l]];
}

__global__ void unpack_codes_kernel(const uint8_t* __restrict__ residuals, int64_t ntokens, int nbits,
                                    const uint8_t* __restrict__ rbm, const uint8_t* __restrict__ lookup,
                                    uint8_t* __restrict__ out) {
    const int keys = 8 / nbits, pd = kDim / keys;
    const int64_t total = ntokens * pd;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = rbm[residuals[i]];
        const int64_t t = i / pd;
        const int k = (int)(i - t * pd);
        for (int l = 0; l < keys; l++) out[t * kDim + k * keys + l] = lookup[y * keys + l];
    }
}

// Decompress the tokens of ONE passage (codes/residual rows [tok0, tok0+len)) with the warps of a CTA.
// A warp pulls 512 B of packed residuals per step (one 128-bit load per lane) into shared memory and then
// emits TWO tokens per iteration: each half-warp owns one token, lane h of the half its dims 8h..8h+7.
// emit(j, v, h, valid) is called by all 32 lanes (it may shuffle); v = the 8 fp32 values, valid = j < len.
template <int NBITS, typename CT, typename Emit>
__device__ __forceinline__ void decompress_passage(int64_t tok0, int len, const uint8_t* __restrict__ residuals,
                                                   const int32_t* __restrict__ codes, const CT* __restrict__ centroids,
                                                   int C, const float* sW, uint8_t* s_stage, Emit emit) {
    constexpr int PB = 16 * NBITS;   // packed bytes per token
    constexpr int TB = 512 / PB;     // tokens per 512-byte warp batch
    const int lane = threadIdx.x & 31, warp = warp_index(), nw = blockDim.x >> 5;
    const int h = lane & 15, half = lane >> 4;
    uint8_t* stage = s_stage + warp * 512;
    for (int t0 = warp * TB; t0 < len; t0 += nw * TB) {
        const int nt = min(TB, len - t0);
        if (lane * 16 < nt * PB)  // one 128-bit streaming load per lane
            reinterpret_cast<int4*>(stage)[lane] = ld_stream_v4(residuals + (tok0 + t0) * PB + lane * 16);
        int code = (lane < nt) ? ld_stream_s32(codes + tok0 + t0 + lane) : 0;
        __syncwarp();
        for (int j0 = 0; j0 < nt; j0 += 2) {
            const int j = min(j0 + half, nt - 1);           // the odd tail re-does the last token, masked out below
            int c = __shfl_sync(0xffffffffu, code, j);
            c = min(max(c, 0), C - 1);
            float w[8], e[8], v[8];
            token_weights8<NBITS>(smem_u32(stage) + j * PB, smem_u32(sW), h, w);
            load_centroid8(centroids + (size_t)c * kDim, h, e);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = w[i] + e[i];
            emit(t0 + j, v, h, j0 + half < nt);
        }
        __syncwarp();
    }
}

template <int NBITS>
__device__ __forceinline__ void load_weight_table(const float* __restrict__ W, float* sW) {
    constexpr int n = 256 * (8 / NBITS);
    for (int i = threadIdx.x; i < n; i += blockDim.x) sW[i] = W[i];
    __syncthreads();
}

// ---- operator form: fp32 out, rows packed in pid order (decompress_residuals.cpp semantics) ----
template <int NBITS, typename CT>
__global__ void __launch_bounds__(kDecWarps * 32)
decompress_packed_kernel(const int32_t* __restrict__ pids, int npids, const int64_t* __restrict__ offsets,
                         const int64_t* __restrict__ out_offsets, const float* __restrict__ W,
                         const uint8_t* __restrict__ residuals, const int32_t* __restrict__ codes,
                         const CT* __restrict__ centroids, int C, float* __restrict__ out) {
    __shared__ __align__(16) float sW[256 * (8 / NBITS)];
    __shared__ __align__(16) uint8_t s_stage[kDecWarps * 512];
    load_weight_table<NBITS>(W, sW);
    for (int i = blockIdx.x; i < npids; i += gridDim.x) {
        const int pid = pids[i];
        const int64_t tok0 = offsets[pid];
        const int len = (int)(offsets[pid + 1] - tok0);
        float* dst = out + out_offsets[i] * kDim;
        decompress_passage<NBITS, CT>(tok0, len, residuals, codes, centroids, C, sW, s_stage,
                                      [&](int j, const float (&v)[8], int h, bool valid) {
                                          if (!valid) return;
                                          float4* o = reinterpret_cast<float4*>(dst + (size_t)j * kDim) + 2 * h;
                                          o[0] = make_float4(v[0], v[1], v[2], v[3]);
                                          o[1] = make_float4(v[4], v[5], v[6], v[7]);
                                      });
    }
}

// ---- pipeline form: + L2 normalise + bf16, rows placed by per-query token offsets ----
template <int NBITS, typename CT>
__global__ void __launch_bounds__(kDecWarps * 32)
decompress_normalize_kernel(const int32_t* __restrict__ pids, const int32_t* __restrict__ counts, int pid_stride,
                            const int32_t* __restrict__ tok_offsets, int tok_stride, const int64_t* __restrict__ offsets,
                            const float* __restrict__ W, const uint8_t* __restrict__ residuals,
                            const int32_t* __restrict__ codes, const CT* __restrict__ centroids, int C,
                            __nv_bfloat16* __restrict__ D) {
    __shared__ __align__(16) float sW[256 * (8 / NBITS)];
    __shared__ __align__(16) uint8_t s_stage[kDecWarps * 512];
    const int b = blockIdx.y;
    const int n = min(counts[b], pid_stride);
    if ((int)blockIdx.x >= n) return;
    load_weight_table<NBITS>(W, sW);
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int pid = pids[(size_t)b * pid_stride + i];
        const int64_t tok0 = offsets[pid];
        const int len = (int)(offsets[pid + 1] - tok0);
        __nv_bfloat16* dst = D + ((size_t)b * tok_stride + tok_offsets[(size_t)b * (pid_stride + 1) + i]) * kDim;
        decompress_passage<NBITS, CT>(tok0, len, residuals, codes, centroids, C, sW, s_stage,
                                      [&](int j, const float (&v)[8], int h, bool valid) {
                                          float ss = 0.f;
#pragma unroll
                                          for (int i = 0; i < 8; i++) ss = fmaf(v[i], v[i], ss);
#pragma unroll
                                          for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                                          // F.normalize: x / max(||x||_2, 1e-12) = x * rsqrt(max(||x||^2, 1e-24))
                                          // (index_storage.py:175)
                                          const float inv = rsqrtf(fmaxf(ss, 1e-24f));
                                          uint32_t pk[4];
#pragma unroll
                                          for (int i = 0; i < 4; i++) {
                                              __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i] * inv, v[2 * i + 1] * inv);
                                              pk[i] = *reinterpret_cast<uint32_t*>(&t);
                                          }
                                          if (valid)
                                              reinterpret_cast<uint4*>(dst + (size_t)j * kDim)[h] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                      });
        // aligned layout: rows between this passage's last token and the next passage's first are zero
        const int span = tok_offsets[(size_t)b * (pid_stride + 1) + i + 1] - tok_offsets[(size_t)b * (pid_stride + 1) + i];
        for (int r = len + (threadIdx.x >> 5); r < span; r += (blockDim.x >> 5))
            reinterpret_cast<uint2*>(dst + (size_t)r * kDim)[threadIdx.x & 31] = make_uint2(0u, 0u);
    }
}


// ---- pipeline form in fp16 (the arithmetic of the fused MaxSim kernel, bit for bit): centroid + weight in half,
// L2 normalise, fp16 rows placed by per-query token offsets; pad rows of the aligned layout are zero ----
template <int NBITS>
__global__ void __launch_bounds__(kDecWarps * 32)
decompress_normalize_f16_kernel(const int32_t* __restrict__ pids, const int32_t* __restrict__ counts, int pid_stride,
                                const int32_t* __restrict__ tok_offsets, int tok_stride,
                                const int64_t* __restrict__ offsets, const float* __restrict__ W,
                                const uint8_t* __restrict__ residuals, const int32_t* __restrict__ codes,
                                const __half* __restrict__ centroids, __half* __restrict__ D) {
    __shared__ __align__(128) uint8_t sLUT[kLutBytes];
    __shared__ __align__(16) uint8_t s_stage[kDecWarps * 512];
    constexpr int PB = 16 * NBITS, TB = 512 / PB;
    const int b = blockIdx.y;
    const int n = min(counts[b], pid_stride);
    if ((int)blockIdx.x >= n) return;
    lut_fill_f16<NBITS>(W, sLUT);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = warp_index(), nw = blockDim.x >> 5;
    const int q = lane & 7, tsub = lane >> 3;            // quarter warp per token, lane q <- chunk q of both k-halves
    const uint32_t lut_sa = lut_lane_base<NBITS>(smem_u32(sLUT), lane);
    const uint32_t stage_sa = smem_u32(s_stage + warp * 512);
    const char* cent_q = reinterpret_cast<const char*>(centroids) + q * 16;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int pid = pids[(size_t)b * pid_stride + i];
        const int64_t tok0 = offsets[pid];
        const int len = (int)(offsets[pid + 1] - tok0);
        __half* dst = D + ((size_t)b * tok_stride + tok_offsets[(size_t)b * (pid_stride + 1) + i]) * kDim;
        for (int t0 = warp * TB; t0 < len; t0 += nw * TB) {
            const int nt = min(TB, len - t0);
            int4 r = make_int4(0, 0, 0, 0);     // rows past nt stay zero: they decode to finite values nobody stores
            if (lane * 16 < nt * PB) r = ld_stream_v4(residuals + (tok0 + t0) * PB + lane * 16);
            sts_v4u32(stage_sa + lane * 16, r.x, r.y, r.z, r.w);
            const int code = (lane < nt) ? ld_stream_s32(codes + tok0 + t0 + lane) : 0;
            __syncwarp();
            for (int j0 = 0; j0 < nt; j0 += 4) {
                const int j = j0 + tsub;
                const unsigned c = (unsigned)__shfl_sync(0xffffffffu, code, j);
                const uint4* crow = centroid_row(cent_q, c);
                const uint4 clo = __ldg(crow), chi = __ldg(crow + 8);
                uint32_t wlo[4], whi[4];
                __half2 v[8];
                token_weights_h8<NBITS>(stage_sa + j * PB, lut_sa, q, wlo);
                token_weights_h8<NBITS>(stage_sa + j * PB, lut_sa, q + 8, whi);
                const float ss = quarter_sum(token_sum_h16(clo, chi, wlo, whi, v));
                uint4 olo, ohi;
                token_scale_h16(v, ss, true, olo, ohi);
                if (j < nt) {
                    uint4* o = reinterpret_cast<uint4*>(dst + (size_t)(t0 + j) * kDim);
                    o[q] = olo;
                    o[q + 8] = ohi;
                }
            }
            __syncwarp();
        }
        const int span = tok_offsets[(size_t)b * (pid_stride + 1) + i + 1] - tok_offsets[(size_t)b * (pid_stride + 1) + i];
        for (int r = len + (threadIdx.x >> 5); r < span; r += (blockDim.x >> 5))
            reinterpret_cast<uint2*>(dst + (size_t)r * kDim)[threadIdx.x & 31] = make_uint2(0u, 0u);
    }
}

// ---- token form (no pid indirection): the operator the reference binds as ResidualCodec.decompress_residuals
// (CB/indexing/codecs/residual.py:115, codecs/decompress_residuals.cu:8-75): out[t, d] = half(weight) + half(centroid),
// ONE half add per element like the reference kernel (`output = bucket_weights[..]; output += centroids[code][d]`).
// Half a warp owns a token, lane h its dimensions 8h..8h+7 = NBITS whole residual bytes; one 128-bit centroid load
// and one 128-bit store per lane.  normalize != 0 adds ResidualCodec.decompress's `F.normalize(..).half()` on top
// (residual.py:272-273: fp32 sum of squares, the norm rounded to half, a half division per element).
template <int NBITS>
__global__ void __launch_bounds__(256)
decompress_tokens_f16_kernel(const uint8_t* __restrict__ residuals, const int32_t* __restrict__ codes, int64_t n,
                             const float* __restrict__ W, const __half* __restrict__ centroids, int C, int normalize,
                             __half* __restrict__ out) {
    constexpr int KEYS = 8 / NBITS;
    __shared__ __half sW[256 * KEYS];
    for (int i = threadIdx.x; i < 256 * KEYS; i += blockDim.x) sW[i] = __float2half_rn(W[i]);   // bucket_weights.half()
    __syncthreads();
    const int lane = threadIdx.x & 31, h = lane & 15, half = lane >> 4;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_even = (n + 1) & ~int64_t(1);           // both halves of a warp run the same number of iterations
    for (int64_t t = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2) + half; t < n_even; t += 2 * warps) {
        const bool valid = t < n;
        const int64_t tt = valid ? t : n - 1;
        int code = codes[tt];
        code = min(max(code, 0), C - 1);
        const uint8_t* src = residuals + tt * (16 * NBITS) + h * NBITS;
        uint8_t x[NBITS];
#pragma unroll
        for (int k = 0; k < NBITS; k++) x[k] = src[k];
        const uint4 craw = __ldg(reinterpret_cast<const uint4*>(centroids + (size_t)code * kDim) + h);
        const __half2* c2 = reinterpret_cast<const __half2*>(&craw);
        __half v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const __half w = sW[x[i / KEYS] * KEYS + (i % KEYS)];
            const __half c = (i & 1) ? __high2half(c2[i >> 1]) : __low2half(c2[i >> 1]);
            v[i] = __hadd(w, c);
        }
        if (normalize) {
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) ss = fmaf(__half2float(v[i]), __half2float(v[i]), ss);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float nrm = __half2float(__float2half_rn(sqrtf(ss)));    // half norm; clamp_min(1e-12) is 0 in half
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __float2half_rn(__half2float(v[i]) / nrm);
        }
        if (valid) {
            uint4 o;
            __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int i = 0; i < 4; i++) o2[i] = __halves2half2(v[2 * i], v[2 * i + 1]);
            reinterpret_cast<uint4*>(out + t * kDim)[h] = o;
        }
    }
}

// ---- per-token scale factors for the fused MaxSim: inv[t] = half(1 / max(||centroid + weights||, 1e-12)) in exactly the
// arithmetic (and reduction order) of the fp16 pipeline above, computed ONCE when an index is loaded.  With the table
// the fused kernel's decompressors skip the sum of squares, its three shuffles and the rsqrt per token and still
// build bit-identical tiles.
template <int NBITS>
__global__ void __launch_bounds__(kDecWarps * 32)
token_inv_norms_kernel(const uint8_t* __restrict__ residuals, const int32_t* __restrict__ codes, int64_t n,
                       const float* __restrict__ W, const __half* __restrict__ centroids, int C, __half* __restrict__ inv) {
    __shared__ __align__(128) uint8_t sLUT[kLutBytes];
    __shared__ __align__(16) uint8_t s_stage[kDecWarps * 512];
    constexpr int PB = 16 * NBITS, TB = 512 / PB;
    lut_fill_f16<NBITS>(W, sLUT);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = warp_index(), nw = blockDim.x >> 5;
    const int q = lane & 7, tsub = lane >> 3;
    const uint32_t lut_sa = lut_lane_base<NBITS>(smem_u32(sLUT), lane);
    const uint32_t stage_sa = smem_u32(s_stage + warp * 512);
    const char* cent_q = reinterpret_cast<const char*>(centroids) + q * 16;
    const int64_t stride = (int64_t)gridDim.x * nw * TB;
    for (int64_t t0 = ((int64_t)blockIdx.x * nw + warp) * TB; t0 < n; t0 += stride) {
        const int nt = (int)min((int64_t)TB, n - t0);
        int4 r = make_int4(0, 0, 0, 0);
        if (lane * 16 < nt * PB) r = ld_stream_v4(residuals + t0 * PB + lane * 16);
        sts_v4u32(stage_sa + lane * 16, r.x, r.y, r.z, r.w);
        int code = (lane < nt) ? ld_stream_s32(codes + t0 + lane) : 0;
        code = min(max(code, 0), C - 1);
        __syncwarp();
        for (int j0 = 0; j0 < nt; j0 += 4) {
            const int j = j0 + tsub;
            const unsigned c = (unsigned)__shfl_sync(0xffffffffu, code, j);
            const uint4* crow = centroid_row(cent_q, c);
            const uint4 clo = __ldg(crow), chi = __ldg(crow + 8);
            uint32_t wlo[4], whi[4];
            __half2 v[8];
            token_weights_h8<NBITS>(stage_sa + j * PB, lut_sa, q, wlo);
            token_weights_h8<NBITS>(stage_sa + j * PB, lut_sa, q + 8, whi);
            const float ss = quarter_sum(token_sum_h16(clo, chi, wlo, whi, v));
            if (j < nt && q == 0) inv[t0 + j] = token_inv_h16(ss);
        }
        __syncwarp();
    }
}

// per query: exclusive prefix sums of passage lengths
__global__ void __launch_bounds__(256)
doc_token_offsets_kernel(const int32_t* __restrict__ pids, const int32_t* __restrict__ counts, int pid_stride,
                         const int64_t* __restrict__ offsets, int align, int32_t* __restrict__ tok_offsets) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = warp_index();
    const int n = min(counts[b], pid_stride);
    int32_t* out = tok_offsets + (size_t)b * (pid_stride + 1);
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < pid_stride; i0 += blockDim.x) {
        const int i = i0 + tid;
        int len = 0;
        if (i < n) {
            const int pid = pids[(size_t)b * pid_stride + i];
            len = (int)(offsets[pid + 1] - offsets[pid]);
            len = (len + align - 1) / align * align;   // every passage starts on an `align`-token boundary
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; w++) wbase += s_warp[w];
        const int excl = s_base + wbase + incl - len;
        if (i < pid_stride) out[i] = excl;
        __syncthreads();
        if (tid == blockDim.x - 1) s_base = excl + len;
        __syncthreads();
    }
    if (tid == 0) out[pid_stride] = s_base;
}

template <typename CT>
static int launch_packed(int nbits, const int32_t* pids, int npids, const int64_t* offsets, const int64_t* out_offsets,
                         const float* W, const uint8_t* residuals, const int32_t* codes, const CT* centroids, int C,
                         float* out, cudaStream_t st) {
    int grid = npids < 148 * 16 ? npids : 148 * 16;
#define PLAID_DEC_CASE(NB)                                                                                              \
    case NB:                                                                                                            \
        decompress_packed_kernel<NB, CT><<<grid, kDecWarps * 32, 0, st>>>(pids, npids, offsets, out_offsets, W, residuals, \
                                                                         codes, centroids, C, out);                   \
        break;
    switch (nbits) {
        PLAID_DEC_CASE(1) PLAID_DEC_CASE(2) PLAID_DEC_CASE(4) PLAID_DEC_CASE(8)
        default:
            set_error("decompress: nbits=%d not in {1,2,4,8}", nbits);
            return PLAID_ERR_UNSUPPORTED;
    }
#undef PLAID_DEC_CASE
    PLAID_LAUNCH_OK("decompress_packed_kernel");
    return PLAID_OK;
}

template <typename CT>
static int launch_normalize(int nbits, const int32_t* pids, const int32_t* counts, int B, int pid_stride,
                            const int32_t* tok_offsets, int tok_stride, const int64_t* offsets, const float* W,
                            const uint8_t* residuals, const int32_t* codes, const CT* centroids, int C,
                            __nv_bfloat16* D, cudaStream_t st) {
    dim3 grid(pid_stride, B);
#define PLAID_DEC_CASE(NB)                                                                                      \
    case NB:                                                                                                    \
        decompress_normalize_kernel<NB, CT><<<grid, kDecWarps * 32, 0, st>>>(pids, counts, pid_stride, tok_offsets, \
                                                                            tok_stride, offsets, W, residuals, codes, \
                                                                            centroids, C, D);                    \
        break;
    switch (nbits) {
        PLAID_DEC_CASE(1) PLAID_DEC_CASE(2) PLAID_DEC_CASE(4) PLAID_DEC_CASE(8)
        default:
            set_error("decompress: nbits=%d not in {1,2,4,8}", nbits);
            return PLAID_ERR_UNSUPPORTED;
    }
#undef PLAID_DEC_CASE
    PLAID_LAUNCH_OK("decompress_normalize_kernel");
    return PLAID_OK;
}

}  // namespace plaid

extern "C" int plaid_build_weight_table(const float* bucket_weights, const uint8_t* reversed_bit_map,
                                        const uint8_t* lookup, int nbits, float* W, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(bucket_weights && reversed_bit_map && lookup && W, PLAID_ERR_ARG, "plaid_build_weight_table: null pointer");
    PLAID_CHECK_ARG(nbits == 1 || nbits == 2 || nbits == 4 || nbits == 8, PLAID_ERR_UNSUPPORTED,
                    "plaid_build_weight_table: nbits=%d not in {1,2,4,8}", nbits);
    build_weight_table_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(bucket_weights, reversed_bit_map, lookup, nbits, W);
    PLAID_LAUNCH_OK("build_weight_table_kernel");
    return PLAID_OK;
}

extern "C" int plaid_unpack_residual_codes(const uint8_t* residuals, int64_t ntokens, int nbits,
                                           const uint8_t* reversed_bit_map, const uint8_t* lookup, uint8_t* out,
                                           void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(residuals && reversed_bit_map && lookup && out && ntokens >= 0, PLAID_ERR_ARG,
                    "plaid_unpack_residual_codes: bad argument");
    PLAID_CHECK_ARG(nbits == 1 || nbits == 2 || nbits == 4 || nbits == 8, PLAID_ERR_UNSUPPORTED,
                    "plaid_unpack_residual_codes: nbits=%d not in {1,2,4,8}", nbits);
    if (ntokens == 0) return PLAID_OK;
    unpack_codes_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(residuals, ntokens, nbits, reversed_bit_map, lookup, out);
    PLAID_LAUNCH_OK("unpack_codes_kernel");
    return PLAID_OK;
}

extern "C" int plaid_decompress_residuals(const int32_t* pids, int npids, const int64_t* offsets,
                                          const int64_t* out_offsets, const float* W, const uint8_t* residuals,
                                          const int32_t* codes, const float* centroids, int C, int nbits, float* out,
                                          void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(pids && offsets && out_offsets && W && residuals && codes && centroids && out, PLAID_ERR_ARG,
                    "plaid_decompress_residuals: null pointer");
    PLAID_CHECK_ARG(npids >= 0 && C > 0, PLAID_ERR_ARG, "plaid_decompress_residuals: bad sizes");
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(residuals) & 15) == 0 && (reinterpret_cast<uintptr_t>(centroids) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                    PLAID_ERR_ARG, "plaid_decompress_residuals: residuals/centroids/out must be 16-byte aligned");
    if (npids == 0) return PLAID_OK;
    return launch_packed<float>(nbits, pids, npids, offsets, out_offsets, W, residuals, codes, centroids, C, out,
                                (cudaStream_t)stream);
}

extern "C" int plaid_doc_token_offsets(const int32_t* pids, const int32_t* counts, int B, int pid_stride,
                                       const int64_t* offsets, int align, int32_t* tok_offsets, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(pids && counts && offsets && tok_offsets && B >= 0 && pid_stride >= 1 && align >= 1, PLAID_ERR_ARG,
                    "plaid_doc_token_offsets: bad argument");
    if (B == 0) return PLAID_OK;
    doc_token_offsets_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(pids, counts, pid_stride, offsets, align, tok_offsets);
    PLAID_LAUNCH_OK("doc_token_offsets_kernel");
    return PLAID_OK;
}

extern "C" int plaid_decompress_normalize_bf16(const int32_t* pids, const int32_t* counts, int B, int pid_stride,
                                               const int32_t* tok_offsets, int tok_stride, const int64_t* offsets,
                                               const float* W, const uint8_t* residuals, const int32_t* codes,
                                               const void* centroids, int centroids_are_f16, int C, int nbits,
                                               void* D_bf16, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(pids && counts && tok_offsets && offsets && W && residuals && codes && centroids && D_bf16, PLAID_ERR_ARG,
                    "plaid_decompress_normalize_bf16: null pointer");
    PLAID_CHECK_ARG(B >= 0 && pid_stride >= 1 && tok_stride >= 1 && C > 0, PLAID_ERR_ARG,
                    "plaid_decompress_normalize_bf16: bad sizes");
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(residuals) & 15) == 0 && (reinterpret_cast<uintptr_t>(centroids) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(D_bf16) & 15) == 0,
                    PLAID_ERR_ARG, "plaid_decompress_normalize_bf16: residuals/centroids/D must be 16-byte aligned");
    if (B == 0) return PLAID_OK;
    PLAID_CHECK_ARG(B <= 65535, PLAID_ERR_UNSUPPORTED, "plaid_decompress_normalize_bf16: B=%d > 65535 per call", B);
    __nv_bfloat16* D = reinterpret_cast<__nv_bfloat16*>(D_bf16);
    if (centroids_are_f16)
        return launch_normalize<__half>(nbits, pids, counts, B, pid_stride, tok_offsets, tok_stride, offsets, W, residuals,
                                        codes, reinterpret_cast<const __half*>(centroids), C, D, (cudaStream_t)stream);
    return launch_normalize<float>(nbits, pids, counts, B, pid_stride, tok_offsets, tok_stride, offsets, W, residuals, codes,
                                   reinterpret_cast<const float*>(centroids), C, D, (cudaStream_t)stream);
}

extern "C" int plaid_decompress_normalize_f16(const int32_t* pids, const int32_t* counts, int B, int pid_stride,
                                              const int32_t* tok_offsets, int tok_stride, const int64_t* offsets,
                                              const float* W, const uint8_t* residuals, const int32_t* codes,
                                              const void* centroids_f16, int C, int nbits, void* D_f16, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(pids && counts && tok_offsets && offsets && W && residuals && codes && centroids_f16 && D_f16, PLAID_ERR_ARG,
                    "plaid_decompress_normalize_f16: null pointer");
    PLAID_CHECK_ARG(B >= 0 && pid_stride >= 1 && tok_stride >= 1 && C > 0, PLAID_ERR_ARG,
                    "plaid_decompress_normalize_f16: bad sizes");
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(residuals) & 15) == 0 && (reinterpret_cast<uintptr_t>(centroids_f16) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(D_f16) & 15) == 0,
                    PLAID_ERR_ARG, "plaid_decompress_normalize_f16: residuals/centroids/D must be 16-byte aligned");
    if (B == 0) return PLAID_OK;
    PLAID_CHECK_ARG(B <= 65535, PLAID_ERR_UNSUPPORTED, "plaid_decompress_normalize_f16: B=%d > 65535 per call", B);
    const __half* cent = reinterpret_cast<const __half*>(centroids_f16);
    __half* D = reinterpret_cast<__half*>(D_f16);
    dim3 grid(pid_stride, B);
    cudaStream_t st = (cudaStream_t)stream;
    switch (nbits) {
        case 1: decompress_normalize_f16_kernel<1><<<grid, kDecWarps * 32, 0, st>>>(pids, counts, pid_stride, tok_offsets, tok_stride, offsets, W, residuals, codes, cent, D); break;
        case 2: decompress_normalize_f16_kernel<2><<<grid, kDecWarps * 32, 0, st>>>(pids, counts, pid_stride, tok_offsets, tok_stride, offsets, W, residuals, codes, cent, D); break;
        case 4: decompress_normalize_f16_kernel<4><<<grid, kDecWarps * 32, 0, st>>>(pids, counts, pid_stride, tok_offsets, tok_stride, offsets, W, residuals, codes, cent, D); break;
        case 8: decompress_normalize_f16_kernel<8><<<grid, kDecWarps * 32, 0, st>>>(pids, counts, pid_stride, tok_offsets, tok_stride, offsets, W, residuals, codes, cent, D); break;
        default:
            set_error("plaid_decompress_normalize_f16: nbits=%d not in {1,2,4,8}", nbits);
            return PLAID_ERR_UNSUPPORTED;
    }
    PLAID_LAUNCH_OK("decompress_normalize_f16_kernel");
    return PLAID_OK;
}

extern "C" int plaid_decompress_tokens_f16(const uint8_t* residuals, const int32_t* codes, int64_t n, const float* W,
                                           const void* centroids_f16, int C, int nbits, int normalize, void* out_f16,
                                           void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(residuals && codes && W && centroids_f16 && out_f16, PLAID_ERR_ARG, "plaid_decompress_tokens_f16: null pointer");
    PLAID_CHECK_ARG(n >= 0 && C > 0, PLAID_ERR_ARG, "plaid_decompress_tokens_f16: bad sizes");
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(centroids_f16) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_f16) & 15) == 0,
                    PLAID_ERR_ARG, "plaid_decompress_tokens_f16: centroids/out must be 16-byte aligned");
    if (n == 0) return PLAID_OK;
    int64_t blocks = (n + 15) / 16;               // 8 warps x 2 tokens per CTA and pass
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    const __half* cent = reinterpret_cast<const __half*>(centroids_f16);
    __half* out = reinterpret_cast<__half*>(out_f16);
    switch (nbits) {
        case 1: decompress_tokens_f16_kernel<1><<<(int)blocks, 256, 0, st>>>(residuals, codes, n, W, cent, C, normalize, out); break;
        case 2: decompress_tokens_f16_kernel<2><<<(int)blocks, 256, 0, st>>>(residuals, codes, n, W, cent, C, normalize, out); break;
        case 4: decompress_tokens_f16_kernel<4><<<(int)blocks, 256, 0, st>>>(residuals, codes, n, W, cent, C, normalize, out); break;
        case 8: decompress_tokens_f16_kernel<8><<<(int)blocks, 256, 0, st>>>(residuals, codes, n, W, cent, C, normalize, out); break;
        default:
            set_error("plaid_decompress_tokens_f16: nbits=%d not in {1,2,4,8}", nbits);
            return PLAID_ERR_UNSUPPORTED;
    }
    PLAID_LAUNCH_OK("decompress_tokens_f16_kernel");
    return PLAID_OK;
}

extern "C" int plaid_token_inv_norms(const uint8_t* residuals, const int32_t* codes, int64_t n, const float* W,
                                     const void* centroids_f16, int C, int nbits, void* inv_f16, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(residuals && codes && W && centroids_f16 && inv_f16, PLAID_ERR_ARG, "plaid_token_inv_norms: null pointer");
    PLAID_CHECK_ARG(n >= 0 && C > 0, PLAID_ERR_ARG, "plaid_token_inv_norms: bad sizes");
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(residuals) & 15) == 0 && (reinterpret_cast<uintptr_t>(centroids_f16) & 15) == 0,
                    PLAID_ERR_ARG, "plaid_token_inv_norms: residuals/centroids must be 16-byte aligned");
    if (n == 0) return PLAID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const __half* cent = reinterpret_cast<const __half*>(centroids_f16);
    __half* inv = reinterpret_cast<__half*>(inv_f16);
    const int grid = 148 * 8;
    switch (nbits) {
        case 1: token_inv_norms_kernel<1><<<grid, kDecWarps * 32, 0, st>>>(residuals, codes, n, W, cent, C, inv); break;
        case 2: token_inv_norms_kernel<2><<<grid, kDecWarps * 32, 0, st>>>(residuals, codes, n, W, cent, C, inv); break;
        case 4: token_inv_norms_kernel<4><<<grid, kDecWarps * 32, 0, st>>>(residuals, codes, n, W, cent, C, inv); break;
        case 8: token_inv_norms_kernel<8><<<grid, kDecWarps * 32, 0, st>>>(residuals, codes, n, W, cent, C, inv); break;
        default:
            set_error("plaid_token_inv_norms: nbits=%d not in {1,2,4,8}", nbits);
            return PLAID_ERR_UNSUPPORTED;
    }
    PLAID_LAUNCH_OK("token_inv_norms_kernel");
    return PLAID_OK;
}
