// filter.cu -- centroid-only passage filtering and (score, pid) selection.
//
// Replaces CB/search/filter_pids.cpp (pthreads, priority queues) and the final
// `scores.sort(descending=True)` of CB/search/index_storage.py:95-96.
//
// approx_scores: one warp per candidate passage, lane = query token (nq <= 32 = warp width).  The
//   passage's centroid codes are read 32 at a time (one coalesced 128 B request), the pruning mask is
//   a per-query bitmap, and every surviving code costs one coalesced 128 B read of the S row
//   S[b, code, 0..31].  The per-passage score is summed sequentially in token order so that it is
//   bit-identical to filter_pids.cpp:59-63.
// select_top: one CTA per query; 64-bit keys (orderable(score) << 32 | pid) reproduce the
//   std::pair<float,int> ordering of filter_pids.cpp:24; an 8-pass MSB radix select finds the
//   keep-th key, survivors are bitonic-sorted in shared memory.
#include <type_traits>

#include "common.cuh"
#include <cuda_fp16.h>

namespace plaid {

// one element of the centroid-score table, fp32 or fp16 storage
__device__ __forceinline__ float load_s(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_s(const __half* p) { return __half2float(__ldg(p)); }

static constexpr int kApproxWarps = 8;        // warps per CTA (stage 2)
static constexpr int kStage1Warps = 16;       // stage-1 scan: 16 warps share one copy of the pruning bitmap (64 KB at C = 2^19, where
                                              // 8-warp CTAs left an SM with 16 warps and ~1 KB of code loads in flight each)
static constexpr int kStage1Dpw = 16;         // stage 1: 256 passages per CTA (amortises the bitmap load)
#ifndef PLAID_S2_ROWS
#define PLAID_S2_ROWS 16
#endif
static constexpr int kS2Rows = PLAID_S2_ROWS;   // stage 2: S rows a warp keeps in flight
static constexpr int kStage2Dpw = 1;          // stage 2: 8 passages per CTA (few queries resident -> S stays in L2)

// Gather the S rows of the codes selected by `mask` (bit l = lane l's `code`), G rows in flight per step.
template <int G, typename ST>
__device__ __forceinline__ float gather_rows(unsigned mask, int code, const ST* __restrict__ Sb, float m) {
    while (mask) {
        int src[G];
        float v[G];
#pragma unroll
        for (int t = 0; t < G; t++) {
            src[t] = mask ? (__ffs(mask) - 1) : -1;
            if (mask) mask &= mask - 1;
        }
#pragma unroll
        for (int t = 0; t < G; t++) {
            const int c = __shfl_sync(0xffffffffu, code, src[t] < 0 ? 0 : src[t]);
            v[t] = (src[t] >= 0) ? load_s(Sb + (size_t)(unsigned)c * PLAID_NQ_MAX) : -9999.0f;
        }
#pragma unroll
        for (int t = 0; t < G; t++) m = fmaxf(m, v[t]);
    }
    return m;
}

// A CTA scores 8*DPW consecutive candidate passages of one query; warp w walks passages DPW*w.. one
// after the other, lane = query token.  Codes are pulled 128 at a time (one 128-bit load per lane).
//   stage 1 (USE_IDX): the query's pruning bitmap (C bits + one zero word for the sentinel code C) sits in
//     shared memory; each lane probes its four codes (4 instructions each), ONE warp vote per 128 codes
//     tells whether anything survived, and only then the surviving codes' S rows are read;
//   stage 2: every code reads its S row (one coalesced 128 B request), eight rows in flight per warp.
// Codes must be < C (the reference asserts the same, filter_pids.cpp:47).
// The 32 per-token maxima of each passage are parked in shared memory; afterwards lane j of warp 0
// adds up passage j's row left to right, which is exactly the sequential fp32 sum of
// filter_pids.cpp:59-63 at 1/32 of the shuffle traffic of doing it inside every warp.
template <bool USE_IDX, int DPW, typename ST, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
approx_scores_kernel(const int32_t* __restrict__ pids, const int32_t* __restrict__ counts, int pid_stride,
                     const ST* __restrict__ S, const int32_t* __restrict__ qlens,
                     const uint32_t* __restrict__ idx_bits, int C, const int32_t* __restrict__ codes,
                     const int64_t* __restrict__ offsets, float* __restrict__ out,
                     const int32_t* __restrict__ only_flagged, int flag_stride) {
    constexpr int kDocs = WARPS * DPW;
    if (only_flagged && only_flagged[(size_t)blockIdx.y * flag_stride] == 0) return;   // hybrid stage 1: not this query
    extern __shared__ __align__(16) uint32_t s_dyn[];
    float (*s_max)[33] = reinterpret_cast<float (*)[33]>(s_dyn);      // [kDocs][33]
    uint32_t* s_bits = s_dyn + kDocs * 33;                            // [C/32] (stage 1 only)
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = warp_index();
    const int n = min(counts[b], pid_stride);
    if ((int)blockIdx.x * kDocs >= n) return;  // whole CTA past the end of this query's list
    const ST* Sb = S + (size_t)b * C * PLAID_NQ_MAX + lane;
    if (USE_IDX) {
        const uint4* src = reinterpret_cast<const uint4*>(idx_bits + (size_t)b * (C >> 5));
        for (int i = threadIdx.x; i < (C >> 7); i += blockDim.x) reinterpret_cast<uint4*>(s_bits)[i] = __ldg(src + i);
        if (threadIdx.x == 0) s_bits[C >> 5] = 0u;   // word of the sentinel code C: "not a survivor"
        __syncthreads();
    }
    // A CTA takes the passage groups blockIdx.x, blockIdx.x + gridDim.x, ... of its query (the regular launches have
    // one group per CTA; the flagged-only fallback launch of the hybrid stage 1 uses a thin grid, because 100 k CTAs
    // that all exit at once still cost 56 us of block scheduling).
  for (int i0 = blockIdx.x * kDocs; i0 < n; i0 += gridDim.x * kDocs) {
    // Passage descriptors of this warp's DPW passages, fetched in one parallel step (lane l <- passage l)
    // so that the scan below has no pid -> offsets -> codes pointer chase per passage.
    int64_t my_off = 0;
    int my_len = 0;
    if (lane < DPW) {
        const int i = i0 + warp * DPW + lane;
        if (i < n) {
            const int pid = pids[(size_t)b * pid_stride + i];
            my_off = offsets[pid];
            my_len = (int)(offsets[pid + 1] - my_off);
        }
    }
    const uint32_t sbits = smem_u32(s_bits);
    int pa[4], pb[4];         // stage 1: the first 256 codes of the next passage, requested one passage ahead
#pragma unroll 1
    for (int d = 0; d < DPW; d++) {
        const int64_t off = __shfl_sync(0xffffffffu, my_off, d);
        const int len = __shfl_sync(0xffffffffu, my_len, d);
        float m = -9999.0f;  // filter_pids.cpp:30-33
        if (USE_IDX) {
            // ---- stage 1: 128 codes per step (one 128-bit load per lane), bitmap probe, one vote ----
            // Elements outside the passage become the sentinel code C, whose bitmap word is zero.
            // The first 256 codes of passage d + 1 are requested before passage d is scanned: the passages of a
            // candidate list are scattered over the index, so every one of them is a DRAM round trip of its own and
            // the walk is bound by how many of them a warp keeps in flight.
            auto load4 = [&](const int32_t* cp, int head, int end, int e0, int (&c4)[4]) {
                const int e = e0 + lane * 4;                    // element index relative to the aligned start
                c4[0] = c4[1] = c4[2] = c4[3] = C;
                if (e + 3 < end) {                              // whole vector before the passage's end
                    const int4 x = ld_stream_v4(cp + e0);
                    c4[0] = e >= head ? x.x : C;                // elements before the passage belong to its neighbour
                    c4[1] = e + 1 >= head ? x.y : C;
                    c4[2] = e + 2 >= head ? x.z : C;
                    c4[3] = x.w;
                } else if (e < end) {                           // the passage's last, partial vector
#pragma unroll
                    for (int u = 0; u < 3; u++)
                        if (e + u >= head && e + u < end) c4[u] = ld_stream_s32(cp + e0 + u);
                }
            };
            auto scan4 = [&](const int (&c4)[4]) {
                unsigned w[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {                   // codes are < C by contract (sentinel = C)
                    asm("ld.shared.u32 %0, [%1];" : "=r"(w[u]) : "r"(sbits + (((unsigned)c4[u] >> 5) << 2)));
                    w[u] >>= ((unsigned)c4[u] & 31u);
                }
                if (__any_sync(0xffffffffu, (w[0] | w[1] | w[2] | w[3]) & 1u)) {
#pragma unroll
                    for (int u = 0; u < 4; u++) m = gather_rows<4>(__ballot_sync(0xffffffffu, w[u] & 1u), c4[u], Sb, m);
                }
            };
            // codes is 16-byte aligned: align the stream down
            const int head = (int)(off & 3);
            const int32_t* cp = codes + (off - head) + lane * 4;
            const int end = len > 0 ? head + len : 0;
            if (d == 0) {
                load4(cp, head, end, 0, pa);
                load4(cp, head, end, 128, pb);
            }
            int ca[4], cb[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { ca[u] = pa[u]; cb[u] = pb[u]; }
            if (d + 1 < DPW) {                                  // next passage of this warp: its loads go out now
                const int64_t noff = __shfl_sync(0xffffffffu, my_off, d + 1);
                const int nlen = __shfl_sync(0xffffffffu, my_len, d + 1);
                const int nhead = (int)(noff & 3);
                const int32_t* ncp = codes + (noff - nhead) + lane * 4;
                const int nend = nlen > 0 ? nhead + nlen : 0;
                load4(ncp, nhead, nend, 0, pa);
                load4(ncp, nhead, nend, 128, pb);
            }
            scan4(ca);
            if (128 < end) scan4(cb);
            for (int e0 = 256; e0 < end; e0 += 256) {           // passages longer than 256 tokens: two vectors per step
                load4(cp, head, end, e0, ca);
                load4(cp, head, end, e0 + 128, cb);
                scan4(ca);
                if (e0 + 128 < end) scan4(cb);
            }
        } else {
            // ---- stage 2: every code reads its S row ----
            const int32_t* cp = codes + off;
            if constexpr (std::is_same<ST, __half>::value) {
                // fp16 table: a row is 64 bytes, so one warp-wide 32-bit load fetches the rows of two tokens (lanes 0-15
                // the even one, 16-31 the odd one) and a lane keeps the running maximum of two query tokens as a half2,
                // exact in fp16.  16 loads = 32 rows in flight per warp; the walk is bound by gather latency.
                const __half2* Sb2 = reinterpret_cast<const __half2*>(S + (size_t)b * C * PLAID_NQ_MAX) + (lane & 15);
                const int hi = lane >> 4;
                __half2 m2 = __half2half2(__ushort_as_half((unsigned short)0xFC00u));   // -inf: "no token yet"
                // the codes of the next 32 tokens are requested before this chunk's rows: one round trip per chunk, not two
                int code = (lane < len) ? ld_stream_s32(cp + lane) : 0;
                for (int t0 = 0; t0 < len; t0 += 32) {
                    const int tn = t0 + 32 + lane;
                    const int ncode = (tn < len) ? ld_stream_s32(cp + tn) : 0;
                    const int cnt = min(32, len - t0);
                    if (cnt == 32) {
                        __half2 v[16];
#pragma unroll
                        for (int u = 0; u < 16; u++) {
                            const unsigned c = (unsigned)__shfl_sync(0xffffffffu, code, 2 * u + hi);
                            v[u] = __ldg(Sb2 + (size_t)c * (PLAID_NQ_MAX / 2));
                        }
#pragma unroll
                        for (int u = 0; u < 16; u++) m2 = __hmax2(m2, v[u]);
                    } else {
                        // the passage's last, partial chunk: sources past the end repeat its last token (max is idempotent)
                        for (int u0 = 0; 2 * u0 < cnt; u0 += 4) {
                            __half2 v[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const unsigned c = (unsigned)__shfl_sync(0xffffffffu, code, min(2 * (u0 + u) + hi, cnt - 1));
                                v[u] = __ldg(Sb2 + (size_t)c * (PLAID_NQ_MAX / 2));
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++) m2 = __hmax2(m2, v[u]);
                        }
                    }
                    code = ncode;
                }
                const unsigned other = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<unsigned*>(&m2), 16);
                m2 = __hmax2(m2, *reinterpret_cast<const __half2*>(&other));
                if (lane < 16) {   // an empty passage keeps the reference's -9999 per query token (filter_pids.cpp:30-33)
                    s_max[warp * DPW + d][2 * lane] = len > 0 ? __low2float(m2) : -9999.0f;
                    s_max[warp * DPW + d][2 * lane + 1] = len > 0 ? __high2float(m2) : -9999.0f;
                }
                continue;
            }
            for (int t0 = 0; t0 < len; t0 += 32) {
                const int t = t0 + lane;
                const int code = (t < len) ? ld_stream_s32(cp + t) : 0;
                if (t0 + 32 <= len) {                           // full chunk: straight-line gather, kS2Rows rows in flight
#pragma unroll
                    for (int g = 0; g < 32; g += kS2Rows) {
                        float v[kS2Rows];
#pragma unroll
                        for (int u = 0; u < kS2Rows; u++) {
                            const unsigned c = (unsigned)__shfl_sync(0xffffffffu, code, g + u);
                            v[u] = load_s(Sb + (size_t)c * PLAID_NQ_MAX);
                        }
#pragma unroll
                        for (int u = 0; u < kS2Rows; u++) m = fmaxf(m, v[u]);
                    }
                } else {
                    m = gather_rows<8>(__ballot_sync(0xffffffffu, t < len), code, Sb, m);
                }
            }
        }
        s_max[warp * DPW + d][lane] = m;
    }
    __syncthreads();
    const int nq = min(qlens[b], PLAID_NQ_MAX);
    for (int j = threadIdx.x; j < kDocs; j += blockDim.x) {
        const int i = i0 + j;
        if (i < n) {
            float s = 0.0f;  // sequential fp32 sum in token order (filter_pids.cpp:59-63)
            for (int k = 0; k < nq; k++) s += s_max[j][k];
            out[(size_t)b * pid_stride + i] = s;
        }
    }
    __syncthreads();     // s_max is reused by the next group
  }
}

// ------------------------------------------------------------------------------------------ select
#ifndef PLAID_SEL_THREADS
#define PLAID_SEL_THREADS 512
#endif
static constexpr int kSelThreads = PLAID_SEL_THREADS;

__device__ __forceinline__ uint64_t make_key(float score, int32_t pid) {
    return ((uint64_t)float_to_ordered(score) << 32) | (uint32_t)pid;
}

// Key sources of the select: a key array in global memory (merge_topk compacts G lists into one), or the (score, pid)
// arrays themselves -- a key is 8 bytes either way, so select_top builds keys on the fly instead of writing a key array
// and reading it back on every pass (its key-building pass was 16 % of select1 on cfg2).
struct KeyArray {
    const uint64_t* __restrict__ k;
    __device__ __forceinline__ uint64_t operator()(int i) const { return k[i]; }
    __device__ __forceinline__ uint32_t hi(int i) const { return reinterpret_cast<const uint32_t*>(k)[2 * i + 1]; }
};
struct ScorePidKeys {
    const float* __restrict__ s;
    const int32_t* __restrict__ p;
    __device__ __forceinline__ uint64_t operator()(int i) const { return make_key(s[i], p[i]); }
    __device__ __forceinline__ uint32_t hi(int i) const { return float_to_ordered(s[i]); }
};

// Block-strided walk over n keys.  (Measured and rejected on cfg2: 4 / 8 independent loads in flight per thread -- 0.23 /
// 0.30 ms vs 0.23; warp-aggregating the histogram atomics with __match_any_sync -- 0.42 ms: same-address shared atomics
// are not what bounds the passes.)
template <typename L, typename F>
__device__ __forceinline__ void for_each_key(const L& load, int n, F f) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) f(load(i));
}

static constexpr int kRankMax = 256;   // buckets up to this size are finished by rank counting instead of more radix passes

// Descending bitonic sort of s_sel[0..sel_cap) (sel_cap a power of two).  Compare-exchange steps with a partner inside the
// warp (stride < 32) run on registers through shuffles, so only strides >= 32 cost a shared-memory round and a barrier
// (sel_cap = 1024: 21 barriers instead of 55).
__device__ void bitonic_sort_desc(uint64_t* s_sel, int sel_cap) {
    const int tid = threadIdx.x;
    if (sel_cap < 32) {
        for (int size = 2; size <= sel_cap; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < (sel_cap >> 1); t += blockDim.x) {
                    const int lo = 2 * t - (t & (stride - 1));
                    const int hi = lo + stride;
                    const bool desc = ((lo & size) == 0);
                    const uint64_t a = s_sel[lo], c = s_sel[hi];
                    if ((a < c) == desc) { s_sel[lo] = c; s_sel[hi] = a; }
                }
                __syncthreads();
            }
        }
        return;
    }
    // whole warps are in or out of every loop below: sel_cap and blockDim.x are multiples of 32
    for (int idx = tid; idx < sel_cap; idx += blockDim.x) {          // sizes 2..32 entirely inside a warp
        uint64_t v = s_sel[idx];
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const uint64_t o = __shfl_xor_sync(0xffffffffu, v, stride);
                const bool want_max = ((idx & stride) == 0) == ((idx & size) == 0);
                v = want_max ? (v > o ? v : o) : (v < o ? v : o);
            }
        }
        s_sel[idx] = v;
    }
    __syncthreads();
    for (int size = 64; size <= sel_cap; size <<= 1) {
        for (int stride = size >> 1; stride >= 32; stride >>= 1) {
            for (int t = tid; t < (sel_cap >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const uint64_t a = s_sel[lo], c = s_sel[hi];
                if ((a < c) == desc) { s_sel[lo] = c; s_sel[hi] = a; }
            }
            __syncthreads();
        }
        for (int idx = tid; idx < sel_cap; idx += blockDim.x) {
            uint64_t v = s_sel[idx];
            const bool desc = (idx & size) == 0;
#pragma unroll
            for (int stride = 16; stride > 0; stride >>= 1) {
                const uint64_t o = __shfl_xor_sync(0xffffffffu, v, stride);
                const bool want_max = ((idx & stride) == 0) == desc;
                v = want_max ? (v > o ? v : o) : (v < o ? v : o);
            }
            s_sel[idx] = v;
        }
        __syncthreads();
    }
}

static constexpr int kBktCap = 2048;   // largest bucket (of the keep-th key) that is collected into shared memory

// Block-wide: among the n keys of `load` pick the `keep` largest into s_sel (sorted descending), return count.
// s_sel has room for sel_cap = next_pow2(keep) keys, s_bkt for kBktCap + kRankMax; s_hist is 256 ints; s_misc 4 ints.
//
// MSB-first radix select of the keep-th largest key.  As soon as the bucket that holds it has at most kBktCap members, ONE
// more walk over the keys finishes the global part: keys above the bucket are selected for certain and go straight to
// s_sel, the bucket's members go to s_bkt, where the remaining digits are resolved from shared memory (radix passes,
// then rank counting once <= kRankMax keys are left) and its top `remaining` keys are appended.  Stage-1 scores tie
// heavily in fp32, so the plain select ran all 8 passes plus the selection walk over the full list (select1 on cfg2:
// 0.30 ms per 1024 queries; 0.24 with the bucket in shared memory; this layout saves the separate selection walk).
template <typename L>
__device__ int select_sorted_desc(const L& load, int n, int keep, uint64_t* s_sel, int sel_cap, uint64_t* s_bkt,
                                  int* s_hist, int* s_misc) {
    const int tid = threadIdx.x;
    const int m = min(n, keep);
    uint64_t* s_rank = s_bkt + kBktCap;
    uint64_t thresh = 0;  // with n <= keep every key is selected
    int n_greater_needed = m;
    bool collected = false;   // s_bkt[0..list_n) = the bucket at collection time, s_sel[0..sure) = the keys above it
    int list_n = 0, sure = 0;
    if (n > keep) {
        uint64_t prefix = 0, prefix_mask = 0;
        int remaining = keep;  // rank (1-based, from the top) still to be located inside the prefix group
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += blockDim.x) s_hist[i] = 0;
            __syncthreads();
            auto count = [&](uint64_t k) {
                if ((k & prefix_mask) == prefix) atomicAdd(&s_hist[(int)((k >> shift) & 0xff)], 1);
            };
            if (collected) {
                for_each_key(KeyArray{s_bkt}, list_n, count);
            } else if (shift >= 32) {
                // digits of the score word: the walk over all keys reads only the scores and works on 32 bits
                const uint32_t p_hi = (uint32_t)(prefix >> 32), m_hi = (uint32_t)(prefix_mask >> 32);
                const int sh = shift - 32;
                for (int i = tid; i < n; i += blockDim.x) {
                    const uint32_t h = load.hi(i);
                    if ((h & m_hi) == p_hi) atomicAdd(&s_hist[(int)((h >> sh) & 0xff)], 1);
                }
            } else {
                for_each_key(load, n, count);
            }
            __syncthreads();
            if (tid < 32) {
                // bucket of the `remaining`-th largest key: warp-parallel scan of the 256 bins from the top
                // (lane l owns bins 8l..8l+7; a serial scan by one thread cost ~2500 cycles per pass)
                int mine = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) mine += s_hist[8 * tid + j];
                int incl = mine;                              // sum over lanes >= tid
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_down_sync(0xffffffffu, incl, o);
                    if (tid + o < 32) incl += v;
                }
                const unsigned reach = __ballot_sync(0xffffffffu, incl >= remaining);   // lanes 0..L
                const int L = reach ? 31 - __clz(reach) : 0;
                if (tid == L) {
                    int acc = incl - mine, d = 8 * L + 7;
                    for (; d > 0; d--) {
                        if (acc + s_hist[d] >= remaining) break;
                        acc += s_hist[d];
                    }
                    s_misc[0] = d;
                    s_misc[1] = remaining - acc;
                    s_misc[3] = s_hist[d];                    // size of that bucket
                }
            }
            __syncthreads();
            prefix |= (uint64_t)s_misc[0] << shift;
            prefix_mask |= (uint64_t)0xff << shift;
            remaining = s_misc[1];
            const int bucket = s_misc[3];
            __syncthreads();
            if (!collected && bucket <= kBktCap) {
                // the last walk over all keys: certain keys -> s_sel, bucket -> s_bkt
                if (tid == 0) { s_misc[2] = 0; s_misc[3] = 0; }
                __syncthreads();
                for_each_key(load, n, [&](uint64_t k) {
                    const uint64_t hi = k & prefix_mask;
                    if (hi > prefix) {
                        const int slot = atomicAdd(&s_misc[2], 1);
                        if (slot < sel_cap) s_sel[slot] = k;
                    } else if (hi == prefix) {
                        s_bkt[atomicAdd(&s_misc[3], 1)] = k;                   // exactly `bucket` keys match
                    }
                });
                __syncthreads();
                collected = true;
                list_n = bucket;
                sure = keep - remaining;      // == s_misc[2]: the keys above the bucket
            }
            if (bucket == remaining) {     // the whole bucket belongs to the selection: no need to split it further
                remaining = 0;
                break;
            }
            if (shift == 0) break;
            if (collected && bucket <= kRankMax) {
                // copy the (narrowed) bucket out of the list; the wanted key is the one with `remaining - 1` greater keys in it
                if (tid == 0) s_misc[2] = 0;
                __syncthreads();
                for_each_key(KeyArray{s_bkt}, list_n, [&](uint64_t k) {
                    if ((k & prefix_mask) == prefix) s_rank[atomicAdd(&s_misc[2], 1)] = k;
                });
                __syncthreads();
                if (tid < bucket) {
                    const uint64_t my = s_rank[tid];
                    int greater = 0, equal = 0;
                    for (int j = 0; j < bucket; j++) {
                        const uint64_t o = s_rank[j];
                        greater += o > my;
                        equal += o == my;
                    }
                    if (greater < remaining && greater + equal >= remaining) {   // duplicates all write the same values
                        s_misc[0] = (int)(uint32_t)(my >> 32);
                        s_misc[1] = (int)(uint32_t)my;
                        s_misc[3] = remaining - greater;
                    }
                }
                __syncthreads();
                prefix = ((uint64_t)(uint32_t)s_misc[0] << 32) | (uint32_t)s_misc[1];
                remaining = s_misc[3];
                __syncthreads();
                break;
            }
        }
        thresh = prefix;              // the keep-th largest key, or (early exit) the smallest key value of its bucket
        n_greater_needed = remaining; // how many copies of `thresh` itself belong to the selection (0 after an early exit)
        if (remaining == 0) {         // every key >= thresh is selected: exactly `keep` of them
            thresh -= 1;              // (prefix > 0 here: bucket 0 of the first pass taken whole would mean n == keep)
        }
    }
    for (int i = tid + sure; i < sel_cap; i += blockDim.x) s_sel[i] = 0;
    if (tid == 0) { s_misc[2] = sure; s_misc[3] = 0; }
    __syncthreads();
    auto select = [&](uint64_t k) {
        bool take = (n <= keep) || (k > thresh);
        if (!take && k == thresh) take = atomicAdd(&s_misc[3], 1) < n_greater_needed;
        if (take) {
            const int slot = atomicAdd(&s_misc[2], 1);
            if (slot < sel_cap) s_sel[slot] = k;
        }
    };
    if (collected) for_each_key(KeyArray{s_bkt}, list_n, select);
    else for_each_key(load, n, select);
    __syncthreads();
    // descending (zero padding sinks to the end; real keys are > 0 because the ordered-float transform never yields 0
    // in the high word for non-NaN scores)
    bitonic_sort_desc(s_sel, sel_cap);
    return m;
}

__device__ void write_selection(const uint64_t* s_sel, int m, int out_stride, int32_t* out_pids, float* out_scores,
                                int32_t* out_count) {
    for (int i = threadIdx.x; i < out_stride; i += blockDim.x) {
        if (i < m) {
            const uint64_t k = s_sel[i];
            out_pids[i] = (int32_t)(uint32_t)(k & 0xffffffffu);
            if (out_scores) out_scores[i] = ordered_to_float((uint32_t)(k >> 32));
        } else {
            out_pids[i] = PLAID_NO_PID;
            if (out_scores) out_scores[i] = -INFINITY;
        }
    }
    if (threadIdx.x == 0 && out_count) *out_count = m;
}

__global__ void __launch_bounds__(kSelThreads)
select_top_kernel(const int32_t* __restrict__ pids, const float* __restrict__ scores, const int32_t* __restrict__ counts,
                  int in_stride, int keep, int sel_cap, int32_t* __restrict__ out_pids, float* __restrict__ out_scores,
                  int32_t* __restrict__ out_counts, int out_stride, uint64_t* __restrict__ ws_keys) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t* s_sel = reinterpret_cast<uint64_t*>(s_raw);
    uint64_t* s_bkt = s_sel + sel_cap;
    int* s_hist = reinterpret_cast<int*>(s_bkt + kBktCap + kRankMax);
    int* s_misc = s_hist + 256;
    const int b = blockIdx.x;
    const int n = min(counts[b], in_stride);
    (void)ws_keys;    // keys are built on the fly from (score, pid); the workspace argument stays in the ABI
    const ScorePidKeys load{scores + (size_t)b * in_stride, pids + (size_t)b * in_stride};
    const int m = select_sorted_desc(load, n, keep, s_sel, sel_cap, s_bkt, s_hist, s_misc);
    write_selection(s_sel, m, out_stride, out_pids + (size_t)b * out_stride,
                    out_scores ? out_scores + (size_t)b * out_stride : nullptr, out_counts ? out_counts + b : nullptr);
}

// gathered lists of G ranks -> per query G*k keys.  List g of query b: scores/pids + g*sp_stride + b*k (k entries),
// count at counts[g*c_stride + b]; pid_bases (optional) turns shard-local pids into global ones.
__global__ void __launch_bounds__(kSelThreads)
merge_topk_kernel(const float* __restrict__ scores, const int32_t* __restrict__ pids, const int32_t* __restrict__ counts,
                  int64_t sp_stride, int64_t c_stride, const int32_t* __restrict__ pid_bases,
                  int G, int B, int k, int sel_cap, int32_t* __restrict__ out_pids, float* __restrict__ out_scores,
                  int32_t* __restrict__ out_counts, uint64_t* __restrict__ ws_keys) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t* s_sel = reinterpret_cast<uint64_t*>(s_raw);
    uint64_t* s_bkt = s_sel + sel_cap;
    int* s_hist = reinterpret_cast<int*>(s_bkt + kBktCap + kRankMax);
    int* s_misc = s_hist + 256;
    const int b = blockIdx.x;
    uint64_t* keys = ws_keys + (size_t)b * G * k;
    // compact the valid entries of the G lists (serial prefix over G <= 64 lists)
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int g = 0; g < G; g++) {
            s_hist[g] = acc;
            acc += min(max(counts[(size_t)g * c_stride + b], 0), k);
        }
        s_hist[G] = acc;
    }
    __syncthreads();
    const int n = s_hist[G];
    for (int t = threadIdx.x; t < G * k; t += blockDim.x) {
        const int g = t / k, i = t - g * k;
        const int cnt = (g + 1 < G ? s_hist[g + 1] : n) - s_hist[g];
        if (i < cnt) {
            const size_t src = (size_t)g * sp_stride + (size_t)b * k + i;
            keys[s_hist[g] + i] = make_key(scores[src], pids[src] + (pid_bases ? pid_bases[g] : 0));
        }
    }
    __syncthreads();
    const int m = select_sorted_desc(KeyArray{keys}, n, k, s_sel, sel_cap, s_bkt, s_hist, s_misc);
    write_selection(s_sel, m, k, out_pids + (size_t)b * k, out_scores + (size_t)b * k, out_counts ? out_counts + b : nullptr);
}

// Exact-global truncation across shards without a merge: every shard's stage list is sorted by (score, global pid)
// descending and the keys of different shards never tie, so the collection's `keep` best keys take a PREFIX of every
// shard's list.  A warp per query finds the length of this shard's prefix by bisection: element i of my list has global
// rank i + (number of keys of the other shards above it), the latter by one bisection per other list (lane g searches
// list g of the all-gathered blocks).  ~100 dependent loads per query instead of a select over G * k keys.
__global__ void __launch_bounds__(256)
prefix_share_kernel(const int32_t* __restrict__ gathered, int G, int B, int rows, int k, int keep,
                    const int32_t* __restrict__ pid_bases, int me, int32_t* __restrict__ out_counts) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= rows) return;
    const int64_t stride = 2 * (int64_t)B * k + B;
    auto key_of = [&](int g, int j) -> uint64_t {
        const int32_t* blk = gathered + (int64_t)g * stride;
        const uint32_t sbits = float_to_ordered(__int_as_float(blk[(int64_t)B * k + (int64_t)b * k + j]));
        return ((uint64_t)sbits << 32) | (uint32_t)(blk[(int64_t)b * k + j] + pid_bases[g]);
    };
    auto len_of = [&](int g) { return min(max(gathered[(int64_t)g * stride + 2 * (int64_t)B * k + b], 0), k); };
    const int n_me = len_of(me);
    int lo = 0, hi = n_me;                       // the answer L lies in [lo, hi]: rank(L - 1) < keep
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const uint64_t e = key_of(me, mid - 1);
        int above = 0;
        for (int g = lane; g < G; g += 32) {
            if (g == me) continue;
            int a = 0, z = len_of(g);            // first index of list g whose key is below e
            while (a < z) {
                const int m = (a + z) >> 1;
                if (key_of(g, m) > e) a = m + 1; else z = m;
            }
            above += a;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
        if (mid - 1 + above < keep) lo = mid; else hi = mid - 1;
    }
    if (lane == 0) out_counts[b] = lo;
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <typename ST>
static int launch_approx_t(const int32_t* pids, const int32_t* counts, int B, int pid_stride, const ST* S,
                           const int32_t* qlens, const uint32_t* idx_bits, int C, const int32_t* codes,
                           const int64_t* offsets, float* out, cudaStream_t st, const int32_t* only_flagged = nullptr,
                           int flag_stride = 0) {
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(codes) & 15) == 0, PLAID_ERR_ARG, "approx_scores: codes must be 16-byte aligned");
    if (idx_bits) {
        constexpr int kDocs = kStage1Warps * kStage1Dpw;
        const size_t smem = (size_t)kDocs * 33 * 4 + (size_t)(C >> 5) * 4 + 16;
        PLAID_CHECK_ARG(smem <= 200 * 1024, PLAID_ERR_UNSUPPORTED, "approx_scores: C=%d pruning bitmap exceeds shared memory", C);
        PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(idx_bits) & 15) == 0 && (C % 128) == 0, PLAID_ERR_ARG,
                        "approx_scores: idx_bits must be 16-byte aligned and C a multiple of 128");
        static int configured[kMaxDevices] = {0};
        if (int rc = ensure_dynamic_smem((const void*)approx_scores_kernel<true, kStage1Dpw, ST, kStage1Warps>, (int)smem, configured)) return rc;
        dim3 grid((pid_stride + kDocs - 1) / kDocs, B);
        // fallback launch (queries flagged for the scan only): few CTAs per query, each walks its groups -- about two
        // waves of resident CTAs in total, so that a batch whose queries ALL take the scan (long inverted lists: the
        // 10M-passage index on one GPU) still fills the machine while the usual nobody-flagged launch stays cheap
        if (only_flagged) {
            const unsigned cap = (unsigned)max(4, (2 * 4 * sm_count() + B - 1) / B);
            if (grid.x > cap) grid.x = cap;
        }
        approx_scores_kernel<true, kStage1Dpw, ST, kStage1Warps><<<grid, kStage1Warps * 32, smem, st>>>(
            pids, counts, pid_stride, S, qlens, idx_bits, C, codes, offsets, out, only_flagged, flag_stride);
    } else {
        constexpr int kDocs = kApproxWarps * kStage2Dpw;
        dim3 grid((pid_stride + kDocs - 1) / kDocs, B);
        approx_scores_kernel<false, kStage2Dpw, ST, kApproxWarps><<<grid, kApproxWarps * 32, kDocs * 33 * 4, st>>>(
            pids, counts, pid_stride, S, qlens, nullptr, C, codes, offsets, out, only_flagged, flag_stride);
    }
    PLAID_LAUNCH_OK("approx_scores_kernel");
    return PLAID_OK;
}

static int launch_approx(const int32_t* pids, const int32_t* counts, int B, int pid_stride, const void* S, int s_is_f16,
                         const int32_t* qlens, const uint32_t* idx_bits, int C, const int32_t* codes,
                         const int64_t* offsets, float* out, cudaStream_t st) {
    if (s_is_f16)
        return launch_approx_t<__half>(pids, counts, B, pid_stride, reinterpret_cast<const __half*>(S), qlens, idx_bits, C,
                                       codes, offsets, out, st);
    return launch_approx_t<float>(pids, counts, B, pid_stride, reinterpret_cast<const float*>(S), qlens, idx_bits, C, codes,
                                  offsets, out, st);
}

// ------------------------------------------------------------------------------------------ stage 1 via the IVF
// Stage 1 only needs, for every candidate passage, the set of SURVIVING centroids (pruning mask set) it
// contains.  Scanning every token of every candidate finds them the hard way; when the mask is sparse the
// inverted file gives them directly: for each survivor c, ivf[c] lists the passages that contain c, the
// candidate bitmap says which of them are candidates and the bitmap's word-prefix counts give their slot.
//   ivf_survivors : per query, compact the mask into a survivor list; estimate the work (sum of list lengths)
//                   and fall back to the token scan for this query when it is not clearly cheaper;
//   ivf_pairs     : warp per survivor, emits (slot, centroid) pairs;
//   ivf_scores    : per query, counting sort of the pairs by slot in shared memory, then per slot the
//                   per-token max over its survivors' S rows and the sequential fp32 sum (filter_pids.cpp:59-63).
// The per-token max is order-independent, so the result is bit-identical to the scan.
static constexpr int kIvfMeta = 4;   // per query: [0] survivors, [1] pairs, [2] use-the-scan flag, [3] visits
static int g_ivf_range_slots = 0;    // plaid_set_ivf_range_slots (test hook): 0 = what shared memory holds

__global__ void __launch_bounds__(256)
ivf_survivors_kernel(const uint32_t* __restrict__ idx_bits, int C, const int64_t* __restrict__ ivf_offsets,
                     const int32_t* __restrict__ counts, int pid_stride, int max_bins, int cap_s, int max_visits,
                     int32_t* __restrict__ surv, int32_t* __restrict__ meta) {
    __shared__ int s_count, s_visits;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) { s_count = 0; s_visits = 0; }
    __syncthreads();
    const uint32_t* bits = idx_bits + (size_t)b * (C >> 5);
    int visits = 0;
    for (int w = threadIdx.x; w < (C >> 5); w += blockDim.x) {
        uint32_t x = bits[w];
        while (x) {
            const int c = (w << 5) + __ffs(x) - 1;
            x &= x - 1;
            const int slot = atomicAdd(&s_count, 1);
            if (slot < cap_s) surv[(size_t)b * cap_s + slot] = c;
            visits += (int)min((int64_t)(1 << 24), ivf_offsets[c + 1] - ivf_offsets[c]);
        }
    }
    if (visits) atomicAdd(&s_visits, min(visits, 1 << 28));
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = min(counts[b], pid_stride);
        int32_t* m = meta + (size_t)b * kIvfMeta;
        m[0] = min(s_count, cap_s);
        m[1] = 0;
        m[2] = (s_count > cap_s || s_visits > max_visits || n > max_bins) ? 1 : 0;
        m[3] = s_visits;
    }
}

#ifndef PLAID_PAIRS_GRIDX
#define PLAID_PAIRS_GRIDX 32
#endif
constexpr int kPairsGridX = PLAID_PAIRS_GRIDX;

__global__ void __launch_bounds__(256)
ivf_pairs_kernel(const int32_t* __restrict__ surv, int cap_s, int32_t* __restrict__ meta,
                 const int32_t* __restrict__ ivf_pids, const int64_t* __restrict__ ivf_offsets,
                 const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ wprefix, int words, int N,
                 int32_t* __restrict__ pair_slot, int32_t* __restrict__ pair_c, int cap_p) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    int32_t* m = meta + (size_t)b * kIvfMeta;
    if (m[2]) return;
    const int wpb = blockDim.x >> 5, ns = min(m[0], cap_s);
    const uint32_t* bm = bitmap + (size_t)b * words;
    const int32_t* wp = wprefix + (size_t)b * words;
    // one warp per surviving centroid, grid-strided so the grid need not cover the cap_s worst case
    for (int si = blockIdx.x * wpb + warp_index(); si < ns; si += gridDim.x * wpb) {
        const int c = surv[(size_t)b * cap_s + si];
        const int64_t lo = ivf_offsets[c], hi = ivf_offsets[c + 1];
        for (int64_t i0 = lo; i0 < hi; i0 += 32) {       // warp-uniform trip count: one atomic per warp and pass
            const int64_t i = i0 + lane;
            const unsigned pid = i < hi ? (unsigned)ld_stream_s32(ivf_pids + i) : 0xFFFFFFFFu;
            uint32_t word = 0;
            bool hit = false;
            if (pid < (unsigned)N) {
                word = __ldg(bm + (pid >> 5));
                hit = (word >> (pid & 31)) & 1u;
            }
            const unsigned hits = __ballot_sync(0xFFFFFFFFu, hit);
            if (hits == 0) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(&m[1], __popc(hits));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (hit) {
                const int idx = base + __popc(hits & ((1u << lane) - 1u));
                if (idx < cap_p) {
                    const int slot = __ldg(wp + (pid >> 5)) + __popc(word & ((1u << (pid & 31)) - 1u));
                    pair_slot[(size_t)b * cap_p + idx] = slot;
                    pair_c[(size_t)b * cap_p + idx] = c;
                } else {
                    m[2] = 1;   // too many pairs for the workspace: this query goes through the scan
                }
            }
        }
    }
}

// Per-warp tile of running maxima in ivf_scores_kernel: [32 passages][32 query tokens] in the table's own precision (the
// maximum of stored values is exact in it).  The fp16 tile marks "no surviving centroid" with -inf and turns it into the
// reference's -9999 (filter_pids.cpp:30-33), which fp16 cannot hold, when the row is summed.
template <typename ST> struct IvfTile;
template <> struct IvfTile<float> {
    static constexpr int kThreads = 512, kBytes = 32 * 33 * 4;
    using Cell = float;                                     // [32][33]
    __device__ static void clear(Cell* t, int lane) {
#pragma unroll
        for (int j = 0; j < 32; j++) t[j * 33 + lane] = -9999.0f;
    }
    // up to 32 (slot, centroid) pairs held one per lane in `mine`; Srow = the query's table
    __device__ static void gather(Cell* t, int2 mine, int cnt, int base, const float* Srow, int lane) {
        const float* Sb = Srow + lane;
        for (int u0 = 0; u0 < cnt; u0 += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const unsigned c = (unsigned)__shfl_sync(0xffffffffu, mine.y, min(u0 + u, cnt - 1));
                v[u] = __ldg(Sb + (size_t)c * PLAID_NQ_MAX);
            }
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int sl = __shfl_sync(0xffffffffu, mine.x, min(u0 + u, cnt - 1)) - base;
                if (u0 + u < cnt) t[sl * 33 + lane] = fmaxf(t[sl * 33 + lane], v[u]);
            }
        }
    }
    __device__ static float sum(const Cell* t, int lane, int nq, float) {
        float acc = 0.0f;  // sequential fp32 sum in token order (filter_pids.cpp:59-63)
        for (int k = 0; k < nq; k++) acc += t[lane * 33 + k];
        return acc;
    }
};
template <> struct IvfTile<__half> {
    // Two query tokens per cell; a 64-byte table row is half a warp's load, so one load instruction fetches the rows of
    // two pairs (lanes 0-15 the even pair, 16-31 the odd one).  Pitch 17 words: the transposed read of sum() is
    // conflict-free.
    static constexpr int kThreads = 1024, kBytes = 32 * 17 * 4;
    using Cell = __half2;                                   // [32][17]
    __device__ static __half2 ninf2() { return __half2half2(__ushort_as_half((unsigned short)0xFC00u)); }
    __device__ static void clear(Cell* t, int lane) {
#pragma unroll
        for (int j = 0; j < 17; j++) t[j * 32 + lane] = ninf2();
    }
    __device__ static void gather(Cell* t, int2 mine, int cnt, int base, const __half* Srow, int lane) {
        const __half2* Sb2 = reinterpret_cast<const __half2*>(Srow) + (lane & 15);
        const int hi = lane >> 4;
        // Split the (slot-sorted) pairs between the half warps at a slot boundary at or past the middle: the halves then
        // never update the same row of the tile, so no exchange or ordering between them is needed.
        const int prev = __shfl_up_sync(0xffffffffu, mine.x, 1);
        const unsigned starts = __ballot_sync(0xffffffffu, lane > 0 && lane < cnt && mine.x != prev);
        const int half = (cnt + 1) >> 1;
        const unsigned above = (starts >> half) << half;
        const int m = above ? __ffs(above) - 1 : cnt;
        const int first = hi ? m : 0, last = hi ? cnt : m;           // this half warp's pairs
        const int steps = max(m, cnt - m);
        const bool active = first < last;                             // the upper half idles when one slot owns the tail
        for (int u0 = 0; u0 < steps; u0 += 16) {
            __half2 v[16];
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (u0 + u < steps) {  // past its last pair a half warp repeats it: max is idempotent
                    const unsigned c = (unsigned)__shfl_sync(0xffffffffu, mine.y, min(first + u0 + u, last - 1));
                    v[u] = __ldg(Sb2 + (size_t)c * (PLAID_NQ_MAX / 2));
                }
            }
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (u0 + u < steps) {
                    const int sl = __shfl_sync(0xffffffffu, mine.x, min(first + u0 + u, last - 1)) - base;
                    if (active) {
                        Cell* cell = t + sl * 17 + (lane & 15);
                        *cell = __hmax2(*cell, v[u]);
                    }
                }
            }
        }
    }
    // -inf marks "no surviving centroid": the reference's -9999 (filter_pids.cpp:30-33), which fp16 cannot hold.  A
    // passage either has no pair at all (every cell empty) or a full table row for every query token.
    __device__ static float sum(const Cell* t, int lane, int nq, float empty_sum) {
        const Cell* row = t + lane * 17;
        if (__half_as_ushort(__low2half(row[0])) == 0xFC00u) return empty_sum;
        float acc = 0.0f;  // sequential fp32 sum in token order (filter_pids.cpp:59-63)
        if (nq == PLAID_NQ_MAX) {
#pragma unroll
            for (int k2 = 0; k2 < PLAID_NQ_MAX / 2; k2++) {
                const float2 f = __half22float2(row[k2]);
                acc += f.x;
                acc += f.y;
            }
        } else {
            for (int k = 0; k < nq; k++) {
                const __half2 h = row[k >> 1];
                acc += __half2float((k & 1) ? __high2half(h) : __low2half(h));
            }
        }
        return acc;
    }
};

template <typename ST>
__global__ void __launch_bounds__(IvfTile<ST>::kThreads)
ivf_scores_kernel(const int32_t* __restrict__ counts, int pid_stride, const int32_t* __restrict__ meta,
                  const int32_t* __restrict__ pair_slot, const int32_t* __restrict__ pair_c, int32_t* __restrict__ sorted_c,
                  int cap_p, const ST* __restrict__ S, int C, const int32_t* __restrict__ qlens, float* __restrict__ out,
                  int range_slots) {
    using Tile = IvfTile<ST>;
    constexpr int kScan = 512;
    extern __shared__ __align__(16) int s_bins[];          // [n + 1] then one tile per warp
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = warp_index(), nw = blockDim.x >> 5;
    const int32_t* m = meta + (size_t)b * kIvfMeta;
    if (m[2]) return;
    // A CTA sorts and reduces the candidate slots [slot0, slot0 + n) of its query: the whole list when it fits the
    // shared-memory bins (gridDim.y = 1, every workload but very large shards), else one range of range_slots slots.
    const int n_all = min(counts[b], pid_stride);
    const int slot0 = blockIdx.y * range_slots;
    if (slot0 >= n_all) return;
    const int n = min(range_slots, n_all - slot0);
    const int np = min(m[1], cap_p);
    typename Tile::Cell* s_max = reinterpret_cast<typename Tile::Cell*>(
        reinterpret_cast<char*>(s_bins + ((n + 1 + 3) & ~3)) + (size_t)warp * Tile::kBytes);
    __shared__ int s_part[kScan];
    __shared__ int s_below;
    const int32_t* ps = pair_slot + (size_t)b * cap_p;
    const int32_t* pc = pair_c + (size_t)b * cap_p;
    for (int i = tid; i <= n; i += blockDim.x) s_bins[i] = 0;
    if (tid == 0) s_below = 0;
    __syncthreads();
    int below = 0;                                          // pairs of the earlier ranges: this range's offset in sorted_c
    for (int i = tid; i < np; i += blockDim.x) {
        const int sl = ps[i] - slot0;
        if ((unsigned)sl < (unsigned)n) atomicAdd(&s_bins[sl + 1], 1);   // count of slot s in bins[s + 1]
        else if (sl < 0) below++;
    }
    if (gridDim.y > 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if (lane == 0 && below) atomicAdd(&s_below, below);
    }
    __syncthreads();
    int2* sc2 = reinterpret_cast<int2*>(sorted_c) + (size_t)b * cap_p + s_below;   // (slot, centroid) of this range, sorted by slot
    // exclusive scan over bins[0..n] by the first 512 threads: thread t owns a contiguous span
    const int span = (n + 1 + kScan - 1) / kScan;
    const int lo = tid < kScan ? min(tid * span, n + 1) : n + 1, hi = min(lo + span, n + 1);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += s_bins[i];
    if (tid < kScan) s_part[tid] = sum;
    __syncthreads();
    if (warp == 0) {   // scan the 512 partial sums: 16 per lane
        int loc[16], tot = 0;
#pragma unroll
        for (int u = 0; u < 16; u++) { loc[u] = tot; tot += s_part[lane * 16 + u]; }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int base = incl - tot;
#pragma unroll
        for (int u = 0; u < 16; u++) s_part[lane * 16 + u] = base + loc[u];
    }
    __syncthreads();
    int run = tid < kScan ? s_part[tid] : 0;
    for (int i = lo; i < hi; i++) { const int v = s_bins[i]; s_bins[i] = run + v; run += v; }
    __syncthreads();
    // now bins[i] = number of pairs with slot < i  (inclusive scan of the shifted counts); start(s) = bins[s]
    // scatter: cursor = start(s); afterwards bins[s] = end(s) and start(s) = (s ? bins[s-1] : 0)
    for (int i = tid; i < np; i += blockDim.x) {
        const int sl = ps[i] - slot0;                       // slots are range-relative from here on
        if ((unsigned)sl < (unsigned)n) sc2[atomicAdd(&s_bins[sl], 1)] = make_int2(sl, pc[i]);
    }
    __syncthreads();
    const int nq = min(qlens[b], PLAID_NQ_MAX);
    const ST* Srow = S + (size_t)b * C * PLAID_NQ_MAX;
    float empty_sum = 0.0f;                 // score of a passage without any surviving centroid
    for (int k = 0; k < nq; k++) empty_sum += -9999.0f;
    // a warp takes 32 passages at a time; their pairs are contiguous in the sorted array.  A group holds ~half a pair per
    // passage on the headline workload, so the walk is a chain of dependent L2 round trips: the first 32 pairs of the
    // next group are fetched while this one is reduced, and 16 loads are in flight per pass.
    int base = warp * 32, beg = 0, end = 0;
    int2 mine = make_int2(base, 0);
    if (base < n) {
        beg = base ? s_bins[base - 1] : 0;
        end = s_bins[min(base + 31, n - 1)];
        if (beg + lane < end) mine = sc2[beg + lane];
    }
    while (base < n) {
        const int nbase = base + nw * 32;
        int nbeg = 0, nend = 0;
        int2 nmine = make_int2(nbase, 0);
        if (nbase < n) {
            nbeg = s_bins[nbase - 1];
            nend = s_bins[min(nbase + 31, n - 1)];
            if (nbeg + lane < nend) nmine = sc2[nbeg + lane];
        }
        Tile::clear(s_max, lane);
        __syncwarp();
        for (int p0 = beg; p0 < end; p0 += 32) {
            if (p0 != beg) mine = (p0 + lane < end) ? sc2[p0 + lane] : make_int2(base, 0);
            Tile::gather(s_max, mine, min(32, end - p0), base, Srow, lane);
            __syncwarp();
        }
        __syncwarp();
        const int sidx = base + lane;
        if (sidx < n) out[(size_t)b * pid_stride + slot0 + sidx] = Tile::sum(s_max, lane, nq, empty_sum);
        __syncwarp();
        base = nbase; beg = nbeg; end = nend; mine = nmine;
    }
}

static int launch_select(const int32_t* pids, const float* scores, const int32_t* counts, int B, int in_stride, int keep,
                         int32_t* out_pids, float* out_scores, int32_t* out_counts, int out_stride, uint64_t* ws_keys,
                         cudaStream_t st) {
    PLAID_CHECK_ARG(keep >= 1 && keep <= 16384, PLAID_ERR_UNSUPPORTED, "select_top: keep=%d outside [1, 16384]", keep);
    PLAID_CHECK_ARG(out_stride >= keep, PLAID_ERR_ARG, "select_top: out_stride=%d < keep=%d", out_stride, keep);
    const int sel_cap = next_pow2(keep < 2 ? 2 : keep);
    const size_t smem = (size_t)(sel_cap + kBktCap + kRankMax) * 8 + 256 * 4 + 16;
    static int configured[kMaxDevices] = {0};
    if (int rc = ensure_dynamic_smem((const void*)select_top_kernel, (int)smem, configured)) return rc;
    // one CTA per query: with more queries in a chunk than 512-thread CTAs fit on the GPU at once (592), half-size CTAs put
    // the whole chunk on the SMs in one wave (1024-query chunks on cfg2: select1 0.165 -> 0.149 ms, select2 / top-k -10 %;
    // 1024-thread CTAs: 0.179)
    const int threads = B > 592 ? kSelThreads / 2 : kSelThreads;
    select_top_kernel<<<B, threads, smem, st>>>(pids, scores, counts, in_stride, keep, sel_cap, out_pids, out_scores,
                                                out_counts, out_stride, ws_keys);
    PLAID_LAUNCH_OK("select_top_kernel");
    return PLAID_OK;
}

}  // namespace plaid

extern "C" int plaid_approx_scores(const int32_t* pids, const int32_t* counts, int B, int pid_stride, const void* S,
                                   int s_is_f16, const int32_t* qlens, const uint32_t* idx_bits, int C, const int32_t* codes,
                                   const int64_t* offsets, float* out_scores, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(pids && counts && S && qlens && codes && offsets && out_scores, PLAID_ERR_ARG,
                    "plaid_approx_scores: null pointer");
    PLAID_CHECK_ARG(B >= 0 && pid_stride >= 0 && C > 0 && (C % 128) == 0, PLAID_ERR_ARG,
                    "plaid_approx_scores: bad sizes (B=%d stride=%d C=%d; C must be a multiple of 128)", B, pid_stride, C);
    if (B == 0 || pid_stride == 0) return PLAID_OK;
    PLAID_CHECK_ARG(B <= 65535, PLAID_ERR_UNSUPPORTED, "plaid_approx_scores: B=%d > 65535 per call", B);
    return launch_approx(pids, counts, B, pid_stride, S, s_is_f16, qlens, idx_bits, C, codes, offsets, out_scores,
                         (cudaStream_t)stream);
}

extern "C" int plaid_filter_stage1_ivf(const int32_t* pids, const int32_t* counts, int B, int pid_stride, const void* S,
                                       int s_is_f16, const int32_t* qlens, const uint32_t* idx_bits, int C,
                                       const int32_t* codes, const int64_t* offsets, const int32_t* ivf_pids,
                                       const int64_t* ivf_offsets, const uint32_t* bitmap, const int32_t* wprefix, int N,
                                       int32_t* ws_surv, int cap_s, int32_t* ws_pair_slot, int32_t* ws_pair_c,
                                       int32_t* ws_sorted_c, int cap_p, int32_t* ws_meta, float* out_scores, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(pids && counts && S && qlens && idx_bits && codes && offsets && ivf_pids && ivf_offsets && bitmap &&
                        wprefix && ws_surv && ws_pair_slot && ws_pair_c && ws_sorted_c && ws_meta && out_scores,
                    PLAID_ERR_ARG, "plaid_filter_stage1_ivf: null pointer");
    PLAID_CHECK_ARG(B >= 0 && pid_stride >= 1 && C > 0 && (C % 128) == 0 && N > 0 && cap_s >= 8 && cap_p >= 32, PLAID_ERR_ARG,
                    "plaid_filter_stage1_ivf: bad sizes");
    if (B == 0) return PLAID_OK;
    PLAID_CHECK_ARG(B <= 65535, PLAID_ERR_UNSUPPORTED, "plaid_filter_stage1_ivf: B=%d > 65535 per call", B);
    cudaStream_t st = (cudaStream_t)stream;
    const int words = (N + 31) / 32;
    // shared memory of ivf_scores: bins [n+1] + one 32 x 32 tile of maxima per warp; passages beyond that use the scan
    const int smem_cap = 200 * 1024;
    const int smax_bytes = s_is_f16 ? (IvfTile<__half>::kThreads / 32) * IvfTile<__half>::kBytes
                                    : (IvfTile<float>::kThreads / 32) * IvfTile<float>::kBytes;
    int max_bins = (smem_cap - smax_bytes) / 4 - 8;
    if (g_ivf_range_slots > 0 && g_ivf_range_slots < max_bins) max_bins = g_ivf_range_slots;
    if (max_bins > pid_stride) max_bins = pid_stride;
    const int smem = ((max_bins + 1 + 3) & ~3) * 4 + smax_bytes;
    // candidate lists longer than the bins (shards of millions of passages) are reduced in ranges of max_bins slots
    const int nranges = (pid_stride + max_bins - 1) / max_bins;
    PLAID_CHECK_ARG(nranges <= 65535, PLAID_ERR_UNSUPPORTED, "plaid_filter_stage1_ivf: pid_stride=%d too large", pid_stride);
    // the scan costs ~4 bytes per candidate token; the IVF route pays ~12 bytes per list entry visited
    const int max_visits = cap_p * 4;
    ivf_survivors_kernel<<<B, 256, 0, st>>>(idx_bits, C, ivf_offsets, counts, pid_stride, pid_stride, cap_s, max_visits,
                                            ws_surv, ws_meta);
    PLAID_LAUNCH_OK("ivf_survivors_kernel");
    ivf_pairs_kernel<<<dim3(min((cap_s + 7) / 8, kPairsGridX), B), 256, 0, st>>>(ws_surv, cap_s, ws_meta, ivf_pids, ivf_offsets, bitmap,
                                                                wprefix, words, N, ws_pair_slot, ws_pair_c, cap_p);
    PLAID_LAUNCH_OK("ivf_pairs_kernel");
    static int configured_f32[kMaxDevices] = {0}, configured_f16[kMaxDevices] = {0};
    if (int rc = ensure_dynamic_smem((const void*)ivf_scores_kernel<float>, smem, configured_f32)) return rc;
    if (int rc = ensure_dynamic_smem((const void*)ivf_scores_kernel<__half>, smem, configured_f16)) return rc;
    if (s_is_f16)
        ivf_scores_kernel<__half><<<dim3(B, nranges), IvfTile<__half>::kThreads, smem, st>>>(counts, pid_stride, ws_meta, ws_pair_slot, ws_pair_c, ws_sorted_c,
                                                        cap_p, reinterpret_cast<const __half*>(S), C, qlens, out_scores, max_bins);
    else
        ivf_scores_kernel<float><<<dim3(B, nranges), IvfTile<float>::kThreads, smem, st>>>(counts, pid_stride, ws_meta, ws_pair_slot, ws_pair_c, ws_sorted_c,
                                                       cap_p, reinterpret_cast<const float*>(S), C, qlens, out_scores, max_bins);
    PLAID_LAUNCH_OK("ivf_scores_kernel");
    // queries flagged for the scan (dense masks, oversized lists): the token-scan kernel, restricted to them
    if (s_is_f16)
        return launch_approx_t<__half>(pids, counts, B, pid_stride, reinterpret_cast<const __half*>(S), qlens, idx_bits, C,
                                       codes, offsets, out_scores, st, ws_meta + 2, kIvfMeta);
    return launch_approx_t<float>(pids, counts, B, pid_stride, reinterpret_cast<const float*>(S), qlens, idx_bits, C, codes,
                                  offsets, out_scores, st, ws_meta + 2, kIvfMeta);
}

extern "C" int plaid_set_ivf_range_slots(int slots) {
    const int prev = plaid::g_ivf_range_slots;
    plaid::g_ivf_range_slots = slots > 0 ? slots : 0;
    return prev;
}

extern "C" int plaid_select_top(const int32_t* pids, const float* scores, const int32_t* counts, int B, int in_stride,
                                int keep, int32_t* out_pids, float* out_scores, int32_t* out_counts, int out_stride,
                                uint64_t* ws_keys, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(pids && scores && counts && out_pids && ws_keys, PLAID_ERR_ARG, "plaid_select_top: null pointer");
    PLAID_CHECK_ARG(B >= 0 && in_stride >= 0, PLAID_ERR_ARG, "plaid_select_top: bad sizes");
    if (B == 0) return PLAID_OK;
    return launch_select(pids, scores, counts, B, in_stride, keep, out_pids, out_scores, out_counts, out_stride, ws_keys,
                         (cudaStream_t)stream);
}

extern "C" int plaid_filter_pids(const int32_t* pids, const int32_t* counts, int B, int pid_stride, const void* S,
                                 int s_is_f16, const int32_t* qlens, const uint32_t* idx_bits, int C, const int32_t* codes,
                                 const int64_t* offsets, int ndocs, float* ws_scores, uint64_t* ws_keys,
                                 int32_t* stage1_pids, float* stage1_scores, int32_t* stage1_counts,
                                 int32_t* stage2_pids, float* stage2_scores, int32_t* stage2_counts, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(pids && counts && S && qlens && idx_bits && codes && offsets && ws_scores && ws_keys && stage1_pids &&
                        stage1_scores && stage1_counts && stage2_pids && stage2_scores && stage2_counts,
                    PLAID_ERR_ARG, "plaid_filter_pids: null pointer");
    PLAID_CHECK_ARG(ndocs >= 4 && B >= 0 && pid_stride >= 0 && C > 0 && (C % 128) == 0, PLAID_ERR_ARG,
                    "plaid_filter_pids: bad sizes (ndocs=%d B=%d stride=%d C=%d)", ndocs, B, pid_stride, C);
    if (B == 0) return PLAID_OK;
    PLAID_CHECK_ARG(B <= 65535, PLAID_ERR_UNSUPPORTED, "plaid_filter_pids: B=%d > 65535 per call", B);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    // stage 1: pruned centroids only (filter_pids.cpp:139-146), keep ndocs
    if (pid_stride > 0 &&
        (rc = launch_approx(pids, counts, B, pid_stride, S, s_is_f16, qlens, idx_bits, C, codes, offsets, ws_scores, st)) !=
            PLAID_OK)
        return rc;
    if ((rc = launch_select(pids, ws_scores, counts, B, pid_stride, ndocs, stage1_pids, stage1_scores, stage1_counts, ndocs,
                            ws_keys, st)) != PLAID_OK)
        return rc;
    // stage 2: all centroids (filter_pids.cpp:148-157), keep ndocs/4.  ws_scores is free again (its content
    // was folded into the keys) and is reused with row stride ndocs.
    if ((rc = launch_approx(stage1_pids, stage1_counts, B, ndocs, S, s_is_f16, qlens, nullptr, C, codes, offsets, ws_scores,
                            st)) !=
        PLAID_OK)
        return rc;
    return launch_select(stage1_pids, ws_scores, stage1_counts, B, ndocs, ndocs / 4, stage2_pids, stage2_scores,
                         stage2_counts, ndocs / 4, ws_keys, st);
}

static int launch_merge(const float* scores, const int32_t* pids, const int32_t* counts, int64_t sp_stride, int64_t c_stride,
                        const int32_t* pid_bases, int G, int B, int k, int32_t* out_pids, float* out_scores,
                        int32_t* out_counts, uint64_t* ws_keys, cudaStream_t st, int rows = -1) {
    using namespace plaid;
    PLAID_CHECK_ARG(G >= 1 && G <= 64 && B >= 0 && k >= 1 && k <= 16384, PLAID_ERR_UNSUPPORTED,
                    "plaid_merge_topk: G=%d (1..64), k=%d (1..16384)", G, k);
    if (rows < 0) rows = B;
    PLAID_CHECK_ARG(rows <= B, PLAID_ERR_ARG, "plaid_merge_topk: rows=%d > B=%d", rows, B);
    if (rows == 0) return PLAID_OK;
    const int sel_cap = next_pow2(k < 2 ? 2 : k);
    const size_t smem = (size_t)(sel_cap + kBktCap + kRankMax) * 8 + 256 * 4 + 16;
    static int configured[kMaxDevices] = {0};
    if (int rc = ensure_dynamic_smem((const void*)merge_topk_kernel, (int)smem, configured)) return rc;
    merge_topk_kernel<<<rows, kSelThreads, smem, st>>>(scores, pids, counts, sp_stride, c_stride, pid_bases, G, B, k, sel_cap,
                                                       out_pids, out_scores, out_counts, ws_keys);
    PLAID_LAUNCH_OK("merge_topk_kernel");
    return PLAID_OK;
}

extern "C" int plaid_merge_topk(const float* scores, const int32_t* pids, const int32_t* counts, int G, int B, int k,
                                int32_t* out_pids, float* out_scores, int32_t* out_counts, uint64_t* ws_keys,
                                void* stream) {
    PLAID_CHECK_ARG(scores && pids && counts && out_pids && out_scores && ws_keys, PLAID_ERR_ARG,
                    "plaid_merge_topk: null pointer");
    return launch_merge(scores, pids, counts, (int64_t)B * k, B, nullptr, G, B, k, out_pids, out_scores, out_counts, ws_keys,
                        (cudaStream_t)stream);
}

extern "C" int plaid_merge_topk_msg(const int32_t* gathered, int G, int B, int k, const int32_t* pid_bases,
                                    int32_t* out_pids, float* out_scores, int32_t* out_counts, uint64_t* ws_keys,
                                    void* stream) {
    PLAID_CHECK_ARG(gathered && out_pids && out_scores && ws_keys, PLAID_ERR_ARG, "plaid_merge_topk_msg: null pointer");
    PLAID_CHECK_ARG(B >= 0 && k >= 1, PLAID_ERR_ARG, "plaid_merge_topk_msg: bad sizes");
    const int64_t stride = 2 * (int64_t)B * k + B;     // one rank's message: pids | score bits | counts
    return launch_merge(reinterpret_cast<const float*>(gathered + (int64_t)B * k), gathered, gathered + 2 * (int64_t)B * k,
                        stride, stride, pid_bases, G, B, k, out_pids, out_scores, out_counts, ws_keys, (cudaStream_t)stream);
}

extern "C" int plaid_prefix_share(const int32_t* gathered, int G, int B, int rows, int k, int keep, const int32_t* pid_bases,
                                  int my_rank, int32_t* out_counts, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(gathered && pid_bases && out_counts, PLAID_ERR_ARG, "plaid_prefix_share: null pointer");
    PLAID_CHECK_ARG(G >= 1 && G <= 1024 && my_rank >= 0 && my_rank < G && B >= 0 && rows >= 0 && rows <= B && k >= 1 && keep >= 1,
                    PLAID_ERR_ARG, "plaid_prefix_share: bad sizes");
    if (rows == 0) return PLAID_OK;
    prefix_share_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(gathered, G, B, rows, k, keep, pid_bases, my_rank, out_counts);
    PLAID_LAUNCH_OK("prefix_share_kernel");
    return PLAID_OK;
}
