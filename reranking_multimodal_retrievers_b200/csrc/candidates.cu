// candidates.cu -- candidate passage generation.
//
// Replaces CandidateGeneration.generate_candidates (CB/search/candidate_generation.py:12-64): the
// `topk(ncells)` over centroid scores (finished here from the per-range partial lists the centroid
// kernel's epilogue produced), `ivf.lookup(cells)` (CB/search/strided_tensor.py:77-99 +
// segmented_lookup.cpp) and `sort` + `unique_consecutive`.  Instead of gather -> sort -> unique, the
// union of the IVF lists is accumulated in a per-query pid bitmap; scanning the bitmap emits the
// candidate pids already sorted and unique.
#include "common.cuh"

namespace plaid {

// One thread per (query, token): merge `nlists` partial top-ncells lists (each ordered by score
// descending, centroid id ascending) into cells[b, k, 0..ncells) under the same total order.
__global__ void merge_cells_kernel(const float* __restrict__ cell_val, const int32_t* __restrict__ cell_idx,
                                   const int32_t* __restrict__ qlens, int B, int ncells, int nlists,
                                   int32_t* __restrict__ cells) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * PLAID_NQ_MAX) return;
    const int b = t / PLAID_NQ_MAX, k = t % PLAID_NQ_MAX;
    int32_t* out = cells + (size_t)t * ncells;
    if (k >= min(qlens[b], PLAID_NQ_MAX)) {
        for (int j = 0; j < ncells; j++) out[j] = -1;
        return;
    }
    float bv[PLAID_NCELLS_MAX];
    int bi[PLAID_NCELLS_MAX];
    int cnt = 0;
    const size_t base = (size_t)t * nlists * ncells;
    for (int s = 0; s < nlists; s++)
        for (int j = 0; j < ncells; j++) {
            const float v = cell_val[base + (size_t)s * ncells + j];
            const int c = cell_idx[base + (size_t)s * ncells + j];
            if (c < 0) continue;
            int p = 0;
            while (p < cnt && !(v > bv[p] || (v == bv[p] && c < bi[p]))) p++;
            if (p >= ncells) continue;
            for (int q = min(cnt, ncells - 1); q > p; q--) { bv[q] = bv[q - 1]; bi[q] = bi[q - 1]; }
            bv[p] = v;
            bi[p] = c;
            if (cnt < ncells) cnt++;
        }
    for (int j = 0; j < ncells; j++) out[j] = (j < cnt) ? bi[j] : -1;
}

// grid (ceil(nq_max*ncells / 8), B), 8 warps per CTA: warp (e, b) expands cell entry e of query b unless an earlier
// entry of the same query names the same centroid (the reference de-duplicates cells with `unique`,
// candidate_generation.py:19).  Each pid of the cell's IVF list sets its bit in the query's bitmap.  A warp per entry
// (lists hold a few hundred pids) keeps 8x more of the cell -> offsets -> pids -> atomic latency chains in flight
// than a 256-thread CTA per entry (candidates 0.234 -> 0.190 ms per 1024 queries on cfg2).
static constexpr int kMarkWarps = 8;
__global__ void __launch_bounds__(kMarkWarps * 32)
mark_candidates_kernel(const int32_t* __restrict__ cells, int ncells, const int32_t* __restrict__ ivf_pids,
                       const int64_t* __restrict__ ivf_offsets, int C, int N, int words,
                       uint32_t* __restrict__ bitmap) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int e = blockIdx.x * kMarkWarps + warp_index();
    if (e >= PLAID_NQ_MAX * ncells) return;
    const int32_t* qc = cells + (size_t)b * PLAID_NQ_MAX * ncells;
    const int c = qc[e];
    if ((unsigned)c >= (unsigned)C) return;
    bool dup = false;
    for (int p = lane; p < e; p += 32) dup |= (qc[p] == c);
    if (__any_sync(0xffffffffu, dup)) return;
    const int64_t lo = ivf_offsets[c], hi = ivf_offsets[c + 1];
    uint32_t* bm = bitmap + (size_t)b * words;
    for (int64_t i0 = lo + lane; i0 < hi; i0 += 4 * 32) {
        int pid[4];
#pragma unroll
        for (int j = 0; j < 4; j++) pid[j] = (i0 + 32 * j < hi) ? ld_stream_s32(ivf_pids + i0 + 32 * j) : -1;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((unsigned)pid[j] < (unsigned)N) atomicOr(bm + (pid[j] >> 5), 1u << (pid[j] & 31));
    }
}

// One CTA per query: scan the bitmap, emit set bits as ascending pids.  A warp takes 128 consecutive words per pass
// (four coalesced loads, lane l the words l, 32 + l, 64 + l, 96 + l of the run) and every warp derives the warp-total
// prefix itself from a double-buffered shared array: one barrier per 4096 words (a 10M-passage shard has 312 k
// words per query; the first version -- one word per thread, four barriers per 1024 words -- spent 5 ms per 1024
// queries there, most of it in barriers).
__global__ void __launch_bounds__(1024)
compact_candidates_kernel(const uint32_t* __restrict__ bitmap, int words, int32_t* __restrict__ cand_pids,
                          int32_t* __restrict__ cand_counts, int cand_stride, int* __restrict__ overflow,
                          int32_t* __restrict__ wprefix) {
    __shared__ int s_warp[2][32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = warp_index(), nw = blockDim.x >> 5;
    const uint32_t* bm = bitmap + (size_t)b * words;
    int32_t* out = cand_pids + (size_t)b * cand_stride;
    int32_t* wp = wprefix ? wprefix + (size_t)b * words : nullptr;
    int base = 0, buf = 0;          // candidates before this pass: the same value in every thread
    for (int w0 = 0; w0 < words; w0 += 128 * nw, buf ^= 1) {
        const int wl = w0 + warp * 128 + lane;
        uint32_t bits[4];
        int before[4], run = 0;     // candidates of this warp's run before the lane's word j
#pragma unroll
        for (int j = 0; j < 4; j++) bits[j] = (wl + 32 * j < words) ? bm[wl + 32 * j] : 0u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int cnt = __popc(bits[j]);
            int incl = cnt;         // inclusive warp scan
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            before[j] = run + incl - cnt;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_warp[buf][warp] = run;
        __syncthreads();
        int tot = lane < nw ? s_warp[buf][lane] : 0, inc2 = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += u;
        }
        const int warp_pos = base + __shfl_sync(0xffffffffu, inc2 - tot, warp);
        base += __shfl_sync(0xffffffffu, inc2, 31);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int w = wl + 32 * j;
            int pos = warp_pos + before[j];
            if (wp && w < words) wp[w] = pos;   // candidates before word w = rank base of its pids
            uint32_t x = bits[j];
            while (x) {
                const int bit = __ffs(x) - 1;
                x &= x - 1;
                if (pos < cand_stride) out[pos] = (w << 5) + bit;
                pos++;
            }
        }
    }
    if (tid == 0) {
        cand_counts[b] = min(base, cand_stride);
        if (base > cand_stride && overflow) atomicExch(overflow, 1);
    }
}

}  // namespace plaid

extern "C" int plaid_candidates(const float* cell_val, const int32_t* cell_idx, const int32_t* qlens, int B, int ncells,
                                int nlists, const int32_t* ivf_pids, const int64_t* ivf_offsets, int C, int N,
                                int32_t* cells, uint32_t* bitmap_ws, int32_t* cand_pids, int32_t* cand_counts,
                                int cand_stride, int* overflow, int32_t* wprefix, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(cell_val && cell_idx && qlens && ivf_pids && ivf_offsets && cells && bitmap_ws && cand_pids &&
                        cand_counts,
                    PLAID_ERR_ARG, "plaid_candidates: null pointer");
    PLAID_CHECK_ARG(ncells >= 1 && ncells <= PLAID_NCELLS_MAX && nlists >= 1 && B >= 0 && C > 0 && N > 0 &&
                        cand_stride >= 1,
                    PLAID_ERR_ARG, "plaid_candidates: bad sizes (ncells=%d nlists=%d B=%d C=%d N=%d stride=%d)", ncells,
                    nlists, B, C, N, cand_stride);
    if (B == 0) return PLAID_OK;
    PLAID_CHECK_ARG(B <= 65535, PLAID_ERR_UNSUPPORTED, "plaid_candidates: B=%d > 65535 per call", B);
    cudaStream_t st = (cudaStream_t)stream;
    const int words = (N + 31) / 32;
    PLAID_CUDA_OK(cudaMemsetAsync(bitmap_ws, 0, (size_t)B * words * sizeof(uint32_t), st));
    const int nt = B * PLAID_NQ_MAX;
    merge_cells_kernel<<<(nt + 127) / 128, 128, 0, st>>>(cell_val, cell_idx, qlens, B, ncells, nlists, cells);
    PLAID_LAUNCH_OK("merge_cells_kernel");
    mark_candidates_kernel<<<dim3((PLAID_NQ_MAX * ncells + kMarkWarps - 1) / kMarkWarps, B), kMarkWarps * 32, 0, st>>>(
        cells, ncells, ivf_pids, ivf_offsets, C, N, words, bitmap_ws);
    PLAID_LAUNCH_OK("mark_candidates_kernel");
    compact_candidates_kernel<<<B, 1024, 0, st>>>(bitmap_ws, words, cand_pids, cand_counts, cand_stride, overflow, wprefix);
    PLAID_LAUNCH_OK("compact_candidates_kernel");
    return PLAID_OK;
}

extern "C" int plaid_merge_cells(const float* cell_val, const int32_t* cell_idx, const int32_t* qlens, int B, int ncells,
                                 int nlists, int32_t* cells, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(cell_val && cell_idx && qlens && cells, PLAID_ERR_ARG, "plaid_merge_cells: null pointer");
    PLAID_CHECK_ARG(B >= 0 && ncells >= 1 && ncells <= PLAID_NCELLS_MAX && nlists >= 1, PLAID_ERR_ARG,
                    "plaid_merge_cells: bad sizes (B=%d ncells=%d nlists=%d)", B, ncells, nlists);
    if (B == 0) return PLAID_OK;
    const int nt = B * PLAID_NQ_MAX;
    merge_cells_kernel<<<(nt + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cell_val, cell_idx, qlens, B, ncells, nlists, cells);
    PLAID_LAUNCH_OK("merge_cells_kernel");
    return PLAID_OK;
}
