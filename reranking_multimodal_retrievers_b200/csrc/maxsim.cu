// maxsim.cu -- late-interaction MaxSim on tcgen05 / TMEM with the per-passage max + sum fused into
// the epilogue, so the [tokens x Lq] similarity matrix never reaches HBM.
//
// Replaces, for a whole batch of queries:
//   packed form  : colbert_score_packed (CB/modeling/colbert.py:289-311) + segmented_maxsim.cpp:22-93
//                  -- zero-initialised running max per (passage, query token), summed over query tokens;
//   padded form  : colbert_score + colbert_score_reduce (CB/modeling/colbert.py:235-286,
//                  src/models/flmr/models/flmr/flmr_utils.py:22-48) -- masked positions count as -9999,
//                  max over passage tokens, sum over query tokens, optional masked matrix output.
//
// Orientation: accumulator lanes (M = 128 per m-tile, MT = ceil(Lq_pad/128) m-tiles) are QUERY tokens,
// accumulator columns (NT per tile) are PASSAGE tokens of consecutive passages.  One epilogue thread
// therefore owns one query token and walks passage tokens left to right: the running max is a
// register, a passage boundary is a flush (clamp, warp-shuffle sum over the 32 query tokens of the
// warp, add into a per-warp shared-memory slot), and a passage that straddles tiles simply keeps its
// register.  Work item = (query, group of 16 consecutive passages); CTAs are persistent over a
// contiguous range of items so the query operand (A, resident in smem) is reloaded only when the
// query changes.  B tiles (passage tokens, bf16 [NT x 128]) stream through an NS-stage TMA ring;
// accumulators are double-buffered in TMEM.
#include "common.cuh"
#include "decompress.cuh"

// ---- build options (A/B variants: python -m reranking_multimodal_retrievers_b200.build --variant NAME -DOPT=V, scripts/ab_variants.sh) ----
#ifndef MS_FUSED_ATMEM
#define MS_FUSED_ATMEM 1       // slim fused kernel: the query (A operand) lives in tensor memory, not in shared memory
#endif
#ifndef MS_FUSED_CB
#define MS_FUSED_CB 2          // fused kernel: steps (of 4 tokens) per centroid batch; two batches are in flight (1: 2.36 ms, 4: 2.17 vs 2.11)
#endif
#ifndef MS_UNIFORM_WARP
#define MS_UNIFORM_WARP 1      // warp index as a lane-0 broadcast (common.cuh: warp_index()); 0 = the pre-round-2b code generation
#endif
#ifndef MS_EPI_MODE
#define MS_EPI_MODE 1          // lean epilogue: 0 = chain of maxima, 1 = tree of 3-input maxima, 2 = tree + two 32-column loads per wait
#endif
#ifndef MS_SKIP_PAD_BATCHES
#define MS_SKIP_PAD_BATCHES 1  // a passage's last 32-row unit decodes only the pairs of centroid batches that hold real rows
#endif
#ifndef MS_SLIM_UT
#define MS_SLIM_UT 32          // slim layout: rows of a tile built by one decompressor warp (32: four groups of 4 warps; 16: two groups of 8)
#endif
#ifndef MS_FUSED_Q16
// Short queries (Lq_pad <= 64): four epilogue warps x 16 query rows, MMAs issued by the decompressor groups.  Correct
// (the whole GPU suite passes with it) and measured 5 % SLOWER than the two-warp epilogue + MMA warp (2.35 vs 2.24 ms on
// cfg2): timing experiments show the epilogue is not what bounds the kernel (skipping its loads and maxima changes
// nothing), and the groups' leaders pay for the issue.  Kept as a build option.
#define MS_FUSED_Q16 0
#endif
// timing experiments only (WRONG results; DESIGN.md section 4, third pass): what the kernel costs without one of its parts
#ifndef MS_DBG_SKIP_EPI
#define MS_DBG_SKIP_EPI 0      // the lean epilogues skip their loads and maxima
#endif
#ifndef MS_DBG_SKIP_STS
#define MS_DBG_SKIP_STS 0      // the decompressors keep their arithmetic but store no tile rows
#endif
#ifndef MS_DBG_SKIP_DEC
#define MS_DBG_SKIP_DEC 0      // the decompressors skip the decoding of their units
#endif
// (-DMS_DBG_STAGE_STRIDE=bytes: overlapping ring stages, i.e. more of them; -DMS_DBG_TIMING: per-warp stage-wait counters)

#ifdef MS_DBG_TIMING
// development aid: per (CTA, decompressor warp) cycles spent waiting for a free stage / in total (scripts/fused_wait_probe.py)
__device__ unsigned long long g_ms_dbg[160 * 16 * 8];
extern "C" int plaid_debug_read_fused_waits(unsigned long long* dst) {
    return (int)cudaMemcpyFromSymbol(dst, g_ms_dbg, sizeof(g_ms_dbg));
}
#endif

namespace plaid {

static constexpr int kMsThreads = 256;   // warps 0-3 epilogue, 4-5 idle, 6 TMA producer, 7 TMEM alloc + MMA issue
#ifndef MS_GD
#define MS_GD 32
#endif
static constexpr int kMsGD = MS_GD;      // passages per work item (<= 32: one lane per passage)
static constexpr int kMsMaxStages = 8;
static constexpr int kMsMaxMT = 4;       // Lq_pad <= 512

struct MsParams {
    const int32_t* qlens;        // [nQ] valid rows per query
    int Lq_pad, MT, NT, NS;
    int na_shift;                // log2 of the accumulator buffers in TMEM (4 when 4 x MT x NT <= 512 columns, else 2)
    int tmem_cols;               // TMEM columns this CTA allocates (256 when two CTAs share an SM, else 512)
    int a_tmem_col;              // > 0: the A operand sits in TMEM from this column on (64 columns); rows written by the epilogue warps
    const void* q_rows;          // fp16 query rows [B_pad * Lq_pad, 128] (A-in-TMEM only)
    int ab_f16;                  // operands (Q and D) are fp16 instead of bf16
    int padded;                  // 0 = packed search form, 1 = padded colbert_score form
    int aligned;                 // packed only: passages start on 32-token boundaries of D, pad rows are zero
    // packed
    const int32_t* tok_offsets;  // [B, pid_stride+1] per-query exclusive prefix of passage lengths
    const int32_t* counts;       // [B]
    int pid_stride, tok_stride;
    // padded
    const uint8_t* mask;         // [n, Ld]
    long long n_docs;
    int Ld, docs_per_query;
    float* scores_raw;           // optional [n, Ld, Lq_out]
    int Lq_out;
    // fused decompression (packed + aligned only): the B tiles are produced from the compressed index
    const int32_t* pids;         // [B, pid_stride] passages to score (local pids)
    const int64_t* doc_offsets;  // [N+1] token offsets of the index
    const int32_t* codes;        // [NE]
    const uint8_t* residuals;    // [NE, 16*nbits]
    const __half* centroids;     // [C, 128] fp16, exactly centroids.pt
    const float* wtable;         // [256, 8/nbits]
    const __half* inv_norms;     // [NE] precomputed per-token scale factors (plaid_token_inv_norms), or NULL
    int C;
    // common
    float* scores;
    int groups_per_query, num_items, items_per_cta;
    int clamp_zero;
    int* watchdog;
};

struct MsItem {
    int q;            // query index (row block of the A operand)
    int nd;           // passages in this item
    long long row0;   // first D row of the item
    int ntok;         // D rows (tokens) in the item
    long long out0;   // index of the first passage's score
    const int32_t* ends;  // packed: &tok_offsets[b][d0] (ends[i+1]-ends[0] = end of passage i); padded: nullptr
};

__device__ __forceinline__ MsItem ms_item(const MsParams& p, int w) {
    MsItem it;
    const int q = w / p.groups_per_query, g = w - q * p.groups_per_query;
    it.q = q;
    if (!p.padded) {
        const int cnt = min(p.counts[q], p.pid_stride);
        const int d0 = g * kMsGD;
        it.nd = max(0, min(kMsGD, cnt - d0));
        const int32_t* to = p.tok_offsets + (size_t)q * (p.pid_stride + 1);
        it.ends = to + d0;
        const int t0 = it.nd > 0 ? to[d0] : 0;
        it.ntok = it.nd > 0 ? to[d0 + it.nd] - t0 : 0;
        it.row0 = (long long)q * p.tok_stride + t0;
        it.out0 = (long long)q * p.pid_stride + d0;
    } else {
        const long long dq0 = (long long)q * p.docs_per_query;
        const long long dq1 = min(dq0 + p.docs_per_query, p.n_docs);
        const long long d0 = dq0 + (long long)g * kMsGD;
        it.nd = (int)max(0ll, min((long long)kMsGD, dq1 - d0));
        it.ends = nullptr;
        it.ntok = it.nd * p.Ld;
        it.row0 = d0 * p.Ld;
        it.out0 = d0;
    }
    return it;
}

struct MsShared {
    uint64_t a_full, a_empty;
    uint64_t full[kMsMaxStages];
    uint64_t empty[kMsMaxStages];
    uint64_t tmem_full[4];
    uint64_t tmem_empty[4];
    uint32_t tmem_base;
    int issue_seq;               // Q16: sequence number of the next tile whose MMAs may be issued (in-order issue token)
    float part[2][4][kMsGD];
};

__device__ __forceinline__ int lds_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_release_s32(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory");
}

// Barrier over the 128 epilogue threads (warps 0-3) that also ORs a predicate, so that a watchdog
// abort seen by one warp stops all four at the same work item (a lone early exit would strand the
// others at the barrier).
__device__ __forceinline__ bool epi_bar_or(bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "bar.red.or.pred q, 1, 128, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r) : "r"((uint32_t)pred) : "memory");
    return r != 0;
}

__device__ __forceinline__ float max32(const uint32_t (&r)[32], float seed) {
    float a = seed;
#pragma unroll
    for (int j = 0; j < 32; j += 2) a = fmaxf(a, fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
    return a;
}

// ===================== MMA issuer (one warp, one issuing lane) =====================
// Walks the CTA's work items in order; for every B tile: wait for the stage to be full and for an
// accumulator buffer to be drained, issue MT x 8 tcgen05.mma (K = 128), commit the stage back to its
// producer and the accumulator to the epilogue.  The A operand (query) is swapped when the query changes.
// The whole warp runs the loop so that its state lives in uniform registers (a loop under `lane == 0` makes
// the compiler re-broadcast every tcgen05 operand, ~300 instructions per tile on the kernel's critical path);
// only lane 0 issues.  Stage / phase counters advance incrementally (no division), and the shared-memory
// descriptors are the base descriptor plus a 16-byte-unit offset.
// With map_q != nullptr the warp also loads the query tile itself (no separate TMA producer warp): it lets the
// MMAs that read the old query retire, issues the TMA and waits for it -- a drain of a few tiles once per query.
// ATMEM (compile time, slim fused kernel only): the A operand is read from tensor memory (p.a_tmem_col).
template <bool ATMEM>
__device__ __forceinline__ void ms_mma_issue(const MsParams& p, MsShared* sh, uint8_t* sA, uint8_t* sB, int b_bytes,
                                             uint32_t tmem_base_in, int item_begin, int item_end, int lane,
                                             const CUtensorMap* map_q = nullptr) {
    // everything the tcgen05 operands derive from is made provably warp-uniform (lane-0 broadcasts), otherwise the
    // compiler wraps every instruction in an elect/broadcast loop
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);
    const int acc_cols = p.MT * p.NT;
    const uint32_t idesc = p.ab_f16 ? umma_idesc_f16(128, p.NT) : umma_idesc_bf16(128, p.NT);
    const uint64_t da_base = umma_smem_desc_sw128(smem_u32(sA)), db_base = umma_smem_desc_sw128(smem_u32(sB));
    const uint32_t b_khalf = (uint32_t)(p.NT * 128) >> 4, b_stage = (uint32_t)b_bytes >> 4;
    constexpr uint32_t a_khalf = (128 * 128) >> 4, a_mtile = (2 * 128 * 128) >> 4;
    int cur_q = -1, a_loads = 0, it_tile = 0;
    int s = 0;                  // stage of the next tile, and the parity its `full` barrier completes with
    uint32_t full_par = 0;
    bool ok = true;
    for (int w = item_begin; ok && w < item_end; w++) {
        const MsItem it = ms_item(p, w);
        const int nd = __shfl_sync(0xffffffffu, it.nd, 0), q = __shfl_sync(0xffffffffu, it.q, 0);
        const int ntok = __shfl_sync(0xffffffffu, it.ntok, 0);
        if (nd == 0) continue;
        if (q != cur_q) {
            if (a_loads > 0 && elect_one()) umma_commit(&sh->a_empty);  // every MMA that read the old A has retired
            if (map_q != nullptr && !ATMEM) {
                if (a_loads > 0 && !__all_sync(0xffffffffu, mbar_wait(&sh->a_empty, (a_loads - 1) & 1, p.watchdog))) break;
                if (elect_one()) {
                    mbar_expect_tx(&sh->a_full, p.MT * 128 * kDim * 2);
                    for (int m = 0; m < p.MT; m++)
                        for (int h = 0; h < 2; h++)
                            tma_load_2d(sA + (m * 2 + h) * (128 * 128), map_q, &sh->a_full, h * 64, q * p.Lq_pad + m * 128);
                }
                __syncwarp();
            }
            if (!__all_sync(0xffffffffu, mbar_wait(&sh->a_full, a_loads & 1, p.watchdog))) break;
            cur_q = q;
            a_loads++;
        }
        const int ntiles = (ntok + p.NT - 1) / p.NT;
        for (int t = 0; t < ntiles; t++, it_tile++) {
            const int acc = it_tile & ((1 << p.na_shift) - 1);
            bool got = mbar_wait(&sh->tmem_empty[acc], ((it_tile >> p.na_shift) & 1) ^ 1, p.watchdog);
            got = got && mbar_wait(&sh->full[s], full_par, p.watchdog);
            if (!__all_sync(0xffffffffu, got)) { ok = false; break; }
            tc_fence_after();
            if (elect_one()) {
                const uint64_t db0 = db_base + (uint64_t)(s * b_stage);
#pragma unroll 1
                for (int m = 0; m < p.MT; m++) {
                    const uint64_t da0 = da_base + (uint64_t)(m * a_mtile);
                    const uint32_t d_tmem = tmem_base + acc * acc_cols + m * p.NT;
                    if constexpr (ATMEM) {
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            umma_f16_ts(d_tmem, tmem_base + p.a_tmem_col + k * 8,
                                        db0 + (uint64_t)((k >> 2) * b_khalf + (k & 3) * 2), idesc, k > 0);
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            umma_bf16(d_tmem, da0 + (uint64_t)((k >> 2) * a_khalf + (k & 3) * 2),
                                      db0 + (uint64_t)((k >> 2) * b_khalf + (k & 3) * 2), idesc, k > 0);
                    }
                }
                umma_commit(&sh->empty[s]);
                umma_commit(&sh->tmem_full[acc]);
            }
            __syncwarp();
            if (++s == p.NS) { s = 0; full_par ^= 1; }
        }
    }
}

// ===================== epilogue: 4 warps, warp = TMEM lane quadrant =====================
// MODE 0: packed + aligned (the search pipeline: every passage starts on a 32-token boundary of D and its
//         pad rows are zero, so a 32-column chunk never straddles two passages and -- with the clamp at
//         0 -- pad columns cannot change a maximum);
// MODE 1: packed, arbitrary passage boundaries (colbert_score_packed operator);
// MODE 2: padded with mask (colbert_score operator), optional masked-matrix output.
template <int MODE>
__device__ __forceinline__ void ms_epilogue(const MsParams& p, MsShared* sh, uint32_t tmem_base, int item_begin,
                                            int item_end, int warp, int lane) {
    const int acc_cols = p.MT * p.NT;
    const int quad = warp;
    const float init = p.clamp_zero ? 0.0f : -INFINITY;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    int it_tile = 0, parity = 0;
    bool ok = true;
    for (int w = item_begin; ok && w < item_end; w++) {
        const MsItem it = ms_item(p, w);
        if (it.nd == 0) continue;
        const int lq = p.qlens[it.q];
        float* part = sh->part[parity][quad];
        float runmax[kMsMaxMT];
        bool rowok[kMsMaxMT], live[kMsMaxMT];
#pragma unroll
        for (int m = 0; m < kMsMaxMT; m++) {
            runmax[m] = init;
            rowok[m] = (m * 128 + quad * 32 + lane) < lq;
            live[m] = m < p.MT && (m * 128 + quad * 32) < lq;   // warp-uniform: some lane has a real query token
        }
        // passage ends of the item (relative to its first token) held one per lane: lane d+1 = end of passage d
        int ends_reg = 0x7fffffff;
        if (lane <= it.nd) ends_reg = it.ends ? (it.ends[lane] - it.ends[0]) : lane * p.Ld;
        // (with 32 passages there is no lane 32 for the item's end: it is it.ntok)
        auto doc_end = [&](int d) -> int {
            const int v = __shfl_sync(0xffffffffu, ends_reg, min(d + 1, 31));
            return d + 1 < 32 ? v : it.ntok;
        };
        int doc = 0;
        int next_end = doc_end(0);

        auto flush_all = [&](int d) {   // one write per passage, all m-tiles folded in
            float v = 0.0f;
#pragma unroll
            for (int m = 0; m < kMsMaxMT; m++) {
                if (m < p.MT && rowok[m]) v += p.clamp_zero ? fmaxf(runmax[m], 0.0f) : runmax[m];
                runmax[m] = init;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) part[d] = v;
        };

        const int ntiles = (it.ntok + p.NT - 1) / p.NT;
        const int nchunks = p.NT >> 5;
        int mnext[4] = {0, 0, 0, 0};
        for (int t = 0; t < ntiles; t++, it_tile++) {
            const int acc = it_tile & ((1 << p.na_shift) - 1);
            // padded form: mask bytes of the tile (lane <- token 32*ch + lane of each chunk).  The epilogue is the
            // slowest role, so the accumulator is usually ready already: the bytes of tile t+1 are requested
            // now and consumed one tile later, keeping their latency off the critical path.
            int mbyte[4] = {mnext[0], mnext[1], mnext[2], mnext[3]};
            if (MODE == 2) {
                if (t == 0) {
#pragma unroll
                    for (int ch = 0; ch < 4; ch++) {
                        const int tk = ch * 32 + lane;
                        mbyte[ch] = (ch * 32 < p.NT && tk < it.ntok) ? p.mask[it.row0 + tk] : 0;
                    }
                }
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const int tk = (t + 1) * p.NT + ch * 32 + lane;
                    mnext[ch] = (ch * 32 < p.NT && tk < it.ntok) ? p.mask[it.row0 + tk] : 0;
                }
            }
            if (!mbar_wait(&sh->tmem_full[acc], (it_tile >> p.na_shift) & 1, p.watchdog)) { ok = false; break; }
            tc_fence_after();
            const uint32_t tmem_acc = tmem_lane + acc * acc_cols;
            for (int ch = 0; ch < nchunks; ch++) {
                const int tk0 = t * p.NT + ch * 32;       // item-relative token of column 0
                if (tk0 >= it.ntok) break;
                if (MODE == 0) {
                    while (tk0 == next_end && doc + 1 < it.nd) {   // this chunk opens the next passage
                        flush_all(doc);
                        doc++;
                        next_end = doc_end(doc);
                    }
#pragma unroll
                    for (int m = 0; m < kMsMaxMT; m++) {
                        if (!live[m]) continue;
                        uint32_t r[32];
                        tmem_ld_32x32(tmem_acc + m * p.NT + ch * 32, r);
                        tc_wait_ld();
                        runmax[m] = max32(r, runmax[m]);
                    }
                } else {
                    const int nv = min(32, it.ntok - tk0);    // valid columns in this chunk
                    while (tk0 == next_end && doc + 1 < it.nd) {     // the previous passage ended exactly at this chunk
                        flush_all(doc);
                        doc++;
                        next_end = doc_end(doc);
                    }
                    uint32_t mword = 0xffffffffu;
                    if (MODE == 2) {
                        const int mb = ch == 0 ? mbyte[0] : ch == 1 ? mbyte[1] : ch == 2 ? mbyte[2] : mbyte[3];
                        mword = __ballot_sync(0xffffffffu, mb != 0);
                    }
                    // whole chunk inside the current passage and no per-element side effects wanted
                    const bool plain = (nv == 32) && (next_end >= tk0 + 32) && (MODE != 2 || p.scores_raw == nullptr);
                    const bool fast = plain && (mword == 0xffffffffu);
                    if (plain && mword == 0u) {           // a fully padded chunk: every column counts as -9999
#pragma unroll
                        for (int m = 0; m < kMsMaxMT; m++)
                            if (live[m]) runmax[m] = fmaxf(runmax[m], -9999.0f);
                        continue;
                    }
                    if (plain) {
#pragma unroll
                    for (int m = 0; m < kMsMaxMT; m++) {
                        if (!live[m]) continue;
                        uint32_t r[32];
                        tmem_ld_32x32(tmem_acc + m * p.NT + ch * 32, r);
                        tc_wait_ld();
                        if (fast) {
                            runmax[m] = max32(r, runmax[m]);
                        } else {                          // partially padded chunk: select, then max
                            float a = runmax[m];
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                a = fmaxf(a, ((mword >> j) & 1u) ? __uint_as_float(r[j]) : -9999.0f);
                            runmax[m] = a;
                        }
                    }
                } else if (nv == 32 && (MODE != 2 || p.scores_raw == nullptr) && doc + 1 < it.nd &&
                           doc_end(doc + 1) >= tk0 + 32) {
                    // exactly one passage boundary inside the chunk, at column s: columns [0, s) close passage `doc`,
                    // columns [s, 32) open passage doc+1.  Two predicated maxima per m-tile, no loop.
                    const int s = next_end - tk0;
                    float open_max[kMsMaxMT];
#pragma unroll
                    for (int m = 0; m < kMsMaxMT; m++) {
                        open_max[m] = init;
                        if (!live[m]) continue;
                        uint32_t r[32];
                        tmem_ld_32x32(tmem_acc + m * p.NT + ch * 32, r);
                        tc_wait_ld();
                        float a = runmax[m], b = init;
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const float v = ((mword >> j) & 1u) ? __uint_as_float(r[j]) : -9999.0f;
                            if (j < s) a = fmaxf(a, v); else b = fmaxf(b, v);
                        }
                        runmax[m] = a;
                        open_max[m] = b;
                    }
                    flush_all(doc);
                    doc++;
                    next_end = doc_end(doc);
#pragma unroll
                    for (int m = 0; m < kMsMaxMT; m++) runmax[m] = open_max[m];
                } else {
                    // generic chunk (several passages end inside it, the item ends inside it, or the masked matrix is
                    // wanted): one column at a time straight from TMEM -- a compact runtime loop, so the rare
                    // path does not bloat the instruction stream of the common one
#pragma unroll 1
                    for (int j = 0; j < nv; j++) {
                        while (tk0 + j == next_end && doc + 1 < it.nd) {     // passage `doc` ends before this token
                            flush_all(doc);
                            doc++;
                            next_end = doc_end(doc);
                        }
                        const bool on = (mword >> j) & 1u;
#pragma unroll
                        for (int m = 0; m < kMsMaxMT; m++) {
                            if (m >= p.MT) break;
                            float v = __uint_as_float(tmem_ld_32x1(tmem_acc + m * p.NT + ch * 32 + j));
                            if (MODE == 2 && !on) v = -9999.0f;              // colbert.py:240-241
                            const int krow = m * 128 + quad * 32 + lane;
                            if (MODE == 2 && p.scores_raw != nullptr && krow < p.Lq_out)
                                p.scores_raw[(size_t)(it.row0 + tk0 + j) * p.Lq_out + krow] = v;
                            runmax[m] = fmaxf(runmax[m], v);
                        }
                    }
                }
            }
        }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->tmem_empty[acc]);
        }
        // close the passages still open (the last one, plus trailing empty ones)
        for (int d = doc; d < it.nd; d++) flush_all(d);
        __syncwarp();
        if (epi_bar_or(!ok)) ok = false;
        const int te = threadIdx.x;  // 0..127 among the epilogue threads
        if (te < it.nd) {
            const float s = ((sh->part[parity][0][te] + sh->part[parity][1][te]) + sh->part[parity][2][te]) +
                            sh->part[parity][3][te];
            p.scores[it.out0 + te] = s;
        }
        parity ^= 1;
    }
}

// max(seed, 32 values) as a four-level tree of 3-input maxima (FMNMX3): the chain of max32 is 16 dependent
// instructions deep, and the lean epilogue's warps are the critical path of the fused kernel (ncu: the decompressors
// spend 21 % of their samples waiting for a free stage, the MMA warp 38 % waiting for a drained accumulator).
__device__ __forceinline__ float max32_tree(const uint32_t (&r)[32], float seed) {
    auto f = [&](int i) { return __uint_as_float(r[i]); };
    auto m3 = [](float a, float b, float c) { return fmaxf(fmaxf(a, b), c); };
    float t[11];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = m3(f(3 * i), f(3 * i + 1), f(3 * i + 2));
    t[10] = m3(f(30), f(31), seed);
    const float u0 = m3(t[0], t[1], t[2]), u1 = m3(t[3], t[4], t[5]), u2 = m3(t[6], t[7], t[8]), u3 = fmaxf(t[9], t[10]);
    return fmaxf(m3(u0, u1, u2), u3);
}

// Barrier over the first `nthreads` epilogue threads (whole warps), ORing a predicate like epi_bar_or.
__device__ __forceinline__ bool epi_bar_or_n(bool pred, int nthreads) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "bar.red.or.pred q, 1, %2, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r) : "r"((uint32_t)pred), "r"(nthreads) : "memory");
    return r != 0;
}

// max(seed, 16 values) as three levels of 3-input maxima instead of one 16-deep chain
__device__ __forceinline__ float max16(const uint32_t (&r)[16], float seed) {
    auto f = [&](int i) { return __uint_as_float(r[i]); };
    const float a0 = fmaxf(fmaxf(seed, f(0)), f(1)), a1 = fmaxf(fmaxf(f(2), f(3)), f(4));
    const float a2 = fmaxf(fmaxf(f(5), f(6)), f(7)), a3 = fmaxf(fmaxf(f(8), f(9)), f(10));
    const float a4 = fmaxf(fmaxf(f(11), f(12)), f(13)), a5 = fmaxf(f(14), f(15));
    return fmaxf(fmaxf(fmaxf(a0, a1), a2), fmaxf(fmaxf(a3, a4), a5));
}

// ===================== epilogue of the search pipeline: aligned passages, one m-tile, NT = 128 =====================
// The common case (Lq_pad <= 128, e.g. FLMR's 32 + 32 query tokens) gets its own lean loop: the only
// epilogue warps that exist for the barriers are the ones whose TMEM lanes can hold query rows
// (n_epi = Lq_pad / 32), the accumulator is read in 16-column pieces with the load of the next piece in flight
// while the current one is reduced, and passage boundaries are looked at once per 32 columns.  With the clamp
// at 0 the running maximum simply starts at 0.
template <bool ATMEM>
__device__ __forceinline__ void ms_epilogue_a1(const MsParams& p, MsShared* sh, uint32_t tmem_base, int item_begin,
                                               int item_end, int warp, int lane, int n_epi) {
    const int quad = warp;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    int it_tile = 0, parity = 0, cur_q = -1;
    bool ok = true;
    for (int w = item_begin; ok && w < item_end; w++) {
        const MsItem it = ms_item(p, w);
        if (it.nd == 0) continue;
        if (ATMEM && it.q != cur_q) {
            // A operand in tensor memory: this warp's 32 query rows (lane = row, 64 columns = 128 fp16).  The MMAs of
            // the previous query have all completed -- this warp has consumed their last accumulator.
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.q_rows) +
                                                              ((size_t)it.q * p.Lq_pad + quad * 32 + lane) * kDim);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t r[16];
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    const uint4 x = __ldg(src + c * 4 + v);
                    r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
                }
                tmem_st_32x16(tmem_lane + p.a_tmem_col + c * 16, r);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->a_full);
            cur_q = it.q;
        }
        const int lq = p.qlens[it.q];
        const bool rowok = (quad * 32 + lane) < lq;
        const bool live = quad * 32 < lq;               // warp-uniform: some lane holds a real query token
        float* part = sh->part[parity][quad];
        // passage ends of the item (relative to its first token) held one per lane: lane d+1 = end of passage d
        int ends_reg = 0x7fffffff;
        if (lane <= it.nd) ends_reg = it.ends[lane] - it.ends[0];
        int doc = 0;
        int next_end = __shfl_sync(0xffffffffu, ends_reg, 1);
        float runmax = 0.0f;
        auto flush = [&](int d) {
            float v = rowok ? runmax : 0.0f;
            runmax = 0.0f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) part[d] = v;
        };
        const int ntiles = (it.ntok + 127) >> 7;
        for (int t = 0; t < ntiles; t++, it_tile++) {
            const int acc = it_tile & ((1 << p.na_shift) - 1);
            if (!mbar_wait(&sh->tmem_full[acc], (it_tile >> p.na_shift) & 1, p.watchdog)) { ok = false; break; }
            tc_fence_after();
            if (live && !MS_DBG_SKIP_EPI) {
                const uint32_t tmem_acc = tmem_lane + acc * 128;
                const int nch = min(4, (it.ntok - t * 128) >> 5);     // 32-column chunks of this tile that hold tokens
                // A tcgen05.ld round trip costs ~200 cycles whatever its width, and this warp is the kernel's critical
                // path when the decompressors are fast (ncu: 97 % busy with one round trip per 16 columns).  One
                // 32-column load per passage chunk halves the round trips with the same 32 registers (deeper schemes
                // -- four 16-column buffers, two 32-column ones -- push the whole kernel over its 96-register cap and
                // spill in the decompressors: 4.5 ms instead of 2.9).
#if MS_EPI_MODE == 2
                // two 32-column loads per round trip: the second chunk's columns arrive while the first are reduced
                uint32_t ra[32], rb[32];
                auto boundary = [&](int tk0) {
                    while (tk0 == next_end && doc + 1 < it.nd) {       // this chunk opens the next passage
                        flush(doc);
                        doc++;
                        const int nx = __shfl_sync(0xffffffffu, ends_reg, min(doc + 1, 31));
                        next_end = doc + 1 < 32 ? nx : it.ntok;         // no lane 32: the 32nd passage ends with the item
                    }
                };
#pragma unroll
                for (int ch = 0; ch < 4; ch += 2) {
                    if (ch < nch) {
                        tmem_ld_32x32(tmem_acc + ch * 32, ra);
                        if (ch + 1 < nch) tmem_ld_32x32(tmem_acc + (ch + 1) * 32, rb);
                        boundary(t * 128 + ch * 32);
                        tc_wait_ld32(ra);
                        runmax = max32_tree(ra, runmax);
                        if (ch + 1 < nch) {
                            boundary(t * 128 + (ch + 1) * 32);
                            tc_wait_ld32(rb);
                            runmax = max32_tree(rb, runmax);
                        }
                    }
                }
#else
                uint32_t r[32];
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    if (ch < nch) {
                        tmem_ld_32x32(tmem_acc + ch * 32, r);
                        const int tk0 = t * 128 + ch * 32;
                        while (tk0 == next_end && doc + 1 < it.nd) {   // this chunk opens the next passage
                            flush(doc);
                            doc++;
                            const int nx = __shfl_sync(0xffffffffu, ends_reg, min(doc + 1, 31));
                            next_end = doc + 1 < 32 ? nx : it.ntok;     // no lane 32: the 32nd passage ends with the item
                        }
                        tc_wait_ld32(r);
#if MS_EPI_MODE == 1
                        runmax = max32_tree(r, runmax);
#else
                        runmax = max32(r, runmax);
#endif
                    }
                }
#endif
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->tmem_empty[acc]);
        }
        for (int d = doc; d < it.nd; d++) flush(d);     // the last passage, plus trailing empty ones
        __syncwarp();
        if (epi_bar_or_n(!ok, n_epi * 32)) ok = false;
        const int te = threadIdx.x;
        if (te < it.nd) {
            float s = sh->part[parity][0][te];
            for (int qd = 1; qd < n_epi; qd++) s += sh->part[parity][qd][te];
            p.scores[it.out0 + te] = s;
        }
        parity ^= 1;
    }
}


// ===================== epilogue for short queries (Lq_pad <= 64): four warps x 16 query rows =====================
// With Lq = 64 only two TMEM lane quadrants hold query rows and their two warps, each reading 128 columns per tile, are
// the fused kernel's critical path (ncu: ~94 % busy, the decompressors wait 21 % of their time for a free stage).  Here
// the A operand puts query rows 16q..16q+15 on lanes 0..15 of quadrant q (lanes 16..31 carry no query row), so that
// ALL FOUR warps own 16 accumulator rows, and a 16-lane load (tcgen05.ld.16x256b) hands a thread half the elements of
// the 32-lane form: thread t holds rows 16q + t/4 and + 8 of the columns 8j + 2(t%4), +1.  Per passage: two running
// maxima per thread, combined over the four threads that share the rows, then summed over the warp's 16 rows.
template <bool ATMEM>
__device__ __forceinline__ void ms_epilogue_q16(const MsParams& p, MsShared* sh, uint32_t tmem_base, int item_begin,
                                                int item_end, int quad, int lane, int n_epi) {
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int row_a = quad * 16 + (lane >> 2), row_b = row_a + 8;
    int it_tile = 0, parity = 0, cur_q = -1;
    bool ok = true;
    for (int w = item_begin; ok && w < item_end; w++) {
        const MsItem it = ms_item(p, w);
        if (it.nd == 0) continue;
        if (ATMEM && it.q != cur_q) {
            // A operand in tensor memory: lane l < 16 of this quadrant <- query row 16 quad + l (64 columns = 128 fp16),
            // lanes 16..31 <- zeros.  The MMAs of the previous query have completed: this warp consumed their last tile.
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.q_rows) +
                                                              ((size_t)it.q * p.Lq_pad + quad * 16 + (lane & 15)) * kDim);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t r[16];
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    uint4 x = make_uint4(0u, 0u, 0u, 0u);
                    if (lane < 16) x = __ldg(src + c * 4 + v);
                    r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
                }
                tmem_st_32x16(tmem_lane + p.a_tmem_col + c * 16, r);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->a_full);
            cur_q = it.q;
        }
        const int lq = p.qlens[it.q];
        const bool ok_a = row_a < lq, ok_b = row_b < lq;
        const bool live = quad * 16 < lq;               // warp-uniform: some lane holds a real query token
        float* part = sh->part[parity][quad];
        int ends_reg = 0x7fffffff;                      // lane d+1 = end of passage d relative to the item's first token
        if (lane <= it.nd) ends_reg = it.ends[lane] - it.ends[0];
        int doc = 0;
        int next_end = __shfl_sync(0xffffffffu, ends_reg, 1);
        float run_a = 0.0f, run_b = 0.0f;
        auto flush = [&](int d) {
            float va = ok_a ? run_a : 0.0f, vb = ok_b ? run_b : 0.0f;
            run_a = run_b = 0.0f;
            va = fmaxf(va, __shfl_xor_sync(0xffffffffu, va, 1));
            vb = fmaxf(vb, __shfl_xor_sync(0xffffffffu, vb, 1));
            va = fmaxf(va, __shfl_xor_sync(0xffffffffu, va, 2));
            vb = fmaxf(vb, __shfl_xor_sync(0xffffffffu, vb, 2));
            float v = va + vb;                          // the same in the four threads of a row group
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) part[d] = v;
        };
        auto boundary = [&](int tk0) {
            while (tk0 == next_end && doc + 1 < it.nd) {       // this chunk opens the next passage
                flush(doc);
                doc++;
                const int nx = __shfl_sync(0xffffffffu, ends_reg, min(doc + 1, 31));
                next_end = doc + 1 < 32 ? nx : it.ntok;         // no lane 32: the 32nd passage ends with the item
            }
        };
        const int ntiles = (it.ntok + 127) >> 7;
        for (int t = 0; t < ntiles; t++, it_tile++) {
            const int acc = it_tile & ((1 << p.na_shift) - 1);
            if (!mbar_wait(&sh->tmem_full[acc], (it_tile >> p.na_shift) & 1, p.watchdog)) { ok = false; break; }
            tc_fence_after();
            if (live && !MS_DBG_SKIP_EPI) {
                const uint32_t tmem_acc = tmem_lane + acc * 128;
                const int nch = min(4, (it.ntok - t * 128) >> 5);     // 32-column chunks of this tile that hold tokens
                uint32_t r[32];
                auto f = [&](int i) { return __uint_as_float(r[i]); };
                auto m3 = [](float a, float b, float c) { return fmaxf(fmaxf(a, b), c); };
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (2 * h < nch) {
                        tmem_ld_16x64(tmem_acc + h * 64, r);       // chunks 2h and 2h + 1 (the second may hold no tokens)
                        boundary(t * 128 + 2 * h * 32);
                        tc_wait_ld32(r);
                        run_a = m3(m3(f(0), f(1), f(4)), m3(f(5), f(8), f(9)), m3(f(12), f(13), run_a));
                        run_b = m3(m3(f(2), f(3), f(6)), m3(f(7), f(10), f(11)), m3(f(14), f(15), run_b));
                        if (2 * h + 1 < nch) {
                            boundary(t * 128 + (2 * h + 1) * 32);
                            run_a = m3(m3(f(16), f(17), f(20)), m3(f(21), f(24), f(25)), m3(f(28), f(29), run_a));
                            run_b = m3(m3(f(18), f(19), f(22)), m3(f(23), f(26), f(27)), m3(f(30), f(31), run_b));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->tmem_empty[acc]);
        }
        for (int d = doc; d < it.nd; d++) flush(d);     // the last passage, plus trailing empty ones
        __syncwarp();
        if (epi_bar_or_n(!ok, n_epi * 32)) ok = false;
        const int te = threadIdx.x;
        if (te < it.nd) {
            float s = sh->part[parity][0][te];
            for (int qd = 1; qd < n_epi; qd++) s += sh->part[parity][qd][te];
            p.scores[it.out0 + te] = s;
        }
        parity ^= 1;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kMsThreads, 2)    // <= 128 registers: two CTAs fit an SM (see ms_launch)
maxsim_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
              const __grid_constant__ CUtensorMap map_q16, const MsParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = p.MT * 128 * kDim * 2;
    const int b_bytes = p.NT * kDim * 2;
    uint8_t* sA = smem;                     // [MT][2 k-halves][128 rows][128 B]
    uint8_t* sB = smem + a_bytes;           // [NS][2 k-halves][NT rows][128 B]
    MsShared* sh = reinterpret_cast<MsShared*>(sB + p.NS * b_bytes);

    // warps 0-3: epilogue (warp = TMEM lane quadrant); 4-5: idle; 6: TMA producer; 7: TMEM alloc + MMA issue.
    // The two single-thread roles sit on SM sub-partitions 2/3, away from the epilogue warps of the first
    // 64 query tokens (the common Lq = 64 case leaves quadrants 2/3 without live rows).
    const int warp = warp_index(), lane = threadIdx.x & 31;
    // the search pipeline's common shape gets the lean epilogue (ms_epilogue_a1), run by the Lq_pad/32 warps whose
    // TMEM lanes can hold query rows
    const bool lean = (MODE == 0) && p.MT == 1 && p.NT == 128;
    // short queries (Lq_pad <= 64): 16 query rows per TMEM lane quadrant, four epilogue warps (ms_epilogue_q16); the
    // A tile is loaded as 16-row boxes onto shared-memory rows 32q..32q+15 (the rows in between are never read back)
    const bool q16 = lean && p.Lq_pad <= 64 && (MS_FUSED_Q16 != 0);   // same row order as the fused kernel: same bits
    const int n_epi = q16 ? (p.Lq_pad >> 4) : lean ? min(4, p.Lq_pad >> 5) : 4;
    const int item_begin = blockIdx.x * p.items_per_cta;
    const int item_end = min(p.num_items, item_begin + p.items_per_cta);

    if (threadIdx.x == 0) {
        mbar_init(&sh->a_full, 1);
        mbar_init(&sh->a_empty, 1);
        for (int s = 0; s < p.NS; s++) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        for (int a = 0; a < 4; a++) { mbar_init(&sh->tmem_full[a], 1); mbar_init(&sh->tmem_empty[a], n_epi); }
        fence_mbar_init();
    }
    if (warp == 7) {
        tmem_alloc(&sh->tmem_base, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_base;
    const int acc_cols = p.MT * p.NT;  // columns of one accumulator buffer

    if (warp == 6) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&map_q);
            tma_prefetch_desc(&map_d);
            tma_prefetch_desc(&map_q16);
            int cur_q = -1, a_loads = 0, it_tile = 0;
            bool ok = true;
            for (int w = item_begin; ok && w < item_end; w++) {
                const MsItem it = ms_item(p, w);
                if (it.nd == 0) continue;
                if (it.q != cur_q) {
                    if (a_loads > 0 && !mbar_wait(&sh->a_empty, (a_loads - 1) & 1, p.watchdog)) break;
                    if (q16) {
                        mbar_expect_tx(&sh->a_full, n_epi * 2 * 16 * 128);
                        for (int qd = 0; qd < n_epi; qd++)
                            for (int h = 0; h < 2; h++)
                                tma_load_2d(sA + h * (128 * 128) + qd * (32 * 128), &map_q16, &sh->a_full, h * 64,
                                            it.q * p.Lq_pad + qd * 16);
                    } else {
                        mbar_expect_tx(&sh->a_full, a_bytes);
                        for (int m = 0; m < p.MT; m++)
                            for (int h = 0; h < 2; h++)
                                tma_load_2d(sA + (m * 2 + h) * (128 * 128), &map_q, &sh->a_full, h * 64,
                                            it.q * p.Lq_pad + m * 128);
                    }
                    cur_q = it.q;
                    a_loads++;
                }
                const int ntiles = (it.ntok + p.NT - 1) / p.NT;
                for (int t = 0; t < ntiles; t++, it_tile++) {
                    const int s = it_tile % p.NS;
                    if (!mbar_wait(&sh->empty[s], ((it_tile / p.NS) & 1) ^ 1, p.watchdog)) { ok = false; break; }
                    mbar_expect_tx(&sh->full[s], b_bytes);
                    uint8_t* dst = sB + s * b_bytes;
                    const int row = (int)(it.row0 + (long long)t * p.NT);
                    tma_load_2d(dst, &map_d, &sh->full[s], 0, row);
                    tma_load_2d(dst + p.NT * 128, &map_d, &sh->full[s], 64, row);
                }
            }
        }
    } else if (warp == 7) {
        ms_mma_issue<false>(p, sh, sA, sB, b_bytes, tmem_base, item_begin, item_end, lane);
    } else if (warp < 4) {
        if (q16) {
            if (warp < n_epi) ms_epilogue_q16<false>(p, sh, tmem_base, item_begin, item_end, warp, lane, n_epi);
        } else if (lean) {
            if (warp < n_epi) ms_epilogue_a1<false>(p, sh, tmem_base, item_begin, item_end, warp, lane, n_epi);
        } else {
            ms_epilogue<MODE>(p, sh, tmem_base, item_begin, item_end, warp, lane);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 7) tmem_dealloc(tmem_base, p.tmem_cols);
}

// =============================================================================================
// Fused form for the search pipeline: decompress -> normalise -> bf16 -> swizzled smem -> tcgen05.
// The B operand never exists in HBM.  16 decompressor warps replace the TMA producer: a tile of NT
// passage tokens is built by NT/32 warps (one 32-row chunk each -- passages are 32-row aligned, so a
// chunk belongs to one passage and its rows past the passage's end are zero), written straight into the
// K-major 128B-swizzled layout the UMMA descriptor expects (16-byte chunk c of row r lands at chunk
// c ^ (r & 7); rows 128 B apart; the two 64-dim k-halves NT*128 B apart), fenced into the async proxy
// and handed to the MMA thread through the stage's mbarrier.  Each group of NT/32 warps owns one stage.
// Per token (quarter warp, 16 dims per lane), in the fp16 arithmetic of the reference's GPU branch: packed
// residual bytes -> fp16 weights from a bank-conflict-free table in shared memory, + fp16 centroid row (two
// 128-bit L2 loads per lane), sum of squares over the 8 lanes, rsqrt, scale, two 128-bit stores into the fp16
// tile.  Centroid rows are requested one batch (8 tokens) ahead, across chunk boundaries too; the shared loads
// of the next 4 tokens are issued before the shuffles of the current ones.  Epilogue = MODE 0 (aligned).
static constexpr int kFusedDecWarps = 16;
// 1- and 2-bit residuals: the packed rows of the next chunk travel global -> shared with per-lane asynchronous
// copies into a second staging buffer (no registers held across the chunk); 4 and 8 bits keep one buffer and
// carry the next chunk in registers, because their staging areas are 2-4x larger.
template <int NBITS> constexpr bool kFusedAsyncStage = (NBITS <= 2);
// Rows of a tile built by one decompressor warp (template parameter UT): 32 for 128-row tiles (4 warps per tile,
// 4 groups, 4 stages); 16 for the 64-row tiles of long queries (Lq_pad > 256: 4 warps per tile again, so that all
// 16 warps have a group -- with 32-row units only half of them would, there being fewer stages than groups).
// Two role layouts.  Regular: warps 0-3 epilogue, 4 Q TMA, 5 MMA, 6-21 decompress (704 threads, 80 registers).
// SLIM (one m-tile and Lq_pad <= 96, i.e. at most three TMEM lane quadrants hold query rows): warps 0-2 epilogue,
// warp 3 -- whose quadrant is empty -- issues the MMAs and loads the queries, 4-19 decompress: 640 threads, which
// lifts the register cap from 80 to 96 per thread for the decompressors (they spill at 80).
static constexpr int kFusedThreads = (6 + kFusedDecWarps) * 32;
static constexpr int kFusedThreadsSlim = (4 + kFusedDecWarps) * 32;

// Q16 (slim layout, Lq_pad <= 64): warps 0-3 are ALL epilogue warps, 16 query rows each (ms_epilogue_q16), and there
// is no MMA warp: the first warp of each decompressor group issues the MMAs of the tiles its group builds, right
// after its own unit -- it waits for its three siblings' arrivals, for a drained accumulator and for the epilogue's A
// rows, and issues.  Tiles may be issued slightly out of order between groups; every wait names the barrier phase of
// its own tile, and a group's leader issues its tiles in order, so the chain of dependencies never closes.
template <int NBITS, bool SLIM, int kFusedUnit, bool PRE, bool Q16 = false>
__global__ void __launch_bounds__(SLIM ? kFusedThreadsSlim : kFusedThreads, 1)
maxsim_fused_kernel(const __grid_constant__ CUtensorMap map_q, const MsParams p) {
    static_assert(!Q16 || (SLIM && MS_FUSED_ATMEM != 0), "Q16 is a variant of the slim layout with A in TMEM");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int PB = 16 * NBITS;                  // packed residual bytes per token
    constexpr bool kATmem = SLIM && (MS_FUSED_ATMEM != 0);     // the host sets p.a_tmem_col exactly for this layout
    // The slim layout exists for 128-token tiles only: with the tile shape a constant the decompressors' stage and
    // barrier addresses need fewer registers (spills of the headline variant: 44 B stored / 108 B loaded per thread,
    // two of the loads once per 32-token unit in the decompressor loop -> 4 B / 8 B, none in that loop; -3 % time).
    const int NT = SLIM ? 128 : p.NT;
    const int a_bytes = kATmem ? 0 : p.MT * 128 * kDim * 2;   // A operand in TMEM: no shared-memory copy
#ifdef MS_DBG_STAGE_STRIDE
    const int b_bytes = SLIM ? MS_DBG_STAGE_STRIDE : NT * kDim * 2;   // timing experiment only: overlapping stages (wrong results)
#else
    const int b_bytes = NT * kDim * 2;
#endif
    uint8_t* sA = smem;
    uint8_t* sB = smem + a_bytes;                   // [NS][2 k-halves][NT rows][128 B]
    uint8_t* sLUT = sB + p.NS * b_bytes;            // fp16 weight table, 256 entries x 128 B (1024-aligned)
    uint8_t* s_stage = sLUT + kLutBytes;            // [kFusedDecWarps][1 or 2 buffers][kFusedUnit tokens * PB]
    MsShared* sh = reinterpret_cast<MsShared*>(s_stage + kFusedDecWarps * (kFusedAsyncStage<NBITS> ? 2 : 1) * kFusedUnit * PB);

#if MS_UNIFORM_WARP
    const int warp = warp_index(), lane = threadIdx.x & 31;   // provably warp-uniform role branches (common.cuh)
#else
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#endif
    const bool lean = SLIM || (p.MT == 1 && NT == 128);      // see maxsim_kernel
    const int n_epi = Q16 ? (p.Lq_pad >> 4) : lean ? min(4, p.Lq_pad >> 5) : 4;
    const int item_begin = blockIdx.x * p.items_per_cta;
    const int item_end = min(p.num_items, item_begin + p.items_per_cta);
    const int wpt = NT / kFusedUnit;                // decompressor warps per tile

    if (threadIdx.x == 0) {
        mbar_init(&sh->a_full, p.a_tmem_col > 0 ? n_epi : 1);
        mbar_init(&sh->a_empty, 1);
        sh->issue_seq = 0;
        for (int s = 0; s < p.NS; s++) { mbar_init(&sh->full[s], wpt); mbar_init(&sh->empty[s], 1); }
        for (int a = 0; a < 4; a++) { mbar_init(&sh->tmem_full[a], 1); mbar_init(&sh->tmem_empty[a], n_epi); }
        fence_mbar_init();
    }
    constexpr int kMmaWarp = Q16 ? 0 : SLIM ? 3 : 5, kFirstDecWarp = SLIM ? 4 : 6;   // Q16: warp 0 only allocates TMEM
    if (warp == kMmaWarp) {
        tmem_alloc(&sh->tmem_base, 512);
        tmem_relinquish();
    }
    lut_fill_f16<NBITS>(p.wtable, sLUT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_base;
    if constexpr (SLIM && (MS_FUSED_ATMEM != 0)) {
        // query rows no TMEM quadrant of a live epilogue warp holds are zero for the whole kernel: written once here
        if (warp < 4 && warp >= n_epi) {
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; i++) z[i] = 0u;
#pragma unroll
            for (int c = 0; c < 4; c++) tmem_st_32x16(tmem_base + ((uint32_t)(warp * 32) << 16) + p.a_tmem_col + c * 16, z);
            tc_wait_st();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (SLIM && !Q16 && warp == kMmaWarp) {
        ms_mma_issue<kATmem>(p, sh, sA, sB, b_bytes, tmem_base, item_begin, item_end, lane, &map_q);
    } else if (warp < 4) {
        if constexpr (Q16) {
            if (warp < n_epi) ms_epilogue_q16<true>(p, sh, tmem_base, item_begin, item_end, warp, lane, n_epi);
        } else if (lean) {
            if (warp < n_epi) ms_epilogue_a1<kATmem>(p, sh, tmem_base, item_begin, item_end, warp, lane, n_epi);
        } else {
            ms_epilogue<0>(p, sh, tmem_base, item_begin, item_end, warp, lane);
        }
    } else if (!SLIM && warp == 4) {
        // ===================== query (A operand) TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&map_q);
            int cur_q = -1, a_loads = 0;
            for (int w = item_begin; w < item_end; w++) {
                const MsItem it = ms_item(p, w);
                if (it.nd == 0 || it.q == cur_q) continue;
                if (a_loads > 0 && !mbar_wait(&sh->a_empty, (a_loads - 1) & 1, p.watchdog)) break;
                mbar_expect_tx(&sh->a_full, a_bytes);
                for (int m = 0; m < p.MT; m++)
                    for (int h = 0; h < 2; h++)
                        tma_load_2d(sA + (m * 2 + h) * (128 * 128), &map_q, &sh->a_full, h * 64, it.q * p.Lq_pad + m * 128);
                cur_q = it.q;
                a_loads++;
            }
        }
    } else if (!SLIM && warp == 5) {
        ms_mma_issue<false>(p, sh, sA, sB, b_bytes, tmem_base, item_begin, item_end, lane);
    } else {
        // ===================== decompressor warps =====================
        // A tile of NT rows is built by wpt = NT / UT warps (one UT-row unit each; passages are 32-row aligned, so
        // a unit belongs to one passage); the 16 warps form G = 16 / wpt groups and group g builds the tiles
        // g, g + G, g + 2G, ... of the CTA's tile sequence, tile i into stage i % NS.  With UT = 16 there are two
        // groups for four stages: a warp that finishes its unit moves straight on to its unit of the group's next
        // tile instead of waiting for the slowest of the 16 warps and for the tensor core (with UT = 32 every
        // stage always has a tile under construction and the whole CTA runs in lock step).
        const int dw = warp - kFirstDecWarp;
        const int group = dw / wpt, cit = dw - group * wpt;   // this warp's group, its unit inside the tile
        // Groups in use: a power of two, at most the number of stages -- a stage's successive users must stay within
        // one mbarrier phase of each other, which a group revisiting a stage after another group's turn would not
        // (with fewer stages than groups the spare warps idle).
        int G = kFusedDecWarps / wpt;
        while (G > p.NS) G >>= 1;
        if (group >= G) goto done;
        // A quarter warp decodes one token: lane q of the quarter owns 16-byte chunk q of both k-halves of the row
        // (dims 8q..8q+7 and 64+8q..64+8q+7); one step of the warp = the 4 tokens 4*step + tsub.
        const int q = lane & 7, tsub = lane >> 3;
        constexpr int UT = kFusedUnit;
        constexpr int CB = kFusedAsyncStage<NBITS> ? MS_FUSED_CB : 1;   // steps per centroid batch (4*CB tokens, 4*CB registers)
        constexpr int NBATCH = (UT / 4) / CB;
        static_assert(NBATCH >= 2 && (NBATCH % 2) == 0, "a unit is an even number of centroid batches");
        constexpr int kStageBufs = kFusedAsyncStage<NBITS> ? 2 : 1;
        constexpr int NPASS = (UT * PB + 511) / 512;           // 512-byte passes (16 B per lane) over a unit's packed rows
        const uint32_t stage0_sa = smem_u32(s_stage + dw * (kStageBufs * UT * PB));   // this warp's residual staging area(s)
        int buf = 0;
        const uint32_t lut_sa = lut_lane_base<NBITS>(smem_u32(sLUT), lane);
        const char* cent_q = reinterpret_cast<const char*>(p.centroids) + q * 16;
        // Row j = 4*step + tsub of the unit; chunk q of a row sits at 16-byte slot q ^ (row & 7) = (q ^ tsub) ^ 4*(step & 1)
        // (units start on multiples of 16 rows).
        const uint32_t tile_lane0 = smem_u32(sB) + (cit * UT + tsub) * 128;
        const uint32_t slot_even = (uint32_t)(q ^ tsub) << 4, slot_odd = (uint32_t)(q ^ tsub ^ 4) << 4;
        const uint32_t khalf = (uint32_t)NT * 128;          // the second k-half of the tile
        int base_g = 0;       // (tiles of the earlier items) % G
        // Q16: the group's first warp issues the MMAs of its group's tiles
        int seq0 = 0;                        // tiles of the earlier items: the CTA's sequence number of this item's tile 0
        int iss_q = -1, iss_loads = 0;       // queries with work so far (the A rows of query n complete phase n of a_full)
        const uint32_t idesc = umma_idesc_f16(128, 128);
        const uint64_t db_base = umma_smem_desc_sw128(smem_u32(sB));
        int st = group;       // stage of this group's next tile and the parity of that use of the stage
        uint32_t st_par = 0;
        while (st >= p.NS) { st -= p.NS; st_par ^= 1; }
        bool ok = true;
#ifdef MS_DBG_TIMING
        long long dbg_wait = 0, dbg_first = 0, dbg_long = 0, dbg_nlong = 0, dbg_n200 = 0, dbg_units = 0;
        const long long dbg_start = clock64();
#endif
        for (int w = item_begin; ok && w < item_end; w++) {
            const MsItem it = ms_item(p, w);
            if (it.nd == 0) continue;
            if (Q16 && it.q != iss_q) { iss_q = it.q; iss_loads++; }
            // passage descriptors of the item, one per lane (lane d <- passage d)
            int64_t my_off = 0;
            int my_len = 0;
            if (lane < it.nd) {
                const int pid = p.pids[it.out0 + lane];
                my_off = p.doc_offsets[pid];
                my_len = (int)(p.doc_offsets[pid + 1] - my_off);
            }
            int ends_reg = 0x7fffffff;   // aligned start of passage `lane` relative to the item (lane nd = item end)
            if (lane <= it.nd) ends_reg = it.ends[lane] - it.ends[0];
            const int ntiles = (it.ntok + NT - 1) / NT;
            // geometry of this warp's unit in tile t: first token in the index arrays and number of real rows
            auto geom = [&](int t, int64_t& tok0, int& valid) {
                const int tk0 = t * NT + cit * UT;            // item-relative first row of the unit
                valid = 0;
                tok0 = 0;
                if (tk0 < it.ntok) {
                    const int d = __popc(__ballot_sync(0xffffffffu, ends_reg <= tk0)) - 1;
                    const int seg = tk0 - __shfl_sync(0xffffffffu, ends_reg, d);
                    tok0 = __shfl_sync(0xffffffffu, my_off, d) + seg;
                    valid = max(0, min(UT, __shfl_sync(0xffffffffu, my_len, d) - seg));
                }
            };
            // packed residuals (valid*PB bytes, 16 per lane and pass) and codes of a unit: rows past `valid` get zero
            // bytes and code 0, so they decode to finite values that the scale step replaces by zeros
            auto fetch = [&](int64_t tok0, int valid, uint32_t st_sa, int4 (&res)[NPASS], int& code, uint32_t& inv) {
#pragma unroll
                for (int v = 0; v < NPASS; v++) {
                    const int byte = v * 512 + lane * 16;
                    const bool real = byte < valid * PB;
                    if constexpr (kFusedAsyncStage<NBITS>) {
                        if (byte < UT * PB) cp_async_16(st_sa + byte, p.residuals + (real ? tok0 * PB + byte : 0), real ? 16 : 0);
                    } else {
                        res[v] = make_int4(0, 0, 0, 0);
                        if (real) res[v] = ld_stream_v4(p.residuals + tok0 * PB + byte);
                    }
                }
                if constexpr (kFusedAsyncStage<NBITS>) cp_async_commit();
                code = (lane < valid) ? ld_stream_s32(p.codes + tok0 + lane) : 0;
                // precomputed scale factor of row `lane` of the unit (0 for the rows past `valid`: they become zeros)
                if constexpr (PRE) inv = (lane < valid) ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p.inv_norms) + tok0 + lane) : 0u;
            };
            // centroid rows of batch bt (steps bt*CB .. bt*CB+CB-1) of a unit whose codes sit one per lane in `code`
            auto load_cents = [&](int code, int bt, uint4 (&clo)[CB], uint4 (&chi)[CB]) {
#pragma unroll
                for (int u = 0; u < CB; u++) {
                    const unsigned c = (unsigned)__shfl_sync(0xffffffffu, code, 4 * (bt * CB + u) + tsub);
                    const uint4* row = centroid_row(cent_q, c);
                    clo[u] = __ldg(row);
                    chi[u] = __ldg(row + 8);
                }
            };
            // decode batch bt into the tile: shared loads of step u+1 are issued before the shuffles of step u
            auto process = [&](int bt, int valid, uint32_t inv, uint32_t stage_sa, uint32_t tile_sa, const uint4 (&clo)[CB],
                               const uint4 (&chi)[CB]) {
                uint32_t wlo[4], whi[4];
                token_weights_h8<NBITS>(stage_sa + (4 * bt * CB + tsub) * PB, lut_sa, q, wlo);
                token_weights_h8<NBITS>(stage_sa + (4 * bt * CB + tsub) * PB, lut_sa, q + 8, whi);
#pragma unroll
                for (int u = 0; u < CB; u++) {
                    const int step = bt * CB + u, j = 4 * step + tsub;
                    __half2 v[8];
                    uint4 olo, ohi;
                    if constexpr (PRE) {
                        token_add_h16(clo[u], chi[u], wlo, whi, v);
                        if (u + 1 < CB) {
                            token_weights_h8<NBITS>(stage_sa + (j + 4) * PB, lut_sa, q, wlo);
                            token_weights_h8<NBITS>(stage_sa + (j + 4) * PB, lut_sa, q + 8, whi);
                        }
                        token_scale_pre_h16(v, __shfl_sync(0xffffffffu, inv, j), olo, ohi);
                    } else {
                        float ss = token_sum_h16(clo[u], chi[u], wlo, whi, v);
                        if (u + 1 < CB) {
                            token_weights_h8<NBITS>(stage_sa + (j + 4) * PB, lut_sa, q, wlo);
                            token_weights_h8<NBITS>(stage_sa + (j + 4) * PB, lut_sa, q + 8, whi);
                        }
                        ss = quarter_sum(ss);
                        token_scale_h16(v, ss, j < valid, olo, ohi);
                    }
                    const uint32_t dst = tile_sa + (((CB & 1) ? (step & 1) : (u & 1)) ? slot_odd : slot_even) + step * 512;
#if MS_DBG_SKIP_STS     // timing experiment only: keep the arithmetic alive, store nothing
                    if ((olo.x ^ olo.y ^ olo.z ^ olo.w ^ ohi.x ^ ohi.y ^ ohi.z ^ ohi.w) == 0x12345678u) sts_v4u32_relaxed(dst, olo.x, olo.y, olo.z, olo.w);
#else
                    sts_v4u32_relaxed(dst, olo.x, olo.y, olo.z, olo.w);
                    sts_v4u32_relaxed(dst + khalf, ohi.x, ohi.y, ohi.z, ohi.w);
#endif
                }
            };
            int4 pres[NPASS];
            int pcode = 0, pvalid = 0, ptile = -1;
            uint32_t pinv = 0;
            uint4 alo[CB], ahi[CB], blo[CB], bhi[CB];
            bool a_ready = false;                               // batch 0 of the coming unit is already in alo/ahi
            const int first = (group - base_g) & (G - 1);       // this group's first tile of the item
            base_g = (base_g + ntiles) & (G - 1);
            for (int t = first; t < ntiles; t += G) {
                const uint32_t stage_sa = stage0_sa + buf * (UT * PB);
                const uint32_t tile_sa = tile_lane0 + st * b_bytes;
                int4 res[NPASS];
                int code, valid;
                uint32_t inv = 0;
                if (ptile == t) {                               // requested while the previous unit was being built
                    if constexpr (!kFusedAsyncStage<NBITS>) {
#pragma unroll
                        for (int v = 0; v < NPASS; v++) res[v] = pres[v];
                    }
                    code = pcode;
                    valid = pvalid;
                    inv = pinv;
                } else {
                    int64_t tok0;
                    geom(t, tok0, valid);
                    fetch(tok0, valid, stage_sa, res, code, inv);
                }
                if (!a_ready && valid > 0) load_cents(code, 0, alo, ahi);
                a_ready = false;
                const bool more = t + G < ntiles;
                if (more) {                                     // this warp's next unit of the item: loads in flight now
                    int64_t ntok0;
                    geom(t + G, ntok0, pvalid);
                    fetch(ntok0, pvalid, stage0_sa + (buf ^ 1) * (UT * PB), pres, pcode, pinv);
                    ptile = t + G;
                }
#ifdef MS_DBG_TIMING
                const long long dbg_t0 = clock64();
#endif
                if (!mbar_wait(&sh->empty[st], st_par ^ 1, p.watchdog)) { ok = false; break; }
#ifdef MS_DBG_TIMING
                {
                    const long long dw_ = clock64() - dbg_t0;
                    dbg_wait += dw_;
                    if (t == first) dbg_first += dw_;                 // first unit of an item
                    if (dw_ > 3000) { dbg_long += dw_; dbg_nlong++; }
                    if (dw_ > 200) dbg_n200++;
                    dbg_units++;
                }
#endif
                if constexpr (kFusedAsyncStage<NBITS>) {        // this unit's rows have landed (the next unit's may not)
                    if (more) cp_async_wait<1>(); else cp_async_wait<0>();
                    __syncwarp();
                    buf ^= 1;
                }
                if (valid == 0 && t * NT + cit * UT < it.ntok) {
                    // the unit lies entirely in the alignment padding of a passage (UT = 16 only): its rows must be zero
#pragma unroll
                    for (int step = 0; step < UT / 4; step++) {
                        const uint32_t dst = tile_sa + ((step & 1) ? slot_odd : slot_even) + step * 512;
                        sts_v4u32_relaxed(dst, 0u, 0u, 0u, 0u);
                        sts_v4u32_relaxed(dst + khalf, 0u, 0u, 0u, 0u);
                    }
                }
                if (valid > 0 && !MS_DBG_SKIP_DEC) {
                    if constexpr (!kFusedAsyncStage<NBITS>) {
#pragma unroll
                        for (int v = 0; v < NPASS; v++)
                            if (v * 512 + lane * 16 < UT * PB)
                                sts_v4u32(stage_sa + v * 512 + lane * 16, res[v].x, res[v].y, res[v].z, res[v].w);
                        __syncwarp();
                    }
                    // A passage's last unit is half empty on average: only the pairs of batches that hold real rows are
                    // decoded, the rows behind them are stored as zeros (what decoding them would have produced).
                    const int nb_used = (MS_SKIP_PAD_BATCHES && UT == 32) ? min(NBATCH, ((valid + 4 * CB - 1) / (4 * CB) + 1) & ~1) : NBATCH;   // 16-row units: measured no gain
#pragma unroll 1
                    for (int bt = 0; bt < nb_used; bt += 2) {   // two batches per trip: A = bt, B = bt + 1
                        load_cents(code, bt + 1, blo, bhi);
                        process(bt, valid, inv, stage_sa, tile_sa, alo, ahi);
                        if (bt + 2 < nb_used) {
                            load_cents(code, bt + 2, alo, ahi);
                        } else if (more && pvalid > 0) {        // batch 0 of the next unit (its codes arrived long ago)
                            load_cents(pcode, 0, alo, ahi);
                            a_ready = true;
                        }
                        process(bt + 1, valid, inv, stage_sa, tile_sa, blo, bhi);
                    }
#pragma unroll 1
                    for (int step = nb_used * CB; step < UT / 4; step++) {
                        const uint32_t dst = tile_sa + ((step & 1) ? slot_odd : slot_even) + step * 512;
                        sts_v4u32_relaxed(dst, 0u, 0u, 0u, 0u);
                        sts_v4u32_relaxed(dst + khalf, 0u, 0u, 0u, 0u);
                    }
                }
                fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh->full[st]);
                if constexpr (Q16) {
                    if (cit == 0) {            // issue this tile: all four units built, accumulator drained, A rows present
                        const int seq = seq0 + t, acc = seq & 1;
                        bool got = mbar_wait(&sh->full[st], st_par, p.watchdog);
                        // Tiles are issued in sequence: a parity wait cannot tell "two uses behind" from "ready", and a
                        // leader that ran ahead of the epilogue by a whole group period would otherwise overwrite a live
                        // accumulator.  The token passes from the leader of tile seq - 1.
                        if (got && lds_acquire_s32(&sh->issue_seq) != seq) {
                            const uint64_t t0 = globaltimer_ns();
                            while (lds_acquire_s32(&sh->issue_seq) != seq) {
                                __nanosleep(20);
                                if (globaltimer_ns() - t0 > 2000000000ull) {
                                    if (p.watchdog) atomicExch(p.watchdog, 1);
                                    got = false;
                                    break;
                                }
                            }
                        }
                        got = got && mbar_wait(&sh->tmem_empty[acc], ((seq >> 1) & 1) ^ 1, p.watchdog);
                        got = got && mbar_wait(&sh->a_full, (iss_loads - 1) & 1, p.watchdog);
                        if (!got) { ok = false; break; }
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t db0 = db_base + (uint64_t)(st * (b_bytes >> 4));
                            const uint32_t d_tmem = tmem_base + acc * 128;
#pragma unroll
                            for (int k = 0; k < 8; k++)
                                umma_f16_ts(d_tmem, tmem_base + p.a_tmem_col + k * 8,
                                            db0 + (uint64_t)((k >> 2) * ((128 * 128) >> 4) + (k & 3) * 2), idesc, k > 0);
                            umma_commit(&sh->empty[st]);
                            umma_commit(&sh->tmem_full[acc]);
                            sts_release_s32(&sh->issue_seq, seq + 1);
                        }
                        __syncwarp();
                    }
                }
                st += G;                       // the group's next tile
                while (st >= p.NS) { st -= p.NS; st_par ^= 1; }
            }
            seq0 += ntiles;
        }
#ifdef MS_DBG_TIMING
        if (lane == 0 && blockIdx.x < 160) {
            unsigned long long* o = g_ms_dbg + (blockIdx.x * 16 + dw) * 8;
            o[0] = (unsigned long long)dbg_wait;
            o[1] = (unsigned long long)(clock64() - dbg_start);
            o[2] = (unsigned long long)dbg_first;
            o[3] = (unsigned long long)dbg_long;
            o[4] = (unsigned long long)dbg_nlong;
            o[5] = (unsigned long long)dbg_n200;
            o[6] = (unsigned long long)dbg_units;
        }
#endif
    }

done:
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

static int ms_configure(MsParams& p, int Lq_pad) {
    p.Lq_pad = Lq_pad;
    p.MT = (Lq_pad + 127) / 128;
    PLAID_CHECK_ARG(p.MT >= 1 && p.MT <= kMsMaxMT, PLAID_ERR_UNSUPPORTED, "maxsim: Lq_pad=%d > 512 query tokens", Lq_pad);
    p.NT = (p.MT <= 2) ? 128 : 64;   // 2 accumulator buffers x MT x NT <= 512 TMEM columns
    p.na_shift = (4 * p.MT * p.NT <= 512) ? 2 : 1;
    p.tmem_cols = 512;
    const int a_bytes = p.MT * 128 * kDim * 2, b_bytes = p.NT * kDim * 2;
    int ns = (200 * 1024 - a_bytes) / b_bytes;
    p.NS = ns > kMsMaxStages ? kMsMaxStages : ns;
    PLAID_CHECK_ARG(p.NS >= 2, PLAID_ERR_UNSUPPORTED, "maxsim: no room for a 2-stage ring");
    return PLAID_OK;
}

static int ms_launch(const void* Qb, int q_rows, const void* D, uint64_t d_rows, MsParams& p, cudaStream_t st) {
    CUtensorMap map_q, map_d, map_q16;
    int rc;
    if ((rc = make_bf16_2d_map(&map_q, Qb, (uint64_t)q_rows, kDim, 128)) != PLAID_OK) return rc;
    if ((rc = make_bf16_2d_map(&map_d, D, d_rows, kDim, p.NT)) != PLAID_OK) return rc;
    if ((rc = make_bf16_2d_map(&map_q16, Qb, (uint64_t)q_rows, kDim, 16)) != PLAID_OK) return rc;   // 16-row boxes (q16 epilogue)
    // One m-tile (Lq_pad <= 128): TWO CTAs per SM, each with 2 B stages, 2 accumulators (256 TMEM columns) and its
    // own TMA / MMA / epilogue warps.  With HBM-resident operands the epilogue warps of one CTA (only Lq_pad/32 of
    // the four hold query rows) cannot keep up with the stream; two independent pipelines double them.
    const bool dual = p.MT == 1;
    if (dual) {
        p.NS = 2;
        p.na_shift = 1;
        p.tmem_cols = 256;
    }
    const int smem = 1024 + p.MT * 128 * kDim * 2 + p.NS * p.NT * kDim * 2 + (int)sizeof(MsShared) + 64;
    const int mode = p.padded ? 2 : (p.aligned ? 0 : 1);
    static int configured[3][kMaxDevices] = {{0}, {0}, {0}};
    const void* fn = mode == 0 ? (const void*)maxsim_kernel<0> : mode == 1 ? (const void*)maxsim_kernel<1>
                                                                           : (const void*)maxsim_kernel<2>;
    if ((rc = ensure_dynamic_smem(fn, smem, configured[mode], true)) != PLAID_OK) return rc;
    int grid = sm_count() * (dual ? 2 : 1);
    if (grid > p.num_items) grid = p.num_items;
    p.items_per_cta = (p.num_items + grid - 1) / grid;
    grid = (p.num_items + p.items_per_cta - 1) / p.items_per_cta;
    if (mode == 0) maxsim_kernel<0><<<grid, kMsThreads, smem, st>>>(map_q, map_d, map_q16, p);
    else if (mode == 1) maxsim_kernel<1><<<grid, kMsThreads, smem, st>>>(map_q, map_d, map_q16, p);
    else maxsim_kernel<2><<<grid, kMsThreads, smem, st>>>(map_q, map_d, map_q16, p);
    PLAID_LAUNCH_OK("maxsim_kernel");
    return PLAID_OK;
}

static int ms_launch_fused(const void* Qb, int q_rows, int nbits, MsParams& p, cudaStream_t st) {
    CUtensorMap map_q;
    int rc;
    if ((rc = make_bf16_2d_map(&map_q, Qb, (uint64_t)q_rows, kDim, 128)) != PLAID_OK) return rc;
    // as many B stages as shared memory allows; the decompressor groups share them in tile order
    const int stage_bufs = nbits <= 2 ? 2 : 1;       // kFusedAsyncStage
    const bool slim_shape = p.MT == 1 && p.NT == 128 && p.Lq_pad <= 96;
    const int unit = (p.NT == 64 || (slim_shape && MS_SLIM_UT == 16)) ? 16 : 32;   // rows per decompressor warp (template parameter UT)
    const bool slim = p.MT == 1 && p.NT == 128 && p.Lq_pad <= 96;
    const bool q16 = slim && p.Lq_pad <= 64 && (MS_FUSED_ATMEM != 0) && (MS_FUSED_Q16 != 0);   // four epilogue warps x 16 query rows
#if MS_FUSED_ATMEM
    if (slim) {                                      // A operand in tensor memory: 2 accumulators (256 columns) + 64 columns of A
        p.na_shift = 1;
        p.a_tmem_col = 256;
        p.q_rows = Qb;
    }
#endif
    const int fixed = 1024 + (p.a_tmem_col > 0 ? 0 : p.MT * 128 * kDim * 2) + kLutBytes + kFusedDecWarps * stage_bufs * unit * 16 * nbits +
                      (int)sizeof(MsShared) + 64;
#ifdef MS_DBG_STAGE_STRIDE
    const int per_stage = slim ? MS_DBG_STAGE_STRIDE : p.NT * kDim * 2;
#else
    const int per_stage = p.NT * kDim * 2;
#endif
    p.NS = (227 * 1024 - fixed) / per_stage;
    if (p.NS > kMsMaxStages) p.NS = kMsMaxStages;
    PLAID_CHECK_ARG(p.NS >= 2, PLAID_ERR_UNSUPPORTED, "maxsim_fused: shared memory too small for Lq_pad=%d, nbits=%d", p.Lq_pad, nbits);
    const int smem = fixed + p.NS * per_stage;
    const bool pre = p.inv_norms != nullptr;
#define PLAID_FUSED_FN2(NB, PRE_)                                                                       \
    (q16 ? (const void*)maxsim_fused_kernel<NB, true, MS_SLIM_UT, PRE_, true>                                   \
         : slim ? (const void*)maxsim_fused_kernel<NB, true, MS_SLIM_UT, PRE_>                                \
          : unit == 16 ? (const void*)maxsim_fused_kernel<NB, false, 16, PRE_> : (const void*)maxsim_fused_kernel<NB, false, 32, PRE_>)
#define PLAID_FUSED_FN(NB) (pre ? PLAID_FUSED_FN2(NB, true) : PLAID_FUSED_FN2(NB, false))
    const void* fn = nbits == 1 ? PLAID_FUSED_FN(1) : nbits == 2 ? PLAID_FUSED_FN(2) : nbits == 4 ? PLAID_FUSED_FN(4) : PLAID_FUSED_FN(8);
#undef PLAID_FUSED_FN
#undef PLAID_FUSED_FN2
    static int configured[8][9][kMaxDevices] = {{{0}}};
    const int variant = (q16 ? 3 : slim ? 1 : unit == 16 ? 2 : 0) + (pre ? 4 : 0);
    if ((rc = ensure_dynamic_smem(fn, smem, configured[variant][nbits])) != PLAID_OK) return rc;
    int grid = sm_count();
    if (grid > p.num_items) grid = p.num_items;
    p.items_per_cta = (p.num_items + grid - 1) / grid;
    grid = (p.num_items + p.items_per_cta - 1) / p.items_per_cta;
    void* args[] = {(void*)&map_q, (void*)&p};
    PLAID_CUDA_OK(cudaLaunchKernel(fn, dim3(grid), dim3(slim ? kFusedThreadsSlim : kFusedThreads), args, (size_t)smem, st));
    PLAID_LAUNCH_OK("maxsim_fused_kernel");
    return PLAID_OK;
}

}  // namespace plaid

extern "C" int plaid_maxsim_fused(const void* Qh_f16, const int32_t* qlens, int B, int B_pad, int Lq_pad,
                                  const int32_t* pids, const int32_t* counts, int pid_stride, const int32_t* tok_offsets,
                                  const int64_t* offsets, const float* W, const uint8_t* residuals, const int32_t* codes,
                                  const void* centroids_f16, int C, int nbits, const void* inv_norms_f16, float* scores,
                                  int* watchdog, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(Qh_f16 && qlens && pids && counts && tok_offsets && offsets && W && residuals && codes && centroids_f16 &&
                        scores,
                    PLAID_ERR_ARG, "plaid_maxsim_fused: null pointer");
    PLAID_CHECK_ARG(B >= 0 && B_pad >= B && pid_stride >= 1 && C > 0 && Lq_pad >= 32 && (Lq_pad % 32) == 0, PLAID_ERR_ARG,
                    "plaid_maxsim_fused: bad sizes");
    PLAID_CHECK_ARG(nbits == 1 || nbits == 2 || nbits == 4 || nbits == 8, PLAID_ERR_UNSUPPORTED,
                    "plaid_maxsim_fused: nbits=%d not in {1,2,4,8}", nbits);
    PLAID_CHECK_ARG((reinterpret_cast<uintptr_t>(residuals) & 15) == 0 && (reinterpret_cast<uintptr_t>(centroids_f16) & 15) == 0,
                    PLAID_ERR_ARG, "plaid_maxsim_fused: residuals/centroids must be 16-byte aligned");
    if (B == 0) return PLAID_OK;
    MsParams p{};
    int rc;
    if ((rc = ms_configure(p, Lq_pad)) != PLAID_OK) return rc;
    p.qlens = qlens;
    p.padded = 0;
    p.aligned = 1;
    p.tok_offsets = tok_offsets;
    p.counts = counts;
    p.pid_stride = pid_stride;
    p.tok_stride = 0;
    p.scores = scores;
    p.clamp_zero = 1;
    p.watchdog = watchdog;
    p.pids = pids;
    p.doc_offsets = offsets;
    p.codes = codes;
    p.residuals = residuals;
    p.centroids = reinterpret_cast<const __half*>(centroids_f16);
    p.wtable = W;
    p.inv_norms = reinterpret_cast<const __half*>(inv_norms_f16);
    p.C = C;
    p.ab_f16 = 1;
    p.groups_per_query = (pid_stride + kMsGD - 1) / kMsGD;
    p.num_items = B * p.groups_per_query;
    return ms_launch_fused(Qh_f16, B_pad * Lq_pad, nbits, p, (cudaStream_t)stream);
}

extern "C" int plaid_maxsim_packed(const void* Qb_bf16, const int32_t* qlens, int B, int B_pad, int Lq_pad,
                                   const void* D_bf16, const int32_t* tok_offsets, const int32_t* counts, int pid_stride,
                                   int tok_stride, int clamp_zero, int aligned32, int operands_f16, float* scores,
                                   int* watchdog, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(Qb_bf16 && qlens && D_bf16 && tok_offsets && counts && scores, PLAID_ERR_ARG,
                    "plaid_maxsim_packed: null pointer");
    PLAID_CHECK_ARG(B >= 0 && B_pad >= B && pid_stride >= 1 && tok_stride >= 1 && Lq_pad >= 32 && (Lq_pad % 32) == 0,
                    PLAID_ERR_ARG, "plaid_maxsim_packed: bad sizes");
    PLAID_CHECK_ARG((long long)B * tok_stride < (1ll << 31), PLAID_ERR_UNSUPPORTED,
                    "plaid_maxsim_packed: B*tok_stride must stay below 2^31 rows per call");
    if (B == 0) return PLAID_OK;
    MsParams p{};
    int rc;
    if ((rc = ms_configure(p, Lq_pad)) != PLAID_OK) return rc;
    p.qlens = qlens;
    p.padded = 0;
    p.tok_offsets = tok_offsets;
    p.counts = counts;
    p.pid_stride = pid_stride;
    p.tok_stride = tok_stride;
    p.scores = scores;
    p.clamp_zero = clamp_zero;
    PLAID_CHECK_ARG(!aligned32 || clamp_zero, PLAID_ERR_ARG,
                    "plaid_maxsim_packed: the aligned layout relies on the clamp at 0 to ignore its zero pad rows");
    p.aligned = aligned32 ? 1 : 0;
    p.ab_f16 = operands_f16 ? 1 : 0;
    p.watchdog = watchdog;
    p.groups_per_query = (pid_stride + kMsGD - 1) / kMsGD;
    p.num_items = B * p.groups_per_query;
    return ms_launch(Qb_bf16, B_pad * Lq_pad, D_bf16, (uint64_t)B * tok_stride, p, (cudaStream_t)stream);
}

extern "C" int plaid_colbert_score_padded(const void* Qb_bf16, const int32_t* qlens, int nQ, int nQ_pad, int Lq_pad,
                                          const void* D_padded_bf16, const uint8_t* D_mask, int64_t n, int Ld,
                                          int docs_per_query, float* scores, float* scores_raw, int Lq_out,
                                          int* watchdog, void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(Qb_bf16 && qlens && D_padded_bf16 && D_mask && scores, PLAID_ERR_ARG,
                    "plaid_colbert_score_padded: null pointer");
    PLAID_CHECK_ARG(nQ >= 1 && nQ_pad >= nQ && n >= 0 && Ld >= 1 && docs_per_query >= 1 && Lq_pad >= 32 && (Lq_pad % 32) == 0,
                    PLAID_ERR_ARG, "plaid_colbert_score_padded: bad sizes");
    PLAID_CHECK_ARG((long long)nQ * docs_per_query >= n, PLAID_ERR_ARG,
                    "plaid_colbert_score_padded: %d queries x %d passages/query < n=%lld", nQ, docs_per_query, (long long)n);
    PLAID_CHECK_ARG(n * Ld < (1ll << 31), PLAID_ERR_UNSUPPORTED,
                    "plaid_colbert_score_padded: n*Ld must stay below 2^31 rows per call");
    if (n == 0) return PLAID_OK;
    MsParams p{};
    int rc;
    if ((rc = ms_configure(p, Lq_pad)) != PLAID_OK) return rc;
    p.qlens = qlens;
    p.padded = 1;
    p.mask = D_mask;
    p.n_docs = n;
    p.Ld = Ld;
    p.docs_per_query = docs_per_query;
    p.scores_raw = scores_raw;
    p.Lq_out = Lq_out;
    p.scores = scores;
    p.clamp_zero = 0;
    p.watchdog = watchdog;
    p.groups_per_query = (docs_per_query + kMsGD - 1) / kMsGD;
    const long long nq_used = (n + docs_per_query - 1) / docs_per_query;
    p.num_items = (int)(nq_used * p.groups_per_query);
    return ms_launch(Qb_bf16, nQ_pad * Lq_pad, D_padded_bf16, (uint64_t)n * Ld, p, (cudaStream_t)stream);
}
