// centroid.cu -- centroid scoring on tcgen05 / TMEM, operands fed by TMA.
//
// Replaces `scores = centroids @ Q.T` + `topk(ncells, dim=0)` of
// CB/search/candidate_generation.py:12-20 and `idx = centroid_scores.max(-1).values >= thr` of
// CB/search/index_storage.py:115, for a whole batch of queries at once.
//
// Tile orientation: M = 128 accumulator lanes = 4 queries x 32 candidate-stage tokens (so each
// epilogue warp owns exactly one query and lane = query token), N = 256 centroids per tile,
// K = 128 = 8 UMMA k-steps of 16.  A CTA keeps its 4 queries' A operand resident in shared memory
// and streams a contiguous range of centroid tiles through a 2-stage TMA ring; two 256-column fp32
// accumulators in TMEM (all 512 columns) let the MMA of tile i+1 overlap the epilogue of tile i.
//
// Epilogue (16 warps; warp = (lane quadrant, 64-column quarter) -- one warp per scheduler is latency-bound on its
// TMEM load -> convert -> store chain, four hide it), straight out of TMEM:
//   * S[b, c, 0..31] -- for one centroid the warp's 32 lanes write 32 consecutive floats, i.e. one
//     full 128-byte line of the reference's [C, nq] layout;
//   * idx bit (b, c) = max over the query's tokens >= threshold, as a warp vote (any token >= threshold),
//     only evaluated per column when some thread's chunk maximum reaches the threshold;
//   * per thread (= query token) a running top-ncells (score desc, centroid id asc) over all
//     columns it sees -- partial per centroid range, merged in candidates.cu.
// The kernel is bound by the S write (4*C*32 bytes per query), not by the tensor pipe.
#include "common.cuh"
#include <cuda_fp16.h>

namespace plaid {

static constexpr int kCsM = 128;            // accumulator rows: 4 queries x 32 tokens
static constexpr int kCsN = 256;            // centroids per tile
static constexpr int kCsStages = 2;         // B-operand ring depth
static constexpr int kCsParts = PLAID_CELL_LISTS_PER_RANGE;   // column parts of a tile, one epilogue warp per (quadrant, part)
static constexpr int kCsPartCols = kCsN / kCsParts;
static constexpr int kCsThreads = (4 + 4 * kCsParts) * 32;    // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. epilogue
static constexpr int kCsABytes = kCsM * kDim * 2;        // 32 KB
static constexpr int kCsBBytes = kCsN * kDim * 2;        // 64 KB per stage
// S leaves through shared memory: every epilogue warp transposes its 32 tokens x 32 centroids block into the table's
// [centroid][token] order in a private staging area (32 rows of 32 values = one contiguous 2 KB (fp16) / 4 KB (fp32)
// run of the table) and one lane hands it to the bulk-copy engine (cp.async.bulk shared -> global), so the table is
// written in full lines by the copy engine instead of 32 two-byte stores per lane and chunk.  fp16: two blocks per
// warp (the copy of one drains while the next is built); fp32: one.
#ifndef PLAID_CS_BULK
#define PLAID_CS_BULK 1
#endif
#ifndef PLAID_CS_PREFETCH
#define PLAID_CS_PREFETCH 0
#endif
static constexpr int kCsStagePerWarp = 4096;
static constexpr int kCsStageBytes = PLAID_CS_BULK ? 4 * kCsParts * kCsStagePerWarp : 0;
static constexpr int kCsSmemBytes = 1024 + kCsABytes + kCsStages * kCsBBytes + kCsStageBytes + 256;

__device__ __forceinline__ void sts_b16(uint32_t addr, unsigned short v) {
    asm volatile("st.shared.b16 [%0], %1;" :: "r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
// shared -> global bulk copy (16-byte aligned, size a multiple of 16), tracked by the issuing thread's bulk groups
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// bit p of the low half -> bit 2p, bit p of the high half -> bit 2p+1 (packed half2 compares see columns 2p / 2p+1 there)
__device__ __forceinline__ uint32_t interleave16(uint32_t v) {
    uint32_t x = v & 0xffffu, y = v >> 16;
    x = (x | (x << 8)) & 0x00ff00ffu; y = (y | (y << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu; y = (y | (y << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u; y = (y | (y << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u; y = (y | (y << 1)) & 0x55555555u;
    return x | (y << 1);
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct CsBarriers {
    uint64_t a_full, a_empty;
    uint64_t full[kCsStages];
    uint64_t empty[kCsStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    int abort_flag;
};

// ST = float: S as the reference's fp32 table; __half: S rounded to fp16 (what the reference's GPU branch computes in,
// candidate_generation.py:52) -- half the bytes to write and to gather.  NC = ncells rounded up to 1/2/4/8 (register-
// resident running top-NC per thread).
template <typename ST, int NC>
__global__ void __launch_bounds__(kCsThreads, 1)
centroid_scores_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                       const int32_t* __restrict__ qlens, int C, int Lq_pad, float threshold, int ncells, int csplit,
                       int num_units, ST* __restrict__ S, uint32_t* __restrict__ idx_bits, float* __restrict__ cell_val,
                       int32_t* __restrict__ cell_idx, int* __restrict__ watchdog) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: required by the 128B swizzle atoms the UMMA descriptors describe
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                          // [2 k-halves][128 rows][128 B]
    uint8_t* sB = smem + kCsABytes;              // [stage][2 k-halves][256 rows][128 B]
    uint8_t* sStage = sB + kCsStages * kCsBBytes;   // [epilogue warp][kCsStagePerWarp]
    CsBarriers* bar = reinterpret_cast<CsBarriers*>(smem + kCsABytes + kCsStages * kCsBBytes + kCsStageBytes);
    __shared__ float s_cut[4][32];   // per (query, token): best NC-th value any of the column-part warps has seen

    const int warp = warp_index(), lane = threadIdx.x & 31;
    // Work unit = (group of 4 queries, centroid range), numbered group-major; the persistent CTAs split the unit
    // sequence into contiguous, equally long runs (a run's units mostly share their queries, so the A operand is
    // reloaded only when the group changes), which keeps all SMs busy whatever the number of query groups.
    const int tiles_total = (C + kCsN - 1) / kCsN;
    const int tiles_per_split = (tiles_total + csplit - 1) / csplit;
    const int unit_begin = (int)((long long)blockIdx.x * num_units / gridDim.x);
    const int unit_end = (int)((long long)(blockIdx.x + 1) * num_units / gridDim.x);
    auto unit_tiles = [&](int u, int& qgroup, int& split, int& tile_begin) -> int {
        qgroup = u / csplit;
        split = u - qgroup * csplit;
        tile_begin = split * tiles_per_split;
        return max(0, min(tiles_total, tile_begin + tiles_per_split) - tile_begin);
    };

    if (threadIdx.x == 0) {
        mbar_init(&bar->a_full, 1);
        mbar_init(&bar->a_empty, 1);
        for (int s = 0; s < kCsStages; s++) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&bar->tmem_full[a], 1); mbar_init(&bar->tmem_empty[a], 4 * kCsParts); }
        bar->abort_flag = 0;
        fence_mbar_init();
    }
    if (threadIdx.x < 128) s_cut[threadIdx.x >> 5][threadIdx.x & 31] = -INFINITY;
    if (warp == 2) {
        tmem_alloc(&bar->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&map_q);
            tma_prefetch_desc(&map_c);
            int cur_g = -1, a_loads = 0, it = 0;
            bool ok = true;
            for (int u = unit_begin; ok && u < unit_end; u++) {
                int qgroup, split, tile_begin;
                const int ntiles = unit_tiles(u, qgroup, split, tile_begin);
                if (qgroup != cur_g) {
                    // the MMAs that read the previous group's queries have retired (the issuer commits a_empty on the switch)
                    if (a_loads > 0 && !mbar_wait(&bar->a_empty, (a_loads - 1) & 1, watchdog)) break;
                    mbar_expect_tx(&bar->a_full, kCsABytes);
                    for (int h = 0; h < 2; h++)
                        for (int q = 0; q < 4; q++)
                            tma_load_2d(sA + h * (kCsM * 128) + q * (32 * 128), &map_q, &bar->a_full, h * 64,
                                        (qgroup * 4 + q) * Lq_pad);
                    cur_g = qgroup;
                    a_loads++;
                }
                for (int t = 0; t < ntiles; t++, it++) {
                    const int s = it % kCsStages;
                    if (!mbar_wait(&bar->empty[s], ((it / kCsStages) & 1) ^ 1, watchdog)) { ok = false; break; }
                    mbar_expect_tx(&bar->full[s], kCsBBytes);
                    uint8_t* dst = sB + s * kCsBBytes;
                    const int row = (tile_begin + t) * kCsN;
                    tma_load_2d(dst, &map_c, &bar->full[s], 0, row);
                    tma_load_2d(dst + kCsN * 128, &map_c, &bar->full[s], 64, row);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kCsM, kCsN);
            int cur_g = -1, a_loads = 0, it = 0;
            bool ok = true;
            for (int u = unit_begin; ok && u < unit_end; u++) {
                int qgroup, split, tile_begin;
                const int ntiles = unit_tiles(u, qgroup, split, tile_begin);
                if (qgroup != cur_g) {
                    if (a_loads > 0) umma_commit(&bar->a_empty);   // every MMA issued so far has read the old queries
                    if (!mbar_wait(&bar->a_full, a_loads & 1, watchdog)) break;
                    cur_g = qgroup;
                    a_loads++;
                }
                for (int t = 0; t < ntiles; t++, it++) {
                    const int s = it % kCsStages, acc = it & 1;
                    if (!mbar_wait(&bar->tmem_empty[acc], ((it >> 1) & 1) ^ 1, watchdog)) { ok = false; break; }
                    if (!mbar_wait(&bar->full[s], (it / kCsStages) & 1, watchdog)) { ok = false; break; }
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + s * kCsBBytes);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        // k-step of 16 bf16 = 32 B inside a 128 B swizzle row; k >= 4 -> second k-half sub-tile
                        const uint64_t da = umma_smem_desc_sw128(a0 + (k >> 2) * (kCsM * 128) + (k & 3) * 32);
                        const uint64_t db = umma_smem_desc_sw128(b0 + (k >> 2) * (kCsN * 128) + (k & 3) * 32);
                        umma_bf16(tmem_base + acc * kCsN, da, db, idesc, k > 0);
                    }
                    umma_commit(&bar->empty[s]);        // smem stage reusable once these MMAs retire
                    umma_commit(&bar->tmem_full[acc]);  // accumulator ready for the epilogue
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int quad = warp & 3, part = (warp - 4) >> 2;
        const __half2 thr2 = __half2half2(__float2half_ru(threshold));
#if PLAID_CS_BULK
        constexpr int kBlkBytes = 32 * PLAID_NQ_MAX * (int)sizeof(ST);          // one staged block: 32 centroids x 32 tokens
        constexpr int kBlkPerCopy = kCsStagePerWarp / kBlkBytes;                // fp16: both blocks of a tile part leave as ONE 4 KB copy
        const uint32_t stage_base = smem_u32(sStage) + (uint32_t)(warp - 4) * kCsStagePerWarp;
        const uint32_t stage_sa = stage_base + lane * (uint32_t)sizeof(ST);
        const bool store = S != nullptr;     // S == NULL: only the top-ncells lists are wanted (index build: argmax)
#endif
        int cur_g = -1, it = 0;
        bool ok = true;
        for (int u = unit_begin; u < unit_end; u++) {
        int qgroup, split, tile_begin;
        const int ntiles = unit_tiles(u, qgroup, split, tile_begin);
        if (qgroup != cur_g) {
            // New queries: the bound the column-part warps share (s_cut) belongs to the old ones.  All 16 epilogue warps
            // have left the old group before it is reset, and nobody reads it before the reset is done.  (A warp that gave
            // up on a barrier wait still walks the units, so the named barrier always sees all of them.)
            if (cur_g >= 0) {
                asm volatile("bar.sync 1, %0;" :: "n"(4 * kCsParts * 32) : "memory");
                if (part == 0) s_cut[quad][lane] = -INFINITY;
                asm volatile("bar.sync 1, %0;" :: "n"(4 * kCsParts * 32) : "memory");
            }
            cur_g = qgroup;
        }
        const int bq = qgroup * 4 + quad;
        const int nq = min(qlens[bq], PLAID_NQ_MAX);
        const bool tok_valid = lane < nq;
#if !PLAID_CS_BULK
        ST* Sq = S + (size_t)bq * C * PLAID_NQ_MAX + lane;
#endif
        uint32_t* bits_q = idx_bits + (size_t)bq * (C >> 5);
        float bv[NC];
        int bi[NC];
#pragma unroll
        for (int p = 0; p < NC; p++) { bv[p] = -INFINITY; bi[p] = -1; }
        float cut = -INFINITY;  // current ncells-th best value of this thread's own list
#if PLAID_CS_BULK
        ST* Sblk = S + (size_t)bq * C * PLAID_NQ_MAX;
#endif
        for (int t = 0; ok && t < ntiles; t++, it++) {
            const int acc = it & 1;
            if (!mbar_wait(&bar->tmem_full[acc], (it >> 1) & 1, watchdog)) { ok = false; break; }
            tc_fence_after();
            const int c_tile = (tile_begin + t) * kCsN + part * kCsPartCols;
#if PLAID_CS_BULK
            int staged = 0;                  // blocks staged and not yet handed to the copy engine
#endif
            uint32_t r[32];
#pragma unroll 1
            for (int ch = 0; ch < kCsPartCols / 32; ch++) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * kCsN + part * kCsPartCols + ch * 32;
                if (!(PLAID_CS_BULK && PLAID_CS_PREFETCH && sizeof(ST) == 2) || ch == 0) tmem_ld_32x32(taddr, r);
                tc_wait_ld32(r);
                const int c0 = c_tile + ch * 32;
                if (c0 >= C) break;  // C is a multiple of 32: a chunk is entirely inside or outside
                // (1) the S rows: for each centroid the warp stores 32 consecutive values (one 128 B / 64 B line).
                //     With fp16 storage everything downstream (mask, cells) is computed from the ROUNDED values,
                //     so that the table in memory is the one and only definition of S.
#if PLAID_CS_BULK
                const uint32_t blk_sa = stage_sa + staged * kBlkBytes;
                auto stage_ready = [&]() {
                    if (store && staged == 0) {
                        // the previous copy out of this warp's staging area must have finished reading it
                        if (lane == 0) bulk_wait_read<0>();
                        __syncwarp();
                    }
                };
#else
                ST* dst = Sq + (size_t)c0 * PLAID_NQ_MAX;
                const bool store = S != nullptr;     // S == NULL: only the top-ncells lists are wanted (index build: argmax)
#endif
                // (2) this thread's best value in the chunk decides whether the rare paths run at all; with fp16 storage
                //     it is a packed maximum of the rounded pairs (no unpacking on the common path)
                __half2 h[sizeof(ST) == 2 ? 16 : 1];
                float mx;
                if constexpr (sizeof(ST) == 2) {
#if PLAID_CS_BULK
#pragma unroll
                    for (int j = 0; j < 16; j++) h[j] = __floats2half2_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
#if PLAID_CS_PREFETCH   // measured slower (1.09 vs 0.99 ms on cfg2): the early load holds 32 more registers through the tail
                    if (ch + 1 < kCsPartCols / 32 && c0 + 32 < C) tmem_ld_32x32(taddr + 32, r);   // the next chunk, behind this one's tail
#endif
                    stage_ready();
#endif
#pragma unroll
                    for (int j = 0; j < 16; j++) {
#if !PLAID_CS_BULK
                        h[j] = __floats2half2_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
#endif
#if PLAID_CS_BULK
                        if (store) {
                            sts_b16(blk_sa + (2 * j) * (PLAID_NQ_MAX * 2), __half_as_ushort(__low2half(h[j])));
                            sts_b16(blk_sa + (2 * j + 1) * (PLAID_NQ_MAX * 2), __half_as_ushort(__high2half(h[j])));
                        }
#else
                        if (store) {
                            __stcs(reinterpret_cast<unsigned short*>(dst + (2 * j) * PLAID_NQ_MAX), __half_as_ushort(__low2half(h[j])));
                            __stcs(reinterpret_cast<unsigned short*>(dst + (2 * j + 1) * PLAID_NQ_MAX), __half_as_ushort(__high2half(h[j])));
                        }
#endif
                    }
                    __half2 m2 = h[0];
#pragma unroll
                    for (int j = 1; j < 16; j++) m2 = __hmax2(m2, h[j]);
                    mx = fmaxf(__low2float(m2), __high2float(m2));
                } else {
#if PLAID_CS_BULK
                    stage_ready();
#endif
#pragma unroll
                    for (int j = 0; j < 32; j++) {
#if PLAID_CS_BULK
                        if (store) sts_b32(blk_sa + j * (PLAID_NQ_MAX * 4), r[j]);
#else
                        if (store) __stcs(dst + j * PLAID_NQ_MAX, __uint_as_float(r[j]));
#endif
                    }
                    mx = __uint_as_float(r[0]);
#pragma unroll
                    for (int j = 1; j < 32; j++) mx = fmaxf(mx, __uint_as_float(r[j]));
                }
#if PLAID_CS_BULK
                if (store) {
                    staged++;
                    // hand the staged run to the copy engine once the area is full or the part / the table ends
                    if (staged == kBlkPerCopy || ch + 1 == kCsPartCols / 32 || c0 + 32 >= C) {
                        fence_proxy_async_smem();    // staged blocks -> visible to the copy engine (async proxy)
                        __syncwarp();
                        if (lane == 0)
                            bulk_store(Sblk + (size_t)(c0 - (staged - 1) * 32) * PLAID_NQ_MAX, stage_base, staged * kBlkBytes);
                        staged = 0;
                    }
                }
#endif
                auto val = [&](int j) -> float {     // the stored (rounded) value of column j
                    if constexpr (sizeof(ST) == 2) return (j & 1) ? __high2float(h[j >> 1]) : __low2float(h[j >> 1]);
                    else return __uint_as_float(r[j]);
                };
                // pruning mask: max_k S[c,k] >= thr  <=>  any valid token has S[c,k] >= thr
                uint32_t word = 0;
                if (__any_sync(0xffffffffu, tok_valid && mx >= threshold)) {
                    uint32_t hit = 0;                // this lane's columns at or above the threshold
                    if (tok_valid) {
                        if constexpr (sizeof(ST) == 2) {
                            // a stored value is a half: v >= thr  <=>  v >= the smallest half that is >= thr
#pragma unroll
                            for (int j = 0; j < 16; j++) hit |= __hge2_mask(h[j], thr2) & (0x00010001u << j);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++) hit |= (val(j) >= threshold ? 1u : 0u) << j;
                        }
                    }
                    word = __reduce_or_sync(0xffffffffu, hit);
                    if constexpr (sizeof(ST) == 2) word = interleave16(word);
                }
                if (lane == 0 && idx_bits != nullptr) bits_q[c0 >> 5] = word;
                // (3) running top-ncells of this query token (score desc, centroid id asc).  A value can only matter if
                //     it beats this list's own ncells-th best AND is not below the ncells-th best the sibling warp (other
                //     128-column half, same query token) has already seen -- that bound is shared through smem (racy
                //     reads/writes only ever make the filter weaker, never wrong).
                //     Few chunks hold such a value and then usually one: the lanes that want an update mark their
                //     candidate columns, and the warp walks the UNION of those columns in ascending order, re-reading one
                //     accumulator column per step (a lane whose own test fails on a column just sits the step out).
                const float other = s_cut[quad][lane];
                const bool want = tok_valid && mx > cut && mx >= other;
                if (__any_sync(0xffffffffu, want)) {
                    uint32_t cols = 0;               // bit j: column j of the chunk may enter this lane's list
                    if (want) {
                        if constexpr (sizeof(ST) == 2) {
                            // packed compares; bit p <-> column 2p, bit 16+p <-> column 2p+1 (un-interleaved below)
                            const __half2 cut2 = __float2half2_rn(cut), oth2 = __float2half2_rn(other);
#pragma unroll
                            for (int j = 0; j < 16; j++)
                                cols |= __hgt2_mask(h[j], cut2) & __hge2_mask(h[j], oth2) & (0x00010001u << j);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                cols |= (val(j) > cut && val(j) >= other ? 1u : 0u) << j;
                        }
                    }
                    uint32_t all = __reduce_or_sync(0xffffffffu, cols);
                    if constexpr (sizeof(ST) == 2) all = interleave16(all);
                    while (all) {
                        const int j = __ffs(all) - 1;
                        all &= all - 1;
                        float cv = __uint_as_float(tmem_ld_32x1(taddr + j));      // this lane's value of column j again
                        if constexpr (sizeof(ST) == 2) cv = __half2float(__float2half_rn(cv));
                        if (want && cv > cut && cv >= other) {
                            int ci = c0 + j;
#pragma unroll
                            for (int p = 0; p < NC; p++) {
                                if (p < ncells) {
                                    const bool ahead = cv > bv[p] || (cv == bv[p] && (unsigned)ci < (unsigned)bi[p]);
                                    if (ahead) {
                                        const float tv = bv[p]; bv[p] = cv; cv = tv;
                                        const int ti = bi[p]; bi[p] = ci; ci = ti;
                                    }
                                    if (p == ncells - 1) cut = bv[p];
                                }
                            }
                        }
                    }
                    if (want && cut > other) s_cut[quad][lane] = cut;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar->tmem_empty[acc]);
        }
        // every (centroid range, column part) keeps its own partial list: slot = split*kCsParts + part
        const size_t base = (((size_t)bq * PLAID_NQ_MAX + lane) * (csplit * kCsParts) + (split * kCsParts + part)) * ncells;
#pragma unroll
        for (int p = 0; p < NC; p++)
            if (p < ncells) {
                cell_val[base + p] = bv[p];
                cell_idx[base + p] = tok_valid ? bi[p] : -1;
            }
        }
#if PLAID_CS_BULK
        if (store && lane == 0) bulk_wait_all();     // the staging area must outlive the copies that read it
#endif
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace plaid

extern "C" int plaid_centroid_scores(const void* centroids_bf16, int C, const void* Qb_bf16, const int32_t* qlens,
                                     int B_pad, int Lq_pad, float threshold, int ncells, int csplit, void* S, int s_is_f16,
                                     uint32_t* idx_bits, float* cell_val, int32_t* cell_idx, int* watchdog,
                                     void* stream) {
    using namespace plaid;
    PLAID_CHECK_ARG(centroids_bf16 && Qb_bf16 && qlens && cell_val && cell_idx, PLAID_ERR_ARG,
                    "plaid_centroid_scores: null pointer");      // S and idx_bits may be NULL (lists only)
    PLAID_CHECK_ARG(C >= 32 && (C % 32) == 0, PLAID_ERR_UNSUPPORTED, "plaid_centroid_scores: C=%d must be a multiple of 32", C);
    PLAID_CHECK_ARG(B_pad >= 0 && (B_pad % 4) == 0 && Lq_pad >= 32 && (Lq_pad % 32) == 0, PLAID_ERR_ARG,
                    "plaid_centroid_scores: B_pad=%d must be a multiple of 4, Lq_pad=%d a multiple of 32", B_pad, Lq_pad);
    PLAID_CHECK_ARG(ncells >= 1 && ncells <= PLAID_NCELLS_MAX && csplit >= 1 && csplit <= 64, PLAID_ERR_ARG,
                    "plaid_centroid_scores: ncells=%d (1..%d), csplit=%d (1..64)", ncells, PLAID_NCELLS_MAX, csplit);
    if (B_pad == 0) return PLAID_OK;
    CUtensorMap map_q, map_c;
    int rc;
    if ((rc = make_bf16_2d_map(&map_q, Qb_bf16, (uint64_t)B_pad * Lq_pad, kDim, 32)) != PLAID_OK) return rc;
    if ((rc = make_bf16_2d_map(&map_c, centroids_bf16, (uint64_t)C, kDim, kCsN)) != PLAID_OK) return rc;
    const int num_units = (B_pad / 4) * csplit;          // (query group, centroid range), group-major
    dim3 grid(min(num_units, sm_count()));
    cudaStream_t st = (cudaStream_t)stream;
#define PLAID_CS_LAUNCH(ST_, NC_)                                                                                       \
    do {                                                                                                                \
        static int configured[kMaxDevices] = {0};                                                                       \
        if ((rc = ensure_dynamic_smem((const void*)centroid_scores_kernel<ST_, NC_>, kCsSmemBytes, configured)) != PLAID_OK) \
            return rc;                                                                                                  \
        centroid_scores_kernel<ST_, NC_><<<grid, kCsThreads, kCsSmemBytes, st>>>(                                       \
            map_q, map_c, qlens, C, Lq_pad, threshold, ncells, csplit, num_units, reinterpret_cast<ST_*>(S), idx_bits,  \
            cell_val,                                                                                                   \
            cell_idx, watchdog);                                                                                        \
    } while (0)
    const int nc = ncells <= 1 ? 1 : ncells <= 2 ? 2 : ncells <= 4 ? 4 : 8;
    if (s_is_f16) {
        if (nc == 1) PLAID_CS_LAUNCH(__half, 1); else if (nc == 2) PLAID_CS_LAUNCH(__half, 2);
        else if (nc == 4) PLAID_CS_LAUNCH(__half, 4); else PLAID_CS_LAUNCH(__half, 8);
    } else {
        if (nc == 1) PLAID_CS_LAUNCH(float, 1); else if (nc == 2) PLAID_CS_LAUNCH(float, 2);
        else if (nc == 4) PLAID_CS_LAUNCH(float, 4); else PLAID_CS_LAUNCH(float, 8);
    }
#undef PLAID_CS_LAUNCH
    PLAID_LAUNCH_OK("centroid_scores_kernel");
    return PLAID_OK;
}
