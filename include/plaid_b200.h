/*
 * plaid_b200.h -- C ABI of libplaid_b200.so: the B200 (sm_100a) kernels behind FLMR's
 * ColBERT/PLAID late-interaction search path.
 *
 * This is the drop-in boundary.  The reference binds its native operators through
 * pybind11/torch (`torch.utils.cpp_extension.load`); a maintainer replaces those loads with a
 * ctypes binding of the functions below (INTEGRATION.md shows the stubs).  Conventions:
 *
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - tensors are dense, row-major, no strides honoured (exactly like the reference operators,
 *     which read `data_ptr<T>()` directly);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only enqueue
 *     work, they never synchronise;
 *   - return value: 0 on success, a negative PLAID_ERR_* otherwise (never aborts, never exits,
 *     unlike filter_pids.cpp:98-101 / the active asserts at filter_pids.cpp:47);
 *     plaid_last_error() returns the message of the calling thread's last failure;
 *   - nothing is allocated behind the caller's back: outputs and workspaces are caller-owned.
 *
 * Paths below are relative to /root/reference/third_party/ColBERT/colbert/ ("CB/").
 */
#ifndef PLAID_B200_H
#define PLAID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLAID_OK 0
#define PLAID_ERR_ARG (-1)         /* bad argument (null pointer, unsupported size) */
#define PLAID_ERR_CUDA (-2)        /* CUDA runtime / driver error (launch, tensor-map encode) */
#define PLAID_ERR_UNSUPPORTED (-3) /* shape outside what the kernels implement */

#define PLAID_DIM 128          /* embedding dim of the path (CB/infra/config/settings.py:101) */
#define PLAID_NQ_MAX 32        /* candidate-stage query tokens = query_maxlen (settings.py:108) */
#define PLAID_NCELLS_MAX 8     /* ncells is 1/2/4 in CB/searcher.py:96-122 */
#define PLAID_NO_PID (-1)      /* filler for unused (pid) slots */

/* ---- library ---------------------------------------------------------------------------- */
int plaid_abi_version(void);
const char* plaid_last_error(void);
/* compiled-for architecture string, e.g. "sm_100a" */
const char* plaid_arch(void);

/* ---- a1: query preparation (CB/searcher.py:124-130, CB/search/index_storage.py:77) -------
 * Q fp32 [B, Lq, 128] -> per query: rows with sum(|q|) > 0 kept in order (remove_zero_rows != 0),
 * converted to bf16 into Qb [B_pad, Lq_pad, 128] (rows past the kept ones and queries >= B are zero),
 * qlens[b] = rows kept (0 for the padding queries b >= B; qlens has B_pad entries).  B_pad must be a
 * multiple of 4, Lq_pad a multiple of 32 and >= Lq.  Qh_f16 (optional, may be NULL) receives the same
 * rows rounded to fp16 -- the operand of the search pipeline's MaxSim, which like the reference's GPU
 * branch (`Q.cuda().half()`, CB/search/candidate_generation.py:52; half D, CB/indexing/codecs/residual.py:273) runs in fp16. */
int plaid_prepare_queries(const float* Q, int B, int Lq, int remove_zero_rows, int B_pad, int Lq_pad,
                          void* Qb_bf16, void* Qh_f16, int32_t* qlens, void* stream);

/* fp32 -> bf16 row conversion used when loading the codebook (round-to-nearest-even). */
int plaid_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* ---- a2 + a4: centroid scoring (CB/search/candidate_generation.py:12-20, index_storage.py:115)
 * S[b, c, k] = <centroids[c], Qb[b, k]> for k < nq_max=32 (bf16 operands, fp32 accumulate on
 * tcgen05; operands fed by TMA).  Fused epilogue products:
 *   idx_bits[b, c/32] bit (c%32)  = max_{k < nq_b} S[b,c,k] >= threshold      (centroid pruning mask)
 *   cell_val/cell_idx[b, k, l, j] = the ncells best (score, centroid) of query token k inside the
 *                                   l-th of PLAID_CELL_LISTS_PER_RANGE*csplit partial lists (csplit centroid
 *                                   ranges x the column parts of a tile), best first, ties -> lowest centroid id;
 *                                   cell_idx = -1 where there is no entry.
 * nq_b = min(qlens[b], 32).  S is laid out [B_pad, C, 32] -- per query exactly the reference's
 * `centroid_scores` [C, nq] tensor (row = centroid) -- in fp32, or with s_is_f16 != 0 rounded to fp16
 * (the precision the reference's GPU branch computes S in); mask and cells are derived from the stored values.  C must be a multiple of 32; the grid is
 * (B_pad/4) x csplit CTAs.  Rows of Qb beyond qlens[b] must be zero (plaid_prepare_queries does it).
 * *watchdog (device int, may be NULL) is set to 1 if an in-kernel pipeline wait ever times out. */
#define PLAID_CELL_LISTS_PER_RANGE 4
int plaid_centroid_scores(const void* centroids_bf16, int C, const void* Qb_bf16, const int32_t* qlens,
                          int B_pad, int Lq_pad, float threshold, int ncells, int csplit,
                          void* S, int s_is_f16, uint32_t* idx_bits, float* cell_val, int32_t* cell_idx, int* watchdog,
                          void* stream);

/* ---- a3: candidate pids (candidate_generation.py:31-37,57-60; strided_tensor.py:77-99;
 *          segmented_lookup.cpp:51-125) --------------------------------------------------------
 * Merges the `nlists` partial lists into cells[b, k, 0..ncells) (i32 [B, 32, ncells], -1 = none),
 * takes the union of their IVF pid lists and emits it sorted + unique through a per-query pid
 * bitmap: bitmap_ws [B, ceil(N/32)] u32 (zeroed by the call), cand_pids [B, cand_stride] i32
 * ascending, cand_counts[b].  If a query would need more than cand_stride slots the list is
 * truncated and *overflow is set to 1.  wprefix (optional, may be NULL): i32 [B, ceil(N/32)], number of
 * candidates before each bitmap word, so that rank(pid) = wprefix[pid/32] + popc(word & ((1 << pid%32) - 1))
 * (used by plaid_filter_stage1_ivf). */
int plaid_candidates(const float* cell_val, const int32_t* cell_idx, const int32_t* qlens, int B, int ncells,
                     int nlists, const int32_t* ivf_pids, const int64_t* ivf_offsets, int C, int N,
                     int32_t* cells, uint32_t* bitmap_ws, int32_t* cand_pids, int32_t* cand_counts,
                     int cand_stride, int* overflow, int32_t* wprefix, void* stream);

/* ---- a5: filter_pids (CB/search/filter_pids.cpp:27-164) ---------------------------------------
 * Approximate score of every listed passage: sum over k < nq_b (sequential fp32, in k order) of
 * max over the passage's codes with idx bit set (all codes when idx_bits == NULL) of S[b, code, k],
 * each per-token max initialised to -9999.  pids [B, pid_stride] with counts[b] valid entries;
 * scores written to the same slots of out_scores.  Lane = query token.  codes must lie in [0, C) (unchecked,
 * as in the reference, which asserts it); C a multiple of 128, codes / idx_bits 16-byte aligned. */
int plaid_approx_scores(const int32_t* pids, const int32_t* counts, int B, int pid_stride, const void* S,
                        int s_is_f16, const int32_t* qlens, const uint32_t* idx_bits, int C, const int32_t* codes,
                        const int64_t* offsets, float* out_scores, void* stream);

/* Stage 1 of the filter driven by the inverted file (same result as plaid_approx_scores with idx_bits, bit for
 * bit): for every surviving centroid (idx bit set) its IVF list names the passages that contain it; the candidate
 * bitmap / word-prefix counts (plaid_candidates) turn those into candidate slots; the (slot, centroid) pairs are
 * counting-sorted per query in shared memory and reduced (candidate lists longer than the ~33 k slots the bins
 * hold -- shards of millions of passages -- in slot ranges, one CTA per query and range).  Queries whose mask is
 * dense (more than cap_s survivors, more than 4*cap_p list entries to visit, more pairs than cap_p) are routed
 * through the token scan instead -- decided per query on the device, no host round trip.
 * Workspaces: ws_surv i32 [B, cap_s], ws_pair_slot / ws_pair_c i32 [B, cap_p], ws_sorted_c i32 [B, 2*cap_p] (8-byte aligned),
 * ws_meta i32 [B, 4]. */
int plaid_filter_stage1_ivf(const int32_t* pids, const int32_t* counts, int B, int pid_stride, const void* S,
                            int s_is_f16, const int32_t* qlens, const uint32_t* idx_bits, int C,
                            const int32_t* codes, const int64_t* offsets, const int32_t* ivf_pids,
                            const int64_t* ivf_offsets, const uint32_t* bitmap, const int32_t* wprefix, int N,
                            int32_t* ws_surv, int cap_s, int32_t* ws_pair_slot, int32_t* ws_pair_c,
                            int32_t* ws_sorted_c, int cap_p, int32_t* ws_meta, float* out_scores, void* stream);

/* Test hook: slots per range of plaid_filter_stage1_ivf's sort (0 = as many as shared memory holds, the default), so
 * that the multi-range path can be exercised on small indexes.  Process-wide; returns the previous value. */
int plaid_set_ivf_range_slots(int slots);

/* Per query the `keep` largest (score, pid) pairs in descending (score, pid) order -- the order of
 * std::pair<float,int> in filter_pids.cpp:24,108-123.  Emits min(counts[b], keep) entries
 * (the rule of the reference's GPU branch, index_storage.py:138-139; the C++ pops an empty heap).
 * Unused output slots get PLAID_NO_PID / -inf.  ws_keys: u64 workspace [B, in_stride]; kept in the ABI and must be
 * non-null, but no longer touched (keys are built on the fly from scores and pids). */
int plaid_select_top(const int32_t* pids, const float* scores, const int32_t* counts, int B, int in_stride,
                     int keep, int32_t* out_pids, float* out_scores, int32_t* out_counts, int out_stride,
                     uint64_t* ws_keys, void* stream);

/* The two-stage filter exactly as IndexScorer.filter_pids is called (index_storage.py:153-156):
 * stage 1 with the pruning mask keeps ndocs, stage 2 with all codes keeps ndocs/4.
 * Workspaces: ws_scores f32 [B, max(pid_stride, ndocs)], ws_keys u64 [B, max(pid_stride, ndocs)],
 * stage1_pids i32 [B, ndocs] / stage1_scores f32 [B, ndocs] / stage1_counts [B] (outputs too, for
 * stage-wise parity), stage2_* likewise with ndocs/4. */
int plaid_filter_pids(const int32_t* pids, const int32_t* counts, int B, int pid_stride, const void* S,
                      int s_is_f16, const int32_t* qlens, const uint32_t* idx_bits, int C, const int32_t* codes,
                      const int64_t* offsets, int ndocs, float* ws_scores, uint64_t* ws_keys,
                      int32_t* stage1_pids, float* stage1_scores, int32_t* stage1_counts,
                      int32_t* stage2_pids, float* stage2_scores, int32_t* stage2_counts, void* stream);

/* ---- a6: decompress_residuals (CB/search/decompress_residuals.cpp:27-155;
 *          tables: CB/indexing/codecs/residual.py:54-89) ---------------------------------------
 * Builds the fused weight table W[x][l] = bucket_weights[lookup[reversed_bit_map[x]][l]],
 * x in 0..255, l in 0..8/nbits, from the reference's three codec tensors. */
int plaid_build_weight_table(const float* bucket_weights, const uint8_t* reversed_bit_map,
                             const uint8_t* lookup, int nbits, float* W, void* stream);

/* out[row, d] = W[residual_byte][l] + centroids[code][d]  (one fp32 add, bit-exact), rows packed in
 * pid order: out_offsets[i] = first output row of pids[i] (exclusive prefix sum of the lengths,
 * npids+1 entries, device).  centroids fp32 [C, 128]. */
int plaid_decompress_residuals(const int32_t* pids, int npids, const int64_t* offsets, const int64_t* out_offsets,
                               const float* W, const uint8_t* residuals, const int32_t* codes,
                               const float* centroids, int C, int nbits, float* out, void* stream);

/* Integer part only: bucket index of every dimension, u8 [ntokens, 128] (parity tap). */
int plaid_unpack_residual_codes(const uint8_t* residuals, int64_t ntokens, int nbits,
                                const uint8_t* reversed_bit_map, const uint8_t* lookup, uint8_t* out,
                                void* stream);

/* The operator the reference binds as ResidualCodec.decompress_residuals on its GPU branch
 * (CB/indexing/codecs/residual.py:115,242-263; codecs/decompress_residuals.cu:8-75): token rows, no pid indirection.
 * out[t, d] = half(W[residuals[t, byte]][l]) + centroids_f16[codes[t], d] -- one half add, bit-exact with the
 * reference kernel -- as fp16 [n, 128].  normalize != 0 applies ResidualCodec.decompress's
 * `F.normalize(x, p=2, dim=-1).half()` on top (residual.py:272-273).  W is plaid_build_weight_table's table. */
int plaid_decompress_tokens_f16(const uint8_t* residuals, const int32_t* codes, int64_t n, const float* W,
                                const void* centroids_f16, int C, int nbits, int normalize, void* out_f16,
                                void* stream);

/* Exclusive per-query prefix sums of the passage lengths of pids [B, pid_stride] (counts[b] valid), each
 * length rounded up to a multiple of `align` tokens (1 = packed back to back; 32 = the layout
 * plaid_maxsim_packed's aligned mode reads): tok_offsets [B, pid_stride+1] (i32, local to the query). */
int plaid_doc_token_offsets(const int32_t* pids, const int32_t* counts, int B, int pid_stride,
                            const int64_t* offsets, int align, int32_t* tok_offsets, void* stream);

/* a6 + a7 for the search pipeline: decompress, L2-normalise (eps 1e-12, index_storage.py:175) and
 * round to bf16 into D [B, tok_stride, 128]; token j of passage i of query b lands on row
 * b*tok_stride + tok_offsets[b, i] + j; rows up to tok_offsets[b, i+1] (alignment padding) are zeroed.
 * centroids [C,128] either fp32 or, with centroids_are_f16 != 0,
 * the fp16 values exactly as stored in centroids.pt (CB/indexing/codecs/residual.py:161; the CPU
 * reference widens them to fp32, residual.py:29 -- the same numbers).  The add is done in fp32
 * before normalisation, as the reference does. */
int plaid_decompress_normalize_bf16(const int32_t* pids, const int32_t* counts, int B, int pid_stride,
                                    const int32_t* tok_offsets, int tok_stride, const int64_t* offsets,
                                    const float* W, const uint8_t* residuals, const int32_t* codes,
                                    const void* centroids, int centroids_are_f16, int C, int nbits,
                                    void* D_bf16, void* stream);

/* Same placement, in the fp16 arithmetic of the reference's GPU branch: half centroid + half bucket
 * weight -> half (CB/indexing/codecs/decompress_residuals.cu:35-37), L2-normalised, fp16 rows
 * (residual.py:273).  This is bit for bit what plaid_maxsim_fused builds in shared memory; the pair
 * plaid_decompress_normalize_f16 + plaid_maxsim_packed(operands_f16 = 1) is the unfused form of it. */
int plaid_decompress_normalize_f16(const int32_t* pids, const int32_t* counts, int B, int pid_stride,
                                   const int32_t* tok_offsets, int tok_stride, const int64_t* offsets,
                                   const float* W, const uint8_t* residuals, const int32_t* codes,
                                   const void* centroids_f16, int C, int nbits, void* D_f16, void* stream);

/* ---- f3 (next row): index-build side of the codec (CB/indexing/codecs/residual.py:169-222) -------------------
 * plaid_merge_cells: the partial top-ncells lists of plaid_centroid_scores -> cells[b, k, 0..ncells) (the first step
 * of plaid_candidates on its own).  With ncells = 1 and S = idx_bits = NULL in plaid_centroid_scores this is
 * ResidualCodec.compress_into_codes (`(centroids @ batch.T).max(dim=0).indices`, residual.py:204-222): 32 embeddings
 * play the role of one query's tokens; ties go to the lowest centroid id, as torch's max does. */
int plaid_merge_cells(const float* cell_val, const int32_t* cell_idx, const int32_t* qlens, int B, int ncells,
                      int nlists, int32_t* cells, void* stream);

/* ResidualCodec.compress minus the argmax (residual.py:176-203): residual = embs[t] - centroids[codes[t]] in fp32,
 * bucket = number of cutoffs strictly below it (torch.bucketize), nbits bits per dimension LSB first, packed MSB
 * first into residuals u8 [n, 16*nbits] -- the byte layout the search path decodes.  *bad_code_flag is set to 1 if
 * a code is outside [0, C). */
int plaid_compress_residuals(const float* embs, const int32_t* codes, const void* centroids_f16,
                             const float* bucket_cutoffs, int64_t n, int C, int nbits, uint8_t* residuals,
                             int* bad_code_flag, void* stream);

/* ResidualCodec.packbits (residual.py:130, codecs/packbits.cu:10-57): nflags u8 flags (non-zero = 1; nflags a
 * multiple of 8, 8-byte aligned) -> nflags/8 bytes, first flag in the most significant bit (np.packbits order). */
int plaid_packbits(const uint8_t* bits, int64_t nflags, uint8_t* packed, void* stream);

/* ---- a8: colbert_score_packed + segmented_maxsim (CB/modeling/colbert.py:289-311,
 *          CB/modeling/segmented_maxsim.cpp:22-93) ---------------------------------------------
 * scores[b, i] = sum_{k < qlens[b]} max(0, max_{t in passage i} <D[b, t], Qb[b, k]>): a tcgen05
 * Qb . D^T tile per 256 passage tokens with the per-passage running max and the sum over query
 * tokens done in the epilogue straight out of TMEM (the similarity matrix never reaches HBM).
 * clamp_zero = 1 reproduces the zero-initialised max buffer of segmented_maxsim.cpp:58-59.
 * aligned32 = 1 promises that every passage starts on a 32-token boundary of D and that its pad rows are
 * zero (what plaid_doc_token_offsets(align=32) + plaid_decompress_normalize_* produce): with the clamp
 * a zero row cannot change a maximum, so the epilogue never has to split a 32-column chunk.
 * operands_f16 = 1: Q and D hold fp16 instead of bf16 values (same layout, same tensor-core rate). */
int plaid_maxsim_packed(const void* Q16, const int32_t* qlens, int B, int B_pad, int Lq_pad,
                        const void* D16, const int32_t* tok_offsets, const int32_t* counts, int pid_stride,
                        int tok_stride, int clamp_zero, int aligned32, int operands_f16, float* scores,
                        int* watchdog, void* stream);

/* a6 + a7 + a8 fused, the form the search pipeline runs: the passages listed in pids [B, pid_stride]
 * (counts[b] valid) are decompressed and normalised in fp16 (see plaid_decompress_normalize_f16) and
 * contracted with the fp16 query Qh [B_pad, Lq_pad, 128] on tcgen05 without the passage embeddings ever
 * existing in HBM (decompressor warps write the B operand straight into the swizzled shared-memory
 * layout of the UMMA descriptor).  tok_offsets is what plaid_doc_token_offsets(align = 32) produced for
 * the same pids.  centroids_f16 [C,128] fp16 as stored in centroids.pt; W the fp32 table of
 * plaid_build_weight_table (rounded to fp16 on load, as `bucket_weights.half()` does, residual.py:40).
 * scores[b, i] as in plaid_maxsim_packed(clamp_zero = 1).  inv_norms_f16: optional table of plaid_token_inv_norms. */
int plaid_maxsim_fused(const void* Qh_f16, const int32_t* qlens, int B, int B_pad, int Lq_pad,
                       const int32_t* pids, const int32_t* counts, int pid_stride, const int32_t* tok_offsets,
                       const int64_t* offsets, const float* W, const uint8_t* residuals, const int32_t* codes,
                       const void* centroids_f16, int C, int nbits, const void* inv_norms_f16, float* scores,
                       int* watchdog, void* stream);

/* Per-token scale factors for plaid_maxsim_fused: inv[t] = half(1 / max(||centroid[code_t] + weights_t||, 1e-12)),
 * computed in exactly the fp16 arithmetic and reduction order of plaid_decompress_normalize_f16, for all n tokens of
 * an index (fp16 [n], 2 bytes per token; derived data, built once when the index is loaded).  Passing the table as
 * inv_norms_f16 lets the fused kernel skip the per-token sum of squares / rsqrt and still build bit-identical
 * operand tiles; NULL keeps the in-kernel normalisation. */
int plaid_token_inv_norms(const uint8_t* residuals, const int32_t* codes, int64_t n, const float* W,
                          const void* centroids_f16, int C, int nbits, void* inv_f16, void* stream);

/* The operator the reference binds as ColBERT.segmented_maxsim (colbert.py:60): scores f32 [T, nq]
 * already computed, lengths i64 [ndocs] -> f32 [ndocs]; zero-initialised running max, then a
 * left-to-right fp32 sum over the nq columns. */
int plaid_segmented_maxsim(const float* scores, int nq, const int64_t* lengths, const int64_t* row_offsets,
                           int ndocs, float* out, void* stream);

/* ---- a9: colbert_score / colbert_score_reduce (CB/modeling/colbert.py:235-286;
 *          FLMR copy src/models/flmr/models/flmr/flmr_utils.py:22-48) ---------------------------
 * Padded MaxSim: D_padded bf16 [n, Ld, 128], D_mask u8 [n, Ld] (non-zero = real token), passage d is
 * scored against query d / docs_per_query of Qb [nQ_pad, Lq_pad, 128]; masked positions count as -9999,
 * max over Ld, sum over the query's qlens rows (no clamp).  scores_raw (optional, may be NULL) receives
 * the masked similarity matrix fp32 [n, Ld, Lq_out] that flmr_utils.colbert_score also returns. */
int plaid_colbert_score_padded(const void* Qb_bf16, const int32_t* qlens, int nQ, int nQ_pad, int Lq_pad,
                               const void* D_padded_bf16, const uint8_t* D_mask, int64_t n, int Ld,
                               int docs_per_query, float* scores, float* scores_raw, int Lq_out,
                               int* watchdog, void* stream);

/* Gradient of the padded MaxSim for the training-time scoring (FLMRModelForRetrieval.score /
 * compute_ib_loss_new, src/models/flmr/models/flmr/modeling_flmr.py:932-947,1089-1125, under autograd in the reference).
 * For every pair p = (query q, passage d) with upstream gradient grad[p] and every query token k < qlens[q]:
 * t* = first arg max_t of the masked similarity (masked positions count as -9999), dQ[q, k, :] += grad[p] * D[d, t*, :],
 * dD[d, t*, :] += grad[p] * Q[q, k, :].  Pairing: all_pairs != 0: p = q * n + d over all nQ x n pairs (the [B, B*n_docs]
 * in-batch score matrix); else passage d pairs with query d / docs_per_query and p = d (colbert_score).
 * Qb / D are the bf16-rounded operands of the forward; ws_argmax i32 [pairs, Lq_pad] (also an output: t* per pair and
 * query token); dQ f32 [nQ, Lq_pad, 128] and dD f32 [n, Ld, 128] are fully written (either may be NULL). */
int plaid_colbert_score_backward(const void* Qb_bf16, const int32_t* qlens, int nQ, int Lq_pad,
                                 const void* D_bf16, const uint8_t* D_mask, int64_t n, int Ld, int docs_per_query,
                                 int all_pairs, const float* grad, int32_t* ws_argmax, float* dQ, float* dD,
                                 void* stream);

/* colbert_score_reduce on an existing fp32 scores_padded [n, Ld, Lq] + mask (colbert.py:237-263,
 * 'colbert' interaction): -9999 fill, max over Ld, sum over Lq (left to right). */
int plaid_colbert_score_reduce(const float* scores_padded, const uint8_t* D_mask, int64_t n, int Ld, int Lq,
                               float* scores, void* stream);

/* ---- a10 + multi-GPU merge (index_storage.py:95-96, searcher.py:136) -----------
 * Merge G per-shard lists (as laid out by an all-gather: [G, B, k] scores / pids, [G, B] counts)
 * into the global top-k per query, (score, pid) descending.  ws_keys: u64 [B, G*k]. */
int plaid_merge_topk(const float* scores, const int32_t* pids, const int32_t* counts, int G, int B, int k,
                     int32_t* out_pids, float* out_scores, int32_t* out_counts, uint64_t* ws_keys,
                     void* stream);

/* The same merge reading the all-gather's receive buffer IN PLACE: every rank's message is one i32 block
 * [B*k pids | B*k score bit patterns | B counts] (what the final plaid_select_top of the rank wrote straight into
 * its send buffer), gathered = [G] such blocks.  pid_bases (i32 [G], may be NULL) = first global pid of every
 * rank's shard, added to the shard-local pids while the keys are built. */
int plaid_merge_topk_msg(const int32_t* gathered, int G, int B, int k, const int32_t* pid_bases,
                         int32_t* out_pids, float* out_scores, int32_t* out_counts, uint64_t* ws_keys,
                         void* stream);

/* Exact-global sharded search (oracle (B)): the reference truncates to ndocs and ndocs/4 over the WHOLE
 * collection (filter_pids.cpp:108-123,148-157), so after each filter stage the shards all-gather their list blocks
 * ([B*k local pids | B*k score bits | B counts] per rank, each list sorted by (score, pid) descending -- what
 * plaid_select_top wrote).  The collection's `keep` best keys take a prefix of every shard's list: out_counts[b] = length
 * of THIS shard's prefix for the first `rows` queries (my_rank = index of this shard's block, pid_bases[g] = first global
 * pid of shard g).  The shard's pid list itself stays as it is. */
int plaid_prefix_share(const int32_t* gathered, int G, int B, int rows, int k, int keep, const int32_t* pid_bases,
                       int my_rank, int32_t* out_counts, void* stream);

/* ---- a11: ragged gather (CB/search/segmented_lookup.cpp:36-125) ------------------------------
 * out rows = concatenation over i of input[offsets[i] .. offsets[i]+lengths[i]) (row_bytes each);
 * out_offsets = exclusive prefix sum of lengths (device, n+1). */
int plaid_segmented_lookup(const uint8_t* input, int64_t row_bytes, const int64_t* lengths,
                           const int64_t* offsets, const int64_t* out_offsets, int n, uint8_t* out,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLAID_B200_H */
