/*
 * oracle/plaid_oracle.c -- CPU restatement of the reference's PLAID search kernels.
 *
 * TEST INFRASTRUCTURE ONLY: may be called from tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs, never from the product path.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the
 * reference's own compiled operators (oracle/_ref, built from /root/reference by
 * oracle/build_ref.py) in tests/test_oracle_vs_reference.py, and against the
 * golden vectors in tests/golden/ that were produced by running the unmodified
 * reference Searcher in the authoring container (tests/golden/make_golden.py).
 * The reference ships no golden vectors of its own for this path.
 *
 * Plain scalar C, single thread, no fast-math: float operations happen exactly in
 * the order the reference performs them.
 *
 * Reference files restated (paths relative to third_party/ColBERT/colbert/):
 *   search/filter_pids.cpp:27-164          -> plaid_oracle_approx_scores, plaid_oracle_filter_pids
 *   search/decompress_residuals.cpp:27-155 -> plaid_oracle_decompress, plaid_oracle_unpack_codes
 *   modeling/segmented_maxsim.cpp:22-93    -> plaid_oracle_segmented_maxsim
 *   search/segmented_lookup.cpp:51-125     -> plaid_oracle_segmented_lookup
 *   indexing/codecs/residual.py:54-89      -> plaid_oracle_codec_tables
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- residual.py:54-73 (reversed_bit_map) and :77-89 (decompression_lookup_table) ---- */
void plaid_oracle_codec_tables(int nbits, uint8_t* reversed_bit_map /*[256]*/,
                               uint8_t* lookup /*[2^8][8/nbits]*/) {
    const int mask = (1 << nbits) - 1;
    const int keys = 8 / nbits;
    for (int i = 0; i < 256; i++) {
        int z = 0;
        for (int j = 8; j > 0; j -= nbits) {
            int x = (i >> (j - nbits)) & mask;
            int y = 0;
            for (int k = nbits - 1; k >= 0; k--) y += ((x >> (nbits - k - 1)) & 1) << k;
            z |= y;
            if (j > nbits) z <<= nbits;
        }
        reversed_bit_map[i] = (uint8_t)z;
    }
    /* itertools.product(range(2^nbits), repeat=keys): row r is the base-2^nbits digits of r,
       most significant digit first. */
    for (int r = 0; r < 256; r++)
        for (int l = 0; l < keys; l++)
            lookup[r * keys + l] = (uint8_t)((r >> (nbits * (keys - 1 - l))) & mask);
}

/* ---- decompress_residuals.cpp:49-72, integer part only: bucket index of every dimension ---- */
void plaid_oracle_unpack_codes(const uint8_t* residuals, int64_t ntokens, int dim, int nbits,
                               const uint8_t* reversed_bit_map, const uint8_t* lookup,
                               uint8_t* out /*[ntokens][dim]*/) {
    const int keys = 8 / nbits, packed_dim = dim / keys;
    for (int64_t t = 0; t < ntokens; t++)
        for (int k = 0; k < packed_dim; k++) {
            uint8_t x = reversed_bit_map[residuals[t * packed_dim + k]];
            for (int l = 0; l < keys; l++) out[t * dim + k * keys + l] = lookup[x * keys + l];
        }
}

/* ---- decompress_residuals.cpp:27-155 ---- */
int64_t plaid_oracle_decompress(const int32_t* pids, int npids, const int64_t* lengths,
                                const int64_t* offsets, const float* bucket_weights,
                                const uint8_t* reversed_bit_map, const uint8_t* lookup,
                                const uint8_t* residuals, const int32_t* codes,
                                const float* centroids, int dim, int nbits, float* out) {
    const int keys = 8 / nbits, packed_dim = dim / keys;
    int64_t row = 0;
    for (int i = 0; i < npids; i++) {
        const int pid = pids[i];
        const int64_t off = offsets[pid];
        for (int64_t j = 0; j < lengths[pid]; j++, row++) {
            const int code = codes[off + j];
            for (int k = 0; k < packed_dim; k++) {
                uint8_t x = reversed_bit_map[residuals[(off + j) * packed_dim + k]];
                for (int l = 0; l < keys; l++) {
                    const int d = k * keys + l;
                    out[row * dim + d] =
                        bucket_weights[lookup[x * keys + l]] + centroids[(int64_t)code * dim + d];
                }
            }
        }
    }
    return row;
}

/* ---- filter_pids.cpp:27-72: per-document approximate score ----
 * idx == NULL means "all centroids kept" (the `ones` array of filter_pids.cpp:148-153). */
void plaid_oracle_approx_scores(const int32_t* pids, int npids, const float* S, int nq,
                                const int32_t* codes, const int64_t* doclens,
                                const int64_t* offsets, const uint8_t* idx, float* out) {
    float* per = (float*)malloc(sizeof(float) * (size_t)(nq > 0 ? nq : 1));
    for (int i = 0; i < npids; i++) {
        const int pid = pids[i];
        for (int k = 0; k < nq; k++) per[k] = -9999.0f;
        for (int64_t j = 0; j < doclens[pid]; j++) {
            const int code = codes[offsets[pid] + j];
            if (idx && !idx[code]) continue;
            /* the reference skips repeated codes via seen_codes; max() is idempotent */
            const float* row = S + (int64_t)code * nq;
            for (int k = 0; k < nq; k++)
                if (row[k] > per[k]) per[k] = row[k];
        }
        float score = 0.0f;
        for (int k = 0; k < nq; k++) score += per[k]; /* sequential fp32 sum, filter_pids.cpp:59-63 */
        out[i] = score;
    }
    free(per);
}

typedef struct { float s; int32_t p; } sp_t;
static int sp_desc(const void* a, const void* b) {
    const sp_t *x = (const sp_t*)a, *y = (const sp_t*)b;
    if (x->s != y->s) return x->s > y->s ? -1 : 1; /* std::pair<float,int> order, descending */
    if (x->p != y->p) return x->p > y->p ? -1 : 1;
    return 0;
}

/* top-`keep` of (score,pid) pairs, descending (filter_pids.cpp:108-123).  When fewer than
 * `keep` pairs exist the reference pops an empty heap (UB); we keep min(n, keep) as the
 * reference's GPU branch does (index_storage.py:138-139).  Returns the count kept. */
int plaid_oracle_select_top(const int32_t* pids, const float* scores, int n, int keep,
                            int32_t* out_pids, float* out_scores) {
    sp_t* v = (sp_t*)malloc(sizeof(sp_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { v[i].s = scores[i]; v[i].p = pids[i]; }
    qsort(v, (size_t)n, sizeof(sp_t), sp_desc);
    const int m = n < keep ? n : keep;
    for (int i = 0; i < m; i++) { out_pids[i] = v[i].p; if (out_scores) out_scores[i] = v[i].s; }
    free(v);
    return m;
}

/* ---- filter_pids.cpp:126-164: two-stage filter.  Outputs both stages for stage-wise parity.
 * stage1: at most ndocs pids, stage2: at most ndocs/4 pids; both ordered (score,pid) desc. */
void plaid_oracle_filter_pids(const int32_t* pids, int npids, const float* S, int nq,
                              const int32_t* codes, const int64_t* doclens,
                              const int64_t* offsets, const uint8_t* idx, int ndocs,
                              int32_t* stage1_pids, float* stage1_scores, int* n1,
                              int32_t* stage2_pids, float* stage2_scores, int* n2) {
    float* sc = (float*)malloc(sizeof(float) * (size_t)(npids > 0 ? npids : 1));
    plaid_oracle_approx_scores(pids, npids, S, nq, codes, doclens, offsets, idx, sc);
    *n1 = plaid_oracle_select_top(pids, sc, npids, ndocs, stage1_pids, stage1_scores);
    free(sc);
    sc = (float*)malloc(sizeof(float) * (size_t)(*n1 > 0 ? *n1 : 1));
    plaid_oracle_approx_scores(stage1_pids, *n1, S, nq, codes, doclens, offsets, NULL, sc);
    *n2 = plaid_oracle_select_top(stage1_pids, sc, *n1, ndocs / 4, stage2_pids, stage2_scores);
    free(sc);
}

/* ---- segmented_maxsim.cpp:22-93: zero-initialised running max per (doc, query token),
 * then a sum over query tokens (torch `max_scores.sum(1)`; we sum left to right in fp32). */
void plaid_oracle_segmented_maxsim(const float* scores /*[T][nq]*/, const int64_t* lengths,
                                   int ndocs, int nq, float* out /*[ndocs]*/,
                                   float* out_max /*[ndocs][nq] or NULL*/) {
    float* m = (float*)malloc(sizeof(float) * (size_t)(nq > 0 ? nq : 1));
    int64_t row = 0;
    for (int i = 0; i < ndocs; i++) {
        for (int k = 0; k < nq; k++) m[k] = 0.0f;
        for (int64_t j = 0; j < lengths[i]; j++, row++)
            for (int k = 0; k < nq; k++) {
                const float v = scores[row * nq + k];
                if (v > m[k]) m[k] = v;
            }
        float s = 0.0f;
        for (int k = 0; k < nq; k++) s += m[k];
        out[i] = s;
        if (out_max) memcpy(out_max + (int64_t)i * nq, m, sizeof(float) * (size_t)nq);
    }
    free(m);
}

/* ---- segmented_lookup.cpp:36-48: ragged row gather ---- */
int64_t plaid_oracle_segmented_lookup(const uint8_t* input, int64_t row_bytes,
                                      const int64_t* pids, int npids, const int64_t* lengths,
                                      const int64_t* offsets, uint8_t* out) {
    int64_t row = 0;
    for (int i = 0; i < npids; i++) {
        memcpy(out + row * row_bytes, input + offsets[i] * row_bytes,
               (size_t)(lengths[i] * row_bytes));
        row += lengths[i];
    }
    (void)pids;
    return row;
}
