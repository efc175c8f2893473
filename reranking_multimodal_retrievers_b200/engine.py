"""Batched PLAID search over one GPU-resident index shard.

This is the hot path of SURVEY.md 8a as one stream of kernel launches per query chunk, with no
host synchronisation between stages (the reference loops over queries one at a time,
CB/searcher.py:80-93, and crosses PCIe several times per query on its GPU branch):

  prepare_queries -> centroid_scores (tcgen05) -> candidates -> filter_pids (2 stages)
  -> doc_token_offsets -> decompress+normalise -> maxsim_packed (tcgen05) -> select_top(k)

Semantics follow the reference's CPU branch (IndexScorer.rank, CB/search/index_storage.py:86-184):
the first `query_maxlen` (32) query tokens drive candidate generation, all tokens the final MaxSim;
ties are ordered (score desc, pid desc).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import torch

from . import _lib, ops
from .index import DeviceIndex
from .ops import _p, _stream

CELL_LISTS_PER_RANGE = 4      # PLAID_CELL_LISTS_PER_RANGE (include/plaid_b200.h)
NQ_MAX = ops.NQ_MAX


def search_defaults(k: int):
    """(ncells, centroid_score_threshold, ndocs) chosen by Searcher.dense_search (CB/searcher.py:96-122)."""
    if k <= 100:
        return 2, 0.45, 1024
    return 4, 0.4, max(k * 4, 4096)


@dataclass
class StageTaps:
    """Device tensors of every stage of the last chunk (parity tests read these)."""
    Qb: torch.Tensor
    qlens: torch.Tensor
    S: torch.Tensor
    idx_bits: torch.Tensor
    cells: torch.Tensor
    cand_pids: torch.Tensor
    cand_counts: torch.Tensor
    stage1_pids: torch.Tensor
    stage1_scores: torch.Tensor
    stage1_counts: torch.Tensor
    stage2_pids: torch.Tensor
    stage2_scores: torch.Tensor
    stage2_counts: torch.Tensor
    tok_offsets: torch.Tensor
    D: torch.Tensor
    tok_stride: int
    scores: torch.Tensor


class _HostFeed:
    """Double-buffered H2D of query chunks: chunk i+1 travels while chunk i is searched.  Only the query
    preparation kernel reads the staging buffer, so it is released right after the chunk is enqueued."""

    def __init__(self, Qh: torch.Tensor, Bc: int, dev, stream):
        self.Qh, self.Bc, self.dev, self.stream = Qh, Bc, dev, stream
        B, Lq, dim = Qh.shape
        self.n_chunks = (B + Bc - 1) // Bc
        self.bufs = [torch.empty(min(Bc, B), Lq, dim, device=dev, dtype=torch.float32) for _ in range(min(2, self.n_chunks))]
        self.ready = [torch.cuda.Event() for _ in self.bufs]
        self.free = [None for _ in self.bufs]
        self._issue(0)

    def _issue(self, ci: int):
        if ci >= self.n_chunks:
            return
        slot = ci % len(self.bufs)
        b0, b1 = ci * self.Bc, min(self.Qh.shape[0], (ci + 1) * self.Bc)
        if self.free[slot] is not None:
            self.stream.wait_event(self.free[slot])          # the chunk that used this buffer has been prepared
        else:
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))   # buffers allocated on the compute stream
        with torch.cuda.stream(self.stream):
            self.bufs[slot][: b1 - b0].copy_(self.Qh[b0:b1], non_blocking=True)
            self.ready[slot].record(self.stream)

    def chunk(self, ci: int) -> torch.Tensor:
        slot = ci % len(self.bufs)
        if ci + 1 < self.n_chunks and len(self.bufs) > 1:
            self._issue(ci + 1)
        torch.cuda.current_stream(self.dev).wait_event(self.ready[slot])
        b0, b1 = ci * self.Bc, min(self.Qh.shape[0], (ci + 1) * self.Bc)
        return self.bufs[slot][: b1 - b0]

    def release(self, ci: int):
        slot = ci % len(self.bufs)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.free[slot] = ev


def pick_csplit(groups: int, tiles: int, sms: int = 148, unit_overhead: float = 8.0) -> int:
    """Centroid ranges per group of 4 queries for plaid_centroid_scores.  Its persistent CTAs take contiguous, equally
    long runs of the (group, range) units, so the kernel's time is that of the longest run: ceil(units / CTAs) units of
    ceil(tiles / ranges) tiles each, plus a per-unit cost (partial top-ncells lists to write and merge, a refill of the
    queries when the group changes).  That cost is measured, and large: on cfg2 (128 groups x 256 tiles) 8 ranges -- runs
    of 7 units x 32 tiles = 224 tile times on paper instead of the 256 of one range with 20 idle SMs -- ran 6 % SLOWER
    (1.07 vs 1.00 ms: the kernel is within reach of the HBM write rate with 128 SMs already), hence ~8 tile times per
    unit.  Where whole waves are at stake it pays: 108 groups x 2048 tiles (second chunk of 1024 queries at C = 2^19)
    take 4 ranges, 3 units per CTA, 1536 tile times instead of 2048."""
    groups, tiles = max(groups, 1), max(tiles, 1)
    best, best_cost = 1, None
    for cs in range(1, min(tiles, 64) + 1):
        units = groups * cs
        run = -(-units // min(sms, units))
        cost = run * (-(-tiles // cs) + unit_overhead)
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = cs, cost
    return best


_SM_COUNTS = {}


def _sm_count(dev) -> int:
    idx = torch.device(dev).index or 0
    if idx not in _SM_COUNTS:
        _SM_COUNTS[idx] = torch.cuda.get_device_properties(idx).multi_processor_count if torch.cuda.is_available() else 148
    return _SM_COUNTS[idx]


class SearchEngine:
    def __init__(self, index: DeviceIndex, s_budget_bytes: int | None = None, max_chunk: int = 512, fused: bool = True,
                 s_dtype: torch.dtype = torch.float16, ivf_stage1: bool = True, query_maxlen: int = NQ_MAX,
                 streams: int | None = None):
        self.index = index
        # number of leading query tokens that drive candidate generation and the two filter stages
        # (`Q[:, :config.query_maxlen]`, CB/search/index_storage.py:77); the kernels hold one token per lane
        if not 1 <= int(query_maxlen) <= NQ_MAX:
            raise _lib.PlaidError(f"query_maxlen={query_maxlen}: the candidate stage supports 1..{NQ_MAX} query tokens")
        self.query_maxlen = int(query_maxlen)
        # stage 1 of the filter through the inverted file (falls back to the token scan per query on the device)
        self.ivf_stage1 = bool(ivf_stage1)
        if os.environ.get("PLAID_IVF_RANGE_SLOTS"):   # development knob: slots per sort range of the inverted-file stage 1
            _lib.lib().plaid_set_ivf_range_slots(int(os.environ["PLAID_IVF_RANGE_SLOTS"]))
        self.cap_s, self.cap_p = 4096, None    # pair capacity: None = sized from the candidate stride (see _workspace)
        # storage precision of the centroid-score table S: fp16 (what the reference's GPU branch computes S in,
        # candidate_generation.py:52; half the bytes to write and to gather) or fp32 (the CPU branch's precision)
        assert s_dtype in (torch.float16, torch.float32)
        self.s_dtype = s_dtype
        # fused = decompression feeds the tensor cores through shared memory (no passage embeddings in HBM);
        # the unfused pair of kernels materialises D (bf16) and is what the stage-wise parity taps read
        self.fused = bool(fused)
        # fused kernel reads the per-token scale factors the index computed at load (same bits, fewer instructions)
        self.use_inv_norms = os.environ.get("PLAID_NO_INV_NORMS", "0") != "1" and getattr(index, "inv_norms", None) is not None
        if s_budget_bytes is None:
            # room for the centroid-score table of one query chunk: a quarter of what is free next to the index, at most
            # 40 GB (the 33 MB-per-query tables of a 2^19-centroid codebook: 1024 queries in ONE chunk when the shard
            # leaves the room -- fewer launches and, when the shards exchange their stage lists, half the
            # synchronisation points -- else chunks of 592)
            free = torch.cuda.mem_get_info(index.device)[0] if torch.cuda.is_available() else 24 << 30
            s_budget_bytes = max(2 << 30, min(40 << 30, free // 4))
        self.s_budget_bytes = int(s_budget_bytes)
        self.max_chunk = int(max_chunk)
        # Query chunks are independent: with streams > 1 chunk i runs on side stream i % streams with its own workspace,
        # so the latency-bound kernels of one chunk (selects, candidate marking) and the tails of the big ones overlap
        # with the other chunk's work
        self.streams = max(1, int(os.environ.get("PLAID_STREAMS", streams or 1)))
        # set by sharded.ShardedSearcher(mode="exact"): exchanges the stage lists between the shards
        self.exchange = None
        self._side_streams = None
        self._ws_slots = {}
        self._ws_key = None
        self._copy_stream = None
        self._qstage = None        # device staging buffer of a host-fed chunk (_run_chunk_host)
        self._qstage_free = None   # event: the staging buffer has been consumed
        self.host_piece = 256      # queries per PCIe piece of a host-fed batch (0: whole chunks behind a double buffer) ...
        self.host_piece_bytes = 8 << 20   # ... and at most this many bytes (256 FLMR queries of 64 tokens; 48 PreFLMR queries of 320)
        self._ws = None
        self.last_taps: StageTaps | None = None
        self.launch_count = 0      # kernels of libplaid_b200 launched so far (bench.py reports the delta)
        self.events = None         # list of (stage, start, end) CUDA events when stage timing is on
        dev = index.device
        self.flags = torch.zeros(2, device=dev, dtype=torch.int32)  # [0] watchdog, [1] candidate overflow

    # ----------------------------------------------------------------------------------- workspace
    def chunk_size(self, B: int, resident: bool = False) -> int:
        """Queries per chunk.  `resident`: the embeddings are already on the device -- there is no H2D copy for a second
        chunk to hide, and one chunk of up to 1024 queries saves the per-chunk tails of ten kernels (cfg2: 5.69 vs 5.82 ms
        per 1024 queries; host-fed batches keep 512-query chunks: 6.18 vs 6.33 ms end to end)."""
        per_query = self.index.num_centroids * NQ_MAX * (2 if self.s_dtype == torch.float16 else 4)
        large = per_query >= (16 << 20)
        if resident and not large and self.max_chunk == 512 and B <= 1024 and ((B + 3) // 4) * 4 * per_query <= self.s_budget_bytes:
            return max(4, ((B + 3) // 4) * 4)
        if large and self.max_chunk >= 512 and ((B + 3) // 4) * 4 * per_query <= self.s_budget_bytes and B <= 2048:
            return max(4, ((B + 3) // 4) * 4)      # the whole batch at once: the centroid kernel balances any number of groups
        bc = max(4, min(max(self.max_chunk, 592) if large and self.max_chunk >= 512 else self.max_chunk, self.s_budget_bytes // per_query))
        if large and B > bc:
            # Large codebooks (C >= 2^18): centroid scoring dominates the step and its throughput is queries x centroid
            # ranges per wave, so take the chunk that fills the 148 SMs exactly (4 * floor(148 / ranges) queries).
            best, best_rate = bc, bc * max(1, 148 // max(bc // 4, 1))
            for cs in range(1, 9):
                cand = 4 * (148 // cs)
                if cand <= bc and cand * cs > best_rate:
                    best, best_rate = cand, cand * cs
            bc = best
        bc = min(bc, ((B + 3) // 4) * 4)
        bc = max(4, (bc // 4) * 4)
        if not large and B > bc:
            # equal chunks: 520 queries are 2 x 260, not 512 + 8 (a short chunk pays the tails of every kernel for nothing)
            n_chunks = -(-B // bc)
            bc = min(bc, ((-(-B // n_chunks) + 3) // 4) * 4)
        return bc

    def _workspace(self, Bc: int, Lq_pad: int, ncells: int, ndocs: int, k: int, slot: int = 0, groups: int | None = None):
        """groups: query groups (of 4) per plaid_centroid_scores launch when a chunk is scored in pieces (host-fed batches),
        default the whole chunk -- it fixes the number of centroid ranges and with it the layout of the partial cell lists."""
        ix = self.index
        groups = Bc // 4 if groups is None else int(groups)
        key = (Bc, Lq_pad, ncells, ndocs, k, groups)
        if slot:
            hit = self._ws_slots.get(slot)
            if hit is not None and hit[0] == key:
                return hit[1]
        elif self._ws_key == key:
            return self._ws
        dev = ix.device
        C, N = ix.num_centroids, ix.num_passages
        tiles = (C + 255) // 256
        csplit = int(os.environ.get("PLAID_CSPLIT", pick_csplit(groups, tiles, _sm_count(dev))))
        nlists = CELL_LISTS_PER_RANGE * csplit
        nd4 = ndocs // 4
        cand_stride = max(ndocs, min(N, NQ_MAX * ncells * max(ix.max_ivf_len, 1)))
        cand_stride = ((cand_stride + 63) // 64) * 64
        fstride = max(cand_stride, ndocs)
        # (slot, centroid) pairs of the inverted-file stage 1: about one per candidate on the bench workloads; twice the
        # longest candidate list, so that shards of millions of passages stay on that route (a query that still
        # overflows is flagged on the device and takes the code scan)
        cap_p = self.cap_p or max(65536, min(1 << 20, 1 << (2 * cand_stride - 1).bit_length()))
        tok_stride = nd4 * (((max(ix.max_doclen, 1) + 31) // 32) * 32)   # passages are 32-token aligned in D
        tok_stride = ((tok_stride + 127) // 128) * 128
        e = lambda *shape, dtype: torch.empty(*shape, device=dev, dtype=dtype)
        # the two filter stages' result lists live in one i32 block each, laid out [pids | score bits | counts]: for the
        # exact-global sharded search the block IS the all-gather send buffer (sharded.StageExchange)
        s1_msg = torch.zeros(2 * Bc * ndocs + Bc, device=dev, dtype=torch.int32)
        s2_msg = torch.zeros(2 * Bc * nd4 + Bc, device=dev, dtype=torch.int32)
        ws = dict(
            csplit=csplit, nlists=nlists, cand_stride=cand_stride, fstride=fstride, tok_stride=tok_stride, nd4=nd4,
            cap_p=cap_p,
            Qb=e(Bc, Lq_pad, 128, dtype=torch.bfloat16), Qh=e(Bc, Lq_pad, 128, dtype=torch.float16),
            qlens=e(Bc, dtype=torch.int32), cqlens=e(Bc, dtype=torch.int32),
            S=e(Bc, C, NQ_MAX, dtype=self.s_dtype), idx_bits=e(Bc, C // 32, dtype=torch.int32),
            cell_val=e(Bc, NQ_MAX, nlists, ncells, dtype=torch.float32),
            cell_idx=e(Bc, NQ_MAX, nlists, ncells, dtype=torch.int32),
            cells=e(Bc, NQ_MAX, ncells, dtype=torch.int32),
            bitmap=e(Bc, (N + 31) // 32, dtype=torch.int32), wprefix=e(Bc, (N + 31) // 32, dtype=torch.int32),
            surv=e(Bc, self.cap_s, dtype=torch.int32), pair_slot=e(Bc, cap_p, dtype=torch.int32),
            pair_c=e(Bc, cap_p, dtype=torch.int32), sorted_c=e(Bc, 2 * cap_p, dtype=torch.int32),
            ivf_meta=e(Bc, 4, dtype=torch.int32),
            cand_pids=e(Bc, cand_stride, dtype=torch.int32), cand_counts=e(Bc, dtype=torch.int32),
            ws_scores=e(Bc, fstride, dtype=torch.float32), ws_keys=e(2, dtype=torch.int64),   # ABI leftover of select_top: must be non-null, never touched

            s1_msg=s1_msg, s1_pids=s1_msg[: Bc * ndocs].view(Bc, ndocs),
            s1_scores=s1_msg[Bc * ndocs: 2 * Bc * ndocs].view(torch.float32).view(Bc, ndocs), s1_counts=s1_msg[2 * Bc * ndocs:],
            s2_msg=s2_msg, s2_pids=s2_msg[: Bc * nd4].view(Bc, nd4),
            s2_scores=s2_msg[Bc * nd4: 2 * Bc * nd4].view(torch.float32).view(Bc, nd4), s2_counts=s2_msg[2 * Bc * nd4:],
            tok_offsets=e(Bc, nd4 + 1, dtype=torch.int32),
            D=None,   # fp16 [Bc * tok_stride, 128], allocated on first use by the unfused path
            scores=e(Bc, nd4, dtype=torch.float32),
            out_pids=e(Bc, k, dtype=torch.int32), out_scores=e(Bc, k, dtype=torch.float32),
            out_counts=e(Bc, dtype=torch.int32),
        )
        if slot:
            self._ws_slots[slot] = (key, ws)
        else:
            self._ws_key, self._ws = key, ws
        return ws

    def _dense_buffer(self, ws, Bc):
        if ws["D"] is None:
            ws["D"] = torch.empty(Bc * ws["tok_stride"], 128, device=self.index.device, dtype=torch.float16)
        return ws["D"]

    # ----------------------------------------------------------------------------------- one chunk
    def _flag_ptrs(self):
        return ctypes.c_void_p(self.flags.data_ptr()), ctypes.c_void_p(self.flags.data_ptr() + 4)

    # kernels launched by each C-ABI entry point (memsets are not kernels)
    _LAUNCHES = {"plaid_prepare_queries": 1, "plaid_centroid_scores": 1, "plaid_candidates": 3, "plaid_approx_scores": 1, "plaid_filter_stage1_ivf": 4,
                 "plaid_doc_token_offsets": 1, "plaid_decompress_normalize_f16": 1, "plaid_maxsim_packed": 1,
                 "plaid_maxsim_fused": 1,
                 "plaid_select_top": 1}

    def _call(self, stage, name, *args):
        """One C-ABI call; with stage timing on, bracketed by CUDA events on the launching stream."""
        self.launch_count += self._LAUNCHES[name]
        if self.events is None:
            _lib.call(name, *args)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call(name, *args)
        e1.record()
        self.events.append((stage, e0, e1))

    def _exchange(self, stage, msg, pids, counts, Bc, rows, keep):
        """Stage-list exchange of the exact-global sharded search (all-gather + merge + keep-own-range), timed like a stage."""
        self.launch_count += 1
        if self.events is None:
            self.exchange.globalize(msg, pids, counts, Bc, rows, keep)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.exchange.globalize(msg, pids, counts, Bc, rows, keep)
        e1.record()
        self.events.append((stage, e0, e1))

    def stage_candidates(self, ws, Qc: torch.Tensor, Lq_pad: int, ncells: int, thr: float,
                         remove_zero_rows: bool, Bc: int):
        """a1-a4: query prep, centroid scoring (+ pruning mask, top-ncells), candidate pids."""
        self.stage_scores(ws, Qc, 0, Bc, Lq_pad, ncells, thr, remove_zero_rows)
        self.stage_candidate_pids(ws, Qc.shape[0], ncells)

    def stage_scores(self, ws, Qc: torch.Tensor, row0: int, rows_pad: int, Lq_pad: int, ncells: int, thr: float,
                     remove_zero_rows: bool):
        """a1-a2, a4 for the workspace rows row0 .. row0 + rows_pad (a multiple of 4; Qc holds the real ones): query prep and
        centroid scoring with its pruning mask and partial top-ncells lists.  A chunk may be scored in several pieces."""
        ix = self.index
        b, Lq, _ = Qc.shape
        st = _stream()
        C = ix.num_centroids
        wd, _ = self._flag_ptrs()
        call = self._call
        r0, r1 = row0, row0 + rows_pad
        call("prepare", "plaid_prepare_queries", _p(Qc), b, Lq, int(remove_zero_rows), rows_pad, Lq_pad, _p(ws["Qb"][r0:r1]),
             _p(ws["Qh"][r0:r1]), _p(ws["qlens"][r0:r1]), st)
        if self.query_maxlen < NQ_MAX:
            torch.clamp(ws["qlens"][r0:r1], max=self.query_maxlen, out=ws["cqlens"][r0:r1])
        cq = ws["qlens"] if self.query_maxlen >= NQ_MAX else ws["cqlens"]
        call("centroid_scores", "plaid_centroid_scores", _p(ix.centroids_bf16), C, _p(ws["Qb"][r0:r1]), _p(cq[r0:r1]), rows_pad,
             Lq_pad, float(thr), ncells, ws["csplit"], _p(ws["S"][r0:r1]), int(self.s_dtype == torch.float16),
             _p(ws["idx_bits"][r0:r1]), _p(ws["cell_val"][r0:r1]), _p(ws["cell_idx"][r0:r1]), wd, st)

    def stage_candidate_pids(self, ws, b: int, ncells: int):
        """a3: merge of the partial cell lists, inverted-file union -> sorted unique candidate pids of the chunk's b queries."""
        ix = self.index
        st = _stream()
        C, N = ix.num_centroids, ix.num_passages
        _, ovf = self._flag_ptrs()
        call = self._call
        cq = self._candidate_qlens(ws, refresh=False)
        call("candidates", "plaid_candidates", _p(ws["cell_val"]), _p(ws["cell_idx"]), _p(cq), b, ncells, ws["nlists"],
             _p(ix.ivf_pids), _p(ix.ivf_offsets), C, N, _p(ws["cells"]), _p(ws["bitmap"]), _p(ws["cand_pids"]),
             _p(ws["cand_counts"]), ws["cand_stride"], ovf, _p(ws["wprefix"]), st)

    def _candidate_qlens(self, ws, refresh: bool = True):
        """Per-query token count of the candidate stage: min(qlens, query_maxlen).  The kernels clamp to 32
        themselves, so the plain qlens serve unless the index was built with a shorter query_maxlen."""
        if self.query_maxlen >= NQ_MAX:
            return ws["qlens"]
        if refresh:
            torch.clamp(ws["qlens"], max=self.query_maxlen, out=ws["cqlens"])
        return ws["cqlens"]

    def stage_rank(self, ws, b: int, Lq_pad: int, ndocs: int, k: int, Bc: int, out=None, out_stride=None):
        """a5-a10: two-stage filter, decompression, exact MaxSim, top-k -- on ws['cand_pids'/'S'/'idx_bits'].
        out = (pids, scores, counts) raw pointers of the rows the final top-k is written to (row stride out_stride);
        default: the workspace's out_* buffers."""
        ix = self.index
        st = _stream()
        C = ix.num_centroids
        wd, _ = self._flag_ptrs()
        call = self._call
        cq = self._candidate_qlens(ws)
        # plaid_filter_pids' four launches issued one by one so each can be timed on its own
        cs, fs, nd4_ = ws["cand_stride"], ws["fstride"], ndocs // 4
        f16 = int(self.s_dtype == torch.float16)
        if self.ivf_stage1:
            call("filter_stage1", "plaid_filter_stage1_ivf", _p(ws["cand_pids"]), _p(ws["cand_counts"]), b, cs, _p(ws["S"]),
                 f16, _p(cq), _p(ws["idx_bits"]), C, _p(ix.codes), _p(ix.offsets), _p(ix.ivf_pids),
                 _p(ix.ivf_offsets), _p(ws["bitmap"]), _p(ws["wprefix"]), ix.num_passages, _p(ws["surv"]), self.cap_s,
                 _p(ws["pair_slot"]), _p(ws["pair_c"]), _p(ws["sorted_c"]), ws["cap_p"], _p(ws["ivf_meta"]),
                 _p(ws["ws_scores"]), st)
        else:
            call("filter_stage1", "plaid_approx_scores", _p(ws["cand_pids"]), _p(ws["cand_counts"]), b, cs, _p(ws["S"]), f16,
                 _p(cq), _p(ws["idx_bits"]), C, _p(ix.codes), _p(ix.offsets), _p(ws["ws_scores"]), st)
        call("select1", "plaid_select_top", _p(ws["cand_pids"]), _p(ws["ws_scores"]), _p(ws["cand_counts"]), b, cs, ndocs,
             _p(ws["s1_pids"]), _p(ws["s1_scores"]), _p(ws["s1_counts"]), ndocs, _p(ws["ws_keys"]), st)
        if self.exchange is not None:     # exact-global sharding: the collection's ndocs best, then this shard's share of them
            self._exchange("exchange_stage1", ws["s1_msg"], ws["s1_pids"], ws["s1_counts"], Bc, b, ndocs)
        call("filter_stage2", "plaid_approx_scores", _p(ws["s1_pids"]), _p(ws["s1_counts"]), b, ndocs, _p(ws["S"]), f16,
             _p(cq), None, C, _p(ix.codes), _p(ix.offsets), _p(ws["ws_scores"]), st)
        call("select2", "plaid_select_top", _p(ws["s1_pids"]), _p(ws["ws_scores"]), _p(ws["s1_counts"]), b, ndocs, nd4_,
             _p(ws["s2_pids"]), _p(ws["s2_scores"]), _p(ws["s2_counts"]), nd4_, _p(ws["ws_keys"]), st)
        nd4 = ws["nd4"]
        if self.exchange is not None:
            self._exchange("exchange_stage2", ws["s2_msg"], ws["s2_pids"], ws["s2_counts"], Bc, b, nd4)
        call("doc_offsets", "plaid_doc_token_offsets", _p(ws["s2_pids"]), _p(ws["s2_counts"]), b, nd4, _p(ix.offsets),
             32, _p(ws["tok_offsets"]), st)
        if self.fused and Lq_pad <= 384:
            call("maxsim_fused", "plaid_maxsim_fused", _p(ws["Qh"]), _p(ws["qlens"]), b, Bc, Lq_pad, _p(ws["s2_pids"]),
                 _p(ws["s2_counts"]), nd4, _p(ws["tok_offsets"]), _p(ix.offsets), _p(ix.weight_table), _p(ix.residuals),
                 _p(ix.codes), _p(ix.centroids_f16), C, ix.nbits, _p(ix.inv_norms) if self.use_inv_norms else None,
                 _p(ws["scores"]), wd, st)
        else:
            D = self._dense_buffer(ws, Bc)
            call("decompress", "plaid_decompress_normalize_f16", _p(ws["s2_pids"]), _p(ws["s2_counts"]), b, nd4,
                 _p(ws["tok_offsets"]), ws["tok_stride"], _p(ix.offsets), _p(ix.weight_table), _p(ix.residuals),
                 _p(ix.codes), _p(ix.centroids_f16), C, ix.nbits, _p(D), st)
            call("maxsim", "plaid_maxsim_packed", _p(ws["Qh"]), _p(ws["qlens"]), b, Bc, Lq_pad, _p(D), _p(ws["tok_offsets"]),
                 _p(ws["s2_counts"]), nd4, ws["tok_stride"], 1, 1, 1, _p(ws["scores"]), wd, st)
        op, os_, oc = out if out is not None else (_p(ws["out_pids"]), _p(ws["out_scores"]), _p(ws["out_counts"]))
        call("topk", "plaid_select_top", _p(ws["s2_pids"]), _p(ws["scores"]), _p(ws["s2_counts"]), b, nd4, k, op, os_, oc,
             out_stride or k, _p(ws["ws_keys"]), st)

    def _run_chunk(self, Qc: torch.Tensor, Lq_pad: int, ncells: int, thr: float, ndocs: int, k: int,
                   remove_zero_rows: bool, Bc: int, out=None, out_stride=None, slot: int = 0):
        """Qc f32 [b, Lq, 128] on device, b <= Bc.  Enqueues the whole pipeline; returns the workspace."""
        ws = self._workspace(Bc, Lq_pad, ncells, ndocs, k, slot)
        self.stage_candidates(ws, Qc, Lq_pad, ncells, thr, remove_zero_rows, Bc)
        self.stage_rank(ws, Qc.shape[0], Lq_pad, ndocs, k, Bc, out, out_stride)
        return ws

    def _run_chunk_host(self, Qh: torch.Tensor, Lq_pad: int, ncells: int, thr: float, ndocs: int, k: int,
                        remove_zero_rows: bool, Bc: int, piece: int, out=None, out_stride=None):
        """One chunk fed from (pinned) host memory in pieces: the embeddings cross PCIe `piece` queries at a time on the copy
        stream, query prep + centroid scoring run per piece as it lands (its launch splits the codebook into ranges, so a
        256-query piece still fills the SMs), and everything behind them runs once on the whole chunk.  Only the first piece's
        copy is exposed (8.4 MB instead of the 16.8 MB of a 512-query first chunk), and the batch is searched as ONE chunk."""
        B, Lq, dim = Qh.shape
        dev = self.index.device
        ws = self._workspace(Bc, Lq_pad, ncells, ndocs, k, 0, groups=piece // 4)
        if self._qstage is None or self._qstage.shape != (Bc, Lq, dim):
            self._qstage = torch.empty(Bc, Lq, dim, device=dev, dtype=torch.float32)
            self._qstage_free = None
        main = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        if self._qstage_free is not None:
            self._copy_stream.wait_event(self._qstage_free)  # the previous call's query prep has read the staging buffer
        else:
            self._copy_stream.wait_stream(main)              # first use: the buffer was allocated on the compute stream
        events = []
        with torch.cuda.stream(self._copy_stream):
            for r0 in range(0, B, piece):
                r1 = min(B, r0 + piece)
                self._qstage[r0:r1].copy_(Qh[r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                events.append(ev)
        for i, r0 in enumerate(range(0, B, piece)):
            r1 = min(B, r0 + piece)
            main.wait_event(events[i])
            rows_pad = (Bc - r0) if r1 == B else (r1 - r0)    # the last piece also zeroes the chunk's padding rows
            self.stage_scores(ws, self._qstage[r0:r1], r0, rows_pad, Lq_pad, ncells, thr, remove_zero_rows)
        self._qstage_free = torch.cuda.Event()
        self._qstage_free.record(main)                       # the next batch may overwrite the staging buffer from here on
        self.stage_candidate_pids(ws, B, ncells)
        self.stage_rank(ws, B, Lq_pad, ndocs, k, Bc, out, out_stride)
        return ws

    # ----------------------------------------------------------------------------------- public
    def search_batch(self, Q: torch.Tensor, k: int = 100, ncells: int | None = None,
                     centroid_score_threshold: float | None = None, ndocs: int | None = None,
                     remove_zero_rows: bool = False, global_pids: bool = True, keep_taps: bool = False,
                     on_chunk=None, out=None):
        """Q f32 [B, Lq, 128] (host or device) -> (pids i32 [B, k], scores f32 [B, k], counts i32 [B]) on
        the device.  Slots past counts[b] hold pid -1 / score -inf.  `out` = preallocated contiguous result tensors
        (the sharded search passes views of its all-gather send buffer); the last kernel of every chunk writes its
        rows straight into them."""
        d_ncells, d_thr, d_ndocs = search_defaults(k)
        ncells = d_ncells if ncells is None else int(ncells)
        thr = d_thr if centroid_score_threshold is None else float(centroid_score_threshold)
        ndocs = d_ndocs if ndocs is None else int(ndocs)
        if ndocs // 4 < 1:
            raise ValueError("ndocs must be >= 4")
        ix = self.index
        dev = ix.device
        B, Lq, dim = Q.shape
        Lq_pad = ((Lq + 31) // 32) * 32
        kk = min(k, ndocs // 4)
        if out is not None:
            out_p, out_s, out_c = out
            assert out_p.shape == (B, k) and out_s.shape == (B, k) and out_c.shape == (B,)
            assert out_p.is_contiguous() and out_s.is_contiguous() and out_c.is_contiguous()
            assert out_p.dtype == torch.int32 and out_s.dtype == torch.float32 and out_c.dtype == torch.int32
        else:
            out_p = torch.empty((B, k), device=dev, dtype=torch.int32)
            out_s = torch.empty((B, k), device=dev, dtype=torch.float32)
            out_c = torch.empty(B, device=dev, dtype=torch.int32)
        if B == 0 or ix.num_passages == 0 or kk < k:       # the kernels fill exactly kk columns of every row
            out_p.fill_(-1)
            out_s.fill_(float("-inf"))
            out_c.zero_()
        if B == 0 or ix.num_passages == 0:
            return out_p, out_s, out_c
        if torch.cuda.current_device() != dev.index:
            # launches go to the current device's stream: the caller must have selected the index's device
            raise _lib.PlaidError(f"search_batch: current CUDA device {torch.cuda.current_device()} != index device {dev.index}; "
                                  "wrap the call in torch.cuda.device(index.device)")
        Bc = self.chunk_size(B, resident=Q.is_cuda and self.exchange is None)
        # host-fed batch that fits one chunk: pieces of `host_piece` queries cross PCIe behind the centroid scoring of the
        # previous piece (_run_chunk_host); larger batches / large codebooks: whole chunks, one ahead of the search
        piece_req = int(os.environ.get("PLAID_HOST_PIECE", self.host_piece))
        if piece_req > 0 and Lq * dim * 4 * piece_req > self.host_piece_bytes:     # long queries: pieces of ~8 MB
            piece_req = max(4, self.host_piece_bytes // (Lq * dim * 4))
        piece = max(4, (piece_req // 4) * 4)
        host_pipe = (not Q.is_cuda and piece_req > 0 and self.exchange is None and self.streams == 1 and B > piece
                     and self.chunk_size(B, resident=True) >= B and self.max_chunk == 512)
        if host_pipe:
            npieces = -(-B // piece)
            piece = ((-(-B // npieces) + 3) // 4) * 4         # equal pieces: a short last one would cost a whole wave of its own
            Bc = self.chunk_size(B, resident=True)
            Qh = Q.to(torch.float32).contiguous()
            if not Qh.is_pinned():
                Qh = Qh.pin_memory()
            rows = (ctypes.c_void_p(out_p.data_ptr()), ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_c.data_ptr()))
            ws = self._run_chunk_host(Qh, Lq_pad, ncells, thr, ndocs, kk, remove_zero_rows, Bc, piece, rows, k)
            if on_chunk is not None:
                on_chunk(ws, B)
            if keep_taps:
                self.last_taps = self._taps(ws)
            if global_pids and ix.pid_base:
                out_p.add_((out_p >= 0).to(torch.int32) * ix.pid_base)
            return out_p, out_s, out_c
        feed = self._host_feed(Q, Bc) if not Q.is_cuda else None
        Qd = Q.to(torch.float32).contiguous() if Q.is_cuda else None
        n_chunks = (B + Bc - 1) // Bc
        nstreams = min(self.streams, n_chunks) if not (keep_taps or on_chunk is not None) else 1
        main = torch.cuda.current_stream(dev)
        if nstreams > 1:
            if self._side_streams is None or len(self._side_streams) < nstreams:
                self._side_streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
            for st in self._side_streams[:nstreams]:
                st.wait_stream(main)                        # inputs, output buffers and the previous call's work
        for ci, b0 in enumerate(range(0, B, Bc)):
            b1 = min(B, b0 + Bc)
            slot = ci % nstreams if nstreams > 1 else 0
            rows = (ctypes.c_void_p(out_p.data_ptr() + b0 * k * 4), ctypes.c_void_p(out_s.data_ptr() + b0 * k * 4),
                    ctypes.c_void_p(out_c.data_ptr() + b0 * 4))
            with torch.cuda.stream(self._side_streams[slot] if nstreams > 1 else main):
                Qc = Qd[b0:b1] if feed is None else feed.chunk(ci)
                ws = self._run_chunk(Qc, Lq_pad, ncells, thr, ndocs, kk, remove_zero_rows, Bc, rows, k, slot)
                if feed is not None:
                    feed.release(ci)
            n = b1 - b0
            if on_chunk is not None:
                on_chunk(ws, n)
            if keep_taps:
                self.last_taps = self._taps(ws)
        if nstreams > 1:
            for st in self._side_streams[:nstreams]:
                main.wait_stream(st)
        if global_pids and ix.pid_base:
            out_p.add_((out_p >= 0).to(torch.int32) * ix.pid_base)
        return out_p, out_s, out_c

    @staticmethod
    def _taps(ws) -> StageTaps:
        return StageTaps(
            Qb=ws["Qb"], qlens=ws["qlens"], S=ws["S"], idx_bits=ws["idx_bits"], cells=ws["cells"],
            cand_pids=ws["cand_pids"], cand_counts=ws["cand_counts"], stage1_pids=ws["s1_pids"],
            stage1_scores=ws["s1_scores"], stage1_counts=ws["s1_counts"], stage2_pids=ws["s2_pids"],
            stage2_scores=ws["s2_scores"], stage2_counts=ws["s2_counts"], tok_offsets=ws["tok_offsets"],
            D=ws["D"], tok_stride=ws["tok_stride"], scores=ws["scores"])

    def _host_feed(self, Q: torch.Tensor, Bc: int):
        """Host query embeddings -> device, one chunk ahead of the compute on a copy stream."""
        Qh = Q.to(torch.float32).contiguous()
        if not Qh.is_pinned():
            Qh = Qh.pin_memory()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.index.device)
        return _HostFeed(Qh, Bc, self.index.device, self._copy_stream)

    def check_flags(self):
        """Host-side check (synchronises): raises if a kernel watchdog fired or candidates overflowed."""
        wd, ovf = self.flags.tolist()
        if wd:
            raise _lib.PlaidError("a tcgen05 pipeline wait timed out inside a kernel (watchdog flag set)")
        if ovf:
            raise _lib.PlaidError("candidate list overflowed its workspace (cand_stride too small)")
