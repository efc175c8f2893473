#!/usr/bin/env python
"""bench.py -- FLMR / PLAID late-interaction search throughput on B200.

A "step" is one pass of the whole search path (centroid scoring -> candidate pids ->
two-stage filter -> decompression -> exact MaxSim -> top-k) over one batch of synthetic queries.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1|cfg2|cfg3]

N = 1  : BASELINE.json configs[1] (OK-VQA GS-112K-shaped index, 1024 FLMR queries, k = 100).
N > 1  : launched by torchrun, one rank per GPU; every rank owns a pid-range shard of the same size
         as the N = 1 index (weak scaling), queries are replicated, and the only exchange is one
         all-gather of the per-shard top-k lists followed by the merge kernel.
`--impl reference` times the reference's CPU implementation of the same path on the host cores
(oracle/ref_search.py) on a bounded sample of the same workload.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: passages, doclen lo/hi, nbits, queries per step, Lq, k   
    "cfg1": dict(N=10_000, lo=120, hi=239, nbits=2, B=256, Lq=64, k=100,
                 desc="10k-passage synthetic PLAID index, nbits=2, 256 FLMR queries (32+32 tokens), k=100"),
    "cfg2": dict(N=112_000, lo=120, hi=239, nbits=2, B=1024, Lq=64, k=100,
                 desc="OK-VQA GS-112K-shaped synthetic index (112k passages, ~180 tok/passage, nbits=2), "
                      "1024 FLMR queries (32+32 tokens), k=100"),
    "cfg4shard": dict(N=1_250_000, lo=120, hi=239, nbits=2, B=1024, Lq=64, k=100, C=524_288,
                      desc="one 1/8 shard (1.25M passages) of the 10M-passage index, C=524288 centroids of the full index, "
                           "nbits=2, 1024 FLMR queries, k=100"),
    "cfg5": dict(B=4096, dpq=100, Ld=240, lo=120, hi=239, Lq=64, rerank=True,
                 desc="exhaustive uncompressed MaxSim rerank of top-100 candidates per query, 4096 queries, bf16 passage "
                      "embeddings [409600, 240, 128] (lengths U{120..239}), Lq=64"),
    "cfg5small": dict(B=512, dpq=100, Ld=240, lo=120, hi=239, Lq=64, rerank=True,
                      desc="cfg5 shape at 512 queries (profiling size)"),
    "codec": dict(n=4_000_000, C=65_536, nbits=2, codec=True,
                  desc="index build (next row 8f-3): ResidualCodec.compress of 4M token embeddings against 65536 centroids, nbits=2"),
    "cfg3": dict(N=100_000, lo=128, hi=512, nbits=4, B=256, Lq=320, k=100,
                 desc="E-VQA/InfoSeek-shaped 100k-passage index, nbits=4, 256 PreFLMR 320-token queries, k=100"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def write_only_peak(dev, nbytes=4 << 30, reps=5):
    """GB/s of a write-only stream (torch fill of `nbytes`), CUDA events, best of `reps`."""
    import torch
    x = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    best = 0.0
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x.zero_()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del x
    return best


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines, self.proc = [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop_after_min_samples(self, step, min_samples=3, max_seconds=1.0):
        """A very short timed region can end before nvidia-smi has delivered a few samples: keep the same load
        running (untimed) until it has, and say so."""
        extra, t0 = 0, time.time()
        while self.proc is not None and len(self.lines) < min_samples and time.time() - t0 < max_seconds:
            step()
            extra += 1
        out = self.stop()
        if extra:
            out["note"] = f"{extra} extra untimed steps of the same load were run to collect the samples"
        return out

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(w, rank, device):
    """Synthetic index (code-space generator) + gold-planted queries, on `device`."""
    import torch
    from reranking_multimodal_retrievers_b200 import synthetic
    sx = synthetic.make_synthetic_index(w["N"], w["lo"], w["hi"], w["nbits"], seed=1234 + 17 * rank, mode="codes",
                                        device=device, num_centroids=w.get("C"))
    Q = synthetic.make_queries(sx, w["B"], w["Lq"], seed=99)      # same queries on every rank
    return sx, Q.to(torch.float32)


def cpu_index_from(sx):
    from oracle import plaid_oracle as po
    c = sx.cpu()
    return po.OracleIndex(centroids=c.centroids, bucket_weights=c.bucket_weights, codes=c.codes, residuals=c.residuals,
                          doclens=c.doclens, ivf=c.ivf, ivf_lengths=c.ivf_lengths, nbits=c.nbits)


def make_cpu_searcher(sx):
    from oracle.ref_search import CpuSearcher
    return CpuSearcher(cpu_index_from(sx))


def run_cpu_sample(cs, Q, k, budget_s, max_queries, warm=True):
    """Reference CPU path on a bounded sample of the step's queries.  Returns dict for `cpu_baseline`."""
    Qc = Q.cpu()
    if warm:
        cs.search_all(Qc[:2], k)                   # warm-up (thread pools, page-in)
    for key in cs.stage_s:
        cs.stage_s[key] = 0.0
    t0 = time.perf_counter()
    n, toks, results = 0, 0, []
    while n < min(max_queries, Qc.shape[0]) and (time.perf_counter() - t0) < budget_s:
        res, t3 = cs.search_all(Qc[n:n + 1], k, remove_zero_tensors=True)
        results.append(res[0])
        n += 1
        toks += t3
    dt = time.perf_counter() - t0
    detail = ("the reference's compiled C++ operators (oracle/_ref: filter_pids / decompress_residuals / segmented_lookup / "
              "segmented_maxsim .cpp, built from /root/reference) under a restatement of IndexScorer.rank's Python glue"
              if cs.kind == "reference" else "single-thread C restatement of the reference operators (oracle/plaid_oracle.c)")
    return dict(queries=n, seconds=dt, tokens=toks, kind=cs.kind, kind_detail=detail, cores=cs.cores, results=results,
                stage_share={s: round(v / max(dt, 1e-9), 3) for s, v in cs.stage_s.items()})


def agreement(ours, ref_results, k):
    """End-to-end agreement when each side uses its OWN centroid-score table: ours = bf16 contraction, fp16
    table, fp16 MaxSim operands; reference = fp32 throughout.  ours = (pids, scores, counts) tensors of the same queries."""
    import torch
    p, s, c = (t.cpu() for t in ours)
    top1 = same_set = 0
    overlap, max_rel, head = [], 0.0, 0
    for b, (rp, rs) in enumerate(ref_results):
        n = int(c[b])
        mine = dict(zip(p[b, :n].tolist(), s[b, :n].tolist()))
        theirs = dict(zip(rp, rs))
        top1 += int(n > 0 and len(rp) > 0 and int(p[b, 0]) == rp[0])
        common = set(mine) & set(theirs)
        overlap.append(len(common) / max(len(rp), 1))
        same_set += int(len(common) == len(rp) == n)
        for pid in common:
            max_rel = max(max_rel, abs(mine[pid] - theirs[pid]) / max(abs(theirs[pid]), 1e-6))
        # rank identity on the well-separated head: reference entries whose gap to the next one exceeds 2e-3 * score
        m = 0
        while m < min(n, len(rp)) - 1 and rs[m] - rs[m + 1] > 2e-3 * abs(rs[m]) and int(p[b, m]) == rp[m]:
            m += 1
        head += m
    nq = max(len(ref_results), 1)
    return {"queries": len(ref_results), "k": k, "top1_agreement": top1 / nq, "overlap_at_k_mean": sum(overlap) / nq,
            "overlap_at_k_min": min(overlap) if overlap else None, "identical_result_sets": same_set / nq,
            "max_rel_score_err_common_pids": max_rel, "mean_identical_well_separated_head": head / nq,
            "note": "each side uses its own centroid-score table (ours: bf16 operands, fp16 table; reference: fp32); integer "
                    "stages are bit-exact only with an injected table (tests), so near-tie passages at the ndocs / ndocs/4 "
                    "cut-offs may differ"}


CFG4 = dict(blocks=8, passages_per_block=1_250_000, lo=120, hi=239, nbits=2, C=524_288, B=1024, Lq=64, k=100,
            desc="10M-passage synthetic PLAID index (8 blocks of 1.25M passages, ~180 tok/passage, 1.8G tokens, C=524288, nbits=2) "
                 "pid-sharded over the ranks, 1024 FLMR queries, k=100")


def bench_cfg4(args, rank, world, dev, peaks, steps=4, warmup=2):
    """BASELINE.json configs[3]: ONE fixed 10M-passage collection, pid-range sharded over `world` ranks (strong scaling:
    every rank holds 10M/world passages), the same 1024 queries on every rank, per-shard search + one all-gather of the
    top-k blocks + merge (oracle (A)).  Returns the `cfg4` block of the bench line (rank 0) or None."""
    import types
    import torch
    import torch.distributed as dist
    from reranking_multimodal_retrievers_b200 import Searcher, sharded, synthetic
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    c = CFG4
    if c["blocks"] % world:
        return {"skipped": f"world size {world} does not divide the collection's {c['blocks']} blocks"}
    free, _ = torch.cuda.mem_get_info(dev)
    per_rank_tokens = c["blocks"] // world * c["passages_per_block"] * (c["lo"] + c["hi"]) / 2
    need = per_rank_tokens * (4 + 16 * c["nbits"] + 2 + 4 + 8) + (24 << 30)      # index + derived arrays + IVF scratch + workspace
    if free < need:
        return {"skipped": f"needs ~{need / 2**30:.0f} GiB per rank, {free / 2**30:.0f} GiB free"}
    t0 = time.perf_counter()
    mine = range(rank * c["blocks"] // world, (rank + 1) * c["blocks"] // world)
    shard = synthetic.make_collection_shard(mine, c["passages_per_block"], c["lo"], c["hi"], c["nbits"], c["C"], c["blocks"],
                                            seed=4000, device=dev)
    index = DeviceIndex(shard, dev)
    B, Lq, k = c["B"], c["Lq"], c["k"]
    Q = torch.empty(B, Lq, 128, device=dev, dtype=torch.float32)
    if rank == 0:      # queries planted in block 0 (rank 0 holds it for every world size): the same queries for every N
        ppb = c["passages_per_block"]
        view = types.SimpleNamespace(codes=shard.codes, residuals=shard.residuals, doclens=shard.doclens[:ppb],
                                     centroids=shard.centroids, bucket_weights=shard.bucket_weights, nbits=shard.nbits,
                                     dim=shard.dim, num_passages=ppb)
        Q.copy_(synthetic.make_queries(view, B, Lq, seed=199))
    if world > 1:
        dist.broadcast(Q, src=0)
    del shard
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    searcher = Searcher(index=index)
    eng = searcher.ranker.engine
    ss = sharded.ShardedSearcher(searcher) if world > 1 else searcher
    Qhost = Q.cpu().pin_memory()
    queries = {i: f"question {i}" for i in range(B)}
    from reranking_multimodal_retrievers_b200 import search_custom_collection

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n

    stats = dict(ncand=0, T1=0, T2=0, T3=0, scan_queries=0)

    def account(ws, n):
        dl = index.doclens
        for key, pk, ck in (("T1", "cand_pids", "cand_counts"), ("T2", "s1_pids", "s1_counts"), ("T3", "s2_pids", "s2_counts")):
            cnt = ws[ck][:n].long()
            for b0 in range(0, n, 16):                       # bounded temporaries: candidate lists are long here
                m = torch.arange(ws[pk].shape[1], device=dev).unsqueeze(0) < cnt[b0:b0 + 16].unsqueeze(1)
                stats[key] += int(dl[torch.where(m, ws[pk][b0:min(n, b0 + 16)], 0).long()].mul(m).sum())
        stats["ncand"] += int(ws["cand_counts"][:n].sum())
        stats["scan_queries"] += int((ws["ivf_meta"][:n, 2] != 0).sum())

    acc = eng.search_batch(Q, k=k, remove_zero_rows=True, on_chunk=account)
    torch.cuda.synchronize()
    eng.check_flags()
    found = int((acc[0][:, 0] >= 0).sum())
    for _ in range(warmup):
        ss.search_batch(Q, k, True)
    eng.events = []
    ms = timed(lambda: ss.search_batch(Q, k, True), steps)
    events, eng.events = eng.events, None
    ms_api = timed(lambda: search_custom_collection(ss, queries, Qhost, num_document_to_retrieve=k, remove_zero_tensors=True), steps)
    eng.check_flags()
    # the same with the stage lists exchanged too (exact-global truncation: the result of ONE index holding the collection)
    exact = None
    if world > 1:
        sx_ = sharded.ShardedSearcher(searcher, mode="exact")
        st2 = dict(T2=0, T3=0)

        def account2(ws, n):
            dl = index.doclens
            for key, pk, ck in (("T2", "s1_pids", "s1_counts"), ("T3", "s2_pids", "s2_counts")):
                cnt = ws[ck][:n].long()
                m = torch.arange(ws[pk].shape[1], device=dev).unsqueeze(0) < cnt.unsqueeze(1)
                st2[key] += int(dl[torch.where(m, ws[pk][:n], 0).long()].mul(m).sum())

        eng.search_batch(Q, k=k, remove_zero_rows=True, on_chunk=account2)
        for _ in range(warmup):
            sx_.search_batch(Q, k, True)
        eng.events = []
        ms_x = timed(lambda: sx_.search_batch(Q, k, True), steps)
        ev_x, eng.events = eng.events, None
        ms_x_api = timed(lambda: search_custom_collection(sx_, queries, Qhost, num_document_to_retrieve=k, remove_zero_tensors=True), steps)
        eng.check_flags()
        t2 = torch.tensor([st2["T2"], st2["T3"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t2)
        kx = {}
        for stage, a, b in ev_x:
            kx[stage] = kx.get(stage, 0.0) + a.elapsed_time(b) / steps
        if os.environ.get("PLAID_BENCH_ALLRANKS"):          # development aid: every rank's stage times (rank skew at the exchanges)
            print(f"[rank {rank}] exact: " + " ".join(f"{s_}={v:.3f}" for s_, v in kx.items()), file=sys.stderr, flush=True)
        exact = {"ms_per_step": ms_x, "queries_per_s": B / (ms_x * 1e-3), "doc_tokens_per_s": float(t2[1]) / (ms_x * 1e-3),
                 "e2e": {"ms_per_step": ms_x_api, "queries_per_s": B / (ms_x_api * 1e-3)},
                 "T2_tokens_per_query_all_shards": float(t2[0]) / B, "T3_tokens_per_query_all_shards": float(t2[1]) / B,
                 "kernels_rank0": {s_: round(v, 3) for s_, v in kx.items()},
                 "semantics": "ShardedSearcher(mode='exact'): two more all-gathers per query chunk (the shards' stage-1 and stage-2 "
                              "(score, pid) lists), merge + keep-own-range kernels; returns exactly what one index holding all "
                              "10M passages returns (tests/nccl_worker.py), and stage 2 / MaxSim shrink with the shard count"}
        eng.exchange = None
    tot = torch.tensor([stats["T1"], stats["T2"], stats["T3"], stats["ncand"], stats["scan_queries"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)                                 # tokens scored by ALL shards
    if rank != 0:
        return None
    stage_ms = {}
    for stage, a, b in events:
        stage_ms[stage] = stage_ms.get(stage, 0.0) + a.elapsed_time(b) / steps
    T1, T2, T3, ncand, scan_q = (float(x) for x in tot.tolist())
    C = c["C"]
    chunks = (B + eng.chunk_size(B) - 1) // eng.chunk_size(B)
    cs_bytes = 2.0 * C * 32 * B + 2.0 * C * 128 * chunks
    kernels = {s_: {"ms_per_step": round(v, 3), "share": round(v / ms, 4)} for s_, v in stage_ms.items()}
    if "centroid_scores" in kernels:
        kernels["centroid_scores"].update(achieved_GBps=round(cs_bytes / (stage_ms["centroid_scores"] * 1e-3) / 1e9, 1),
                                          frac_hbm=round(cs_bytes / (stage_ms["centroid_scores"] * 1e-3) / 1e9 / peaks["hbm"], 4),
                                          write_only_peak_GBps=round(peaks.get("write_only", 0.0), 1),
                                          frac_write_only_peak=round(cs_bytes / (stage_ms["centroid_scores"] * 1e-3) / 1e9 / max(peaks.get("write_only", 0.0), 1.0), 4),
                                          note="replicated on every rank: the codebook is global and every shard's filter needs "
                                               "the whole score table of every query")
    return {
        "workload": "cfg4: " + c["desc"], "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup,
        "passages_total": c["blocks"] * c["passages_per_block"], "passages_per_rank": index.num_passages,
        "tokens_per_rank": index.num_embeddings, "index_bytes_per_rank": index.bytes(), "centroids": C,
        "ms_per_step": ms, "queries_per_s": B / (ms * 1e-3), "doc_tokens_per_s": T3 / (ms * 1e-3),
        "e2e": {"ms_per_step": ms_api, "queries_per_s": B / (ms_api * 1e-3),
                "call": "search_custom_collection(ShardedSearcher, ...) -> Ranking, host embeddings in"},
        "queries_per_chunk": eng.chunk_size(B),
        "candidates_per_query_all_shards": ncand / B, "T1_tokens_per_query_all_shards": T1 / B,
        "T2_tokens_per_query_all_shards": T2 / B, "T3_tokens_per_query_all_shards": T3 / B,
        "stage1_scan_fallback_queries_all_shards": scan_q, "queries_with_results_rank0": found,
        "semantics": "per-shard truncation to ndocs / ndocs/4 (SURVEY.md 8e, oracle (A)): every shard exact-scores its own 256 "
                     "passages per query, so T2/T3 grow with the number of shards while T1 stays the collection's",
        "kernels_rank0": kernels, "exact_global": exact, "index_build_s": round(build_s, 1),
    }


def bench_codec(args, w, peaks, rank, world, local_rank):
    """Index-build codec: argmax over the centroids on tcgen05 (no score table) + residual/bucketize/pack kernel."""
    import torch
    from reranking_multimodal_retrievers_b200 import codec
    from oracle import plaid_oracle as po
    dev = torch.device("cuda", local_rank)
    n, C, nbits = w["n"], w["C"], w["nbits"]
    g = torch.Generator(device=dev)
    g.manual_seed(777 + rank)
    cent = torch.nn.functional.normalize(torch.randn(C, 128, generator=g, device=dev), dim=-1).half()
    assign = torch.randint(0, C, (n,), generator=g, device=dev)
    embs = torch.empty(n, 128, device=dev)
    for i in range(0, n, 1 << 20):
        a = assign[i:i + (1 << 20)]
        embs[i:i + a.numel()] = torch.nn.functional.normalize(
            cent[a].float() + 0.05 * torch.randn(a.numel(), 128, generator=g, device=dev), dim=-1)
    cut = torch.tensor([-0.0304, 0.0, 0.0302], device=dev) if nbits == 2 else torch.linspace(-0.06, 0.06, (1 << nbits) - 1, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, out

    for _ in range(args.warmup):
        codes = codec.compress_into_codes(embs, cent)
        codec.compress_residuals(embs, codes, cent, cut, nbits)
    ms_codes, codes = timed(lambda: codec.compress_into_codes(embs, cent), args.steps)
    ms_res, res = timed(lambda: codec.compress_residuals(embs, codes, cent, cut, nbits), args.steps)
    clocks = sampler.stop_after_min_samples(lambda: codec.compress_residuals(embs, codes, cent, cut, nbits))
    assert float((codes == assign.to(torch.int32)).float().mean()) > 0.999
    m = 4096                                              # parity on a slice: bit-exact vs the oracle given the codes
    _, want = po.codec_compress(cent[:].cpu(), cut.cpu(), nbits, embs[:m].cpu(), codes=codes[:m].cpu())
    assert torch.equal(res[:m].cpu(), want), "compress_residuals differs from the oracle"
    torch.set_num_threads(os.cpu_count() or 1)
    mc = 20000
    t0 = time.perf_counter()
    po.codec_compress(cent.cpu(), cut.cpu(), nbits, embs[:mc].cpu())
    dt = time.perf_counter() - t0
    ms = ms_codes + ms_res
    flops = 2.0 * C * 128 * n
    res_bytes = (512.0 + 4 + 16 * nbits) * n
    return {
        "metric": "compressed_tokens_per_s", "value": world * n / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 argmax (fp32 accumulate), fp32 residual", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "tokens_per_step": n, "centroids": C},
        "e2e": {"value": world * n / (ms * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "index build works on device-resident encoder output"},
        "gpu_launches": None, "clocks": clocks,
        "kernels": {"compress_into_codes": {"ms_per_step": round(ms_codes, 3), "achieved_TFLOPs": round(flops / (ms_codes * 1e-3) / 1e12, 1),
                                            "frac_tensor": round(flops / (ms_codes * 1e-3) / 1e12 / peaks["tf_sustained"], 4)},
                    "compress_residuals": {"ms_per_step": round(ms_res, 3), "achieved_GBps": round(res_bytes / (ms_res * 1e-3) / 1e9, 1),
                                           "frac_hbm": round(res_bytes / (ms_res * 1e-3) / 1e9 / peaks["hbm"], 4)}},
        "roofline": {"kernel": "centroid_scores_kernel (argmax only)", "bound": "tensor", "achieved": round(flops / (ms_codes * 1e-3) / 1e12, 1),
                     "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "peak_source": peaks["source"],
                     "frac": round(flops / (ms_codes * 1e-3) / 1e12 / peaks["tf_sustained"], 4), "traffic": None},
        "cpu_baseline": {"value": mc / dt, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{mc} tokens, torch CPU ops as residual.py:169-222 (fp32 matmul + argmax, bucketize, packbits)"},
    }


def bench_rerank(args, w, peaks, rank, world, local_rank):
    """cfg5: exhaustive padded MaxSim (`colbert_score`, CB/modeling/colbert.py:268-286) over the top-100 candidates
    of 4096 queries, bf16 passage embeddings resident in HBM.  Every rank scores its own copy-sized slab (weak)."""
    import torch
    import reranking_multimodal_retrievers_b200 as pkg
    from oracle import plaid_oracle as po
    dev = torch.device("cuda", local_rank)
    nQ, dpq, Ld, Lq = w["B"], w["dpq"], w["Ld"], w["Lq"]
    n = nQ * dpq
    g = torch.Generator(device=dev)
    g.manual_seed(4321 + rank)
    Q = torch.nn.functional.normalize(torch.randn(nQ, Lq, 128, generator=g, device=dev), dim=-1)
    D = torch.empty(n, Ld, 128, device=dev, dtype=torch.bfloat16)
    for i in range(0, n, 16384):
        D[i:i + 16384] = torch.nn.functional.normalize(
            torch.randn(min(16384, n - i), Ld, 128, generator=g, device=dev), dim=-1).bfloat16()
    lens = torch.randint(w["lo"], w["hi"] + 1, (n,), generator=g, device=dev)
    mask = torch.arange(Ld, device=dev).unsqueeze(0) < lens.unsqueeze(1)
    tokens = int(lens.sum())
    Qhost = Q.cpu().pin_memory()
    out_host = torch.empty(n, dtype=torch.float32).pin_memory()

    def step(q):
        return pkg.colbert_score(q, D, mask, docs_per_query=dpq)

    def step_e2e():
        s = step(Qhost)          # host (pinned) query matrices in: colbert_score copies them in pieces behind the MaxSim
        out_host.copy_(s, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local_rank)
    sampler.start()                                     # before the warm-up (nvidia-smi needs ~100 ms for its first sample)
    for _ in range(args.warmup):
        step(Q)
    ms = timed(lambda: step(Q), args.steps) / args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    clocks = sampler.stop_after_min_samples(lambda: step(Q))
    # check a slice against the CPU restatement of the reference's colbert_score (same bf16 operands)
    got = step(Q)[:dpq].cpu()
    ref = po.colbert_score(Q[:1].cpu().bfloat16().float(), D[:dpq].cpu().float(), mask[:dpq].cpu())
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-3), "padded MaxSim differs from the oracle"
    # CPU baseline: the reference's colbert_score is plain torch (colbert.py:268-286) -> oracle port, all host threads
    torch.set_num_threads(os.cpu_count() or 1)
    nq_cpu = 16
    Qc, Dc, Mc = Q[:nq_cpu].cpu(), D[:nq_cpu * dpq].cpu().float(), mask[:nq_cpu * dpq].cpu()
    t0 = time.perf_counter()
    for i in range(nq_cpu):
        po.colbert_score(Qc[i:i + 1], Dc[i * dpq:(i + 1) * dpq], Mc[i * dpq:(i + 1) * dpq])
    dt = time.perf_counter() - t0
    cpu_tok = int(lens[:nq_cpu * dpq].sum())
    alg_bytes = 2.0 * 128 * tokens
    line = {
        "metric": "scored_doc_tokens_per_s", "value": world * tokens / (ms * 1e-3), "unit": "doc-tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "queries_per_s": nQ / (ms * 1e-3),
        "pairs_per_s": n / (ms * 1e-3),
        "config": {"workload": f"{args.workload}: {w['desc']}", "queries_per_step": nQ, "passages_per_query": dpq,
                   "Ld_padded": Ld, "Lq": Lq, "valid_tokens_per_step": tokens,
                   "l2": "inputs larger than L2 (25 GB of passage embeddings), no explicit flush"},
        "e2e": {"value": world * tokens / (ms_e2e * 1e-3), "unit": "doc-tokens/s", "queries_per_s": nQ / (ms_e2e * 1e-3),
                "ms_per_step": ms_e2e, "h2d_bytes_per_step": nQ * Lq * 128 * 4, "d2h_bytes_per_step": n * 4},
        "gpu_launches": 2 * args.steps, "clocks": clocks,
        "roofline": {"kernel": "maxsim_kernel<2> (padded)", "bound": "hbm", "achieved": round(alg_bytes / (ms * 1e-3) / 1e9, 1),
                     "peak": peaks["hbm"], "unit": "GB/s", "peak_source": peaks["source"] + " (burst: kernel timed alone)",
                     "frac": round(alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm"], 4), "traffic": None,
                     "algorithmic_bytes_per_launch": alg_bytes, "padded_bytes_per_launch": 256.0 * n * Ld,
                     "achieved_TFLOPs": round(2.0 * Lq * 128 * tokens / (ms * 1e-3) / 1e12, 1)},
        "cpu_baseline": {"value": cpu_tok / dt, "unit": "doc-tokens/s", "cores": os.cpu_count(), "kind": "port",
                         "queries_per_s": nq_cpu / dt,
                         "sample": f"{nq_cpu} queries x {dpq} passages of the step, fp32, torch CPU ops as colbert.py:268-286"},
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="materialise bf16 passage embeddings (unfused decompress + MaxSim)")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the 10M-passage strong-scaling block")
    ap.add_argument("--streams", type=int, default=None, help="query chunks in flight on separate streams (engine default if unset)")
    ap.add_argument("--max-chunk", type=int, default=None, help="queries per chunk (engine default 512)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    peaks = load_peaks()

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        sx, Q = build_workload(w, 0, dev)
        sample_q = 8
        cs = make_cpu_searcher(sx)
        tot_q, tot_s, tot_tok, info = 0, 0.0, 0, None
        for step in range(args.warmup + args.steps):
            lo = (step * sample_q) % max(1, Q.shape[0] - sample_q)
            r = run_cpu_sample(cs, Q[lo:lo + sample_q], w["k"], 1e9, sample_q, warm=(step == 0))
            r.pop("results")
            info = r
            if step >= args.warmup:
                tot_q += r["queries"]; tot_s += r["seconds"]; tot_tok += r["tokens"]
        tok_s = tot_tok / max(tot_s, 1e-9)
        line = {
            "impl": "reference", "metric": "scored_doc_tokens_per_s", "value": tok_s, "unit": "doc-tokens/s",
            "queries_per_s": tot_q / max(tot_s, 1e-9), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "step": f"{sample_q}-query sample of the batch"},
            "cpu_baseline": {"value": tok_s, "unit": "doc-tokens/s", "cores": info["cores"], "kind": info["kind"],
                             "kind_detail": info["kind_detail"],
                             "sample": f"{tot_q} queries of the step's {w['B']} (reference CPU path, all host threads)",
                             "queries_per_s": tot_q / max(tot_s, 1e-9), "stage_share": info["stage_share"]},
            "e2e": {"value": tok_s, "unit": "doc-tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if w.get("rerank") or w.get("codec"):
        line = (bench_codec if w.get("codec") else bench_rerank)(args, w, peaks, rank, world, local_rank)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        if rank == 0:
            print(json.dumps(line), flush=True)
        return
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from reranking_multimodal_retrievers_b200 import sharded
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex

    peaks["write_only"] = write_only_peak(dev)          # before the index takes its memory
    sx, Qdev = build_workload(w, rank, dev)
    index = DeviceIndex(sx, dev)
    index.pid_base = rank * w["N"]                      # this rank's shard of the N*world passage collection
    # the reference-facing objects: Searcher over the resident shard, wrapped for the exchange when there are several
    from reranking_multimodal_retrievers_b200 import Searcher, search_custom_collection
    searcher = Searcher(index=index)
    if args.no_fused:
        searcher.ranker.engine.fused = False
    eng = searcher.ranker.engine
    if args.streams:
        eng.streams = args.streams
    if args.max_chunk:
        eng.max_chunk = args.max_chunk
    ss = sharded.ShardedSearcher(searcher) if world > 1 else searcher
    B, Lq, k = w["B"], w["Lq"], w["k"]
    Qhost = Qdev.cpu().pin_memory()
    queries = {i: f"question {i}" for i in range(B)}
    out_host = (torch.empty(B, k, dtype=torch.int32).pin_memory(), torch.empty(B, k, dtype=torch.float32).pin_memory(),
                torch.empty(B, dtype=torch.int32).pin_memory())

    def step(q):
        # Searcher.search_batch: the whole path on one shard; with several ranks + one all-gather of the top-k blocks and
        # the merge kernel (ShardedSearcher).  remove_zero_tensors=True as in the plugin call (src/models/flmr/searching.py:54-61)
        out = ss.search_batch(q, k, True)
        if world > 1:
            eng.launch_count += 1
        return out

    def step_e2e_engine():
        # HOST (pinned) query embeddings in, device lists out + D2H: the engine copies the queries on a copy stream in pieces,
        # ahead of the centroid scoring of the previous piece (H2D of this step's inputs is inside the timed region)
        p, s, c = step(Qhost)
        out_host[0].copy_(p, non_blocking=True); out_host[1].copy_(s, non_blocking=True)
        out_host[2].copy_(c, non_blocking=True)         # D2H of the step's result
        torch.cuda.current_stream().synchronize()

    api_result = [None]

    def step_api():
        # the plugin call of the reference: search_custom_collection -> _search_all_Q -> Ranking
        # (src/models/flmr/searching.py:43-63, CB/searcher.py:80-93), host embeddings in, Ranking object out
        api_result[0] = search_custom_collection(ss, queries, Qhost, num_document_to_retrieve=k, remove_zero_tensors=True)
        if world > 1:
            eng.launch_count += 1

    def timed(fn, steps, sampler=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)   # max over ranks
        return float(ms.item()), clocks

    # accounting pass (untimed): tokens entering each stage, summed over the batch
    stats = dict(ncand=0, T1=0, T2=0, T3=0, T3_padded=0, found=0, ivf_visits=0, ivf_pairs=0, ivf_survivors=0, scan_queries=0)

    def account(ws, n):
        dl = index.doclens
        cc = ws["cand_counts"][:n].long()
        m = torch.arange(ws["cand_stride"], device=dev).unsqueeze(0) < cc.unsqueeze(1)
        stats["ncand"] += int(cc.sum())
        stats["T1"] += int(dl[torch.where(m, ws["cand_pids"][:n], 0).long()].mul(m).sum())
        c1 = ws["s1_counts"][:n].long()
        m1 = torch.arange(ws["s1_pids"].shape[1], device=dev).unsqueeze(0) < c1.unsqueeze(1)
        stats["T2"] += int(dl[torch.where(m1, ws["s1_pids"][:n], 0).long()].mul(m1).sum())
        c2 = ws["s2_counts"][:n].long()
        m2 = torch.arange(ws["s2_pids"].shape[1], device=dev).unsqueeze(0) < c2.unsqueeze(1)
        stats["T3"] += int(dl[torch.where(m2, ws["s2_pids"][:n], 0).long()].mul(m2).sum())     # REAL passage tokens
        stats["T3_padded"] += int(ws["tok_offsets"][:n].gather(1, c2.unsqueeze(1)).sum())       # + 32-token alignment rows
        meta = ws["ivf_meta"][:n].long()                # per query: survivors, pairs, scan flag, list entries visited
        use = meta[:, 2] == 0
        stats["ivf_survivors"] += int(meta[use, 0].sum()); stats["ivf_pairs"] += int(meta[use, 1].sum())
        stats["ivf_visits"] += int(meta[use, 3].sum()); stats["scan_queries"] += int((~use).sum())

    acc_p, acc_s, acc_c = eng.search_batch(Qdev, k=k, remove_zero_rows=True, on_chunk=account)
    stats["found"] = int(acc_c.sum())
    torch.cuda.synchronize()
    eng.check_flags()

    sampler = ClockSampler(local_rank)
    sampler.start()                                     # nvidia-smi needs ~100 ms to deliver its first sample: start it
    for _ in range(args.warmup):                        # before the warm-up; it runs through both timed regions
        step(Qdev)
    # timed region 1: inputs resident in HBM, per-kernel CUDA events on the launching stream
    eng.events = []
    l0 = eng.launch_count
    ms_total, _ = timed(lambda: step(Qdev), args.steps)
    launches = eng.launch_count - l0
    events, eng.events = eng.events, None
    stage_ms = {}
    for stage, a, b in events:
        stage_ms.setdefault(stage, []).append(a.elapsed_time(b))
    eng.check_flags()
    # timed region 2: end to end through the reference-facing plugin call (host embeddings in, Ranking out)
    for _ in range(2):
        step_api()
    l1 = eng.launch_count
    ms_e2e, _ = timed(step_api, args.steps)
    launches_api = eng.launch_count - l1
    # timed region 3: the same with the engine-level call (device lists + D2H, no Ranking object)
    for _ in range(2):
        step_e2e_engine()
    ms_e2e_engine, _ = timed(step_e2e_engine, args.steps)
    clocks = sampler.stop_after_min_samples(lambda: step(Qdev))
    # what a consumer pays on top when it turns the whole Ranking into Python tuples (untimed extra, host only)
    t0 = time.perf_counter()
    rk_dict = api_result[0].todict()
    todict_ms = (time.perf_counter() - t0) * 1e3
    assert len(rk_dict) == B and (world > 1 or sum(len(v) for v in rk_dict.values()) == stats["found"])

    def build_line():
        ms_step = ms_total / args.steps
        C, nbits = index.num_centroids, index.nbits
        bc_dev = eng.chunk_size(B, resident=eng.exchange is None)      # the timed region feeds device-resident embeddings
        chunks = (B + bc_dev - 1) // bc_dev
        T1, T2, T3, ncand = stats["T1"], stats["T2"], stats["T3"], stats["ncand"]
        T3p = stats["T3_padded"]
        # ALGORITHMIC bytes / flops per step, per stage
        s_row = 64.0 if eng.s_dtype == torch.float16 else 128.0
        alg = {
            "centroid_scores": dict(bytes=(2.0 if eng.s_dtype == torch.float16 else 4.0) * C * 32 * B + 2.0 * C * 128 * chunks,
                                    flops=2.0 * C * 128 * 32 * B),
            # inverted-file route: every visited IVF entry costs its pid (4 B) + the candidate-bitmap word and
            # the word-prefix count that turn it into a slot (4 + 4 B); every (slot, centroid) pair is written and read back
            # (2 x 8 B) and gathers one S row; one score per candidate goes out.  Queries routed to the token scan instead
            # read 4 B per candidate token.
            "filter_stage1": dict(bytes=12.0 * stats["ivf_visits"] + (16.0 + s_row) * stats["ivf_pairs"] + 4.0 * ncand
                                  + (4.0 * T1 / max(B, 1)) * stats["scan_queries"], flops=0.0,
                                  note="issue / latency bound (warp-level list walks and a shared-memory counting sort); the byte "
                                       "model is the inverted-file route's, not the token scan's"),
            # every token gathers one score row, but a query has only C distinct rows and re-reads are L2 hits: the
            # compulsory HBM traffic is each touched row once
            "filter_stage2": dict(bytes=4.0 * T2 + 16.0 * 1024 * B + s_row * min(T2, float(C) * B), flops=0.0),
            "candidates": dict(bytes=4.0 * ncand + (w["N"] / 8.0) * B * 2, flops=0.0),
            # codes + residual in, fp16 row out, and the fp16 centroid table once per launch (it stays in L2)
            "decompress": dict(bytes=(4.0 + 16 * nbits + 256.0) * T3 + 256.0 * min(T3, float(C) * chunks), flops=0.0),
            "maxsim": dict(bytes=256.0 * T3, flops=2.0 * Lq * 128 * T3,
                           note="reads D right after decompress wrote it: part of it is still in L2"),
            # K4' of SURVEY 8d on REAL passage tokens (the 32-token alignment rows the tiles also carry are not counted)
            "maxsim_fused": dict(bytes=(4.0 + 16 * nbits) * T3, flops=2.0 * Lq * 128 * T3),
        }
        kernels = {}
        for stage, v in stage_ms.items():
            per_step = sum(v) / args.steps
            d = {"ms_per_step": round(per_step, 4), "share": round(per_step / ms_step, 4), "launches_per_step": len(v) // args.steps}
            if stage in alg:
                d["achieved_GBps"] = round(alg[stage]["bytes"] / (per_step * 1e-3) / 1e9, 1)
                d["frac_hbm"] = round(d["achieved_GBps"] / peaks["hbm"], 4)
                if alg[stage]["flops"]:
                    d["achieved_TFLOPs"] = round(alg[stage]["flops"] / (per_step * 1e-3) / 1e12, 2)
                    d["frac_tensor"] = round(d["achieved_TFLOPs"] / peaks["tf_sustained"], 4)
                if "note" in alg[stage]:
                    d["note"] = alg[stage]["note"]
            kernels[stage] = d
        if "centroid_scores" in kernels:
            # centroid_scores only WRITES (4.2 GB of table per step on cfg2); MEASURED_PEAKS' figure is a copy (read + write).
            # A write-only stream has its own, lower ceiling: measured here with a device fill of 4 GB.
            wp = peaks["write_only"]
            kernels["centroid_scores"].update(
                write_only_peak_GBps=round(wp, 1), frac_write_only_peak=round(kernels["centroid_scores"]["achieved_GBps"] / wp, 4),
                note="write-only kernel: frac_hbm is against the copy (read+write) peak of MEASURED_PEAKS.json, "
                     "frac_write_only_peak against a fill of 4 GB timed in this run")
        dom = max((s for s in kernels if s in alg), key=lambda s: kernels[s]["ms_per_step"])
        nl = kernels[dom]["launches_per_step"]
        per_launch_s = kernels[dom]["ms_per_step"] * 1e-3 / nl
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")        # dram bytes per launch from `ncu --set full`
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload, {}).get(dom)
        if dom == "maxsim_fused":    # intensity 2*Lq*128/(4+16*nbits) flop/B >> ridge: tensor roofline
            roofline = {"kernel": dom, "bound": "tensor", "achieved": round(alg[dom]["flops"] / nl / per_launch_s / 1e12, 2),
                        "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                        "peak_source": peaks["source"] + " (sustained: timed inside the step)", "traffic": traffic,
                        "algorithmic_flops_per_launch": alg[dom]["flops"] / nl, "avg_launch_ms": round(per_launch_s * 1e3, 4)}
        else:
            roofline = {"kernel": dom, "bound": "hbm", "achieved": round(alg[dom]["bytes"] / nl / per_launch_s / 1e9, 1),
                        "peak": peaks["hbm"], "unit": "GB/s",
                        "peak_source": peaks["source"] + " (sustained: timed inside the step)", "traffic": traffic,
                        "algorithmic_bytes_per_launch": alg[dom]["bytes"] / nl, "avg_launch_ms": round(per_launch_s * 1e3, 4)}
        roofline["frac"] = round(roofline["achieved"] / roofline["peak"], 4)

        tok_s = world * T3 / (ms_step * 1e-3)               # every rank exact-scores ~T3 (real) tokens of its own shard
        e2e_ms = ms_e2e / args.steps
        e2e_engine_ms = ms_e2e_engine / args.steps
        line = {
            "metric": "scored_doc_tokens_per_s", "value": tok_s, "unit": "doc-tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 (centroid contraction) / fp16 (MaxSim operands), fp32 accumulate",
            "data": "synthetic", "queries_per_s": B / (ms_step * 1e-3),
            "config": {"workload": f"{args.workload}: {w['desc']}", "passages_per_gpu": w["N"], "tokens_per_gpu": index.num_embeddings,
                       "centroids": C, "queries_per_step": B, "parallelism": f"pid-range shards x{world}, all-gather top-k merge",
                       "queries_per_chunk": bc_dev,
                       "queries_per_chunk_host_fed": (bc_dev if (eng.host_piece and eng.exchange is None and bc_dev >= B and world == 1)
                                                      else eng.chunk_size(B)),
                       "host_fed_pcie_piece_queries": eng.host_piece if world == 1 else None, "chunk_streams": eng.streams,
                       "l2": "inputs larger than L2 (index + centroid-score table >> 126 MB), no explicit flush",
                       "candidates_per_query": ncand / B, "T1_tokens_per_query": T1 / B, "T2_tokens_per_query": T2 / B,
                       "T3_tokens_per_query": T3 / B, "T3_padded_tokens_per_query": T3p / B,
                       "token_accounting": "T1/T2/T3 count real passage tokens (doclens of the listed pids), as the reference arm does; "
                                           "T3_padded adds the 32-token alignment rows of the MaxSim tiles and is not used in any rate",
                       "results_per_query": stats["found"] / B},
            "e2e": {"value": world * T3 / (e2e_ms * 1e-3), "unit": "doc-tokens/s", "queries_per_s": B / (e2e_ms * 1e-3),
                    "ms_per_step": e2e_ms,
                    # several ranks: each copies its 1/N slice of the batch from the host, one NVLink all-gather replicates it
                    "h2d_bytes_per_step": (-(-B // world) if world > 1 else B) * Lq * 128 * 4,
                    "nvlink_allgather_bytes_per_step": (B * Lq * 128 * 4 if world > 1 else 0),
                    "d2h_bytes_per_step": B * k * 8 + B * 4,
                    "call": "search_custom_collection(searcher, queries, Q_host, k, remove_zero_tensors=True) -> Ranking "
                            "(src/models/flmr/searching.py:43-63 -> CB/searcher.py:80-93); pinned host embeddings in, Ranking over "
                            "host arrays out, rows become Python tuples on access",
                    "gpu_launches": launches_api,
                    "ranking_todict_ms": round(todict_ms, 3),
                    "ranking_todict_note": "host-only cost of turning all B x k results into Python tuples (Ranking.todict()), "
                                           "outside the timed region; the reference pays the same per-query tolist/zip inside its loop",
                    "engine_level": {"ms_per_step": e2e_engine_ms, "queries_per_s": B / (e2e_engine_ms * 1e-3),
                                     "value": world * T3 / (e2e_engine_ms * 1e-3),
                                     "call": "Searcher.search_batch(Q_host) + D2H of (pids, scores, counts) into pinned buffers"}},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu_sample(make_cpu_searcher(sx), Qdev, k, args.cpu_budget_s, 64)
            line["agreement"] = agreement((acc_p, acc_s, acc_c), r["results"], k)
            line["cpu_baseline"] = {"value": r["tokens"] / r["seconds"], "unit": "doc-tokens/s", "cores": r["cores"],
                                    "kind": r["kind"], "kind_detail": r["kind_detail"], "queries_per_s": r["queries"] / r["seconds"],
                                    "sample": f"first {r['queries']} of the step's {B} queries, same index, {r['seconds']:.1f} s",
                                    "stage_share": r["stage_share"]}
        return line

    line = build_line() if rank == 0 else None
    # BASELINE.json configs[3] in the same run: the fixed 10M-passage collection sharded over the ranks (strong scaling)
    if args.workload == "cfg2" and not args.no_cfg4:
        del searcher, ss, eng, index, sx, rk_dict, acc_p, acc_s, acc_c, Qdev, Qhost, build_line, step, step_api, step_e2e_engine, account
        api_result[0] = None
        torch.cuda.empty_cache()
        try:
            cfg4 = bench_cfg4(args, rank, world, dev, peaks)
        except Exception as e:          # the headline line must still come out
            cfg4 = {"failed": f"{type(e).__name__}: {e}"[:300]}
        if rank == 0:
            line["cfg4"] = cfg4
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
