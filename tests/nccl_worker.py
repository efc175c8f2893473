"""Worker of tests/test_gpu_multirank.py (launched by torchrun, one rank per GPU): pid-range sharded search over NCCL.

Every rank searches its shard through the product path (ShardedSearcher: per-shard pipeline, the final top-k written
into the all-gather send block, one all_gather_into_tensor, merge kernel reading the receive buffer in place) and
ALSO runs the oracle on its shard with the shard's own centroid-score table injected (oracle (A)).
Rank 0 gathers the oracle's per-shard lists, merges them with the oracle's (score, pid) selection and compares.
The exact-global mode (stage lists exchanged too) is checked against the search of the unsharded index."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main(out_path):
    from oracle import plaid_oracle as po
    import reranking_multimodal_retrievers_b200 as pkg
    from reranking_multimodal_retrievers_b200 import sharded, synthetic
    from reranking_multimodal_retrievers_b200.index import shard_bounds, slice_host_index
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    try:
        sx = synthetic.make_synthetic_index(4000, 10, 60, 2, seed=77, num_centroids=1024, mode="codes")   # same on every rank
        Q = synthetic.make_queries(sx, 24, 64, seed=78)
        k, ndocs = 20, 128
        p0, p1 = shard_bounds(sx.num_passages, world, rank)
        host = slice_host_index(sx, p0, p1)
        searcher = pkg.Searcher(index=host, device=dev)
        searcher.configure(ndocs=ndocs)
        ss = sharded.ShardedSearcher(searcher)
        mp, ms, mc = ss.search_batch(Q, k, False)
        mp, ms, mc = mp.clone(), ms.clone(), mc.clone()
        rk = pkg.search_custom_collection(ss, {i: "q" for i in range(Q.shape[0])}, Q, num_document_to_retrieve=k,
                                          remove_zero_tensors=False)
        # the shard alone, with taps, for the injected-table oracle
        eng = searcher.ranker.engine
        lp, ls, lc = eng.search_batch(Q, k=k, ndocs=ndocs, keep_taps=True)
        eng.check_flags()
        t = eng.last_taps
        ix = po.OracleIndex(centroids=host.centroids, bucket_weights=host.bucket_weights, codes=host.codes,
                            residuals=host.residuals, doclens=host.doclens, ivf=searcher.ranker.index.ivf_pids.cpu(),
                            ivf_lengths=searcher.ranker.index.ivf_lengths.cpu(), nbits=2)
        mine = []
        for b in range(Q.shape[0]):
            S = t.S[b].float().cpu().contiguous()
            r = po.rank(ix, Q[b], 2, 0.45, ndocs, S_override=S, taps=True)
            n2 = int(t.stage2_counts[b])
            assert torch.equal(t.stage2_pids[b, :n2].cpu(), r["stage2_pids"]), "shard stage-2 pids differ from the oracle"
            sc = t.scores[b, :n2].cpu()
            assert ((sc - r["scores_unsorted"]).abs() <= 1e-3 * r["scores_unsorted"].abs() + 1e-5).all()
            op, os_ = po.select_top(r["stage2_pids"], sc, k)          # the oracle's order on OUR scores (exact compare)
            assert torch.equal(lp[b, :int(lc[b])].cpu(), op + p0) and torch.equal(ls[b, :int(lc[b])].cpu(), os_)
            mine.append((op + p0, os_))
        # exact-global mode (oracle (B)): the stage lists are exchanged as well, and the sharded search must
        # return EXACTLY what one index holding the whole collection returns -- pids, scores and counts, bit for bit
        whole = pkg.Searcher(index=sx, device=dev)
        whole.configure(ndocs=ndocs)
        wp_, ws_, wc_ = whole.search_batch(Q, k, False)
        wp_, ws_, wc_ = wp_.clone(), ws_.clone(), wc_.clone()
        searcher_x = pkg.Searcher(index=host, device=dev)
        searcher_x.configure(ndocs=ndocs)
        sx_exact = sharded.ShardedSearcher(searcher_x, mode="exact")
        for max_chunk in (512, 8):                      # one chunk, and three (the last one ragged)
            searcher_x.ranker.engine.max_chunk = max_chunk
            ep, es, ec = sx_exact.search_batch(Q, k, False)
            assert torch.equal(ec, wc_) and torch.equal(ep, wp_) and torch.equal(es, ws_), "exact-global sharded search differs from the single index"
        searcher_x.ranker.engine.check_flags()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            for b in range(Q.shape[0]):
                allp = torch.cat([gathered[r][b][0] for r in range(world)])
                alls = torch.cat([gathered[r][b][1] for r in range(world)])
                wp, ws = po.select_top(allp, alls, k)
                n = int(mc[b])
                assert n == wp.numel()
                assert torch.equal(mp[b, :n].cpu(), wp) and torch.equal(ms[b, :n].cpu(), ws), f"merged list of query {b} differs"
                assert [e[0] for e in rk.data[b]] == wp.tolist()
            assert len({int(x) * world // sx.num_passages for x in mp[:, :k].flatten().tolist() if x >= 0}) == world   # every shard contributes
            with open(out_path, "w") as f:
                f.write(f"ok world={world} queries={Q.shape[0]}\n")
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
