"""Development aid: where the end-to-end overhead of cfg2 goes (host feed, result handoff, streams)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from reranking_multimodal_retrievers_b200 import Searcher, search_custom_collection
from reranking_multimodal_retrievers_b200.index import DeviceIndex

w = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda", 0)
sx, Qdev = bench.build_workload(w, 0, dev)
searcher = Searcher(index=DeviceIndex(sx, dev))
eng = searcher.ranker.engine
Qhost = Qdev.cpu().pin_memory()
B, k = w["B"], w["k"]
queries = {i: "q" for i in range(B)}


def timed(fn, steps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


for streams in (1, 2):
    for chunk in (512, 256):
        eng.streams, eng.max_chunk = streams, chunk
        r = {}
        r["batch_dev"] = timed(lambda: searcher.search_batch(Qdev, k, True))
        r["batch_dev_sync"] = timed(lambda: (searcher.search_batch(Qdev, k, True), torch.cuda.synchronize()))
        r["batch_host_sync"] = timed(lambda: (searcher.search_batch(Qhost, k, True), torch.cuda.synchronize()))
        r["api_dev"] = timed(lambda: search_custom_collection(searcher, queries, Qdev, k, True))
        r["api_host"] = timed(lambda: search_custom_collection(searcher, queries, Qhost, k, True))
        print(f"streams={streams} chunk={chunk}: " + "  ".join(f"{a}={b:.3f}" for a, b in r.items()), flush=True)
