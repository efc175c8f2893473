import torch, sys
sys.path.insert(0,'.')
from reranking_multimodal_retrievers_b200 import synthetic
from reranking_multimodal_retrievers_b200.engine import SearchEngine
from reranking_multimodal_retrievers_b200.index import DeviceIndex
for (N,lo,hi,C,B,Lq,k,nb) in [(50,1,5,128,3,64,100,2),(7,3,9,128,1,32,10,4),(300,10,60,256,5,96,1000,2),(2000,20,60,1024,9,128,100,2),(2000,20,60,1024,2,160,50,1)]:
    sx = synthetic.make_synthetic_index(N, lo, hi, nb, seed=3, num_centroids=C, mode="codes", device="cuda")
    Q, gold = synthetic.make_queries(sx, B, Lq, seed=4, return_gold=True)
    eng = SearchEngine(DeviceIndex(sx))
    p,s,c = eng.search_batch(Q, k=k)
    torch.cuda.synchronize(); eng.check_flags()
    ok = all(int(p[b,0])==int(gold[b]) for b in range(B))
    print((N,lo,hi,C,B,Lq,k,nb), 'counts', c.tolist(), 'gold top1', ok, 'min score', float(s[s>-1e30].min()))
