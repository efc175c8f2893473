"""GPU-box probe: the reference's OWN CUDA codec kernels (oracle/_ref/*_gpu_cpp.so, built from /root/reference by
`python oracle/build_ref.py --gpu`) timed beside ours on the same B200 and the same tokens, CUDA events, after warm-up.
Evidence for DESIGN.md; not part of bench.py, not collected by pytest.  Lives under tests/ because it executes oracle/ (test infrastructure).  Usage: python tests/ref_cuda_ab.py > gpurun_out/x.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref                      # noqa: E402  (test infrastructure: the checker / baseline, never the product)
from oracle import plaid_oracle as po             # noqa: E402
from reranking_multimodal_retrievers_b200 import ops   # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda", 0)
    ref_dec = build_ref.load("decompress_residuals_gpu_cpp").decompress_residuals_cpp
    ref_pack = build_ref.load("packbits_gpu_cpp").packbits_cpp
    g = torch.Generator().manual_seed(1)
    C = 65536
    cent = torch.nn.functional.normalize(torch.randn(C, 128, generator=g), dim=-1).half().to(dev)
    rows = []
    for nbits in (2, 4):
        rbm, lut = po.codec_tables(nbits)
        rbm_d, lut_d = rbm.to(dev), lut.to(dev)
        bw = torch.linspace(-0.05, 0.05, 1 << nbits).half().to(dev)
        for n in (1 << 15, 1 << 22):
            res = torch.randint(0, 256, (n, 16 * nbits), generator=g, dtype=torch.uint8).to(dev)
            codes = torch.randint(0, C, (n,), generator=g, dtype=torch.int32).to(dev)
            a = ref_dec(res, bw, rbm_d, lut_d, codes, cent, 128, nbits)
            b = ops.codec_decompress_residuals(res, bw, rbm_d, lut_d, codes, cent, 128, nbits)
            same = bool(torch.equal(a.view(torch.int16), b.view(torch.int16)))
            t_ref = timed(lambda: ref_dec(res, bw, rbm_d, lut_d, codes, cent, 128, nbits))
            t_our = timed(lambda: ops.codec_decompress_residuals(res, bw, rbm_d, lut_d, codes, cent, 128, nbits))
            gb = n * (4 + 16 * nbits + 256) / 1e9
            rows.append({"op": "decompress_residuals (GPU form)", "nbits": nbits, "tokens": n, "bit_identical": same,
                         "reference_cuda_ms": round(t_ref, 4), "ours_ms": round(t_our, 4), "speedup": round(t_ref / t_our, 2),
                         "reference_GBps": round(gb / t_ref * 1e3, 1), "ours_GBps": round(gb / t_our * 1e3, 1),
                         "note": "both calls include their output allocation (the reference's torch::zeros fill as well)"})
    for n in (1 << 22, 1 << 28):
        flags = torch.randint(0, 2, (n,), generator=g, dtype=torch.uint8).to(dev)
        same = bool(torch.equal(ref_pack(flags), ops.packbits(flags)))
        t_ref = timed(lambda: ref_pack(flags))
        t_our = timed(lambda: ops.packbits(flags))
        gb = n * (1 + 1 / 8) / 1e9
        rows.append({"op": "packbits", "flags": n, "bit_identical": same, "reference_cuda_ms": round(t_ref, 4),
                     "ours_ms": round(t_our, 4), "speedup": round(t_ref / t_our, 2),
                     "reference_GBps": round(gb / t_ref * 1e3, 1), "ours_GBps": round(gb / t_our * 1e3, 1)})
    print(json.dumps({"device": torch.cuda.get_device_name(0), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
