"""The real multi-rank path: torchrun, one rank per GPU, NCCL all-gather of the per-shard top-k blocks, merge kernel.
Skipped when the box has fewer than two GPUs (the single-GPU emulation of the shards and the gloo exchange test cover
the logic there)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_two_rank_nccl_sharded_search_equals_oracle_per_shard_merge(tmp_path):
    world = 2 if torch.cuda.device_count() < 4 else 4
    out = tmp_path / "result.txt"
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py"), str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert out.read_text().startswith(f"ok world={world}")
