"""N > 1 on real GPUs (skipped on a one-GPU box; the CPU suite covers the same host logic over gloo):
two NCCL ranks -- the gathered library answers k-NN queries sharded over ranks like one GPU does, and the sharded
Lloyd loop (one all-reduce of [k, d] sums + [k] counts per iteration) reaches am_kmeans_fit's fixed point."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_knn_and_kmeans():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:   # pytest abbreviates long assertion messages: print the workers' own traceback
        sys.stderr.write(r.stdout[-4000:] + "\n" + r.stderr[-8000:] + "\n")
    assert r.returncode == 0, "a rank failed (its traceback is in the captured stderr above)"
    assert "MULTI_OK rank 0/2" in r.stdout and "MULTI_OK rank 1/2" in r.stdout
