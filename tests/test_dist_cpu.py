"""N>1 host logic on CPU: world_size-2 gloo processes exercise sharding + the embedding all-gather."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
from audiomuse_ai_b200 import dist as d
rank, local, world = d.init_process_group("gloo")
n_total, dim = 11, 4
lo, hi = d.shard_bounds(n_total, rank, world)
full = np.arange(n_total * dim, dtype=np.float32).reshape(n_total, dim)
got = d.all_gather_embeddings(torch.from_numpy(full[lo:hi].copy()), n_total).numpy()
assert np.array_equal(got, full), (rank, got)
assert d.max_over_ranks(float(rank)) == world - 1
assert d.sum_over_ranks(1.0) == world
t = torch.ones(3) * (rank + 1)
d.all_reduce_sum_(t)
assert float(t[0]) == sum(range(1, world + 1))
print("RANK_OK", rank)
"""


def test_shard_bounds_cover_everything():
    from audiomuse_ai_b200 import dist as d
    for n in (0, 1, 7, 8, 100_000, 1_000_003):
        for w in (1, 2, 3, 4, 8):
            edges = [d.shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == d.padded_shard_len(n, w) or n % w == 0


def test_two_rank_gloo_allgather():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % ROOT], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for r, (p, (so, se)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, se[-2000:]
        assert f"RANK_OK {r}" in so
