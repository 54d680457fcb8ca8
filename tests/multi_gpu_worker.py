"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.  Checks, on every rank:
  1. all_gather_embeddings -> index built from the gathered DEVICE buffer answers like a single-GPU index;
  2. sharded_knn_query (replicated library, queries round-robin over ranks) == the full single-GPU answer;
  3. kmeans_lloyd_sharded on row shards == am_kmeans_fit on the whole matrix from the same initial centres."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiomuse_ai_b200 import clustering_gpu as cg, corpus, dist as amdist, voyager_compat as vc  # noqa: E402

rank, local, world = amdist.init_process_group("nccl")
dev = torch.device("cuda", local)

# 1 + 2: k-NN
lib = corpus.knn_library(20000, 512, 1234)
lo, hi = amdist.shard_bounds(len(lib), rank, world)
full = amdist.all_gather_embeddings(torch.from_numpy(lib[lo:hi]).to(dev), len(lib))
assert torch.equal(full.cpu(), torch.from_numpy(lib))
idx = vc.Index.from_device(full, vc.Space.Cosine)
q = corpus.knn_queries(lib, 301, 50, 4321)
ids, dd = amdist.sharded_knn_query(idx, q, 50)
ref = vc.Index(vc.Space.Cosine, num_dimensions=512)
ref.add_items(lib)
ids_ref, dd_ref = ref.query(q, 50)
assert np.array_equal(ids, np.asarray(ids_ref, dtype=np.int64)), "sharded query ids differ"
assert np.allclose(dd, dd_ref, atol=1e-6)

# 3: k-means
x, _, _ = corpus.kmeans_library(60000, 128, 24, 5)
init = x[np.random.default_rng(1).choice(len(x), 24, replace=False)]
lo, hi = amdist.shard_bounds(len(x), rank, world)
xs, initd = torch.from_numpy(x[lo:hi]).to(dev), torch.from_numpy(init).to(dev)
# the same 12 Lloyd iterations on both sides (tol = 0): centres agree to rounding (+ the odd near-tie row)
c, lab, inertia, it = amdist.kmeans_lloyd_sharded(xs, initd, max_iter=12, tol=0.0)
c_ref, lab_ref, inertia_ref, it_ref = cg.kmeans_fit(x, 24, init_centers=init, max_iter=12, tol=0.0)
print(f"MULTI_DIAG rank {rank} fixed-12: inertia {inertia} vs {inertia_ref}, labels equal "
      f"{(lab.cpu().numpy() == lab_ref[lo:hi]).mean():.6f}, max |dc| {np.abs(c.cpu().numpy() - c_ref).max():.3e}, it {it}/{it_ref}", flush=True)
assert abs(inertia - inertia_ref) <= 1e-4 * inertia_ref, (inertia, inertia_ref)
assert (lab.cpu().numpy() == lab_ref[lo:hi]).mean() > 0.9995
assert np.abs(c.cpu().numpy() - c_ref).max() <= 5e-4, np.abs(c.cpu().numpy() - c_ref).max()
tm = {}
c, lab, inertia, it = amdist.kmeans_lloyd_sharded(xs, initd, timing=tm)
c_ref, lab_ref, inertia_ref, it_ref = cg.kmeans_fit(x, 24, init_centers=init)
print(f"MULTI_DIAG rank {rank} tol-1e-4: inertia {inertia} vs {inertia_ref}, labels equal "
      f"{(lab.cpu().numpy() == lab_ref[lo:hi]).mean():.6f}, max |dc| {np.abs(c.cpu().numpy() - c_ref).max():.3e}, it {it}/{it_ref}", flush=True)
assert abs(inertia - inertia_ref) <= 1e-3 * inertia_ref, (inertia, inertia_ref)
assert (lab.cpu().numpy() == lab_ref[lo:hi]).mean() > 0.999
# the all-reduce sums the partial sums in another order than one GPU does: the stopping test (shift <= tol * var) can
# fire a few iterations apart (late iterations move one or two near-tie rows, shift^2 hovers at tol * var), so the centres agree to the tolerance's scale, not to rounding
assert it < 300 and it_ref < 300, (it, it_ref)   # both converged
assert np.abs(c.cpu().numpy() - c_ref).max() <= 2e-3, np.abs(c.cpu().numpy() - c_ref).max()
torch.distributed.barrier()
print(f"MULTI_OK rank {rank}/{world} iters {it} assign_ms {tm.get('assign_ms', 0):.2f} allreduce_ms {tm.get('allreduce_ms', 0):.2f}", flush=True)
torch.distributed.destroy_process_group()
