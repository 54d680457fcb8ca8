import sys, time, numpy as np
sys.path.insert(0, ".")
from audiomuse_ai_b200 import corpus, voyager_compat as vc
x = corpus.knn_library(100_000, 512, 1234); q = corpus.knn_queries(x, 10_000, 1_000, 4321)
idx = vc.Index(vc.Space.Cosine, num_dimensions=512); idx.add_items(x)
for nq in (4096, 4095, 2048, 4096):
    ts = []
    for r in range(6):
        t0 = time.perf_counter(); idx.query(q[:nq], 50); ts.append((time.perf_counter() - t0) * 1e3)
    print(nq, [round(t, 2) for t in ts], flush=True)
