"""Wall time of GPUPCA / GPUDBSCAN (host API, copies included) at clustering-task sizes, next to scikit-learn on the box's
CPU cores.  One JSON line.  python tools/cluster_extra_bench.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiomuse_ai_b200 import clustering_gpu as cg, corpus  # noqa: E402

out = {}
x = corpus.knn_library(100_000, 512, 7)
cg.GPUPCA(50).fit_transform(x[:2000])
t0 = time.perf_counter(); p = cg.GPUPCA(50); y = p.fit_transform(x); out["pca_100k_x_512_k50_s"] = round(time.perf_counter() - t0, 4)
from sklearn.decomposition import PCA  # noqa: E402
t0 = time.perf_counter(); PCA(n_components=50, svd_solver="full").fit_transform(x[:20_000]); out["sklearn_pca_20k_x_512_k50_s"] = round(time.perf_counter() - t0, 4)

rng = np.random.default_rng(1)
centres = rng.standard_normal((40, 64)).astype(np.float32) * 3
xb = (centres[rng.integers(0, 40, 50_000)] + 0.35 * rng.standard_normal((50_000, 64))).astype(np.float32)
cg.GPUDBSCAN(4.2, 8).fit_predict(xb[:2000])
t0 = time.perf_counter(); m = cg.GPUDBSCAN(4.2, 8); lab = m.fit_predict(xb); out["dbscan_50k_x_64_s"] = round(time.perf_counter() - t0, 4)
out["dbscan_clusters"] = int(m.n_clusters_)
from sklearn.cluster import DBSCAN  # noqa: E402
t0 = time.perf_counter(); ref = DBSCAN(eps=4.2, min_samples=8, n_jobs=-1).fit_predict(xb[:10_000]); out["sklearn_dbscan_10k_x_64_s"] = round(time.perf_counter() - t0, 4)
out["labels_equal_on_first_10k_subset"] = bool(np.array_equal(cg.GPUDBSCAN(4.2, 8).fit_predict(xb[:10_000]), ref.astype(np.int32)))
xl = corpus.knn_library(100_000, 512, 9)
t0 = time.perf_counter(); cg.GPUDBSCAN(0.9, 5).fit_predict(xl); out["dbscan_100k_x_512_s"] = round(time.perf_counter() - t0, 4)
print(json.dumps(out))
