"""k-means parity probe: GPU Lloyd vs sklearn from identical initial centres on the config-4 subsample (GPU box)."""
import sys, numpy as np
sys.path.insert(0, ".")
import audiomuse_ai_b200 as am
from audiomuse_ai_b200 import clustering_gpu as cg, corpus
from sklearn.cluster import KMeans
x, lab, centers = corpus.kmeans_library(1_000_000, 512, 128, 7)
x = np.ascontiguousarray(x[:100_000])
init = x[np.random.default_rng(3).choice(100_000, 128, replace=False)].copy()
c_g, l_g, in_g, it_g = cg.kmeans_fit(x, 128, n_init=1, max_iter=300, tol=1e-4, seed=0, init_centers=init)
km = KMeans(n_clusters=128, init=init, n_init=1, max_iter=300, tol=1e-4, algorithm="lloyd").fit(x.astype(np.float64))
def inertia(c, l): return float(((x.astype(np.float64) - c.astype(np.float64)[l]) ** 2).sum())
print("gpu reported", in_g, "numpy(gpu c,l)", inertia(c_g, l_g), "iters", it_g)
print("sklearn reported", km.inertia_, "numpy(sk c,l)", inertia(km.cluster_centers_, km.labels_), "iters", km.n_iter_)
print("label agreement", float(np.mean(l_g == km.labels_)), "max center diff", float(np.abs(c_g - km.cluster_centers_).max()))
print("counts gpu min", np.bincount(l_g, minlength=128).min(), "sk min", np.bincount(km.labels_, minlength=128).min())
