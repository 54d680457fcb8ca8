"""Oracle: Lloyd k-means.  TEST INFRASTRUCTURE ONLY.

Restates what ``tasks/clustering_gpu.py:96-148`` asks of its backends
(cuml.cluster.KMeans on GPU, sklearn.cluster.KMeans on CPU; both third-party):
``fit_predict(X) -> labels`` plus ``cluster_centers_``.  The reference has no numeric
test for it (test_clustering_helper.py patches USE_GPU_CLUSTERING False), so the bar
is the sklearn branch itself: same inertia (within 1 %) and label agreement from an
identical initialisation.  ``lloyd`` below is the plain algorithm (float64) used for
small deterministic cases; ``sklearn_fit`` wraps the reference's CPU branch.
"""
from __future__ import annotations

import numpy as np


def assign(X, C):
    X = np.asarray(X, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    d2 = (X * X).sum(1)[:, None] - 2.0 * X @ C.T + (C * C).sum(1)[None, :]
    labels = d2.argmin(1)
    return labels.astype(np.int32), float(np.maximum(d2[np.arange(len(X)), labels], 0).sum())


def lloyd(X, init_centers, max_iter=300, tol=1e-4):
    """Lloyd iterations from fixed initial centers.  Empty clusters keep their center.
    Stops when the squared center shift <= tol * mean feature variance (sklearn's rule)."""
    X = np.asarray(X, dtype=np.float64)
    C = np.array(init_centers, dtype=np.float64)
    thr = tol * X.var(axis=0).mean()
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        labels, _ = assign(X, C)
        newC = C.copy()
        for j in range(len(C)):
            m = labels == j
            if m.any():
                newC[j] = X[m].mean(0)
        shift = ((newC - C) ** 2).sum()
        C = newC
        if shift <= thr:
            break
    labels, inertia = assign(X, C)
    return C, labels, inertia, n_iter


def sklearn_fit(X, k, init, n_init=1, max_iter=300, tol=1e-4, random_state=0):
    """The reference's CPU branch (clustering_gpu.py:135-142) with an explicit init."""
    from sklearn.cluster import KMeans

    km = KMeans(n_clusters=k, init=init, n_init=n_init, max_iter=max_iter, tol=tol,
                random_state=random_state, algorithm="lloyd")
    labels = km.fit_predict(X)
    return km.cluster_centers_, labels.astype(np.int32), float(km.inertia_), km.n_iter_
