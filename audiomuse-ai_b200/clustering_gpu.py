"""Mirror of ``tasks/clustering_gpu.py`` (KMeans rows of SURVEY 8(a)) on the B200 library.

    check_gpu_available()                      tasks/clustering_gpu.py:26-79
    GPUKMeans(n_clusters, init, n_init, random_state).fit_predict(X)   :82-148
        -> labels; sets cluster_centers_, labels_, using_gpu, inertia_, n_iter_
    get_clustering_model('kmeans', {'n_clusters': k}, use_gpu)          :338-404

cuML is replaced by am_kmeans_fit (k-means++ D^2 seeding, n_init restarts, Lloyd with sklearn's
tol rule).  The reference silently falls back to sklearn when the GPU path raises (:130-148);
here that fallback exists only when ``B200_ALLOW_SKLEARN_FALLBACK=1`` (default: fail loudly,
so a missing CUDA library can never masquerade as the GPU path).  DBSCAN / PCA / GMM / spectral
are out of scope and are handed to scikit-learn exactly like the reference's
``use_gpu=False`` branch.
"""
from __future__ import annotations

import ctypes as C
import logging
import os

import numpy as np

from . import _lib

logger = logging.getLogger("tasks.clustering_gpu")


def check_gpu_available() -> bool:
    try:
        return _lib.load().am_init(-1) == _lib.AM_OK
    except Exception as e:
        logger.info(f"B200 library unavailable: {e}")
        return False


def kmeans_fit(X, k, n_init=10, max_iter=300, tol=1e-4, seed=0, init_centers=None):
    """-> (centers f32[k,d], labels i32[N], inertia, n_iter) via am_kmeans_fit."""
    lib = _lib.load()
    X = np.ascontiguousarray(X, dtype=np.float32)
    if X.ndim != 2:
        raise ValueError("X must be [N, d]")
    N, d = X.shape
    centers = np.empty((k, d), dtype=np.float32)
    labels = np.empty((N,), dtype=np.int32)
    inertia, n_iter = C.c_float(0), C.c_int(0)
    init = None
    if init_centers is not None:
        init = np.ascontiguousarray(init_centers, dtype=np.float32)
        if init.shape != (k, d):
            raise ValueError(f"init_centers must be ({k}, {d})")
    _lib.check(lib.am_kmeans_fit(_lib.ptr(X), N, d, int(k), int(n_init), int(max_iter), float(tol),
                                 int(seed) & 0xFFFFFFFFFFFFFFFF, None if init is None else _lib.ptr(init),
                                 _lib.ptr(centers), _lib.ptr(labels), C.byref(inertia), C.byref(n_iter)))
    return centers, labels, float(inertia.value), int(n_iter.value)


class GPUKMeans:
    def __init__(self, n_clusters, init="k-means++", n_init=10, random_state=None, max_iter=300, tol=1e-4):
        self.n_clusters = n_clusters
        self.init = init
        self.n_init = n_init
        self.random_state = random_state
        self.max_iter = max_iter
        self.tol = tol
        self.model = None
        self.cluster_centers_ = None
        self.labels_ = None
        self.inertia_ = None
        self.n_iter_ = None
        self.using_gpu = False

    def fit_predict(self, X):
        try:
            init_centers = None if isinstance(self.init, str) else np.asarray(self.init, dtype=np.float32)
            seed = 0 if self.random_state is None else int(self.random_state)
            c, l, inertia, it = kmeans_fit(X, int(self.n_clusters), n_init=int(self.n_init),
                                           max_iter=self.max_iter, tol=self.tol, seed=seed,
                                           init_centers=init_centers)
            self.cluster_centers_, self.labels_, self.inertia_, self.n_iter_ = c, l, inertia, it
            self.using_gpu = True
            logger.debug(f"GPU KMeans completed: {self.n_clusters} clusters")
            return l
        except Exception as e:
            if os.environ.get("B200_ALLOW_SKLEARN_FALLBACK", "0") != "1":
                raise
            logger.warning(f"GPU KMeans failed, falling back to CPU: {e}")
        from sklearn.cluster import KMeans
        self.model = KMeans(n_clusters=self.n_clusters, init=self.init, n_init=self.n_init,
                            random_state=self.random_state)
        labels = self.model.fit_predict(X)
        self.cluster_centers_ = self.model.cluster_centers_
        self.labels_ = labels
        self.using_gpu = False
        return labels

    def fit(self, X):
        self.fit_predict(X)
        return self

    def predict(self, X):
        X = np.asarray(X, dtype=np.float32)
        c = self.cluster_centers_
        d2 = (X * X).sum(1)[:, None] - 2.0 * X @ c.T + (c * c).sum(1)[None, :]
        return d2.argmin(1).astype(np.int32)


def get_clustering_model(method, params, use_gpu=False):
    if method == "kmeans" and use_gpu:
        return GPUKMeans(n_clusters=params["n_clusters"], init="k-means++", n_init=10)
    from sklearn.cluster import DBSCAN, KMeans
    if method == "kmeans":
        return KMeans(n_clusters=params["n_clusters"], init="k-means++", n_init=10)
    if method == "dbscan":
        return DBSCAN(eps=params["eps"], min_samples=params["min_samples"])
    raise ValueError(f"Unsupported clustering method: {method}")
