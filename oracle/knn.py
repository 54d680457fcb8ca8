"""Oracle: exact cosine / inner-product / Euclidean k-NN.  TEST INFRASTRUCTURE ONLY.

Restates the index-object contract the reference uses (a ``voyager.Index`` duck type;
voyager==2.1.0 is a third-party C++ HNSW wheel absent from /root/reference):
  * ``tests/unit/test_clap_text_search.py:11-24`` (DummyVoyagerIndex): the brute-force
    statement of ``query`` -- similarities = E @ q, descending, distance = 1 - sim;
  * ``tasks/voyager_manager.py:99-142``: exact distance helpers (cosine distance in
    [0, 2], zero vector -> +inf, None -> +inf; Euclid = ||a-b||);
  * ``tasks/voyager_manager.py:1447,1580`` / ``clap_text_search.py:493``: call sites,
    ``query(vec, k) -> (ids, distances)`` ascending by distance.
voyager's Cosine space stores unit-normalised vectors and normalises the query
(upstream behaviour), so distance = 1 - cos(x, q).

Ordering rule of this oracle (and of the CUDA index): ascending distance computed in
float64 from the float32 inputs, ties broken by LOWER id first.
Pinned by tests/golden/knn_distance_golden.json (made by importing the reference's
helpers) and the reference's known-answer values.
"""
from __future__ import annotations

import numpy as np

COSINE, EUCLIDEAN, INNER_PRODUCT = 0, 1, 2


def normalize_rows(x):
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    n[n == 0] = 1.0
    return (x / n).astype(np.float32)


def direct_cosine_distance(v1, v2):
    """voyager_manager.py:111-135."""
    if v1 is None or v2 is None:
        return float("inf")
    a = np.asarray(v1).astype(np.float32)
    b = np.asarray(v2).astype(np.float32)
    denom = np.linalg.norm(a) * np.linalg.norm(b)
    if denom == 0:
        return float("inf")
    return 1.0 - float(np.clip(np.dot(a, b) / denom, -1.0, 1.0))


def direct_euclidean_distance(v1, v2):
    """voyager_manager.py:99-108."""
    if v1 is None or v2 is None:
        return float("inf")
    return float(np.linalg.norm(np.asarray(v1).astype(np.float32) - np.asarray(v2).astype(np.float32)))


def filter_by_distance(vectors, order, threshold, lookback=1, metric=COSINE, batch=50):
    """voyager_manager.py:526-617 (_filter_by_distance) with :487-524 (_compute_distance_batch), minus the SQL
    look-ups that only feed the log line.  `order` is the result list (row ids, closest first); returns the kept
    ids in order.  Lists of <= `batch` (BATCH_SIZE_VECTOR_OPS, :63) items compare each item with the last
    `lookback` kept ones (:572-599); longer lists go batch by batch, an item being compared with the lookback
    window as of the batch start plus everything already kept inside its batch (:601-615, :502)."""
    dist = direct_euclidean_distance if metric == EUCLIDEAN else direct_cosine_distance
    vec = lambda i: vectors[i] if 0 <= i < len(vectors) else None  # noqa: E731
    if lookback <= 0:
        return list(order)
    kept = []
    order = list(order)
    if len(order) <= batch:
        for cur in order:
            v = vec(cur)
            if v is None:
                continue
            if not any(dist(v, vec(r)) < threshold for r in kept[-lookback:]):
                kept.append(cur)
        return kept
    for b0 in range(0, len(order), batch):
        window = kept[-lookback:] if kept else []
        batch_kept = []
        for cur in order[b0:b0 + batch]:
            v = vec(cur)
            if v is None:
                continue
            if not any(dist(v, vec(r)) < threshold for r in window + batch_kept):
                batch_kept.append(cur)
        kept.extend(batch_kept)
    return kept


def exact_scores_f64(stored, queries, metric=COSINE):
    """Distances f64[nq, N] from the STORED float32 matrix (unit rows for cosine)."""
    x = np.asarray(stored, dtype=np.float32).astype(np.float64)
    q = np.asarray(queries, dtype=np.float32).astype(np.float64)
    if metric == COSINE:
        qn = np.sqrt((q * q).sum(axis=1, keepdims=True))
        qn[qn == 0] = 1.0
        return 1.0 - (q @ x.T) / qn
    if metric == INNER_PRODUCT:
        return 1.0 - q @ x.T
    d2 = (q * q).sum(1)[:, None] - 2.0 * (q @ x.T) + (x * x).sum(1)[None, :]
    return np.maximum(d2, 0.0)


def topk(stored, queries, k, metric=COSINE):
    """Exact top-k: (ids int64[nq,k], dist f32[nq,k]); ascending distance, lower id first."""
    d = exact_scores_f64(stored, queries, metric)
    n = d.shape[1]
    k = min(k, n)
    ids = np.empty((d.shape[0], k), dtype=np.int64)
    out = np.empty((d.shape[0], k), dtype=np.float32)
    ar = np.arange(n)
    for i in range(d.shape[0]):
        order = np.lexsort((ar, d[i]))[:k]
        ids[i] = order
        out[i] = d[i, order].astype(np.float32)
    return ids, out


class BruteForceIndex:
    """The DummyVoyagerIndex contract (test_clap_text_search.py:11-24) with the stable
    ordering rule above; ``embeddings`` are stored as given (callers pass unit rows)."""

    def __init__(self, embeddings):
        self.embeddings = np.asarray(embeddings, dtype=np.float32)

    def __len__(self):
        return len(self.embeddings)

    def get_vector(self, i):
        return self.embeddings[int(i)]

    def query(self, query_vector, k):
        sims = self.embeddings.astype(np.float64) @ np.asarray(query_vector, dtype=np.float32).astype(np.float64)
        order = np.lexsort((np.arange(len(sims)), -sims))[:k]
        return list(order), (1.0 - sims[order]).astype(np.float32)
