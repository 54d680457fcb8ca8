"""Drop-in for the ``voyager`` module surface the reference uses (voyager==2.1.0, Spotify HNSW):

    voyager.Index(space, num_dimensions, M=, ef_construction=)   voyager_manager.py:341-346,
    index.add_items(vectors, ids=)                                clap_text_search.py:242,263
    index.query(vector, k) -> (ids, distances)                    voyager_manager.py:1447,1580,1681
    index.get_vector(id), len(index), index.num_elements, index.ef
    index.save(path | file), Index.load(file)                     voyager_manager.py:183,375-384
    voyager.Space.{Cosine, Euclidean, InnerProduct}, voyager.RecallError

Instead of an HNSW graph the index is the flat [N, d] matrix in HBM and ``query`` is the exact
brute-force top-k of libaudiomuse_b200 (am_knn_*): same return convention (ascending distance,
distance = 1 - cos for Cosine), exact instead of approximate, and ``query`` also accepts a
[nq, d] batch (voyager does too).  M / ef_construction / ef are accepted and ignored.

Monkey-patch point in the reference: ``sys.modules['voyager'] = audiomuse_ai_b200.voyager_compat``
before importing tasks.voyager_manager / tasks.clap_text_search.
"""
from __future__ import annotations

import ctypes as C
import enum
import io
import struct
import threading
from typing import Optional

import numpy as np

from . import _lib


class Space(enum.IntEnum):
    Euclidean = 0
    InnerProduct = 1
    Cosine = 2


class StorageDataType(enum.IntEnum):
    Float8 = 16
    Float32 = 32
    E4M3 = 48


class RecallError(RuntimeError):
    """Raised when fewer than k neighbours can be returned (voyager.RecallError)."""


_METRIC = {Space.Cosine: 0, Space.Euclidean: 1, Space.InnerProduct: 2}
_MAGIC = b"AMIX"


class Index:
    def __init__(self, space: Space = Space.Cosine, num_dimensions: int = 0, M: int = 12,
                 ef_construction: int = 200, random_seed: int = 1, max_elements: int = 1,
                 storage_data_type: StorageDataType = StorageDataType.Float32):
        if num_dimensions <= 0:
            raise ValueError("num_dimensions must be positive")
        self.space = Space(space)
        self.num_dimensions = int(num_dimensions)
        self.M = M
        self.ef_construction = ef_construction
        self.ef = 10
        self._rows = np.zeros((0, self.num_dimensions), dtype=np.float32)
        self._ids = np.zeros((0,), dtype=np.int64)
        self._identity_ids = True
        self._id_to_row = {}
        self._handle: Optional[C.c_void_p] = None
        self._dirty = False
        self._mu = threading.RLock()

    # ------------------------------------------------------------------ building
    def add_items(self, vectors, ids=None, num_threads: int = -1):
        """Appends rows; an id that is already present has its vector REPLACED in place (voyager / hnswlib update
        the stored vector of an existing label; appending a second row would return the id twice from query)."""
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        if v.ndim == 1:
            v = v[np.newaxis, :]
        if v.ndim != 2 or v.shape[1] != self.num_dimensions:
            raise ValueError(f"expected vectors of dimension {self.num_dimensions}, got {v.shape}")
        with self._mu:
            self._materialise_rows()
            start = len(self._ids)
            if ids is None:
                nxt = start if self._identity_ids else (int(self._ids.max()) + 1 if start else 0)
                new_ids = np.arange(nxt, nxt + len(v), dtype=np.int64)
            else:
                new_ids = np.asarray(list(ids), dtype=np.int64)
            if len(new_ids) != len(v):
                raise ValueError("ids and vectors differ in length")
            fresh = np.ones(len(v), dtype=bool)
            if start and ids is not None:
                for a, i in enumerate(new_ids):
                    row = self._lookup(int(i))
                    if row >= 0:
                        self._rows[row] = v[a]
                        fresh[a] = False
            if len(new_ids) > 1 and ids is not None:  # duplicates inside this call: the last one wins
                last = {}
                for a, i in enumerate(new_ids):
                    if fresh[a]:
                        if int(i) in last:
                            fresh[last[int(i)]] = False
                        last[int(i)] = a
            if fresh.any():
                base = len(self._ids)
                self._rows = np.concatenate([self._rows, v[fresh]], axis=0)
                self._ids = np.concatenate([self._ids, new_ids[fresh]])
                self._identity_ids = bool(self._identity_ids and np.array_equal(new_ids[fresh], np.arange(base, base + int(fresh.sum()))))
                if not self._identity_ids:
                    self._id_to_row = {int(i): r for r, i in enumerate(self._ids)}
            self._dirty = True
        return [int(i) for i in new_ids]

    def add_item(self, vector, id=None):
        return self.add_items(np.asarray(vector, dtype=np.float32)[np.newaxis, :],
                              None if id is None else [id])[0]

    @classmethod
    def from_device(cls, rows_dev, space: Space = Space.Cosine) -> "Index":
        """Index over rows already in HBM (a torch.cuda f32[N, d] tensor, e.g. the all-gathered embedding matrix):
        am_knn_build_dev takes a device-to-device copy (cosine rows are stored unit-normalised, so the index owns its
        rows; 0.06 ms per 100 k x 512) -- no host round trip.  Ids are 0..N-1."""
        import torch

        rows_dev = rows_dev.contiguous()
        n, d = rows_dev.shape
        idx = cls(space, int(d))
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.am_knn_build_dev(C.c_void_p(rows_dev.data_ptr()), int(n), int(d), _METRIC[idx.space],
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(h)))
        torch.cuda.current_stream().synchronize()
        idx._handle = h
        idx._rows = None                       # fetched from the device only if someone asks (save / add_items)
        idx._ids = np.arange(n, dtype=np.int64)
        idx._dirty = False
        return idx

    def _materialise_rows(self):
        if self._rows is None:
            n = len(self._ids)
            rows = np.empty((n, self.num_dimensions), dtype=np.float32)
            if n:
                all_rows = np.arange(n, dtype=np.int64)
                _lib.check(_lib.load().am_knn_get_vectors(self._handle, _lib.ptr(all_rows), n, _lib.ptr(rows)))
            self._rows = rows

    def _ensure_built(self):
        with self._mu:
            if self._handle is not None and not self._dirty:
                return self._handle
            lib = _lib.load()
            self._free()
            h = C.c_void_p()
            rows = np.ascontiguousarray(self._rows, dtype=np.float32)
            _lib.check(lib.am_knn_build(_lib.ptr(rows), rows.shape[0], self.num_dimensions,
                                        _METRIC[self.space], C.byref(h)))
            self._handle = h
            self._dirty = False
            return h

    def _free(self):
        if self._handle is not None:
            _lib.load().am_knn_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    # ------------------------------------------------------------------ inspection
    def __len__(self):
        return int(len(self._ids))

    @property
    def num_elements(self):
        return len(self)

    @property
    def ids(self):
        return [int(i) for i in self._ids]

    def _lookup(self, id: int) -> int:
        """row of `id`, or -1: O(1) (identity ids, else a dict kept beside the id array)"""
        if self._identity_ids:
            return id if 0 <= id < len(self._ids) else -1
        return self._id_to_row.get(id, -1)

    def __contains__(self, id):
        return self._lookup(int(id)) >= 0

    def _row_of(self, id) -> int:
        row = self._lookup(int(id))
        if row < 0:
            raise KeyError(f"id {int(id)} not in index")
        return row

    def get_vector(self, id) -> np.ndarray:
        """The STORED vector: unit-normalised for Space.Cosine, like voyager.  One row gathered on the device and
        copied back (no host mirror of the library)."""
        return self.get_vectors([id])[0]

    def get_vectors(self, ids) -> np.ndarray:
        """Stored vectors of several ids: one device gather + one copy (no host mirror of the library)."""
        rows = np.array([self._row_of(i) for i in ids], dtype=np.int64)
        out = np.empty((len(rows), self.num_dimensions), dtype=np.float32)
        if len(rows):
            h = self._ensure_built()
            _lib.check(_lib.load().am_knn_get_vectors(h, _lib.ptr(rows), len(rows), _lib.ptr(out)))
        return out

    def pairwise_distances(self, ids) -> np.ndarray:
        """f32[n, n] direct distances (voyager_manager.get_direct_distance for this index's metric: cosine / inner
        product 1 - cos, euclidean ||a - b||) between the stored vectors of `ids`; +inf for unknown ids."""
        rows = np.empty((len(ids),), dtype=np.int64)
        for a, i in enumerate(ids):
            try:
                rows[a] = self._row_of(i)
            except Exception:
                rows[a] = -1
        out = np.empty((len(rows), len(rows)), dtype=np.float32)
        if len(rows):
            h = self._ensure_built()
            _lib.check(_lib.load().am_knn_pairwise(h, _lib.ptr(rows), len(rows), _lib.ptr(out)))
        return out

    # ------------------------------------------------------------------ query
    def query(self, vectors, k: int = 1, num_threads: int = -1, query_ef: int = -1, mode: int = 0):
        q = np.ascontiguousarray(vectors, dtype=np.float32)
        single = q.ndim == 1
        if single:
            q = q[np.newaxis, :]
        if q.ndim != 2 or q.shape[1] != self.num_dimensions:
            raise ValueError(f"query must have dimension {self.num_dimensions}, got {q.shape}")
        k = int(k)
        if k > len(self):
            raise RecallError(f"Fewer than expected results were retrieved; only found {len(self)} of {k} "
                              "requested neighbors.")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float32)
        if k > 0 and nq > 0:
            h = self._ensure_built()
            st = _lib.load().am_knn_query_ex(h, _lib.ptr(q), nq, k, int(mode), _lib.ptr(ids), _lib.ptr(dist))
            if st == _lib.AM_ERR_RECALL:
                raise RecallError(_lib.last_error())
            _lib.check(st)
            if not self._identity_ids:
                ids = self._ids[ids]
        ids = ids.astype(np.uint64)
        return (ids[0], dist[0]) if single else (ids, dist)

    # ------------------------------------------------------------------ duplicate filter
    def filter_by_distance(self, ids, threshold: float, lookback: int = 1, batch: int = 50):
        """Device version of voyager_manager._filter_by_distance (voyager_manager.py:526-617) on the vectors
        already in HBM: `ids` is one result list (closest first) or an array [n_lists, n] of them; returns a
        boolean keep mask of the same shape.  Ids that are not in the index are dropped, like the reference
        drops items whose vector is missing.  threshold / lookback: config.DUPLICATE_DISTANCE_* (config.py:550-552);
        batch: BATCH_SIZE_VECTOR_OPS (voyager_manager.py:63)."""
        arr = np.asarray(ids)
        single = arr.ndim == 1
        lists = arr[np.newaxis, :] if single else arr
        rows = np.full(lists.shape, -1, dtype=np.int64)
        for a in range(lists.shape[0]):
            for b in range(lists.shape[1]):
                try:
                    rows[a, b] = self._row_of(int(lists[a, b]))
                except Exception:
                    rows[a, b] = -1
        keep = np.zeros(lists.shape, dtype=np.uint8)
        if lists.size:
            h = self._ensure_built()
            _lib.check(_lib.load().am_knn_filter_by_distance(h, _lib.ptr(rows), int(lists.shape[0]), int(lists.shape[1]),
                                                             float(threshold), int(lookback), int(batch), _lib.ptr(keep)))
        keep = keep.astype(bool)
        return keep[0] if single else keep

    # ------------------------------------------------------------------ persistence
    def as_bytes(self) -> bytes:
        with self._mu:
            self._materialise_rows()
            head = _MAGIC + struct.pack("<IIIQ", 1, int(self.space), self.num_dimensions, len(self._ids))
            return head + self._ids.astype("<i8").tobytes() + self._rows.astype("<f4").tobytes()

    def save(self, output_path_or_file):
        data = self.as_bytes()
        if hasattr(output_path_or_file, "write"):
            output_path_or_file.write(data)
        else:
            with open(output_path_or_file, "wb") as f:
                f.write(data)

    @classmethod
    def load(cls, file_or_path, space: Optional[Space] = None, num_dimensions: Optional[int] = None,
             storage_data_type=None) -> "Index":
        if hasattr(file_or_path, "read"):
            data = file_or_path.read()
        else:
            with open(file_or_path, "rb") as f:
                data = f.read()
        if data[:4] != _MAGIC:
            raise RuntimeError("not an audiomuse-b200 flat index (a voyager HNSW blob cannot be read here: "
                               "rebuild from the embedding table, voyager_manager.build_and_store_voyager_index)")
        ver, sp, d, n = struct.unpack_from("<IIIQ", data, 4)
        if ver != 1:
            raise RuntimeError(f"unsupported flat-index version {ver}")
        off = 4 + struct.calcsize("<IIIQ")
        ids = np.frombuffer(data, dtype="<i8", count=n, offset=off)
        rows = np.frombuffer(data, dtype="<f4", count=n * d, offset=off + 8 * n).reshape(n, d)
        idx = cls(Space(sp), d)
        idx.add_items(rows, ids=ids)
        return idx


def loads(data: bytes) -> Index:
    return Index.load(io.BytesIO(data))
