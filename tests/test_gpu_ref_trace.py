"""The reference's own query functions over the shim, by trace replay.

tests/golden/ref_trace.npz holds every index call that tasks.voyager_manager.find_nearest_neighbors_by_vector /
find_nearest_neighbors_by_id (standard branch and radius walk) / get_max_distance_for_id and
tasks.clap_text_search.search_by_text made when they ran UNMODIFIED over a recording brute-force index
(tests/golden/make_ref_trace.py; 973 calls: k = n + 4n, n + 0.2n, n + max(20, 3n) + 1, k = len(index), get_vector).
Here each call is replayed against audiomuse_ai_b200.voyager_compat.Index on the GPU: identical ids, distances and
vectors for every call means those functions -- deterministic given the index's returns -- give the recorded answers
over the shim too (the reference tree itself is not on the GPU box)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_replay_reference_call_trace(golden_dir):
    from audiomuse_ai_b200 import voyager_compat as vc
    from tests.golden import make_ref_trace as gen
    tr = np.load(os.path.join(golden_dir, "ref_trace.npz"))
    with open(os.path.join(golden_dir, "ref_trace.json")) as f:
        meta = json.load(f)
    libs = {"music": gen.music_library(), "clap": gen.clap_library()}
    idx = {}
    for name, rows in libs.items():
        idx[name] = vc.Index(vc.Space.Cosine, num_dimensions=rows.shape[1], M=64, ef_construction=1024)
        idx[name].add_items(rows, ids=np.arange(len(rows)))
        idx[name].ef = 1024
    assert int(tr["n_calls"]) == len(meta["calls"]) >= 900
    n_query = n_vec = 0
    ks = set()
    for i, c in enumerate(meta["calls"]):
        index = idx[c["index"]]
        if c["op"] == "get_vector":
            np.testing.assert_allclose(index.get_vector(c["id"]), tr[f"vec_{i}"], atol=1.2e-7, err_msg=str(c))
            n_vec += 1
        else:
            ids, dist = index.query(tr[f"vec_{i}"], k=c["k"])
            np.testing.assert_array_equal(np.asarray(ids, dtype=np.int64), tr[f"ids_{i}"], err_msg=str(c))
            np.testing.assert_allclose(dist, tr[f"dist_{i}"], atol=2.4e-7, err_msg=str(c))
            n_query += 1
            ks.add(c["k"])
    print(f"[ref trace] replayed {n_query} queries (k in {sorted(ks)}) and {n_vec} get_vector calls over "
          f"{sorted(set(c['scenario'] for c in meta['calls']))}")
    assert len(libs["music"]) in ks and 500 in ks      # the k = len(index) scan and the n + 4n expansion are in the trace
