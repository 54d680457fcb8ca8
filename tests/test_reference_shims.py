"""The reference's OWN functions running over the shims, in this container (skipped where /root/reference is absent,
i.e. on the GPU box): the index builder / loader pair with `voyager` resolving to voyager_compat
(tasks/voyager_manager.py:145-460: flat AMIX blob in the `voyager_index_data` rows, one row or <name>_<i>_<n> segments
of <= VOYAGER_MAX_PART_SIZE bytes, id_map_json in part 1 only), and the INTEGRATION.md section-3 patch applied to the
stub-imported modules.  Queries need the GPU: tests/test_gpu_ref_trace.py covers them by trace replay."""
import io
import json

import numpy as np
import pytest

from tests import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference is not present (GPU box)")


@pytest.fixture()
def ref():
    from audiomuse_ai_b200 import voyager_compat as vc
    db = rh.FakeDB()
    return rh.load_reference(vc, db), db, vc


def _fill(db, n, d, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    db.embeddings = [(f"item{i}", x[i].tobytes()) for i in range(n)]
    db.embeddings.insert(5, ("broken", None))                               # NULL blob: skipped by the builder (:351)
    db.embeddings.insert(9, ("short", np.zeros(d - 1, np.float32).tobytes()))   # wrong dimension: skipped (:357)
    return x


def test_build_store_load_single_row(ref):
    r, db, vc = ref
    vm, d = r.vm, r.config.EMBEDDING_DIMENSION
    x = _fill(db, 500, d)
    vm.build_and_store_voyager_index(db)
    assert list(db.index_rows) == [r.config.INDEX_NAME] and db.commits == 1
    blob, id_map_json, dim = db.index_rows[r.config.INDEX_NAME]
    assert blob[:4] == b"AMIX" and dim == d and len(json.loads(id_map_json)) == 500
    vm.voyager_index = None
    vm.load_voyager_index_for_querying(force_reload=True)
    assert isinstance(vm.voyager_index, vc.Index) and len(vm.voyager_index) == 500
    assert vm.voyager_index.ef == r.config.VOYAGER_QUERY_EF
    assert vm.id_map[0] == "item0" and vm.reverse_id_map["item499"] == 499 and "broken" not in vm.reverse_id_map
    np.testing.assert_array_equal(vm.voyager_index._rows, x)             # float32 rows survive the round trip bit for bit


def test_build_store_load_segmented_rows(ref):
    """An index larger than VOYAGER_MAX_PART_SIZE is stored as <INDEX_NAME>_<part>_<total> rows (:410-436) and
    reassembled by the loader (:186-283)."""
    r, db, vc = ref
    vm, d = r.vm, r.config.EMBEDDING_DIMENSION
    x = _fill(db, 4000, d)
    vm.VOYAGER_MAX_PART_SIZE = 1 << 20                                    # 1 MiB parts instead of 50 MB
    vm.build_and_store_voyager_index(db)
    names = sorted(db.index_rows, key=lambda s: int(s.split("_")[-2]))
    total = len(names)
    assert total == -(-len(vc.loads(b"".join(db.index_rows[n][0] for n in names)).as_bytes()) // (1 << 20)) >= 3
    assert names == [f"{r.config.INDEX_NAME}_{i}_{total}" for i in range(1, total + 1)]
    assert all(len(db.index_rows[n][0]) <= (1 << 20) for n in names)
    assert db.index_rows[names[0]][1] and all(db.index_rows[n][1] == "" for n in names[1:])
    vm.voyager_index = None
    vm.load_voyager_index_for_querying(force_reload=True)
    assert len(vm.voyager_index) == 4000 == len(vm.id_map)
    np.testing.assert_array_equal(vm.voyager_index._rows, x)
    # a missing segment aborts the load instead of serving a corrupt index (:224-227)
    del db.index_rows[names[1]]
    vm.load_voyager_index_for_querying(force_reload=True)
    assert vm.voyager_index is None


def test_an_old_hnsw_blob_is_refused_and_the_loader_survives(ref):
    r, db, vc = ref
    vm = r.vm
    db.index_rows[r.config.INDEX_NAME] = (b"VOYA" + b"\x00" * 64, json.dumps({"0": "item0"}), r.config.EMBEDDING_DIMENSION)
    vm.load_voyager_index_for_querying(force_reload=True)                  # logs, leaves the cache empty: rebuild path
    assert vm.voyager_index is None
    with pytest.raises(RuntimeError):
        vc.Index.load(io.BytesIO(b"VOYA" + b"\x00" * 64))


def test_not_loaded_errors_match_the_reference_contract(ref):
    """tests/unit/test_voyager_manager.py:420-473 of the reference: querying without a loaded index raises."""
    r, db, vc = ref
    vm = r.vm
    vm.voyager_index = vm.id_map = vm.reverse_id_map = None
    with pytest.raises(RuntimeError):
        vm.find_nearest_neighbors_by_vector(np.zeros(r.config.EMBEDDING_DIMENSION, np.float32))
    with pytest.raises(RuntimeError):
        vm.find_nearest_neighbors_by_id("item0")
    with pytest.raises(RuntimeError):
        vm.get_max_distance_for_id("item0")


def test_integration_patch_applies_to_the_reference_modules(ref):
    import importlib.util
    import os
    import sys
    import types
    from audiomuse_ai_b200 import clap_analyzer as b200_clap, clustering_gpu as b200_cg, integration
    r, db, vc = ref
    integration.install_voyager_shim()
    assert sys.modules["voyager"] is vc
    # the reference's clustering module imports cleanly here (its GPU imports are inside try blocks)
    spec = importlib.util.spec_from_file_location("tasks.clustering_gpu", os.path.join(rh.REF, "tasks", "clustering_gpu.py"))
    ref_cg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_cg)
    ref_clap = types.ModuleType("tasks.clap_analyzer")   # (importing the real one needs librosa / onnxruntime at call time only)
    clap_src = open(os.path.join(rh.REF, "tasks", "clap_analyzer.py")).read()
    for name in integration.CLAP_NAMES:
        assert f"def {name}(" in clap_src, name            # every patched name exists upstream with that spelling
        setattr(ref_clap, name, object())
    old = os.environ.pop("B200_ALLOW_SKLEARN_FALLBACK", None)
    try:
        integration.apply(clap=ref_clap, voyager_manager=r.vm, clustering=ref_cg)
        assert os.environ.get("B200_ALLOW_SKLEARN_FALLBACK") == "1"        # the reference's silent-fallback contract
    finally:
        os.environ.pop("B200_ALLOW_SKLEARN_FALLBACK", None)
        if old is not None:
            os.environ["B200_ALLOW_SKLEARN_FALLBACK"] = old
    assert all(getattr(ref_clap, n) is getattr(b200_clap, n) for n in integration.CLAP_NAMES)
    assert ref_cg.GPUKMeans is b200_cg.GPUKMeans and ref_cg.check_gpu_available is b200_cg.check_gpu_available
    assert r.vm._filter_by_distance.__name__ == "_filter_by_distance_b200"
    # get_clustering_model of the REFERENCE now hands out the B200 class (clustering_gpu.py:338-404)
    m = ref_cg.get_clustering_model("kmeans", {"n_clusters": 7}, use_gpu=True)
    assert isinstance(m, b200_cg.GPUKMeans) and m.n_clusters == 7
    # ... and the B200 DBSCAN / PCA classes (clustering_gpu.py:151-278, 407-421)
    m = ref_cg.get_clustering_model("dbscan", {"eps": 0.5, "min_samples": 4}, use_gpu=True)
    assert isinstance(m, b200_cg.GPUDBSCAN) and m.eps == 0.5 and m.min_samples == 4
    m = ref_cg.get_pca_model(12, use_gpu=True)
    assert isinstance(m, b200_cg.GPUPCA) and m.n_components == 12
