"""PCA and DBSCAN on the device (csrc/cluster_extra.cu) against scikit-learn, the reference's own fallback and therefore
the bar for GPUPCA / GPUDBSCAN (tasks/clustering_gpu.py:151-278)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blobs(n, d, k, seed, spread=0.35, noise=0.03):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((k, d)).astype(np.float32) * 3
    lab = rng.integers(0, k, n)
    x = centres[lab] + spread * rng.standard_normal((n, d)).astype(np.float32)
    m = rng.random(n) < noise
    x[m] = rng.uniform(-8, 8, (int(m.sum()), d)).astype(np.float32)
    return x.astype(np.float32)


@pytest.mark.parametrize("n,d,k", [(3000, 64, 10), (2000, 200, 50), (5000, 512, 32)])
def test_pca_matches_sklearn(n, d, k):
    from sklearn.decomposition import PCA
    from audiomuse_ai_b200 import clustering_gpu as cg
    rng = np.random.default_rng(n + d)
    basis = rng.standard_normal((d, d)).astype(np.float32)
    x = (rng.standard_normal((n, d)).astype(np.float32) * np.linspace(4.0, 0.05, d, dtype=np.float32)) @ basis + 3.0
    ref = PCA(n_components=k, svd_solver="full")
    y_ref = ref.fit_transform(x.astype(np.float64))
    got = cg.GPUPCA(k)
    y = got.fit_transform(x)
    assert got.using_gpu and got.n_components_ == k
    np.testing.assert_allclose(got.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-5, atol=1e-8)
    # components agree up to sign (both sides normalise it: the largest-magnitude coordinate is positive)
    cos = np.abs(np.sum(got.components_ * ref.components_, axis=1))
    assert cos.min() > 1 - 1e-6, cos.min()
    sign = np.sign(np.sum(got.components_ * ref.components_, axis=1))
    scale = np.abs(y_ref).max()
    assert np.abs(y * sign[None, :] - y_ref).max() <= 2e-5 * scale, np.abs(y * sign[None, :] - y_ref).max() / scale
    back = got.inverse_transform(y)
    assert np.abs(back - ref.inverse_transform(y_ref)).max() <= 5e-5 * np.abs(x).max()
    np.testing.assert_allclose(got.transform(x[:17]), y[:17], rtol=0, atol=1e-6 * scale)


def test_pca_variance_fraction_and_factory():
    from sklearn.decomposition import PCA
    from audiomuse_ai_b200 import clustering_gpu as cg
    x = _blobs(4000, 96, 12, 5)
    ref = PCA(n_components=0.9, svd_solver="full").fit(x.astype(np.float64))
    got = cg.get_pca_model(0.9, use_gpu=True)
    got.fit_transform(x)
    assert got.n_components_ == ref.n_components_
    assert isinstance(cg.get_pca_model(8, use_gpu=False), PCA)


@pytest.mark.parametrize("n,d,eps,min_samples", [(4000, 16, 1.9, 5), (6000, 64, 4.2, 8), (3000, 200, 7.6, 4), (2500, 8, 0.9, 10)])
def test_dbscan_labels_equal_sklearn(n, d, eps, min_samples):
    from sklearn.cluster import DBSCAN
    from audiomuse_ai_b200 import clustering_gpu as cg
    x = _blobs(n, d, 9, n + d)
    ref = DBSCAN(eps=eps, min_samples=min_samples, algorithm="brute").fit_predict(x.astype(np.float64))
    model = cg.get_clustering_model("dbscan", {"eps": eps, "min_samples": min_samples}, use_gpu=True)
    got = model.fit_predict(x)
    assert model.using_gpu
    n_ref = len(set(ref.tolist()) - {-1})
    print(f"n={n} d={d}: {n_ref} clusters, {(ref == -1).sum()} noise points")
    assert n_ref >= 2 and (ref == -1).sum() > 0, "the case must have clusters and noise to mean anything"
    assert model.n_clusters_ == n_ref
    assert np.array_equal(got, ref.astype(np.int32))


def test_dbscan_chain_needs_many_propagation_rounds():
    """points on a line, each within eps of its neighbours only: one cluster whose label has to travel the whole chain"""
    from sklearn.cluster import DBSCAN
    from audiomuse_ai_b200 import clustering_gpu as cg
    n = 3000
    x = np.zeros((n, 4), dtype=np.float32)
    x[:, 0] = np.arange(n, dtype=np.float32) * 0.5
    x = x[np.random.default_rng(0).permutation(n)]
    ref = DBSCAN(eps=0.6, min_samples=2, algorithm="brute").fit_predict(x.astype(np.float64))
    got = cg.GPUDBSCAN(0.6, 2).fit_predict(x)
    assert len(set(ref.tolist())) == 1 and np.array_equal(got, ref.astype(np.int32))
