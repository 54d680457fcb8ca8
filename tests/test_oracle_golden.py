"""Pin the oracle against golden vectors produced by the reference's own code
(tests/golden/make_golden.py) and the reference's known-answer tests."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import SEGMENT_CASE_LENGTHS, golden_waveform, standin_encoder, standin_mel  # noqa: E402

from oracle import knn, segments


@pytest.fixture(scope="module")
def seg_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "segments_golden.npz"))


@pytest.mark.parametrize("ci", range(len(SEGMENT_CASE_LENGTHS)))
def test_round_trip_segmentation_pooling_match_reference(seg_gold, ci):
    """oracle == reference analyze_audio_file (clap_analyzer.py:467-574) on the same waveform."""
    L = SEGMENT_CASE_LENGTHS[ci]
    audio, _ = segments.int16_round_trip(golden_waveform(100 + ci, L))
    segs = segments.segment_audio(audio)
    assert len(segs) == int(seg_gold[f"nseg_{ci}"]) == int(seg_gold[f"nseg_student_{ci}"])
    assert abs(len(audio) / 48000 - float(seg_gold[f"dur_{ci}"])) == 0
    np.testing.assert_array_equal(np.stack([s[:8] for s in segs]), seg_gold[f"seg_head_{ci}"])
    np.testing.assert_array_equal(np.stack([s[-8:] for s in segs]), seg_gold[f"seg_tail_{ci}"])
    np.testing.assert_array_equal(np.array([s.astype(np.float64).sum() for s in segs]),
                                  seg_gold[f"seg_sum_{ci}"])
    embs = np.vstack([standin_encoder(standin_mel(s)) for s in segs])
    np.testing.assert_array_equal(segments.pool_segments(embs), seg_gold[f"emb_{ci}"])


@pytest.mark.parametrize("ci", range(len(SEGMENT_CASE_LENGTHS)))
def test_segment_starts_match_student_positions(seg_gold, ci):
    L = SEGMENT_CASE_LENGTHS[ci]
    pos = seg_gold[f"positions_{ci}"]
    starts = segments.segment_starts(L)
    assert len(starts) == len(pos)
    if L > segments.SEGMENT_LENGTH:
        assert starts == [int(p[0]) for p in pos]


def test_duplicate_tail_quirk():
    """L - SEG an exact multiple of HOP: the tail window duplicates the last regular one."""
    assert segments.segment_starts(720_000) == [0, 240_000, 240_000]
    assert segments.segment_starts(1_200_000) == [0, 240_000, 480_000, 720_000, 720_000]


def test_int16_truncation():
    x = np.array([0.5, -0.5, 1.5, -1.5, 1e-5, 32766.9 / 32767.0], dtype=np.float32)
    y, q = segments.int16_round_trip(x)
    assert q.tolist() == [16383, -16383, 32767, -32767, 0, 32766]
    assert y[2] == np.float32(1.0)


def test_distance_helpers_match_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "knn_distance_golden.json")))
    assert gold["metric"] == "angular"
    for c in gold["cases"]:
        a = None if c["a"] is None else np.array(c["a"], np.float32)
        b = None if c["b"] is None else np.array(c["b"], np.float32)
        for key, fn in (("cosine", knn.direct_cosine_distance), ("euclidean", knn.direct_euclidean_distance)):
            want = float("inf") if c[key] == "inf" else c[key]
            got = fn(a, b)
            assert got == want or abs(got - want) <= 1e-6, (key, got, want)


def test_known_answers_from_reference_unit_tests():
    """tests/unit/test_voyager_manager.py:23-166."""
    f32 = np.float32
    assert knn.direct_euclidean_distance(np.array([1, 2, 3], f32), np.array([1, 2, 3], f32)) == 0.0
    assert abs(knn.direct_euclidean_distance(np.zeros(3, f32), np.array([3, 4, 0], f32)) - 5.0) < 1e-5
    assert knn.direct_euclidean_distance(None, np.ones(2, f32)) == float("inf")
    assert abs(knn.direct_cosine_distance(np.array([1, 2, 3], f32), np.array([1, 2, 3], f32))) < 1e-5
    assert abs(knn.direct_cosine_distance(np.array([1, 0], f32), np.array([0, 1], f32)) - 1.0) < 1e-5
    assert abs(knn.direct_cosine_distance(np.array([1, 0], f32), np.array([-1, 0], f32)) - 2.0) < 1e-5
    assert knn.direct_cosine_distance(np.zeros(2, f32), np.ones(2, f32)) == float("inf")
    assert abs(knn.direct_cosine_distance(np.ones(2, f32), 10 * np.ones(2, f32))) < 1e-5


def test_bruteforce_index_matches_dummy_voyager_index(golden_dir):
    """oracle.knn == the reference tests' DummyVoyagerIndex on seeded data."""
    g = np.load(os.path.join(golden_dir, "dummy_index_golden.npz"))
    rng = np.random.default_rng(int(g["seed"]))
    E = rng.standard_normal((500, 512)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    Q = rng.standard_normal((8, 512)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    idx = knn.BruteForceIndex(E)
    assert len(idx) == int(g["n"])
    for qi, q in enumerate(Q):
        ids, d = idx.query(q, 50)
        np.testing.assert_array_equal(np.array(ids), g["ids"][qi])
        np.testing.assert_allclose(d, g["dists"][qi], atol=2e-7)
    ids2, d2 = knn.topk(E, Q, 50)
    np.testing.assert_array_equal(ids2, g["ids"])


def test_filter_by_distance_matches_reference(golden_dir):
    """oracle.knn.filter_by_distance vs the kept lists produced by the reference's own
    voyager_manager._filter_by_distance (tests/golden/make_golden.py): lists of 40 / 50 (sequential
    branch, voyager_manager.py:572-599), 51 / 130 / 237 items (batched branch, :601-615), look-back 1 and 3,
    exact and near duplicates, one item with a missing vector."""
    g = np.load(os.path.join(golden_dir, "filter_golden.npz"))
    F, thr, B = g["vectors"], float(g["threshold"]), int(g["batch"])
    for ci in range(int(g["n_cases"])):
        order = [int(i) for i in g[f"order_{ci}"]]
        for lb in (1, 3):
            got = knn.filter_by_distance(F, order, thr, lb, knn.COSINE, B)
            assert got == g[f"kept_{ci}_lb{lb}"].tolist(), (ci, lb)
    assert knn.filter_by_distance(F, [3, 1, 2], thr, 0) == [3, 1, 2]      # look-back 0: unchanged (:531-532)


def test_pcm16_scaling_sequence():
    """The mel kernel scales PCM16 with x * kInv, e = fma(-32767, r0, x), r = fma(e, kInv, r0) instead of a
    division (csrc/mel.cu pcm16_to_f32).  For every int16 value that equals the reference's
    (q / 32767.0).astype(float32) (clap_analyzer.py:505); checked here with exact rational arithmetic."""
    from fractions import Fraction

    def rnd32(fr):  # round-to-nearest-even of an exact rational to float32
        c = np.float32(float(fr))
        best, bd = c, abs(Fraction(float(c)) - fr)
        for cand in (np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(-np.inf))):
            d = abs(Fraction(float(cand)) - fr)
            if d < bd or (d == bd and (int(np.float32(cand).view(np.int32)) & 1) == 0):
                best, bd = np.float32(cand), d
        return np.float32(best)

    q = np.arange(-32768, 32768, dtype=np.int64)
    ref = (q / 32767.0).astype(np.float32)
    kinv = Fraction(float(np.float32(1.0) / np.float32(32767.0)))
    for v in range(-32768, 32768, 7):     # every 7th value here (the exhaustive run is on the GPU, test_gpu_mel.py)
        x = Fraction(v)
        r0 = rnd32(x * kinv)
        e = rnd32(x - 32767 * Fraction(float(r0)))
        r = rnd32(Fraction(float(e)) * kinv + Fraction(float(r0)))
        assert r == ref[v + 32768], v
