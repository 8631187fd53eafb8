"""Arithmetic of the oracle's distance (the Lacaml ssqr_diff stand-in, ohnsw.ml:899).

Lacaml's own output is unpinned (no reference test, third-party, unversioned); what is pinned
here is that the oracle's fast path equals its stated definition bit for bit, and how far the
two documented summation orders can differ (this bounds the 'ties within 1e-5' tolerance).
"""
import numpy as np
import pytest

from oracle import oracle as O


def _team8_numpy(a, b, dot=False):
    """Independent statement of SUM_TEAM8 in numpy float32 (fma emulated in float64: the product
    of two fp32 is exact in fp64 and one fp64 add of an fp32 is rounded once more to fp32 —
    double rounding can differ from a true fma only when the fp64 sum is a tie, which the
    comparison below tolerates by falling back to the C scalar definition)."""
    pp = np.zeros((2, 8), np.float32)          # accumulator pairs: components x,z -> [0], y,w -> [1]
    for i in range(len(a)):
        t = (i >> 2) & 7
        if dot:
            prod = np.float64(a[i]) * np.float64(b[i])
        else:
            x = np.float32(a[i] - b[i])
            prod = np.float64(x) * np.float64(x)
        pp[i & 1, t] = np.float32(prod + np.float64(pp[i & 1, t]))
    p = (pp[0] + pp[1]).astype(np.float32)
    for m in (4, 2, 1):
        p = np.array([np.float32(p[t] + p[t ^ m]) for t in range(8)], np.float32)
    return float(p[0])


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 7, 31, 32, 33, 96, 100, 128, 200, 784, 960])
@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_ANGULAR, O.METRIC_IP])
def test_avx2_path_equals_scalar_definition(dim, metric):
    rng = np.random.default_rng(dim * 7 + metric)
    for _ in range(50):
        a = (rng.random(dim, dtype=np.float32) * 2 - 1) * np.float32(rng.choice([1, 100, 1e-3]))
        b = (rng.random(dim, dtype=np.float32) * 2 - 1)
        fast = O.work_distance(a, b, metric, O.SUM_TEAM8)
        ref = O.work_distance_scalar(a, b, metric)
        assert np.float32(fast).tobytes() == np.float32(ref).tobytes()


def test_scalar_definition_matches_numpy_statement():
    rng = np.random.default_rng(5)
    mism = 0
    for _ in range(200):
        d = int(rng.integers(1, 140))
        a = rng.random(d, dtype=np.float32) * 2 - 1
        b = rng.random(d, dtype=np.float32) * 2 - 1
        mism += np.float32(O.work_distance_scalar(a, b)) != np.float32(_team8_numpy(a, b))
        mism += np.float32(O.work_distance_scalar(a, b, O.METRIC_IP)) != np.float32(-_team8_numpy(a, b, dot=True))
    assert mism <= 2          # fp64-emulated fma double-rounding ties only


def test_integer_valued_data_is_order_independent():
    # SIFT-like data: integer coordinates in [0,218] -> every partial sum is an exact integer
    # < 2^24, so both summation orders (and the GPU) agree exactly.
    rng = np.random.default_rng(3)
    for _ in range(100):
        a = rng.integers(0, 219, 128).astype(np.float32)
        b = rng.integers(0, 219, 128).astype(np.float32)
        s1 = O.work_distance(a, b, O.METRIC_L2, O.SUM_TEAM8)
        s2 = O.work_distance(a, b, O.METRIC_L2, O.SUM_SEQUENTIAL)
        assert s1 == s2 == float(((a.astype(np.int64) - b.astype(np.int64)) ** 2).sum())


def test_summation_orders_differ_by_rounding_only():
    rng = np.random.default_rng(4)
    worst = 0.0
    for _ in range(500):
        a = rng.random(128, dtype=np.float32) * 2 - 1
        b = rng.random(128, dtype=np.float32) * 2 - 1
        s1 = O.work_distance(a, b, O.METRIC_L2, O.SUM_TEAM8)
        s2 = O.work_distance(a, b, O.METRIC_L2, O.SUM_SEQUENTIAL)
        worst = max(worst, abs(s1 - s2) / s2)
    assert worst < 1e-5       # the tolerance north_star names for summation-order ties


def test_distance_is_double_sqrt_of_fp32_sum():   # ohnsw.ml:899
    rng = np.random.default_rng(6)
    a = rng.random(128, dtype=np.float32)
    b = rng.random(128, dtype=np.float32)
    o = O.VecOracle(128)
    assert o.distance(a, b) == float(np.sqrt(np.float64(np.float32(O.work_distance(a, b)))))


def test_recall_compute_semantics():              # dataset.ml:105-127
    exp = np.array([[1, 2, 3], [1, 2, 3]], np.float32)
    got = np.array([[1, 2, 3.5], [np.nan, 1, 3]], np.float32)
    assert O.recall(exp, got) == pytest.approx((2 / 3 + 2 / 3) / 2)
    with pytest.raises(ValueError, match="unequal shapes"):
        O.recall(exp, got[:, :2])


def test_bruteforce_matches_numpy():              # dataset.ml:15-30
    rng = np.random.default_rng(8)
    X = rng.random((500, 16), dtype=np.float32)
    Q = rng.random((7, 16), dtype=np.float32)
    ids, d = O.bruteforce(X, Q, 5)
    ref = np.sqrt(((X[None].astype(np.float64) - Q[:, None].astype(np.float64)) ** 2).sum(-1))
    assert (ids == np.argsort(ref, axis=1, kind="stable")[:, :5]).all()
    assert np.allclose(d, np.sort(ref, axis=1)[:, :5], rtol=1e-6)
