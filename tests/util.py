"""Shared helpers for the parity tests (seeded inputs; SURVEY.md section 8d seeds)."""
import numpy as np


def draw_levels(n, M, seed=7):
    """The level the reference would draw (lib/ohnsw.ml:781): round_nearest(-ln U * 1/ln M)."""
    u = 1.0 - np.random.default_rng(seed).random(n)          # (0, 1]
    return np.floor(-np.log(u) / np.log(M) + 0.5).astype(np.int32)


def uniform(n, dim, seed):
    """Lacaml.S.Mat.random: uniform [-1, 1) (benchmark/dataset.ml:48)."""
    return (np.random.default_rng(seed).random((n, dim), dtype=np.float32) * 2 - 1)


def grid36(seed=0):
    """test/test.ml:88-101: shuffled 6x6 grid of 2-D points."""
    pts = np.array([[i, j] for i in range(6) for j in range(6)], np.float32)
    return pts[np.random.default_rng(seed).permutation(36)]


def assert_same_results(ids_a, d_a, ids_b, d_b):
    """Bit-exact comparison of two [nq][k] result sets (ids, and distances incl. NaN padding)."""
    assert ids_a.shape == ids_b.shape
    bad = np.nonzero((ids_a != ids_b).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} queries differ, first {bad[:5]}: {ids_a[bad[0]]} vs {ids_b[bad[0]]}"
    assert np.array_equal(d_a.view(np.uint32), d_b.view(np.uint32)), "distances differ bitwise"
