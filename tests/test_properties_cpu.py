"""CPU property tests of the oracle and the host solver (hypothesis; bounded example counts)."""
import itertools

import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import pytorch3d_ops as p3d  # noqa: E402
from oracle import ume_oracle as orc  # noqa: E402

FAST = settings(max_examples=25, deadline=None)


@FAST
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 60), st.integers(1, 12), st.integers(1, 70), st.floats(0.05, 3.0))
def test_ball_query_c_equals_numpy_restatement(seed, P2, P1, K, radius):
    rng = np.random.default_rng(seed)
    # a lattice with duplicates: exact ties at the radius and repeated points
    p2 = (rng.integers(-3, 4, size=(1, P2, 3)) * 0.5).astype(np.float32)
    p1 = (rng.integers(-3, 4, size=(1, P1, 3)) * 0.5).astype(np.float32)
    a, b = p3d.ball_query_np(p1, p2, K, radius), p3d.ball_query_c(p1, p2, K, radius)
    assert np.array_equal(a.idx, b.idx) and np.array_equal(a.dists, b.dists) and np.array_equal(a.knn, b.knn)
    # semantics: ascending row order, strictly inside, -1 padding at the end only
    for i in range(P1):
        rows = a.idx[0, i]
        valid = rows[rows >= 0]
        assert np.all(np.diff(valid) > 0) and np.all(rows[len(valid):] == -1)
        d2 = ((p2[0, valid] - p1[0, i]) ** 2).sum(-1)
        assert np.all(d2 < np.float32(radius) * np.float32(radius))


@FAST
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 400), st.floats(0.05, 2.0))
def test_sparse_quantize_properties(seed, N, q):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-3, 3, (N, 3)).astype(np.float32)
    vox, idx = orc.sparse_quantize(pts, q)
    assert np.all(np.diff(idx) > 0) and len(np.unique(vox, axis=0)) == len(vox)
    disc = np.floor(pts / np.float32(q)).astype(np.int32)
    kept = {tuple(v): i for v, i in zip(map(tuple, vox), idx)}
    for i in range(N):                                       # every row's voxel is represented by its FIRST row
        assert kept[tuple(disc[i])] <= i
    assert np.array_equal(orc.sparse_quantize(pts[idx], q)[1], np.arange(len(idx)))     # idempotent


@FAST
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 6), st.integers(1, 6))
def test_assignment_is_optimal_against_brute_force(seed, nr, nc):
    import umeregrobust_b200 as ume
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 6, (nr, nc)).astype(np.float32)      # small integers: many ties
    r, k = ume.linear_sum_assignment(c)
    n = min(nr, nc)
    assert len(r) == n and len(set(k.tolist())) == n and np.all(np.diff(r) > 0)
    if nr <= nc:
        best = min(sum(c[i, p[i]] for i in range(nr)) for p in itertools.permutations(range(nc), nr))
    else:
        best = min(sum(c[p[j], j] for j in range(nc)) for p in itertools.permutations(range(nr), nc))
    assert c[r, k].sum() == best


@FAST
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40), st.floats(0.5, 4.0))
def test_moments_backward_is_the_adjoint_of_the_raw_moments(seed, K, radius):
    # <F_raw(feat), w> == <feat, backward(w)> for every feat, w: the raw moment build is linear in
    # the features and the backward pass is its transpose (fp64 oracle)
    rng = np.random.default_rng(seed)
    N, n, C = 150, 9, 4
    pts = rng.uniform(-4, 4, (1, N, 3)).astype(np.float32)
    kp = pts[:, rng.choice(N, n, replace=False)].copy()
    feat = rng.normal(size=(1, N, C))
    w = rng.normal(size=(1, n, C, 4))
    idx = p3d.ball_query_c(kp, pts, K, radius, return_nn=False).idx
    F = np.zeros((1, n, C, 4))
    for i in range(n):
        rows = idx[0, i][idx[0, i] >= 0]
        F[0, i, :, 0] = feat[0, rows].sum(0)
        F[0, i, :, 1:] = feat[0, rows].T @ pts[0, rows].astype(np.float64)
    g = orc.ume_moments_backward(pts, kp, w, K, radius)
    assert abs((F * w).sum() - (feat * g).sum()) <= 1e-9 * max(1.0, abs((F * w).sum()))
