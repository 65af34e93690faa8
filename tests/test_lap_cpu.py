"""The host assignment solver behind the Hungarian option (evaluate.py:216-222) against
scipy.optimize.linear_sum_assignment — the function the reference calls there."""
import numpy as np
import pytest

scipy_opt = pytest.importorskip("scipy.optimize")


@pytest.fixture(scope="module")
def ume():
    import umeregrobust_b200 as u
    from umeregrobust_b200 import _lib
    _lib.lib()
    return u


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (5, 5), (7, 12), (12, 7), (1, 9), (9, 1), (128, 128), (200, 333), (333, 200)])
def test_matches_scipy(ume, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    c = rng.uniform(0.0, 2.0, shape).astype(np.float32)               # the range of ume_cdist
    r, k = ume.linear_sum_assignment(c)
    r2, k2 = scipy_opt.linear_sum_assignment(c.astype(np.float64))
    assert np.array_equal(r, r2) and np.array_equal(k, k2)


def test_ties_inf_and_errors(ume):
    rng = np.random.default_rng(4)
    c = rng.integers(0, 4, (50, 50)).astype(np.float32)               # many ties: the optimum value is what is defined
    r, k = ume.linear_sum_assignment(c)
    r2, k2 = scipy_opt.linear_sum_assignment(c)
    assert sorted(k.tolist()) == list(range(50)) and c[r, k].sum() == c[r2, k2].sum()
    c = rng.uniform(0, 1, (6, 6)).astype(np.float32)
    c[np.arange(6), np.arange(6)] = np.inf                             # forbidden pairs
    r, k = ume.linear_sum_assignment(c)
    r2, k2 = scipy_opt.linear_sum_assignment(c)
    assert np.array_equal(k, k2) and (k != np.arange(6)).all()
    with pytest.raises(ValueError):                                    # scipy raises ValueError for both, too
        ume.linear_sum_assignment(np.full((3, 3), np.inf, np.float32))  # infeasible
    with pytest.raises(ValueError):
        ume.linear_sum_assignment(np.array([[0.0, np.nan], [1.0, 2.0]], np.float32))
    assert ume.linear_sum_assignment(np.zeros((0, 4), np.float32))[0].shape == (0,)


def test_hungarian_match_layout(ume):
    # evaluate.py:216-222: m[b,:,0] = row indices, m[b,:,1] = column indices, int64
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    D = torch.from_numpy(rng.uniform(0, 2, (3, 40, 40)).astype(np.float32))
    m = ume.hungarian_match(D)
    assert m.shape == (3, 40, 2) and m.dtype == torch.int64
    for b in range(3):
        r2, k2 = scipy_opt.linear_sum_assignment(D[b].numpy())
        assert np.array_equal(m[b, :, 0].numpy(), r2) and np.array_equal(m[b, :, 1].numpy(), k2)
