"""The oracle's kd-tree (libnabo restatement) against scipy.spatial.cKDTree and the golden file."""
import numpy as np
import pytest
from scipy.spatial import cKDTree


def _scipy_knn(ref, q, k, dim, r):
    tree = cKDTree(ref[:, :dim].astype(np.float64))
    d, i = tree.query(q[:, :dim].astype(np.float64), k=k, distance_upper_bound=r)
    d = d.reshape(len(q), k)
    i = i.reshape(len(q), k)
    return np.where(np.isinf(d), -1, i), d ** 2


@pytest.mark.parametrize("n,nq,k,r,dim", [(20000, 3000, 1, np.inf, 3), (20000, 3000, 6, 2.0, 3), (20000, 3000, 10, np.inf, 3),
                                           (5000, 2000, 8, 0.5, 2), (300, 100, 32, np.inf, 3), (7, 50, 3, np.inf, 2)])
def test_knn_matches_scipy(oracle, n, nq, k, r, dim):
    rng = np.random.default_rng(n + k)
    ref = np.c_[rng.uniform(-20, 20, (n, dim)), np.ones(n)].astype(np.float32)
    q = np.c_[rng.uniform(-25, 25, (nq, dim)), np.ones(nq)].astype(np.float32)
    ids, d2 = oracle.knn(ref, q, k, dim=dim, max_radius=r)
    sid, sd2 = _scipy_knn(ref, q, k, dim, r)
    assert np.array_equal(np.isinf(d2), np.isinf(sd2))
    fin = np.isfinite(d2)
    np.testing.assert_allclose(d2[fin], sd2[fin], rtol=1e-5, atol=1e-9)
    assert (ids == sid).mean() > 0.999  # fp32 vs fp64 can swap near-ties
    assert np.all(np.diff(np.where(fin, d2, np.float32(3e38)), axis=1) >= 0), "results must ascend"
    assert np.all(ids[~fin] == -1)


def test_knn_fewer_points_than_k(oracle):
    ref = np.array([[0, 0, 0, 1], [1, 0, 0, 1]], np.float32)
    q = np.array([[0.1, 0, 0, 1]], np.float32)
    ids, d2 = oracle.knn(ref, q, 4, dim=3)
    assert ids[0].tolist() == [0, 1, -1, -1]
    assert np.isinf(d2[0, 2:]).all() and d2[0, 0] == np.float32(0.1) ** 2


def test_knn_radius_is_inclusive_and_squared(oracle):
    ref = np.array([[3, 4, 0, 1]], np.float32)
    q = np.array([[0, 0, 0, 1]], np.float32)
    ids, d2 = oracle.knn(ref, q, 1, dim=3, max_radius=5.0)  # dist2 == maxRadius^2 is accepted (nabo: <=)
    assert ids[0, 0] == 0 and d2[0, 0] == 25.0
    ids, d2 = oracle.knn(ref, q, 1, dim=3, max_radius=4.999)
    assert ids[0, 0] == -1 and np.isinf(d2[0, 0])


def test_knn_duplicates_and_self_match(oracle):
    rng = np.random.default_rng(5)
    base = rng.uniform(-1, 1, (500, 3)).astype(np.float32)
    ref = np.c_[np.r_[base, base], np.ones(1000)].astype(np.float32)  # every point twice
    ids, d2 = oracle.knn(ref, ref, 2, dim=3)
    assert np.all(d2 == 0.0)  # ALLOW_SELF_MATCH: the point and its duplicate
    assert np.all(np.sort(ids, axis=1) % 500 == (np.arange(1000) % 500)[:, None])


def test_golden_knn(oracle, golden):
    for k, r in ((1, np.inf), (6, 2.0)):
        ids, d2 = oracle.knn(golden["map"], golden["reading"], k, dim=3, max_radius=r)
        gd2, gid = golden[f"knn{k}_d2"], golden[f"knn{k}_ids"]
        assert np.array_equal(np.isinf(d2), np.isinf(gd2))
        fin = np.isfinite(gd2)
        np.testing.assert_allclose(d2[fin], gd2[fin], rtol=2e-5, atol=1e-7)
        assert (ids == gid).mean() > 0.999
