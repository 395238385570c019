"""Parity of the CUDA k-NN (through the C-ABI) with the oracle: squared distances bit for bit,
ids wherever the distance is unique; edge cases; size-independent properties at full size."""
import numpy as np
import pytest

from norlab_icp_mapper_b200._abi import make_config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from norlab_icp_mapper_b200.icp import ICP
    g = ICP(make_config(dim=3, knn=1, outliers=(), minimizer="point_to_point"))
    yield g
    g.close()


def _check(gpu, oracle, ref, q, k, dim, r):
    ids, d2 = gpu.knn(ref, q, k, dim=dim, max_radius=r)
    oi, od = oracle.knn(ref, q, k, dim=dim, max_radius=r)
    assert np.array_equal(d2, od), "squared distances must be bit-identical"
    # ids may differ only inside groups of equal distance
    diff = ids != oi
    if diff.any():
        rows = np.unique(np.nonzero(diff)[0])
        for i in rows:
            for dval in np.unique(d2[i][diff[i]]):
                a = np.sort(ids[i][d2[i] == dval])
                b = np.sort(oi[i][od[i] == dval])
                if not np.array_equal(a, b):  # a tie straddling the k-th place: both must be true neighbours at that distance
                    assert len(a) == len(b)
                    for cand in (a, b):  # same fp32 expression as both implementations: fma(dz,dz, fma(dy,dy, dx*dx))
                        assert (cand >= 0).all() and len(np.unique(cand)) == len(cand)
                        dlt = ref[cand, :dim].astype(np.float32) - q[i, :dim].astype(np.float32)
                        dd = (dlt[:, 0] * dlt[:, 0]).astype(np.float32)
                        for c in range(1, dim):
                            dd = (dlt[:, c].astype(np.float64) * dlt[:, c].astype(np.float64) + dd.astype(np.float64)).astype(np.float32)
                        assert np.array_equal(dd, np.full(len(cand), dval, np.float32)), (i, dval, cand, dd)
    return ids, d2


@pytest.mark.parametrize("n,nq,k,r,dim", [(50000, 5000, 1, np.inf, 3), (50000, 5000, 6, 2.0, 3), (50000, 5000, 10, np.inf, 3),
                                           (20000, 3000, 8, 0.5, 2), (20000, 3000, 32, np.inf, 3), (1000, 500, 1, 0.05, 3),
                                           (3, 100, 5, np.inf, 3), (1, 10, 1, np.inf, 2), (200000, 20000, 1, 1.0, 3), (200000, 2000, 16, 3.0, 3)])
def test_knn_parity(gpu, oracle, n, nq, k, r, dim):
    rng = np.random.default_rng(1000 + n + 7 * k)
    ref = np.c_[rng.uniform(-20, 20, (n, dim)), np.ones(n)].astype(np.float32)
    q = np.c_[rng.uniform(-25, 25, (nq, dim)), np.ones(nq)].astype(np.float32)
    if dim == 3:  # flat-ish cloud like a lidar map
        ref[:, 2] *= 0.1
        q[:, 2] *= 0.1
    ids, d2 = _check(gpu, oracle, ref, q, k, dim, r)
    fin = np.isfinite(d2)
    assert np.all(ids[~fin] == -1) and np.all(ids[fin] >= 0)
    assert np.all(np.diff(np.where(fin, d2, np.float32(3e38)), axis=1) >= 0)
    if np.isfinite(r):
        assert np.all(d2[fin] <= np.float32(r) * np.float32(r))


def test_knn_surface_world(gpu, oracle):
    from norlab_icp_mapper_b200 import synth
    d = synth.make_pair_3d(n_map=300_000, n_scan=30_000)
    _check(gpu, oracle, d["map"], d["reading"], 1, 3, 1.0)
    _check(gpu, oracle, d["map"], d["reading"][:5000], 6, 3, 2.0)


def test_knn_queries_far_outside_and_degenerate_clouds(gpu, oracle):
    rng = np.random.default_rng(9)
    ref = np.c_[rng.uniform(0, 1, (5000, 3)), np.ones(5000)].astype(np.float32)
    q = np.array([[1000, 1000, 1000, 1], [-500, 0.5, 0.5, 1], [0.5, 0.5, 0.5, 1], [1e6, -1e6, 0, 1]], np.float32)
    _check(gpu, oracle, ref, q, 3, 3, np.inf)      # unbounded radius: far queries still get their true neighbours
    ids, d2 = _check(gpu, oracle, ref, q, 3, 3, 0.2)
    assert np.all(ids[[0, 1, 3]] == -1)
    line = ref.copy()
    line[:, 1:3] = 0.25  # all points on one line: two grid axes have zero extent
    _check(gpu, oracle, line, q, 4, 3, np.inf)
    same = np.tile(np.array([[1, 2, 3, 1]], np.float32), (100, 1))  # all points identical
    ids, d2 = gpu.knn(same, q, 2, dim=3)
    oi, od = oracle.knn(same, q, 2, dim=3)
    assert np.array_equal(d2, od)


def test_knn_duplicates_and_self_match(gpu, oracle):
    rng = np.random.default_rng(5)
    base = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    ref = np.c_[np.r_[base, base], np.ones(10000)].astype(np.float32)
    ids, d2 = gpu.knn(ref, ref, 2, dim=3)
    assert np.all(d2 == 0.0)
    assert np.all(np.sort(ids, axis=1) % 5000 == (np.arange(10000) % 5000)[:, None])


def test_knn_nan_and_inf_queries_have_no_neighbours(gpu):
    ref = np.c_[np.random.default_rng(0).uniform(0, 1, (100, 3)), np.ones(100)].astype(np.float32)
    q = np.array([[np.nan, 0, 0, 1], [np.inf, 0, 0, 1], [0.5, 0.5, 0.5, 1]], np.float32)
    ids, d2 = gpu.knn(ref, q, 2, dim=3)
    assert np.all(ids[:2] == -1) and np.all(np.isinf(d2[:2])) and np.all(ids[2] >= 0)


def test_match_uses_centred_map_like_set_map(oracle):
    """b200icp_match == KDTreeMatcher::findClosests on the mean-centred map (bit-exact), and the
    mean (T_refIn_refMean) equals the oracle's."""
    from norlab_icp_mapper_b200 import synth
    from norlab_icp_mapper_b200.icp import ICP
    d = synth.make_pair_3d(n_map=120_000, n_scan=12_000)
    for k, r in ((1, 1.0), (6, 2.0)):
        cfg = make_config(dim=3, knn=k, max_dist=r, outliers=(), minimizer="point_to_point")
        g = ICP(cfg)
        g.set_map(d["map"], d["normals"])
        o = oracle.OracleICP(cfg)
        o.set_map(d["map"], d["normals"])
        assert np.array_equal(g.map_mean(), o.mean())
        ids, d2 = g.match(d["reading"])
        rc, oi, od = o.match(d["reading"])
        assert np.array_equal(d2, od)
        assert (ids == oi).mean() > 0.9999
        g.close()


def test_full_size_properties():
    """BASELINE config 2 sizes (2M-point map): properties that need no oracle -- every map point is
    its own nearest neighbour at distance 0, results ascend, and a second build gives identical output."""
    from norlab_icp_mapper_b200 import synth
    from norlab_icp_mapper_b200.icp import ICP
    d = synth.make_pair_3d()
    cfg = make_config(dim=3, knn=4, max_dist=1.0, outliers=(), minimizer="point_to_point")
    g = ICP(cfg)
    g.set_map(d["map"], d["normals"])
    sub = np.random.default_rng(1).choice(len(d["map"]), 200_000, replace=False)
    ids, d2 = g.match(d["map"][sub])
    assert np.all(d2[:, 0] == 0.0)
    hit = ids[:, 0] == sub
    # exact duplicates in the map are the only way the first neighbour is not the point itself
    assert hit.mean() > 0.9999 and np.all(d["map"][ids[~hit, 0], :3] == d["map"][sub[~hit], :3])
    assert np.all(np.diff(d2, axis=1) >= 0)
    g.set_map(d["map"], d["normals"])
    ids2, d22 = g.match(d["map"][sub])
    assert np.array_equal(ids, ids2) and np.array_equal(d2, d22)
    g.close()


def test_knn_ties_straddling_the_kth_place(gpu, oracle):
    """Points on an integer lattice: many exact ties, also across the k-th place -- whichever member of a tie group an
    implementation returns must be a true neighbour at exactly that distance (and never twice the same point)."""
    g = np.arange(-4, 5, dtype=np.float32)
    ref = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    ref = np.c_[ref, np.ones(len(ref))].astype(np.float32)
    q = np.c_[np.array([[0, 0, 0], [0.5, 0.5, 0.5], [1, 0.5, 0], [3.5, -2, 1]], np.float32), np.ones(4)].astype(np.float32)
    for k in (1, 3, 5, 7, 12):
        _check(gpu, oracle, ref, q, k, 3, np.inf)
        _check(gpu, oracle, ref, q, k, 3, 1.0)
