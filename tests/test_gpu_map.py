"""Device-resident map update steps against the oracle: PointDistance insert (bit-exact keep mask),
SurfaceNormal post filter, the 20 m cell window, and that ICP after an update sees the new map."""
import numpy as np
import pytest

from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200._abi import make_config

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu():
    from norlab_icp_mapper_b200.icp import ICP
    g = ICP(make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=15))
    yield g
    g.close()


@pytest.fixture(scope="module")
def pair():
    return synth.make_pair_3d(n_map=150_000, n_scan=15_000)


def test_point_distance_insert_matches_oracle(gpu, oracle, pair):
    gpu.set_map(pair["map"], pair["normals"])
    inp = pair["reading"].copy()
    inp[:, :3] = synth.apply_T(pair["correction_true"], pair["reading"])  # the input arrives in the map frame
    inp = inp.astype(np.float32)
    for min_dist in (0.05, 0.15, 0.5):
        gpu.set_map(pair["map"], pair["normals"])
        added, keep = gpu.map_insert_point_distance(inp, min_dist, want_keep=True)
        kept, okeep = oracle.point_distance_keep(pair["map"], inp, min_dist)
        assert np.array_equal(keep, okeep), (min_dist, (keep != okeep).sum())
        assert added == kept == keep.sum()
        n_local, n_global = gpu.map_counts()
        assert n_local == n_global == len(pair["map"]) + kept
        feat, nrm = gpu.map_download()
        assert nrm is None  # DataPoints::concatenate drops the map's normals when the input has none
        assert np.array_equal(feat[:len(pair["map"])], pair["map"])
        assert np.array_equal(feat[len(pair["map"]):], inp[keep])  # appended in input order


def test_create_map_on_empty(gpu, pair):
    added, keep = gpu.map_insert_point_distance(pair["reading"], 0.15, want_keep=True)
    assert added == len(pair["reading"]) and keep.all()  # createMap copies the input
    gpu.map_commit()
    assert gpu.has_map()
    ids, d2 = gpu.match(pair["reading"][:100])
    assert np.all(d2 == 0) and np.array_equal(ids[:, 0], np.arange(100))


def test_surface_normals_match_oracle(gpu, oracle, pair):
    sub = pair["map"][:60_000]
    gpu.set_map(sub, None)
    gpu.map_surface_normals(10)
    _, nrm = gpu.map_download()
    rc, onrm = oracle.surface_normals(sub, 10)
    assert rc == 0 and nrm is not None
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
    cosang = np.abs(np.einsum("ij,ij->i", nrm, onrm))
    # identical neighbour sets; the eigenvector differs only where the two smallest eigenvalues are nearly equal
    assert np.median(cosang) > 0.999999 and (cosang > 0.9999).mean() > 0.99, (np.median(cosang), (cosang > 0.9999).mean())
    # and they are real surface normals: close to the analytic ones of the synthetic world
    truth = np.abs(np.einsum("ij,ij->i", nrm, pair["normals"][:60_000]))
    assert np.median(truth) > 0.95


def test_icp_after_update_uses_new_normals(gpu, oracle, pair):
    """Full Map::updateLocalPointCloud sequence on the device, then icp(): same pose as the oracle
    run on the downloaded map."""
    gpu.set_map(pair["map"][:100_000], None)
    extra = pair["map"][100_000:]
    added, _ = gpu.map_insert_point_distance(extra, 0.05)
    gpu.map_commit()
    gpu.map_surface_normals(10)
    feat, nrm = gpu.map_download()
    assert len(feat) == 100_000 + added and nrm is not None
    T_g = gpu(pair["reading"])
    cfg = gpu.cfg
    o = oracle.OracleICP(cfg)
    o.set_map(feat, nrm)
    rc, T_o, res, _, _ = o.register(pair["reading"])
    er, et = synth.pose_error(T_g, T_o)
    assert rc == 0 and er <= 1e-4 and et <= 1e-3, (er, et)


def test_cell_window_load_unload(gpu, pair):
    m = pair["map"]
    gpu.set_map(m, pair["normals"])
    cell = np.floor(m[:, :3] / 20.0).astype(int)
    # unload the slab rows [-5, -2] (x in [-100, -20)), everything in y / z
    big = 10 ** 6
    changed = gpu.map_window(False, [-5, -2, -big, big, -big, big])
    inside = (m[:, 0] >= -100.0) & (m[:, 0] < -20.0)
    assert changed == inside.sum()
    n_local, n_global = gpu.map_counts()
    assert n_local == len(m) - inside.sum() and n_global == len(m)
    gpu.map_commit()
    feat, nrm = gpu.map_download()
    assert np.array_equal(feat, m[~inside]) and np.array_equal(nrm, pair["normals"][~inside])
    gfeat, _ = gpu.map_download(global_map=True)
    assert np.array_equal(gfeat, m)
    # queries in the unloaded region no longer find neighbours within maxDist; others are untouched
    q = m[inside][:2000]
    ids, d2 = gpu.match(q)
    assert (ids[:, 0] == -1).mean() > 0.9
    # load back two of the four rows
    changed = gpu.map_window(True, [-4, -3, -big, big, -big, big])
    back = inside & (cell[:, 0] >= -4) & (cell[:, 0] <= -3)
    assert changed == back.sum()
    gpu.map_commit()
    ids, d2 = gpu.match(m[back][:2000])
    assert np.all(d2[:, 0] == 0)
    assert gpu.map_counts()[0] == len(m) - inside.sum() + back.sum()


def test_incremental_surface_normals_match_a_full_recompute(monkeypatch):
    """After an append-only map update the SurfaceNormal pass recomputes only the new points and the old points that got
    a new point within their k-th neighbour distance; everybody else keeps neighbours and normal.  Same normals as a
    full recompute (up to the arbitrary sign and fp32 noise from the moved mean-centring), far fewer points touched."""
    import ctypes
    from norlab_icp_mapper_b200 import synth
    from norlab_icp_mapper_b200.icp import ICP, make_config
    d = synth.make_pair_3d(n_map=300_000, n_scan=40_000, seed=5)
    inp = synth.homog(synth.apply_T(d["correction_true"], d["reading"]))
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=5)
    out = {}
    for mode in ("incremental", "full"):
        if mode == "full":
            monkeypatch.setenv("B200ICP_FULL_NORMALS", "1")
        else:
            monkeypatch.delenv("B200ICP_FULL_NORMALS", raising=False)
        g = ICP(cfg)
        g._L.b200icp_debug_normals_recomputed.restype = ctypes.c_int64
        g._L.b200icp_debug_normals_recomputed.argtypes = [ctypes.c_void_p]
        g.set_map(d["map"], None)
        g.map_surface_normals(10)
        first = g._L.b200icp_debug_normals_recomputed(g._h)
        added, _ = g.map_insert_point_distance(inp, 0.05)
        assert not g._L.b200icp_map_has_normals(g._h)  # concatenate with a scan that has none: formally gone until the post filter runs
        g.map_surface_normals(10)
        second = g._L.b200icp_debug_normals_recomputed(g._h)
        feat, nrm = g.map_download()
        T = g(d["reading"])  # point-to-plane on the refreshed normals
        g.map_surface_normals(10)  # nothing changed since: nothing to do
        third = g._L.b200icp_debug_normals_recomputed(g._h)
        out[mode] = (feat, nrm, first, second, third, added, T)
        g.close()
    fi, ni, first_i, second_i, third_i, added_i, Ti = out["incremental"]
    ff, nf, first_f, second_f, third_f, added_f, Tf = out["full"]
    assert added_i == added_f > 1000 and np.array_equal(fi, ff)
    assert first_i == first_f == len(d["map"]) and second_f == len(ff) and third_f == len(ff)
    assert added_i < second_i < 0.6 * len(fi) and third_i == 0  # new points + their neighbourhoods only
    dots = np.abs((ni * nf).sum(axis=1))
    assert np.isfinite(ni).all() and (dots > 0.9999).mean() > 0.999, (dots > 0.9999).mean()
    er, et = synth.pose_error(Ti, Tf)
    assert er <= 1e-5 and et <= 1e-4, (er, et)


def test_incremental_normals_across_a_window_move(monkeypatch):
    """Unloading / loading 20 m cells (Map::updatePose) changes neighbourhoods at the window's edge: the points whose
    loaded flag flipped are treated like appended / removed points, the pass stays incremental and agrees with a full
    recompute."""
    import ctypes
    from norlab_icp_mapper_b200 import synth
    from norlab_icp_mapper_b200.icp import ICP, make_config
    d = synth.make_pair_3d(n_map=300_000, n_scan=10_000, seed=9)
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(), minimizer="point_to_plane", max_iteration_count=3)
    out = {}
    for mode in ("incremental", "full"):
        if mode == "full":
            monkeypatch.setenv("B200ICP_FULL_NORMALS", "1")
        else:
            monkeypatch.delenv("B200ICP_FULL_NORMALS", raising=False)
        g = ICP(cfg)
        g._L.b200icp_debug_normals_recomputed.restype = ctypes.c_int64
        g._L.b200icp_debug_normals_recomputed.argtypes = [ctypes.c_void_p]
        g.set_map(d["map"], None)
        g.map_surface_normals(8)
        slab = (-2, -2, -10, 10, -10, 10)  # rows -2..-2 (x in [-40, -20)): a 20 m wide strip of the 200 m world
        changed = g.map_window(False, slab)
        g.map_commit()
        g.map_surface_normals(8)
        n_after_unload = g._L.b200icp_debug_normals_recomputed(g._h)
        local1, _ = g.map_counts()
        changed2 = g.map_window(True, slab)
        g.map_commit()
        g.map_surface_normals(8)
        n_after_load = g._L.b200icp_debug_normals_recomputed(g._h)
        feat, nrm = g.map_download()
        out[mode] = (changed, changed2, n_after_unload, n_after_load, local1, feat, nrm)
        g.close()
    ci, ci2, a_i, b_i, l1_i, fi, ni = out["incremental"]
    cf, cf2, a_f, b_f, l1_f, ff, nf = out["full"]
    assert ci == cf == ci2 == cf2 > 1000 and l1_i == l1_f == len(d["map"]) - ci and np.array_equal(fi, ff)
    assert 0 < a_i < 0.2 * l1_i and a_f == l1_f            # only the strip's neighbours
    assert ci <= b_i < ci + 0.2 * len(fi) and b_f == len(ff)  # the strip itself + its neighbours
    dots = np.abs((ni * nf).sum(axis=1))
    assert np.isfinite(ni).all() and (dots > 0.9999).mean() > 0.999, (dots > 0.9999).mean()


def test_cloud_surface_normals_for_the_input_chain(gpu, oracle, pair):
    """SurfaceNormalDataPointsFilter on a reading (input chain): same normals as the oracle's filter, in the cloud's order,
    and the map held by the context is left alone."""
    gpu.set_map(pair["map"][:30_000], None)
    before = gpu.map_counts()
    scan = pair["reading"]
    nrm = gpu.cloud_surface_normals(scan, 12)
    rc, onrm = oracle.surface_normals(scan, 12)
    assert rc == 0 and nrm.shape == (len(scan), 3)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
    cosang = np.abs(np.einsum("ij,ij->i", nrm, onrm))
    assert np.median(cosang) > 0.999999 and (cosang > 0.9999).mean() > 0.99, (np.median(cosang), (cosang > 0.9999).mean())
    assert gpu.map_counts() == before


@pytest.mark.parametrize("case", ["surface3d", "dense2d", "sparse3d"])
def test_staged_self_knn_equals_the_shell_walk(case):
    """SurfaceNormal's neighbour search with TMA-staged candidate tiles (selfknn.cu, nn_variant bit 20) against the per-query shell
    walk (the default): same neighbours in the same order, hence bit-identical normals and k-th distances -- on a surface map, on a dense
    2-D map (regions larger than a stage, tiles with more points than threads) and on a sparse cloud where most queries
    leave their halo and are redone by the shell walk."""
    from norlab_icp_mapper_b200.icp import ICP
    rng = np.random.default_rng(12)
    if case == "surface3d":
        d = synth.make_pair_3d(n_map=200_000, n_scan=1000)
        pts, dim, knn = d["map"], 3, 10
    elif case == "dense2d":
        d = synth.make_pair_2d(n_map=120_000, n_scan=1000)
        pts, dim, knn = d["map"], 2, 8
    else:
        pts = synth.homog(rng.uniform(-30, 30, (20_000, 3)))
        dim, knn = 3, 12
    outs = {}
    for variant in (0, 0x100000, 0x200000):  # default (speculative one-cell bound + rerun list), staged tiles, plain shell walk
        g = ICP(make_config(dim=dim, knn=1, max_dist=1.0, outliers=(), minimizer="point_to_point", max_iteration_count=5, nn_variant=variant))
        g.set_map(pts, None)
        g.map_surface_normals(knn)
        _, nrm = g.map_download()
        outs[variant] = (nrm, g.debug_selfknn_redone())
        # the same search behind the `input:` chain's SurfaceNormal filter
        outs[(variant, "cloud")] = g.cloud_surface_normals(pts[:50_000], knn)
        g.close()
    assert outs[0x200000][1] == -1 and outs[0x100000][1] >= 0 and outs[0][1] >= -1
    assert np.array_equal(outs[0][0], outs[0x200000][0]) and np.array_equal(outs[(0, "cloud")], outs[(0x200000, "cloud")])
    if case != "sparse3d":  # (sparse: most queries leave their halo and are redone, or the truncated list falls back wholesale)
        assert outs[0x100000][1] < 0.25 * len(pts), outs[0x100000][1]
    assert np.isfinite(outs[0][0]).all()
    assert np.array_equal(outs[0][0], outs[0x100000][0])
    assert np.array_equal(outs[(0, "cloud")], outs[(0x100000, "cloud")])
