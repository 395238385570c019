"""The oracle's ICP loop (libpointmatcher restatement): closed-form checks, the numpy second opinion,
the golden file, the checkers and the error behaviour."""
import numpy as np
import pytest

import numpy_icp
from norlab_icp_mapper_b200 import _abi, synth
from norlab_icp_mapper_b200._abi import make_config

TOL_RAD, TOL_M = 1e-4, 1e-3  # BASELINE.json north_star tolerance


@pytest.fixture(scope="module")
def pair3d():
    return synth.make_pair_3d(n_map=60_000, n_scan=6_000, world_size=(80.0, 80.0), n_boxes=12, scan_radius=35.0)


@pytest.fixture(scope="module")
def pair2d():
    return synth.make_pair_2d(n_map=40_000, n_scan=4_000)


def _run(oracle, cfg, d, **kw):
    o = oracle.OracleICP(cfg)
    assert o.set_map(d["map"], d["normals"]) == _abi.OK
    return o, o.register(d["reading"], **kw)


def test_point_to_plane_recovers_truth(oracle, pair3d):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
    o, (rc, T, res, trace, secs) = _run(oracle, cfg, pair3d, want_trace=True)
    assert rc == _abi.OK and res.iterations == 30 and res.max_iter_reached == 1
    er, et = synth.pose_error(T, pair3d["correction_true"])
    assert er < 5e-4 and et < 5e-3  # 1 cm sensor noise bounds what any ICP can recover
    assert abs(res.overlap - 0.85) < 2e-3  # weightedPointUsedRatio of the trimmed filter
    assert len(trace) == 30


@pytest.mark.parametrize("minimizer,knn,outliers", [("point_to_plane", 1, (("trimmed", 0.85),)), ("point_to_plane", 6, (("max_dist", 0.7),)),
                                                    ("point_to_point", 1, (("trimmed", 0.7),)), ("point_to_plane", 1, (("median", 3.0),)),
                                                    ("point_to_plane", 2, (("min_dist", 0.001), ("max_dist", 0.8)))])
def test_oracle_agrees_with_numpy_second_opinion_3d(oracle, pair3d, minimizer, knn, outliers):
    cfg = make_config(dim=3, knn=knn, max_dist=1.0, outliers=outliers, minimizer=minimizer, max_iteration_count=12)
    _, (rc, T, res, _, _) = _run(oracle, cfg, pair3d)
    assert rc == _abi.OK
    T_np = numpy_icp.icp(pair3d["map"][:, :3], pair3d["normals"], pair3d["reading"][:, :3], knn_k=knn, max_dist=1.0,
                         outliers=outliers, minimizer=minimizer, iterations=12)
    er, et = synth.pose_error(T, T_np)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)


@pytest.mark.parametrize("minimizer,knn", [("point_to_point", 8), ("point_to_plane", 1)])
def test_oracle_agrees_with_numpy_second_opinion_2d(oracle, pair2d, minimizer, knn):
    cfg = make_config(dim=2, knn=knn, max_dist=0.5, outliers=(), minimizer=minimizer, max_iteration_count=15)
    _, (rc, T, res, _, _) = _run(oracle, cfg, pair2d)
    assert rc == _abi.OK and T.shape == (3, 3)
    T_np = numpy_icp.icp(pair2d["map"][:, :2], pair2d["normals"], pair2d["reading"][:, :2], knn_k=knn, max_dist=0.5,
                         outliers=(), minimizer=minimizer, iterations=15)
    er, et = synth.pose_error(T, T_np)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    assert res.overlap == pytest.approx(res.point_used_ratio)


def test_golden_corrections(oracle, golden):
    d = dict(map=golden["map"], normals=golden["normals"], reading=golden["reading"])
    for key, kw in (("T_plane_k6_it10", dict(knn=6, outliers=(), minimizer="point_to_plane", max_iteration_count=10)),
                    ("T_plane_trim_it30", dict(knn=1, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)),
                    ("T_point_trim_it30", dict(knn=1, outliers=(("trimmed", 0.85),), minimizer="point_to_point", max_iteration_count=30)),
                    ("T_plane_robust_cauchy_mad_it10", dict(knn=1, outliers=(("robust", dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),),
                                                            minimizer="point_to_plane", max_iteration_count=10))):
        cfg = make_config(dim=3, max_dist=2.0, **kw)
        _, (rc, T, res, _, _) = _run(oracle, cfg, d)
        assert rc == _abi.OK
        er, et = synth.pose_error(T, golden[key])
        assert er <= TOL_RAD and et <= TOL_M, (key, er, et)


def test_trimmed_quantile_index_semantics(oracle):
    """limit = sorted(finite dists)[size_t(n * ratio)], weights = dist <= limit (ties included)."""
    m = np.array([[float(i), 0, 0, 1] for i in range(10)], np.float32)
    n = np.tile(np.array([[0, 1, 0]], np.float32), (10, 1))
    off = np.array([0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.10], np.float32)
    reading = m.copy()
    reading[:, 1] += off
    for ratio, expected_pairs in ((0.5, 6), (0.85, 9), (1.0, 10), (0.0, 1)):
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", ratio),), minimizer="identity", max_iteration_count=1)
        o = oracle.OracleICP(cfg)
        o.set_map(m, n)
        rc, T, res, _, _ = o.register(reading)
        assert rc == _abi.OK and res.pairs_last_iter == expected_pairs, (ratio, res.pairs_last_iter)
        assert np.allclose(T, np.eye(4))  # IdentityErrorMinimizer (examples/config.yaml:62-63)


def test_counter_and_differential_checkers(oracle, pair3d):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=7)
    _, (rc, T, res, _, _) = _run(oracle, cfg, pair3d)
    assert res.iterations == 7 and res.max_iter_reached == 1
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=40,
                      differential=(1e-3, 1e-3, 3))
    _, (rc, T, res, _, _) = _run(oracle, cfg, pair3d)
    assert rc == _abi.OK and 3 <= res.iterations < 40 and res.max_iter_reached == 0


def test_bound_checker_raises(oracle, pair3d):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30,
                      bound=(1.0, 0.05))  # the 0.37 m initial error exceeds a 5 cm bound
    o, (rc, T, res, _, _) = _run(oracle, cfg, pair3d)
    assert rc == _abi.ERR_BOUND and "bound" in o.last_error()


def test_error_paths(oracle, pair3d):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=5)
    o = oracle.OracleICP(cfg)
    rc, T, res, _, _ = o.register(pair3d["reading"])  # no map: LPM returns identity
    assert rc == _abi.ERR_NO_MAP and np.allclose(T, np.eye(4))
    o.set_map(pair3d["map"], None)  # point-to-plane without normals: InvalidField
    rc, *_ = o.register(pair3d["reading"])
    assert rc == _abi.ERR_INVALID_FIELD
    o.set_map(pair3d["map"], pair3d["normals"])
    far = pair3d["reading"].copy()
    far[:, :3] += 1000.0  # nothing within maxDist: "no outlier to filter"
    rc, *_ = o.register(far)
    assert rc == _abi.ERR_CONVERGENCE
    bad = np.eye(4, dtype=np.float32)
    bad[0, 0] = 1.1  # not a rotation: TransformationError
    rc, *_ = o.register(pair3d["reading"], T_init=bad)
    assert rc == _abi.ERR_TRANSFORM


def test_rigid_transform(oracle):
    rng = np.random.default_rng(2)
    pts = synth.homog(rng.normal(size=(100, 3)))
    nrm = rng.normal(size=(100, 3)).astype(np.float32)
    T = synth.make_T((1, 2, 3), (10, 20, 30))
    rc, out, on = oracle.transform(pts, T, nrm)
    assert rc == _abi.OK
    np.testing.assert_allclose(out[:, :3], synth.apply_T(T, pts), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(on, nrm @ T[:3, :3].T, rtol=1e-5, atol=1e-5)
    assert np.all(out[:, 3] == 1.0)
    T[0, 0] += 0.01
    rc, _, _ = oracle.transform(pts, T)
    assert rc == _abi.ERR_TRANSFORM


def test_point_distance_and_normals(oracle, golden):
    m = golden["map"]
    rng = np.random.default_rng(3)
    inp = synth.homog(m[rng.choice(len(m), 500), :3] + rng.normal(0, 0.2, (500, 3)))
    kept, keep = oracle.point_distance_keep(m, inp, 0.15)
    ids, d2 = numpy_icp.knn(m[:, :3], inp[:, :3], 1)
    expect = d2[:, 0] >= 0.15 ** 2
    assert kept == keep.sum() and (keep == expect).mean() > 0.995  # fp32/fp64 threshold ties
    rc, nrm = oracle.surface_normals(m, 10)
    g = golden["normals"]
    cosang = np.abs(np.einsum("ij,ij->i", nrm, g))
    assert rc == _abi.OK and np.median(cosang) > 0.9999 and (cosang > 0.999).mean() > 0.97


ROBUST_CASES = [
    ("point_to_plane", 1, dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),                      # libpointmatcher's defaults
    ("point_to_plane", 2, dict(robustFct="huber", tuning=1.5, scaleEstimator="mad", nbIterationForScale=3)),
    ("point_to_point", 1, dict(robustFct="tukey", tuning=3.0, scaleEstimator="none", approximation=0.9)),
    ("point_to_plane", 1, dict(robustFct="cauchy", tuning=0.05, scaleEstimator="berg")),
    ("point_to_plane", 1, dict(robustFct="welsch", tuning=2.0, scaleEstimator="mad", distanceType="point2plane")),
    ("point_to_point", 3, dict(robustFct="gm", tuning=1.0, scaleEstimator="mad")),
    ("point_to_plane", 1, dict(robustFct="sc", tuning=1.0, scaleEstimator="mad", distanceType="point2plane")),
    ("point_to_plane", 1, dict(robustFct="student", tuning=2.0, scaleEstimator="mad")),
    ("point_to_plane", 1, dict(robustFct="L1", tuning=1.0, scaleEstimator="none")),  # (1 / |e| is unbounded: point2plane would hinge on near-zero residuals)
]


@pytest.mark.parametrize("minimizer,knn,rp", ROBUST_CASES)
def test_robust_outlier_filter_agrees_with_numpy_second_opinion(oracle, pair3d, minimizer, knn, rp):
    """RobustOutlierFilter: every weight function, scale estimator and distance type against the fp64 numpy statement."""
    outliers = (("robust", rp),)
    cfg = make_config(dim=3, knn=knn, max_dist=1.0, outliers=outliers, minimizer=minimizer, max_iteration_count=10)
    _, (rc, T, res, _, _) = _run(oracle, cfg, pair3d)
    assert rc == _abi.OK
    T_np = numpy_icp.icp(pair3d["map"][:, :3], pair3d["normals"], pair3d["reading"][:, :3], knn_k=knn, max_dist=1.0,
                         outliers=outliers, minimizer=minimizer, iterations=10)
    er, et = synth.pose_error(T, T_np)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    if rp["robustFct"] != "L1":  # weights <= 1 for the others: the weighted ratio is below the kept ratio
        assert 0.0 < res.overlap < res.point_used_ratio <= 1.0


def test_robust_scale_estimators(oracle, pair3d):
    """mad / std / berg scales of the first iteration against numpy on the oracle's own first matches."""
    base = dict(dim=3, knn=2, max_dist=float("inf"), minimizer="point_to_plane", max_iteration_count=1)
    o = oracle.OracleICP(make_config(outliers=(), **base))
    o.set_map(pair3d["map"], pair3d["normals"])
    rc, ids, d2 = o.match(pair3d["reading"])
    fin = np.sort(d2[np.isfinite(d2)].astype(np.float64))
    med = fin[len(fin) // 2]
    want = {"mad": np.sqrt(np.sort(np.abs(fin - med))[len(fin) // 2]), "std": np.sqrt(np.std(d2.astype(np.float64), ddof=1)),
            "berg": 1.9 * np.sqrt(med), "none": 1.0}
    for est, expect in want.items():
        o = oracle.OracleICP(make_config(outliers=(("robust", dict(scaleEstimator=est, tuning=0.05)),), **base))
        o.set_map(pair3d["map"], pair3d["normals"])
        rc = o.register(pair3d["reading"])[0]
        assert rc == _abi.OK
        assert o.last_robust_scale() == pytest.approx(expect, rel=2e-5), est
    # berg from the second iteration on: scale <- 0.85 (scale - target) + target
    o = oracle.OracleICP(make_config(outliers=(("robust", dict(scaleEstimator="berg", tuning=0.05)),), **dict(base, max_iteration_count=3)))
    o.set_map(pair3d["map"], pair3d["normals"])
    assert o.register(pair3d["reading"])[0] == _abi.OK
    s = want["berg"]
    for _ in range(2):
        s = 0.85 * (s - 0.05) + 0.05
    assert o.last_robust_scale() == pytest.approx(s, rel=2e-5)


def test_robust_point2plane_needs_reference_normals(oracle, pair3d):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("robust", dict(distanceType="point2plane")),), minimizer="point_to_point",
                      max_iteration_count=3)
    o = oracle.OracleICP(cfg)
    o.set_map(pair3d["map"], None)
    assert o.register(pair3d["reading"])[0] == _abi.ERR_INVALID_FIELD


@pytest.mark.parametrize("minimizer,knn,rp", [("point_to_point", 4, dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),
                                              ("point_to_plane", 1, dict(robustFct="huber", tuning=1.5, scaleEstimator="mad", distanceType="point2plane"))])
def test_robust_outlier_filter_2d(oracle, pair2d, minimizer, knn, rp):
    outliers = (("robust", rp),)
    cfg = make_config(dim=2, knn=knn, max_dist=0.5, outliers=outliers, minimizer=minimizer, max_iteration_count=12)
    _, (rc, T, res, _, _) = _run(oracle, cfg, pair2d)
    assert rc == _abi.OK and T.shape == (3, 3)
    T_np = numpy_icp.icp(pair2d["map"][:, :2], pair2d["normals"], pair2d["reading"][:, :2], knn_k=knn, max_dist=0.5,
                         outliers=outliers, minimizer=minimizer, iterations=12)
    er, et = synth.pose_error(T, T_np)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)


# ---- round 2: solve fallback, minimiser options, checker order, conventions ---------------------------------------------------
def planar_pair(n_map=20_000, n_scan=3_000, seed=7):
    """A perfectly flat map (all normals = +z): roll, pitch and z are observable, yaw, x and y are not -> the point-to-plane
    normal matrix has rank 3 and LPM's solvePossiblyUnderdeterminedLinearSystem takes its minimum-norm branch."""
    rng = np.random.default_rng(seed)
    P = np.c_[rng.uniform(-20, 20, (n_map, 2)), np.zeros(n_map)]
    N = np.tile([0.0, 0.0, 1.0], (n_map, 1))
    S = np.c_[rng.uniform(-15, 15, (n_scan, 2)), np.zeros(n_scan)]
    T_off = synth.make_T((0.0, 0.0, 0.08), (0.6, -0.4, 0.0))  # what ICP can see: z, roll, pitch
    reading = synth.apply_T(T_off, S)
    return dict(map=synth.homog(P), normals=np.ascontiguousarray(N, np.float32), reading=synth.homog(reading), T_off=T_off)


def test_rank_deficient_system_takes_the_minimum_norm_solution(oracle):
    d = planar_pair()
    cfg = make_config(dim=3, knn=1, max_dist=2.0, outliers=(), minimizer="point_to_plane", max_iteration_count=8)
    _, (rc, T, res, trace, _) = _run(oracle, cfg, d, want_trace=True)
    assert rc == _abi.OK and np.isfinite(T).all() and res.iterations == 8
    T_np = numpy_icp.icp(d["map"][:, :3], d["normals"], d["reading"][:, :3], knn_k=1, max_dist=2.0, outliers=(), iterations=8)
    er, et = synth.pose_error(T, T_np)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    # the observable part is recovered, the unobservable part (yaw, x, y: the null space) stays where the minimum norm puts it: 0
    fixed = T @ d["T_off"]
    assert abs(fixed[2, 3]) < 1e-4 and abs(fixed[2, 0]) < 1e-5 and abs(fixed[2, 1]) < 1e-5
    assert abs(T[0, 3]) < 2e-3 and abs(T[1, 3]) < 2e-3 and abs(np.arctan2(T[1, 0], T[0, 0])) < 1e-4
    # the very first step already is the min-norm step of a rank-3 system: no in-plane translation
    assert abs(trace[0][0, 3]) < 5e-3 and abs(trace[0][1, 3]) < 5e-3


@pytest.mark.parametrize("opt", ["force2D", "force4DOF"])
def test_point_to_plane_options_agree_with_numpy(oracle, pair3d, opt):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=10,
                      **{opt: True})
    _, (rc, T, res, _, _) = _run(oracle, cfg, pair3d)
    assert rc == _abi.OK
    T_np = numpy_icp.icp(pair3d["map"][:, :3], pair3d["normals"], pair3d["reading"][:, :3], knn_k=1, max_dist=1.0,
                         outliers=(("trimmed", 0.85),), iterations=10, **{opt: True})
    er, et = synth.pose_error(T, T_np)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    # the correction is a rotation about z (+ x, y[, z]): the third row / column of the rotation block is untouched
    assert np.allclose(T[2, :3], [0, 0, 1], atol=1e-7) and np.allclose(T[:3, 2], [0, 0, 1], atol=1e-7)
    if opt == "force2D":
        assert abs(T[2, 3]) < 1e-6
    full = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=10)
    _, (_, T_full, _, _, _) = _run(oracle, full, pair3d)
    assert synth.pose_error(T, T_full)[0] > 1e-4  # (the option really changes the answer on this pair: roll / pitch errors stay)


def test_force2d_with_force4dof_is_a_configuration_error(oracle):
    with pytest.raises(ValueError):
        oracle.OracleICP(make_config(dim=3, force2D=True, force4DOF=True))


def test_counter_throws_and_skips_the_checkers_listed_after_it(oracle, pair3d):
    """LPM runs the checkers in YAML order and the Counter reports its limit by throwing: a Bound violation on the last
    iteration is only seen when the Bound checker is listed BEFORE the Counter (checker_order bit 1)."""
    kw = dict(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane")
    free = make_config(max_iteration_count=3, **kw)
    _, (rc, T, res, trace, _) = _run(oracle, free, pair3d, want_trace=True)
    t_norms = [float(np.linalg.norm(t[:3, 3])) for t in trace]  # translation of T_iter after each iteration (refMean frame)
    assert rc == _abi.OK and len(t_norms) == 3
    limit = 0.5 * (max(t_norms[:2]) + t_norms[2]) if t_norms[2] > max(t_norms[:2]) else None
    if limit is None:
        pytest.skip("the translation does not grow on the last iteration for this pair")
    counter_first = make_config(max_iteration_count=3, bound=(10.0, limit), checker_order=0, **kw)
    _, (rc, _, res, _, _) = _run(oracle, counter_first, pair3d)
    assert rc == _abi.OK and res.max_iter_reached == 1 and res.iterations == 3
    bound_first = make_config(max_iteration_count=3, bound=(10.0, limit), checker_order=2, **kw)
    _, (rc, _, res, _, _) = _run(oracle, bound_first, pair3d)
    assert rc == _abi.ERR_BOUND


def test_conventions_switches(oracle, pair3d):
    # bit 0: '<' instead of '<=' at maxDist -- a neighbour at exactly maxDist
    ref = np.array([[1, 0, 0, 1], [-1, 0, 0, 1], [0, 2, 0, 1], [0, -2, 0, 1]], np.float32)
    q = np.array([[0, 0, 0, 1]], np.float32)
    ids, d2 = oracle.knn(ref, q, 2, dim=3, max_radius=1.0)
    assert np.array_equal(d2, [[1.0, 1.0]]) and set(ids[0]) == {0, 1}
    ids, d2 = oracle.knn(ref, q, 2, dim=3, max_radius=1.0, strict=True)
    assert np.isinf(d2).all() and (ids == -1).all()
    ids, d2 = oracle.knn(ref, np.repeat(q, 2, 0), 1, dim=3, max_radii=[0.5, 2.5])  # per-point radii replace maxDist
    assert np.isinf(d2[0, 0]) and d2[1, 0] == 1.0
    # bit 1: Median factor applied to the distance (factor^2 on the squared distances) -- a different, larger inlier set
    a = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("median", 1.5),), minimizer="point_to_plane", max_iteration_count=1)
    b = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("median", 1.5),), minimizer="point_to_plane", max_iteration_count=1, conventions=2)
    c = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("median", 2.25),), minimizer="point_to_plane", max_iteration_count=1)
    _, (_, Ta, ra, _, _) = _run(oracle, a, pair3d)
    _, (_, Tb, rb, _, _) = _run(oracle, b, pair3d)
    _, (_, Tc, rc_, _, _) = _run(oracle, c, pair3d)
    assert rb.pairs_last_iter > ra.pairs_last_iter and rb.pairs_last_iter == rc_.pairs_last_iter and np.array_equal(Tb, Tc)
