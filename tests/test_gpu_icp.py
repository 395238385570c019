"""Parity of the CUDA ICP loop (through the C-ABI) with the oracle: final pose within the BASELINE
tolerance (1e-4 rad, 1e-3 m), same iteration counts / pair counts, same error behaviour."""
import numpy as np
import pytest

from norlab_icp_mapper_b200 import _abi, synth
from norlab_icp_mapper_b200._abi import make_config

pytestmark = pytest.mark.gpu
TOL_RAD, TOL_M = 1e-4, 1e-3


@pytest.fixture(scope="module")
def pair3d():
    return synth.make_pair_3d(n_map=200_000, n_scan=20_000)


@pytest.fixture(scope="module")
def pair2d():
    return synth.make_pair_2d()


def _both(oracle, cfg, d, T_init=None):
    from norlab_icp_mapper_b200.icp import ICP
    o = oracle.OracleICP(cfg)
    o.set_map(d["map"], d["normals"])
    rc, T_o, res_o, tr_o, _ = o.register(d["reading"], T_init=T_init, want_trace=True)
    g = ICP(cfg)
    g.set_trace(True)
    g.set_map(d["map"], d["normals"])
    T_g = g(d["reading"], T_init=T_init)
    res_g, tr_g = g.last_result, g.trace()
    g.close()
    assert rc == _abi.OK
    return T_g, res_g, tr_g, T_o, res_o, tr_o


CHAINS_3D = [
    dict(knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30),
    dict(knn=6, max_dist=2.0, outliers=(("max_dist", 1.0),), minimizer="point_to_plane", max_iteration_count=10),
    dict(knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_point", max_iteration_count=40, differential=(1e-3, 1e-3, 3)),
    dict(knn=1, max_dist=1.0, outliers=(("median", 3.0),), minimizer="point_to_plane", max_iteration_count=15),
    dict(knn=3, max_dist=1.5, outliers=(("min_dist", 0.001), ("trimmed", 0.9)), minimizer="point_to_plane", max_iteration_count=15),
    dict(knn=1, max_dist=float("inf"), outliers=(("trimmed", 0.7),), minimizer="point_to_plane", max_iteration_count=20),
    dict(knn=1, max_dist=1.0, outliers=(), minimizer="point_to_plane", max_iteration_count=40, differential=(1e-3, 1e-3, 3)),
    dict(knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=20, sort_reading=0),
]


@pytest.mark.parametrize("chain", CHAINS_3D)
def test_pose_parity_3d(oracle, pair3d, chain):
    cfg = make_config(dim=3, **chain)
    T_g, res_g, tr_g, T_o, res_o, tr_o = _both(oracle, cfg, pair3d)
    er, et = synth.pose_error(T_g, T_o)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    assert res_g.iterations == res_o.iterations and res_g.max_iter_reached == res_o.max_iter_reached
    assert abs(res_g.overlap - res_o.overlap) < 2e-3
    assert abs(res_g.pairs_last_iter - res_o.pairs_last_iter) <= 0.002 * max(res_o.pairs_last_iter, 1) + 2
    n = min(len(tr_g), len(tr_o))
    for i in range(n):  # the whole trajectory of T_iter agrees, not only the end point
        e = synth.pose_error(tr_g[i], tr_o[i])
        assert e[0] <= 5 * TOL_RAD and e[1] <= 5 * TOL_M, (i, e)


@pytest.mark.parametrize("chain", [dict(knn=8, max_dist=0.5, outliers=(), minimizer="point_to_point", max_iteration_count=30),
                                   dict(knn=1, max_dist=0.5, outliers=(("trimmed", 0.9),), minimizer="point_to_plane", max_iteration_count=30),
                                   dict(knn=2, max_dist=0.5, outliers=(("median", 3.0),), minimizer="point_to_plane", max_iteration_count=20)])
def test_pose_parity_2d(oracle, pair2d, chain):
    cfg = make_config(dim=2, **chain)
    T_g, res_g, _, T_o, res_o, _ = _both(oracle, cfg, pair2d)
    assert T_g.shape == (3, 3)
    er, et = synth.pose_error(T_g, T_o)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    assert res_g.iterations == res_o.iterations


def test_unconverged_trajectory_parity(oracle):
    """SURVEY.md 8d's original perturbation (3 deg yaw at 80 m range) does not converge in 30
    iterations; the two implementations must still walk the same path."""
    d = synth.make_pair_3d(n_map=200_000, n_scan=20_000, drpy_deg=(1.0, -1.0, 3.0))
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
    T_g, res_g, _, T_o, res_o, _ = _both(oracle, cfg, d)
    er, et = synth.pose_error(T_g, T_o)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)


def test_initial_transform_argument(oracle, pair3d):
    """icp(cloud, T_init): registering the raw scan with T_init = T_est equals registering the
    pre-transformed reading with identity (Mapper.cpp:197,213)."""
    from norlab_icp_mapper_b200.icp import ICP
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=20)
    g = ICP(cfg)
    g.set_map(pair3d["map"], pair3d["normals"])
    T_a = g(pair3d["reading"])
    T_b = g(pair3d["scan"], T_init=pair3d["T_est"].astype(np.float32))
    o = oracle.OracleICP(cfg)
    o.set_map(pair3d["map"], pair3d["normals"])
    rc, T_ob, *_ = o.register(pair3d["scan"], T_init=pair3d["T_est"].astype(np.float32))
    g.close()
    e = synth.pose_error(T_b, T_ob)
    assert e[0] <= TOL_RAD and e[1] <= TOL_M
    e = synth.pose_error(T_b, T_a @ pair3d["T_est"])
    assert e[0] <= 5e-4 and e[1] <= 5e-3


def test_golden_corrections(golden):
    from norlab_icp_mapper_b200.icp import ICP
    for key, kw in (("T_plane_k6_it10", dict(knn=6, outliers=(), minimizer="point_to_plane", max_iteration_count=10)),
                    ("T_plane_trim_it30", dict(knn=1, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)),
                    ("T_point_trim_it30", dict(knn=1, outliers=(("trimmed", 0.85),), minimizer="point_to_point", max_iteration_count=30)),
                    ("T_plane_robust_cauchy_mad_it10", dict(knn=1, outliers=(("robust", dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),),
                                                            minimizer="point_to_plane", max_iteration_count=10))):
        g = ICP(make_config(dim=3, max_dist=2.0, **kw))
        g.set_map(golden["map"], golden["normals"])
        T = g(golden["reading"])
        ids, d2 = g.match(golden["reading"])
        g.close()
        er, et = synth.pose_error(T, golden[key])
        assert er <= TOL_RAD and et <= TOL_M, (key, er, et)
        k = kw["knn"]
        gd2 = golden[f"knn{k}_d2"]
        gd2 = np.where(gd2 > 4.0, np.inf, gd2)  # match() applies the chain's maxDist 2.0
        fin = np.isfinite(gd2)
        assert np.array_equal(np.isfinite(d2), fin)


def test_run_to_run_determinism(pair3d):
    """Each execution path is bitwise reproducible; the search strategy (warm ball vs cold shells) does
    not change the result at all; the persistent-kernel and kernel-per-step paths slice the error
    sums differently and therefore agree to rounding only."""
    from norlab_icp_mapper_b200.icp import ICP
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
    outs = {}
    for variant in (0, 4, 4 | 2):  # 0: persistent loop kernel; 4: kernel per step; 4|2: kernel per step, cold search every iteration
        cfg.nn_variant = variant
        g = ICP(cfg)
        g.set_map(pair3d["map"], pair3d["normals"])
        a, b = g(pair3d["reading"]), g(pair3d["reading"])
        g.close()
        assert np.array_equal(a, b), variant
        outs[variant] = a
    assert np.array_equal(outs[4], outs[4 | 2])
    er, et = synth.pose_error(outs[0], outs[4])
    assert er <= 1e-6 and et <= 1e-5, (er, et)


def test_quantile_fallback_path_is_identical(pair3d):
    """nn_variant bit 3 forces the loop kernel's global-pass fallback of the quantile select (used when
    the quantile's bucket is larger than the candidate list); the limit, hence the pose, is the same."""
    from norlab_icp_mapper_b200.icp import ICP
    outs = []
    for variant in (0, 8):
        for chain in ((("trimmed", 0.85),), (("median", 3.0),)):
            cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=chain, minimizer="point_to_plane", max_iteration_count=12, nn_variant=variant)
            g = ICP(cfg)
            g.set_map(pair3d["map"], pair3d["normals"])
            outs.append((g(pair3d["reading"]), g.last_result.pairs_last_iter))
            g.close()
    # same limit, hence exactly the same kept pairs; the two paths slice the error sums differently (the windowed iteration adds
    # its ~0.1 % of candidate pairs on a fixed-point grid), so the poses agree to rounding
    for a, b in ((0, 2), (1, 3)):
        assert outs[a][1] == outs[b][1]
        er, et = synth.pose_error(outs[a][0], outs[b][0])
        assert er <= 1e-6 and et <= 1e-5, (a, b, er, et)


@pytest.mark.parametrize("chain", [
    dict(outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30),
    dict(outliers=(("trimmed", 0.6),), minimizer="point_to_point", max_iteration_count=25),
    dict(outliers=(("min_dist", 0.01), ("trimmed", 0.9)), minimizer="point_to_plane", max_iteration_count=20),
    dict(outliers=(("max_dist", 0.5),), minimizer="point_to_plane", max_iteration_count=20),
    dict(outliers=(), minimizer="point_to_point", max_iteration_count=15),
    dict(outliers=(("median", 3.0),), minimizer="point_to_plane", max_iteration_count=12),
    dict(outliers=(("median", 0.8), ("max_dist", 0.9)), minimizer="point_to_plane", max_iteration_count=12),
])
def test_one_barrier_iterations(pair3d, chain, monkeypatch):
    """The loop kernel's one-barrier iteration (quantile window predicted from the previous limit), its
    two-barrier iteration (window = the quantile's radix bucket, nn_variant bit 4 forces it) and the general
    three-barrier iteration (bits 4 + 6) must find the same quantile limit -- hence exactly the same kept
    pairs -- also when the window policy is made so narrow that most predictions fail and the kernel falls
    back mid-iteration; poses agree to summation-order rounding."""
    from norlab_icp_mapper_b200.icp import ICP
    has_quantile = any(k in ("trimmed", "median") for k, _ in chain["outliers"])  # (Median: the limit's window is the median's, scaled)
    outs = {}
    for name, variant, window in (("fast", 0, None), ("general", 16, None), ("narrow", 0, "0.0,0.00001,1.0"), ("wide", 0, "8.0,0.05,0.5"),
                                  ("always_search", 32, None), ("three_barrier", 16 | 64, None)):
        if window is None:
            monkeypatch.delenv("B200ICP_WINDOW", raising=False)
        else:
            monkeypatch.setenv("B200ICP_WINDOW", window)
        cfg = make_config(dim=3, knn=1, max_dist=1.0, nn_variant=variant, **chain)
        g = ICP(cfg)
        g.set_map(pair3d["map"], pair3d["normals"])
        T = g(pair3d["reading"])
        T2 = g(pair3d["reading"])
        outs[name] = (T, g.last_result.pairs_last_iter, g.last_result.overlap, g.timing().loop_fast_iterations, g.last_result.iterations,
                      g.timing().loop_searched_queries, g.timing().loop_two_barrier_iterations)
        g.close()
        assert np.array_equal(T, T2), name  # each policy is run-to-run deterministic
    assert outs["general"][3] == 0 and outs["three_barrier"][3] == 0 and outs["three_barrier"][6] == 0
    if has_quantile:  # without the prediction every iteration finds the quantile's bucket with a histogram pass: two barriers
        assert outs["general"][6] == outs["general"][4], outs["general"]
    # the match cache only skips searches whose outcome is proven: with it disabled (bit 5) every bit is the same
    assert np.array_equal(outs["fast"][0], outs["always_search"][0]) and outs["fast"][1:5] == outs["always_search"][1:5]
    assert outs["always_search"][5] >= (outs["fast"][4] - 1) * len(pair3d["reading"]) > outs["fast"][5] > 0
    assert outs["fast"][3] + outs["fast"][6] > (5 if has_quantile else 0), outs["fast"]  # windowed iterations (one or two barriers)
    for name in ("fast", "narrow", "wide", "three_barrier"):
        er, et = synth.pose_error(outs[name][0], outs["general"][0])
        assert er <= 1e-6 and et <= 1e-5, (name, er, et)
        assert outs[name][1] == outs["general"][1] and outs[name][2] == outs["general"][2], (name, outs[name], outs["general"])
        assert outs[name][4] == outs["general"][4]


def test_large_reading_spills_the_match_cache(oracle):
    """A CTA of the loop kernel keeps 2048 queries in shared memory and walks its slice 1024 at a time: a 360 k-point
    reading (2433 per CTA) exercises the multi-block walk and the global-memory spill of the cache.  Same pose as the
    oracle, identical to the general path and to the exhaustive search."""
    from norlab_icp_mapper_b200.icp import ICP
    d = synth.make_pair_3d(n_map=300_000, n_scan=360_000, seed=77)
    outs = {}
    for variant in (0, 32, 16 | 64):
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.8),), minimizer="point_to_plane", max_iteration_count=12, nn_variant=variant)
        g = ICP(cfg)
        g.set_map(d["map"], d["normals"])
        outs[variant] = (g(d["reading"]), g.last_result.pairs_last_iter, g.last_result.overlap)
        g.close()
    assert np.array_equal(outs[0][0], outs[32][0]) and outs[0][1:] == outs[32][1:]
    assert outs[0][1:] == outs[16 | 64][1:]
    er, et = synth.pose_error(outs[0][0], outs[16 | 64][0])
    assert er <= 1e-6 and et <= 1e-5, (er, et)
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.8),), minimizer="point_to_plane", max_iteration_count=12)
    o = oracle.OracleICP(cfg)
    o.set_map(d["map"], d["normals"])
    rc, T_o, res_o, _, _ = o.register(d["reading"])
    er, et = synth.pose_error(outs[0][0], T_o)
    assert rc == _abi.OK and er <= TOL_RAD and et <= TOL_M, (er, et)
    assert outs[0][1] == res_o.pairs_last_iter


def test_k_gt_1_spills_the_match_cache(oracle):
    """k > 1 in the loop kernel keeps one cache entry per (reading point, neighbour) pair: 130 k points x knn 3 are
    2 637 entries per CTA, beyond the 2 048 that fit in shared memory, so entries (and some points' k entries only in
    part) live in the global spill arrays.  Same pose and kept pairs as the kernel-per-step path and as the oracle;
    disabling the verification changes nothing."""
    from norlab_icp_mapper_b200.icp import ICP
    d = synth.make_pair_3d(n_map=200_000, n_scan=130_000, seed=78)
    base = dict(dim=3, knn=3, max_dist=1.0, outliers=(("trimmed", 0.8),), minimizer="point_to_plane", max_iteration_count=8)
    outs = {}
    for variant in (0, 32, 4):
        g = ICP(make_config(nn_variant=variant, **base))
        g.set_map(d["map"], d["normals"])
        outs[variant] = (g(d["reading"]), g.last_result.pairs_last_iter, g.last_result.overlap)
        g.close()
    assert outs[0][1:] == outs[32][1:]  # same path with and without the verification: identical
    assert abs(outs[0][1] - outs[4][1]) <= 2 and abs(outs[0][2] - outs[4][2]) <= 1e-5  # other summation order: a boundary pair may flip
    for v in (32, 4):
        er, et = synth.pose_error(outs[0][0], outs[v][0])
        assert er <= 1e-6 and et <= 1e-5, (v, er, et)
    o = oracle.OracleICP(make_config(**base))
    o.set_map(d["map"], d["normals"])
    rc, T_o, res_o, _, _ = o.register(d["reading"])
    er, et = synth.pose_error(outs[0][0], T_o)
    assert rc == _abi.OK and er <= TOL_RAD and et <= TOL_M, (er, et)
    assert abs(outs[0][1] - res_o.pairs_last_iter) <= 2


def test_surface_normal_outlier_filter(oracle, pair3d):
    """SurfaceNormalOutlierFilter{maxAngle} (a staple of norlab's configurations, next to TrimmedDist): pairs whose
    reading normal -- carried through icp(input) and rotated with the reading -- and map normal disagree by more than
    maxAngle get weight 0.  Reading normals: the true surface normals plus noise, sign made consistent with the map's;
    the filter must actually drop pairs, and poses / kept pairs must match the oracle on the persistent loop (k = 1,
    one- and two-barrier iterations), the general path and the kernel-per-step path (k = 3)."""
    from norlab_icp_mapper_b200.icp import ICP
    d = pair3d
    rng = np.random.default_rng(11)
    # reading normals: normal of the nearest map point (after the true correction), perturbed, expressed in the reading's frame
    o0 = oracle.OracleICP(make_config(dim=3, knn=1, max_dist=5.0, outliers=(), minimizer="point_to_point", max_iteration_count=1))
    o0.set_map(d["map"], d["normals"])
    truth = synth.homog(synth.apply_T(d["correction_true"], d["reading"]))
    _, ids, _ = o0.match(truth)
    R = np.asarray(d["correction_true"], np.float64)[:3, :3]
    rn = d["normals"][np.maximum(ids[:, 0], 0)].astype(np.float64) + rng.normal(0, 0.25, (len(truth), 3))
    rn = (rn / np.linalg.norm(rn, axis=1, keepdims=True)) @ R  # map frame -> reading frame (R^T applied to rows)
    rn = rn.astype(np.float32)
    for k, minimizer, variants in ((1, "point_to_plane", (0, 16 | 64, 4)), (3, "point_to_point", (0,))):
        chain = (("trimmed", 0.9), ("surface_normal", 0.35))
        cfg = make_config(dim=3, knn=k, max_dist=1.0, outliers=chain, minimizer=minimizer, max_iteration_count=12)
        o = oracle.OracleICP(cfg)
        o.set_map(d["map"], d["normals"])
        rc, T_o, res_o, _, _ = o.register(d["reading"], reading_normals=rn)
        assert rc == _abi.OK
        rc, T_plain, res_plain, _, _ = o.register(d["reading"])
        assert res_o.pairs_last_iter < 0.9 * res_plain.pairs_last_iter  # the filter bites
        for variant in variants:
            cfg.nn_variant = variant
            g = ICP(cfg)
            g.set_map(d["map"], d["normals"])
            T_g = g(d["reading"], reading_normals=rn)
            res_g = g.last_result
            T_non = g(d["reading"])  # without reading normals the filter is skipped, like upstream
            res_non = g.last_result
            g.close()
            er, et = synth.pose_error(T_g, T_o)
            assert er <= TOL_RAD and et <= TOL_M, (k, variant, er, et)
            assert abs(res_g.pairs_last_iter - res_o.pairs_last_iter) <= 0.002 * res_o.pairs_last_iter + 2, (k, variant)
            assert res_non.pairs_last_iter == res_plain.pairs_last_iter


def test_var_trimmed_dist_outlier_filter(oracle, pair3d):
    """VarTrimmedDistOutlierFilter (LPM defaults minRatio 0.05, maxRatio 0.99, lambda 0.95): the ratio is tuned every
    iteration from the sorted distances.  The device sums in fp64 where upstream (and the oracle) sum sequentially in
    fp32, so the tuned ratio may differ in its last digits: poses must agree to the BASELINE tolerance, the kept
    pairs to a fraction of a percent."""
    for k, minimizer in ((1, "point_to_plane"), (3, "point_to_point")):
        cfg = make_config(dim=3, knn=k, max_dist=1.0, outliers=(("var_trimmed", 0.05, 0.99, 0.95),), minimizer=minimizer, max_iteration_count=15)
        T_g, res_g, tr_g, T_o, res_o, tr_o = _both(oracle, cfg, pair3d)
        er, et = synth.pose_error(T_g, T_o)
        assert er <= TOL_RAD and et <= TOL_M, (k, er, et)
        assert res_g.iterations == res_o.iterations == 15
        assert abs(res_g.pairs_last_iter - res_o.pairs_last_iter) <= 0.01 * res_o.pairs_last_iter, (res_g.pairs_last_iter, res_o.pairs_last_iter)
        assert 0.3 * k * len(pair3d["reading"]) < res_g.pairs_last_iter < 0.995 * k * len(pair3d["reading"])


ROBUST_CASES = [
    ("point_to_plane", 1, 1.0, dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),  # libpointmatcher's defaults
    ("point_to_plane", 2, 1.0, dict(robustFct="huber", tuning=1.5, scaleEstimator="mad", nbIterationForScale=3)),
    ("point_to_point", 1, 1.0, dict(robustFct="tukey", tuning=3.0, scaleEstimator="none", approximation=0.9)),
    ("point_to_plane", 1, 1.0, dict(robustFct="cauchy", tuning=0.05, scaleEstimator="berg")),
    ("point_to_plane", 1, 1.0, dict(robustFct="welsch", tuning=2.0, scaleEstimator="mad", distanceType="point2plane")),
    ("point_to_point", 3, 1.0, dict(robustFct="gm", tuning=1.0, scaleEstimator="mad")),
    ("point_to_plane", 1, 1.0, dict(robustFct="sc", tuning=1.0, scaleEstimator="mad", distanceType="point2plane")),
    ("point_to_plane", 1, 1.0, dict(robustFct="student", tuning=2.0, scaleEstimator="mad")),
    ("point_to_plane", 1, 1.0, dict(robustFct="L1", tuning=1.0, scaleEstimator="none")),
    ("point_to_plane", 1, float("inf"), dict(robustFct="cauchy", tuning=2.0, scaleEstimator="std")),  # (std needs every point matched)
]


@pytest.mark.parametrize("minimizer,knn,max_dist,rp", ROBUST_CASES)
def test_robust_outlier_filter(oracle, pair3d, minimizer, knn, max_dist, rp):
    """RobustOutlierFilter: soft M-estimator weights with a per-iteration scale estimate (two device sorts for mad).
    Same pose, same kept pairs and the same weighted ratio (getOverlap) as the oracle."""
    cfg = make_config(dim=3, knn=knn, max_dist=max_dist, outliers=(("robust", rp), ("max_dist", 0.9)), minimizer=minimizer,
                      max_iteration_count=10)
    T_g, res_g, tr_g, T_o, res_o, tr_o = _both(oracle, cfg, pair3d)
    er, et = synth.pose_error(T_g, T_o)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    assert res_g.iterations == res_o.iterations == 10
    assert abs(res_g.pairs_last_iter - res_o.pairs_last_iter) <= 0.001 * res_o.pairs_last_iter + 2
    assert res_g.overlap == pytest.approx(res_o.overlap, rel=2e-3)


ROBUST_LOOP_CASES = [
    ("point_to_plane", 1, dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),
    ("point_to_plane", 3, dict(robustFct="huber", tuning=1.5, scaleEstimator="mad", nbIterationForScale=3)),
    ("point_to_point", 1, dict(robustFct="welsch", tuning=2.0, scaleEstimator="mad", distanceType="point2plane")),
    ("point_to_plane", 1, dict(robustFct="cauchy", tuning=0.05, scaleEstimator="berg")),
    ("point_to_point", 1, dict(robustFct="tukey", tuning=3.0, scaleEstimator="none", approximation=0.9)),
]


@pytest.mark.parametrize("minimizer,knn,rp", ROBUST_LOOP_CASES)
def test_robust_scale_in_the_loop_kernel(pair3d, minimizer, knn, rp):
    """RobustOutlierFilter inside the persistent loop kernel (mad / berg scale through exact radix selects between grid barriers,
    loop.cu loop_exact_median) against the kernel-per-step path (two device sorts, outlier.cu; nn_variant bit 26): the same
    order statistics, so the same scale bit for bit -- poses agree to the fixed-point rounding of the loop's sums."""
    from norlab_icp_mapper_b200.icp import ICP
    outs = {}
    for name, variant in (("loop", 0), ("steps", 0x4000000)):
        cfg = make_config(dim=3, knn=knn, max_dist=1.0, outliers=(("robust", rp),), minimizer=minimizer, max_iteration_count=12, nn_variant=variant)
        g = ICP(cfg)
        g.set_trace(True)
        g.set_map(pair3d["map"], pair3d["normals"])
        T = g(pair3d["reading"])
        outs[name] = (T, g.last_result, g.trace(), g.timing())
        g.close()
    assert outs["loop"][3].loop_iterations > 0 and outs["steps"][3].loop_iterations == 0  # (each really took its path)
    assert outs["loop"][1].iterations == outs["steps"][1].iterations == 12
    assert outs["loop"][1].pairs_last_iter == outs["steps"][1].pairs_last_iter
    for a, b in zip(outs["loop"][2], outs["steps"][2]):  # every iteration's T_iter
        er, et = synth.pose_error(a, b)
        assert er <= 1e-6 and et <= 1e-5, (er, et)
    assert outs["loop"][1].overlap == pytest.approx(outs["steps"][1].overlap, rel=1e-5)


def test_robust_next_to_a_quantile_filter_in_the_loop_kernel(oracle, pair3d):
    """TrimmedDist / MedianDist + RobustOutlierFilter in one chain: the windowed iterations recompute the M-estimator weight of
    their candidate pairs after the barrier (robust_candidate_weight).  Same poses as the kernel-per-step path and the oracle."""
    from norlab_icp_mapper_b200.icp import ICP
    chains = [
        ("point_to_plane", (("trimmed", 0.8), ("robust", dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")))),
        ("point_to_plane", (("robust", dict(robustFct="tukey", tuning=2.5, scaleEstimator="mad", distanceType="point2plane")), ("median", 3.0))),
        ("point_to_point", (("trimmed", 0.9), ("robust", dict(robustFct="huber", tuning=1.5, scaleEstimator="berg")))),
    ]
    for minimizer, chain in chains:
        outs = {}
        for name, variant in (("loop", 0), ("steps", 0x4000000)):
            cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=chain, minimizer=minimizer, max_iteration_count=14, nn_variant=variant)
            g = ICP(cfg)
            g.set_trace(True)
            g.set_map(pair3d["map"], pair3d["normals"])
            T = g(pair3d["reading"])
            outs[name] = (T, g.last_result, g.trace(), g.timing())
            g.close()
        tm = outs["loop"][3]
        assert tm.loop_iterations > 0 and outs["steps"][3].loop_iterations == 0
        assert tm.loop_fast_iterations > 0  # (one-barrier iterations with candidates took part)
        assert abs(outs["loop"][1].pairs_last_iter - outs["steps"][1].pairs_last_iter) <= 2
        for a, b in zip(outs["loop"][2], outs["steps"][2]):
            er, et = synth.pose_error(a, b)
            assert er <= 2e-6 and et <= 2e-5, (chain, er, et)
        o = oracle.OracleICP(make_config(dim=3, knn=1, max_dist=1.0, outliers=chain, minimizer=minimizer, max_iteration_count=14))
        o.set_map(pair3d["map"], pair3d["normals"])
        rc, T_o, res_o, _, _ = o.register(pair3d["reading"])
        er, et = synth.pose_error(outs["loop"][0], T_o)
        assert rc == _abi.OK and er <= TOL_RAD and et <= TOL_M, (chain, er, et)


def test_robust_scale_in_the_loop_kernel_three_level_select():
    """A reading large enough that the median's radix bucket overflows the in-kernel candidate list (4096 keys): the select
    falls back to its two further global levels (the stamped development build prints it: buckets of 4 900-7 300 keys, twelve
    such selects per registration here).  Same poses as the kernel-per-step path."""
    from norlab_icp_mapper_b200.icp import ICP
    d = synth.make_pair_3d(n_map=400_000, n_scan=400_000, seed=77)
    outs = {}
    for name, variant in (("loop", 0), ("steps", 0x4000000)):
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("robust", dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),),
                          minimizer="point_to_plane", max_iteration_count=6, nn_variant=variant)
        g = ICP(cfg)
        g.set_trace(True)
        g.set_map(d["map"], d["normals"])
        for rep in range(2):  # (twice: the select's histogram regions must come back clean)
            T = g(d["reading"])
        outs[name] = (T, g.last_result, g.trace())
        g.close()
    assert outs["loop"][1].pairs_last_iter == outs["steps"][1].pairs_last_iter
    for a, b in zip(outs["loop"][2], outs["steps"][2]):
        er, et = synth.pose_error(a, b)
        assert er <= 1e-6 and et <= 1e-5, (er, et)


def test_robust_scale_in_the_loop_kernel_edge_cases(oracle):
    """The in-kernel scale selects on the shapes the loop kernel treats specially: k > 1 with entries spilled out of the shared
    memory cache (130 k points x knn 3), a 2-D reading, a reading smaller than the grid (one point per CTA, CTAs without any)
    and a reading without a single neighbour within maxDist (ConvergenceError, as the kernel-per-step path and the oracle)."""
    from norlab_icp_mapper_b200.icp import ICP
    from norlab_icp_mapper_b200._lib import B200ICPError
    rp = dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")
    cases = [
        ("spill", synth.make_pair_3d(n_map=200_000, n_scan=130_000, seed=78), dict(dim=3, knn=3, max_dist=1.0, minimizer="point_to_plane")),
        ("2d", synth.make_pair_2d(n_map=50_000, n_scan=4_000, seed=3100), dict(dim=2, knn=2, max_dist=0.5, minimizer="point_to_point")),
        ("tiny", synth.make_pair_3d(n_map=60_000, n_scan=100, seed=79), dict(dim=3, knn=1, max_dist=1.0, minimizer="point_to_plane")),
    ]
    for name, d, kw in cases:
        outs = {}
        for variant in (0, 0x4000000):
            g = ICP(make_config(outliers=(("robust", rp),), max_iteration_count=8, nn_variant=variant, **kw))
            g.set_map(d["map"], d.get("normals"))
            for rep in range(2):
                T = g(d["reading"])
            outs[variant] = (T, g.last_result.pairs_last_iter, g.last_result.overlap, g.timing().loop_iterations)
            g.close()
        assert outs[0][3] > 0 and outs[0x4000000][3] == 0, name
        assert outs[0][1] == outs[0x4000000][1], name
        er, et = synth.pose_error(outs[0][0], outs[0x4000000][0])
        assert er <= 1e-6 and et <= 1e-5, (name, er, et)
        o = oracle.OracleICP(make_config(outliers=(("robust", rp),), max_iteration_count=8, **kw))
        o.set_map(d["map"], d.get("normals"))
        rc, T_o, res_o, _, _ = o.register(d["reading"])
        er, et = synth.pose_error(outs[0][0], T_o)
        assert rc == _abi.OK and er <= TOL_RAD and et <= TOL_M, (name, er, et)
    # nothing within maxDist: the select finds no finite distance
    d = cases[2][1]
    far = d["reading"].copy()
    far[:, :3] += 500.0
    for variant in (0, 0x4000000):
        g = ICP(make_config(dim=3, knn=1, max_dist=1.0, outliers=(("robust", rp),), minimizer="point_to_plane", max_iteration_count=8, nn_variant=variant))
        g.set_map(d["map"], d["normals"])
        with pytest.raises(B200ICPError) as ei:
            g(far)
        assert ei.value.status == _abi.ERR_CONVERGENCE
        T = g(d["reading"])  # (and the context is usable afterwards: buffers left clean)
        er, et = synth.pose_error(T, d["correction_true"])
        assert er < 2e-2 and et < 0.3
        g.close()


def test_robust_point2plane_needs_reference_normals(pair3d):
    from norlab_icp_mapper_b200.icp import ICP, B200ICPError
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("robust", dict(distanceType="point2plane")),), minimizer="point_to_point",
                      max_iteration_count=3)
    g = ICP(cfg)
    g.set_map(pair3d["map"], None)
    with pytest.raises(B200ICPError) as e:
        g(pair3d["reading"])
    assert e.value.status == _abi.ERR_INVALID_FIELD
    g.close()
    with pytest.raises(B200ICPError) as e:  # one scale estimate per chain
        ICP(make_config(dim=3, outliers=(("robust", {}), ("robust", {}))))
    assert e.value.status == _abi.ERR_NOT_IMPLEMENTED


def test_error_behaviour_matches_libpointmatcher(oracle, pair3d):
    from norlab_icp_mapper_b200.icp import ICP, B200ICPError
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=5)
    g = ICP(cfg)
    with pytest.raises(B200ICPError) as e:  # no map yet
        g(pair3d["reading"])
    assert e.value.status == _abi.ERR_NO_MAP
    assert g.set_map(np.zeros((0, 4), np.float32)) is False and not g.has_map()  # LPM ignores an empty cloud
    g.set_map(pair3d["map"], None)
    with pytest.raises(B200ICPError) as e:  # point-to-plane needs reference normals
        g(pair3d["reading"])
    assert e.value.status == _abi.ERR_INVALID_FIELD
    g.set_map(pair3d["map"], pair3d["normals"])
    far = pair3d["reading"].copy()
    far[:, :3] += 1000.0
    with pytest.raises(B200ICPError) as e:  # nothing within maxDist: ConvergenceError
        g(far)
    assert e.value.status == _abi.ERR_CONVERGENCE
    bad = np.eye(4, dtype=np.float32)
    bad[0, 0] = 1.1
    with pytest.raises(B200ICPError) as e:  # TransformationError
        g(pair3d["reading"], T_init=bad)
    assert e.value.status == _abi.ERR_TRANSFORM
    with pytest.raises(B200ICPError) as e:
        g(np.zeros((0, 4), np.float32))
    assert e.value.status == _abi.ERR_CONVERGENCE
    T = g(pair3d["reading"])  # the context survives errors
    assert np.isfinite(T).all()
    g.close()
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30,
                      bound=(1.0, 0.05))
    g = ICP(cfg)
    g.set_map(pair3d["map"], pair3d["normals"])
    with pytest.raises(B200ICPError) as e:  # BoundTransformationChecker
        g(pair3d["reading"])
    assert e.value.status == _abi.ERR_BOUND
    g.close()


def test_ragged_readings_match_the_oracle(oracle, pair3d):
    """Edge cases of the reading the loop kernel's match cache has to survive: points with no map neighbour within
    maxDist (id -1 from iteration 0 on, some of them drifting into range later), NaN / inf coordinates (never matched),
    readings far smaller than the grid of CTAs.  Poses, iteration counts and kept pairs as the oracle."""
    from norlab_icp_mapper_b200.icp import ICP
    rng = np.random.default_rng(17)
    base = pair3d["reading"]
    cases = {}
    r = base.copy()
    r[::7, :3] += rng.normal(0, 1.0, (len(r[::7]), 3)).astype(np.float32)      # a seventh of the points 1 m off: mostly unmatched
    r[::31, 2] += 50.0                                                           # hopeless ones
    cases["unmatched"] = r
    r = base.copy()
    r[5, 0] = np.nan
    r[77, 1] = np.inf
    r[1234, :3] = np.nan
    cases["nan"] = r
    cases["tiny"] = base[:37].copy()
    cases["one_chunk"] = base[:1000].copy()
    for name, reading in cases.items():
        for chain in ((("trimmed", 0.8),), ()):
            cfg = make_config(dim=3, knn=1, max_dist=0.7, outliers=chain, minimizer="point_to_plane", max_iteration_count=14)
            o = oracle.OracleICP(cfg)
            o.set_map(pair3d["map"], pair3d["normals"])
            rc, T_o, res_o, _, _ = o.register(reading)
            g = ICP(cfg)
            g.set_map(pair3d["map"], pair3d["normals"])
            T_g = g(reading)
            T_g2 = g(reading)
            res_g = g.last_result
            g.close()
            assert rc == _abi.OK and np.array_equal(T_g, T_g2)
            er, et = synth.pose_error(T_g, T_o)
            assert er <= TOL_RAD and et <= TOL_M, (name, chain, er, et)
            assert res_g.iterations == res_o.iterations
            assert abs(res_g.pairs_last_iter - res_o.pairs_last_iter) <= 0.002 * res_o.pairs_last_iter + 2, (name, chain, res_g.pairs_last_iter, res_o.pairs_last_iter)


def test_rigid_transform_bit_exact(oracle):
    from norlab_icp_mapper_b200.icp import ICP, B200ICPError
    rng = np.random.default_rng(2)
    g = ICP(make_config())
    for dim in (3, 2):
        pts = synth.homog(rng.normal(scale=50, size=(10_000, dim)))
        nrm = rng.normal(size=(10_000, dim)).astype(np.float32)
        T = synth.make_T((1, 2, 3), (10, 20, 30)) if dim == 3 else synth.make_T2((1, 2), 33.0)
        out, on = g.transform(pts, T, nrm)
        rc, oo, onn = oracle.transform(pts, T, nrm)
        assert np.array_equal(out, oo) and np.array_equal(on, onn)
    T = synth.make_T((0, 0, 0))
    T[1, 1] = 0.9
    with pytest.raises(B200ICPError) as e:
        g.transform(pts if dim == 3 else synth.homog(rng.normal(size=(4, 3))), T)
    assert e.value.status == _abi.ERR_TRANSFORM
    g.close()


def test_full_size_config2(oracle):
    """BASELINE config 2 at full size: pose parity with the oracle and with the known truth."""
    from norlab_icp_mapper_b200.icp import ICP
    d = synth.make_pair_3d()
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
    g = ICP(cfg)
    g.set_map(d["map"], d["normals"])
    T_g = g(d["reading"])
    res = g.last_result
    g.close()
    o = oracle.OracleICP(cfg)
    o.set_map(d["map"], d["normals"])
    rc, T_o, res_o, _, _ = o.register(d["reading"])
    er, et = synth.pose_error(T_g, T_o)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    assert res.iterations == 30 and res.pairs_last_iter == res_o.pairs_last_iter
    er, et = synth.pose_error(T_g, d["correction_true"])
    assert er <= 1e-4 and et <= 1e-3


# ---- round 2: solve fallback, minimiser options, checker order, conventions, per-point search radius ---------------------------
def _planar_pair():
    import test_oracle_icp
    return test_oracle_icp.planar_pair()


@pytest.mark.parametrize("variant", [0, 4])  # persistent loop kernel / kernel-per-step path: both end in solve6_warp's fallback
def test_rank_deficient_system_matches_the_oracle(oracle, variant):
    """A flat map with parallel normals: the 6 x 6 normal matrix has rank 3, LLT fails and the minimum-norm solution is taken
    (LPM solvePossiblyUnderdeterminedLinearSystem; csrc/icp_device.cuh solve_min_norm)."""
    d = _planar_pair()
    cfg = make_config(dim=3, knn=1, max_dist=2.0, outliers=(), minimizer="point_to_plane", max_iteration_count=8, nn_variant=variant)
    T_g, res_g, tr_g, T_o, res_o, tr_o = _both(oracle, cfg, d)
    er, et = synth.pose_error(T_g, T_o)
    assert np.isfinite(T_g).all() and er <= TOL_RAD and et <= TOL_M, (er, et)
    assert res_g.iterations == res_o.iterations == 8
    fixed = T_g @ d["T_off"]
    assert abs(fixed[2, 3]) < 1e-4 and abs(fixed[2, 0]) < 1e-5 and abs(fixed[2, 1]) < 1e-5  # z, roll, pitch recovered
    assert abs(T_g[0, 3]) < 2e-3 and abs(T_g[1, 3]) < 2e-3 and abs(np.arctan2(T_g[1, 0], T_g[0, 0])) < 1e-4  # null space untouched
    for a, b in zip(tr_g, tr_o):
        e = synth.pose_error(a, b)
        assert e[0] <= 5 * TOL_RAD and e[1] <= 5 * TOL_M


@pytest.mark.parametrize("opt", ["force2D", "force4DOF"])
@pytest.mark.parametrize("variant", [0, 4])
def test_point_to_plane_options(oracle, pair3d, opt, variant):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=12,
                      nn_variant=variant, **{opt: True})
    T_g, res_g, tr_g, T_o, res_o, tr_o = _both(oracle, cfg, pair3d)
    er, et = synth.pose_error(T_g, T_o)
    assert er <= TOL_RAD and et <= TOL_M, (er, et)
    assert res_g.iterations == res_o.iterations and abs(res_g.pairs_last_iter - res_o.pairs_last_iter) <= 0.002 * res_o.pairs_last_iter + 2
    assert np.allclose(T_g[2, :3], [0, 0, 1], atol=1e-6) and np.allclose(T_g[:3, 2], [0, 0, 1], atol=1e-6)
    if opt == "force2D":
        assert abs(T_g[2, 3]) < 1e-5


def test_force2d_with_force4dof_is_rejected():
    from norlab_icp_mapper_b200.icp import ICP, B200ICPError
    with pytest.raises(B200ICPError) as e:
        ICP(make_config(dim=3, force2D=True, force4DOF=True))
    assert e.value.status == _abi.ERR_INVALID_ARG


@pytest.mark.parametrize("variant", [0, 4])
def test_checker_order_counter_throws(oracle, pair3d, variant):
    """Counter listed first (LPM setDefault order): its MaxNumIterationsReached skips a Bound violation of the last iteration;
    Bound listed first: the violation is reported."""
    from norlab_icp_mapper_b200.icp import ICP, B200ICPError
    kw = dict(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", nn_variant=variant)
    o = oracle.OracleICP(make_config(max_iteration_count=3, **kw))
    o.set_map(pair3d["map"], pair3d["normals"])
    _, _, _, trace, _ = o.register(pair3d["reading"], want_trace=True)
    t_norms = [float(np.linalg.norm(t[:3, 3])) for t in trace]
    assert t_norms[2] > max(t_norms[:2])
    limit = 0.5 * (max(t_norms[:2]) + t_norms[2])
    for order, want in ((0, _abi.OK), (2, _abi.ERR_BOUND)):
        cfg = make_config(max_iteration_count=3, bound=(10.0, limit), checker_order=order, **kw)
        o = oracle.OracleICP(cfg)
        o.set_map(pair3d["map"], pair3d["normals"])
        rc_o, T_o, res_o, _, _ = o.register(pair3d["reading"])
        g = ICP(cfg)
        g.set_map(pair3d["map"], pair3d["normals"])
        if want == _abi.OK:
            T_g = g(pair3d["reading"])
            assert rc_o == _abi.OK and g.last_result.max_iter_reached == res_o.max_iter_reached == 1 and g.last_result.iterations == 3
            er, et = synth.pose_error(T_g, T_o)
            assert er <= TOL_RAD and et <= TOL_M
        else:
            with pytest.raises(B200ICPError) as e:
                g(pair3d["reading"])
            assert e.value.status == rc_o == _abi.ERR_BOUND
        g.close()


def test_conventions_switches(oracle, pair3d):
    from norlab_icp_mapper_b200.icp import ICP
    # bit 0: '<' instead of '<=' at maxDist.  A lattice map whose mean is exactly 0, a query at exactly maxDist from six points.
    g1 = np.array([-2, -1, 0, 1, 2], np.float32)
    ref = np.stack(np.meshgrid(g1, g1, g1, indexing="ij"), -1).reshape(-1, 3)
    ref = ref[np.abs(ref).sum(1) > 0]  # (no point at the origin)
    ref = np.c_[ref, np.ones(len(ref))].astype(np.float32)
    q = np.array([[0, 0, 0, 1], [0.25, 0, 0, 1]], np.float32)
    for conv in (0, 1):
        cfg = make_config(dim=3, knn=2, max_dist=1.0, outliers=(), minimizer="point_to_point", conventions=conv)
        g = ICP(cfg)
        g.set_map(ref)
        assert np.array_equal(g.map_mean(), [0, 0, 0])
        ids, d2 = g.match(q)
        o = oracle.OracleICP(cfg)
        o.set_map(ref)
        _, oi, od = o.match(q)
        g.close()
        assert np.array_equal(d2, od)
        if conv == 0:
            assert np.array_equal(d2[0], [1.0, 1.0])
        else:
            assert np.isinf(d2[0]).all() and (ids[0] == -1).all()
        assert d2[1, 0] == np.float32(0.75) ** 2 and np.isinf(d2[1, 1])  # (1,0,0) at 0.75; everything else beyond 1
    # bit 1: Median factor on the distance
    outs = {}
    for name, factor, conv in (("sq", 1.5, 0), ("root", 1.5, 2), ("sq225", 2.25, 0)):
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("median", factor),), minimizer="point_to_plane", max_iteration_count=6,
                          conventions=conv)
        T_g, res_g, _, T_o, res_o, _ = _both(oracle, cfg, pair3d)
        er, et = synth.pose_error(T_g, T_o)
        assert er <= TOL_RAD and et <= TOL_M and abs(res_g.pairs_last_iter - res_o.pairs_last_iter) <= 0.002 * res_o.pairs_last_iter + 2
        outs[name] = (T_g, res_g.pairs_last_iter)
    assert outs["root"][1] > outs["sq"][1] and outs["root"][1] == outs["sq225"][1] and np.array_equal(outs["root"][0], outs["sq225"][0])


def test_max_search_dist_descriptor(oracle, pair3d):
    """A reading with a `maxSearchDist` descriptor: every point is searched within its own radius, the matcher's maxDist is not
    used (LPM KDTreeMatcher / libnabo's vector-of-radii knn)."""
    from norlab_icp_mapper_b200.icp import ICP
    rng = np.random.default_rng(3)
    radii = rng.choice(np.array([0.05, 0.3, 2.0, np.inf], np.float32), len(pair3d["reading"]))
    for knn, minimizer in ((1, "point_to_plane"), (4, "point_to_point")):
        cfg = make_config(dim=3, knn=knn, max_dist=0.01, outliers=(("trimmed", 0.9),), minimizer=minimizer, max_iteration_count=10)
        o = oracle.OracleICP(cfg)
        o.set_map(pair3d["map"], pair3d["normals"])
        o.set_reading_max_search_dist(radii)
        rc, T_o, res_o, _, _ = o.register(pair3d["reading"])
        g = ICP(cfg)
        g.set_map(pair3d["map"], pair3d["normals"])
        T_g = g(pair3d["reading"], reading_max_search_dist=radii)
        res_g = g.last_result
        T_plain_ok = True
        try:
            g(pair3d["reading"])  # without the descriptor maxDist 0.01 leaves (almost) nothing to minimise
        except Exception:
            T_plain_ok = False
        g.close()
        er, et = synth.pose_error(T_g, T_o)
        assert rc == _abi.OK and er <= TOL_RAD and et <= TOL_M, (knn, er, et)
        assert res_g.iterations == res_o.iterations and abs(res_g.pairs_last_iter - res_o.pairs_last_iter) <= 0.002 * res_o.pairs_last_iter + 2
        assert res_g.pairs_last_iter > 0.3 * knn * len(radii) or not T_plain_ok
