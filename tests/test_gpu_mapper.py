"""The host-side C++ mirror of Mapper::processInput (libb200mapper.so over libb200icp.so) against a
step-by-step restatement of the reference's sequence on the CPU oracle (tests/reference_mapper.py)."""
import numpy as np
import pytest

from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200._abi import make_config

pytestmark = pytest.mark.gpu


def _scans(n_scans=6, n_pts=20_000, step=0.8, seed=77):
    world = synth.World3D(seed=1234, size=(120.0, 120.0), n_boxes=16)
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_scans):
        T_true = synth.make_T((-10.0 + step * i, 2.0 + 0.3 * i, 1.5), (0.0, 0.0, 10.0 + 3.0 * i))
        S, _ = world.sample(n_pts, np.random.default_rng(seed + 100 + i), noise=0.01, center=T_true[:3, 3], radius=40.0)
        scan = synth.homog(synth.apply_T(np.linalg.inv(T_true), S))
        noise = synth.make_T(rng.normal(0, 0.05, 3), rng.normal(0, 0.3, 3))
        out.append((scan, T_true, (T_true @ noise).astype(np.float32)))
    return out


def test_process_input_sequence_matches_reference_sequence(oracle):
    from norlab_icp_mapper_b200.mapper import Mapper
    from reference_mapper import RefMapper
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=20)
    gpu = Mapper(cfg, True, False, True, False, updateCondition=("distance", 1.0), minDistNewPoint=0.15, surfaceNormalKnn=10)
    ref = RefMapper(cfg, update_distance=1.0, min_dist_new_point=0.15, surface_normal_knn=10)
    updates = []
    for i, (scan, T_true, T_est) in enumerate(_scans()):
        filtered = gpu.applyInputFilters(scan)
        assert len(filtered) == len(scan)  # everything is within sensorMaxRange (200 m)
        gpu.processInput(filtered, T_est, 0.1 * i)
        ref.process_input(scan, T_est, 0.1 * i)
        er, et = synth.pose_error(gpu.getPose(), ref.pose)
        assert er <= 1e-4 and et <= 1e-3, (i, er, et)
        st = gpu.stats()
        assert bool(st.map_updated) == ref.updated, i
        assert abs(st.n_local - len(ref.map)) <= max(2, 0.002 * len(ref.map)), (i, st.n_local, len(ref.map))
        updates.append(ref.updated)
        if i == 0:
            pose0, true0 = gpu.getPose().astype(np.float64), T_true
        else:  # localisation really works: motion relative to the first scan (whose pose defines the map frame)
            rel_gpu = np.linalg.inv(pose0) @ gpu.getPose().astype(np.float64)
            rel_true = np.linalg.inv(true0) @ T_true
            e = synth.pose_error(rel_gpu, rel_true)
            assert e[0] < 2e-3 and e[1] < 0.03, (i, e)
    assert updates[0] and any(updates[1:]) and not all(updates[1:])  # the distance condition throttles updates
    poses, stamps = gpu.getTrajectory()
    assert len(poses) == 6 and np.allclose(stamps, 0.1 * np.arange(6))
    feat, nrm = gpu.getMap()
    assert nrm is not None and len(feat) == gpu.stats().n_global
    cosang = np.abs(np.einsum("ij,ij->i", nrm[:len(ref.normals)], ref.normals[:len(nrm)]))
    assert np.median(cosang) > 0.9999
    gpu.close()


def test_is_mapping_false_only_localises(oracle):
    from norlab_icp_mapper_b200.mapper import Mapper
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=15)
    m = Mapper(cfg, True, False, True, False, surfaceNormalKnn=10)
    scans = _scans(4)
    m.processInput(scans[0][0], scans[0][2], 0.0)
    n0 = m.stats().n_global
    m.setIsMapping(False)
    assert m.getIsMapping() is False
    for i in (1, 2, 3):
        m.processInput(scans[i][0], scans[i][2], 0.1 * i)
        assert m.stats().map_updated == 0 and m.stats().n_global == n0
    m.close()


def test_get_new_local_map_flag_is_consumed():
    """Mapper::getNewLocalMap (Mapper.cpp / python/src/mapper.cpp:16): true with the local map once after an update, false until
    the next one; a scan that only localises (isMapping off) does not raise the flag."""
    from norlab_icp_mapper_b200.mapper import Mapper
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=15)
    m = Mapper(cfg, True, False, True, False, surfaceNormalKnn=10)
    scans = _scans(3)
    m.processInput(scans[0][0], scans[0][2], 0.0)
    ok, feat, nrm = m.getNewLocalMap()
    assert ok and len(feat) == m.stats().n_local and nrm is not None and nrm.shape == (len(feat), 3)
    local, _ = m.getLocalMap()
    assert np.array_equal(feat, local)
    assert m.getNewLocalMap() == (False, None, None)
    m.setIsMapping(False)
    m.processInput(scans[1][0], scans[1][2], 0.1)
    assert m.getNewLocalMap()[0] is False
    m.setIsMapping(True)
    m.processInput(scans[2][0], scans[2][2], 0.2)
    assert (m.stats().map_updated == 1) == m.getNewLocalMap()[0]
    m.close()


def test_update_conditions_and_validation():
    from norlab_icp_mapper_b200.mapper import Mapper
    from norlab_icp_mapper_b200._lib import B200ICPError
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=10)
    for bad in (dict(updateCondition=("overlap", 1.5)), dict(updateCondition=("distance", -1.0)), dict(updateCondition=("nope", 1.0)),
                dict(sensorMaxRange=-5.0)):
        with pytest.raises(B200ICPError):
            Mapper(cfg, True, False, True, False, **bad)
    scans = _scans(3)
    m = Mapper(cfg, True, False, True, False, updateCondition=("delay", 0.15), surfaceNormalKnn=10)
    flags = []
    for i in range(3):
        m.processInput(scans[i][0], scans[i][2], 0.1 * i)
        flags.append(m.stats().map_updated)
    assert flags == [1, 0, 1]  # (t - lastUpdate) > 0.15 s: 0.1 no, 0.2 yes
    m.close()
    m = Mapper(cfg, True, False, True, False, updateCondition=("overlap", 0.99), surfaceNormalKnn=10)
    for i in range(2):
        m.processInput(scans[i][0], scans[i][2], 0.1 * i)
    assert m.stats().map_updated == 1 and m.stats().overlap < 0.99  # trimmed ratio 0.85 < 0.99 -> always update
    m.close()


def test_cell_window_follows_reference_state_machine():
    """sensorMaxRange 30 m and a 400 m drive: the slabs Map::updatePose loads / unloads equal the
    reference's per-axis code, and the local map is exactly the global points inside the window."""
    from norlab_icp_mapper_b200.mapper import Mapper
    from reference_mapper import RefWindow, BUFFER
    cfg = make_config(dim=3, knn=1, max_dist=2.0, outliers=(("trimmed", 0.85),), minimizer="identity", max_iteration_count=1)
    m = Mapper(cfg, True, False, False, False, sensorMaxRange=30.0)  # not mapping: the map is what setMap gave
    rng = np.random.default_rng(3)
    pts = synth.homog(np.c_[rng.uniform(-250, 250, 200_000), rng.uniform(-250, 250, 200_000), rng.uniform(-30, 30, 200_000)])
    m.setMap(pts, None)
    win = RefWindow(30.0)
    scan = synth.homog(rng.normal(0, 5, (500, 3)))
    loaded = np.ones(len(pts), bool)
    cell = np.floor(pts[:, :3] / np.float32(20.0)).astype(int)
    path = [(-200 + 13.0 * k, -150 + 9.0 * k, 3.0 * np.sin(k)) for k in range(32)] + [(216 - 25.0 * k, 138 - 7.0 * k, -45.0 + 4 * k) for k in range(12)]
    n_slabs = 0
    for k, pos in enumerate(path):
        T = synth.make_T(pos, (0, 0, 5.0 * k)).astype(np.float32)
        m.processInput(scan, T, 0.1 * k)
        expect = win.update(np.float32(T[:3, 3]))
        got = m.windowUpdates()
        if expect and expect[0] == "unload-all":
            assert len(got) == 2 and got[0][6] == 0 and tuple(got[1]) == expect[1]
            loaded[:] = False
            expect = expect[1:]
        else:
            assert [tuple(g) for g in got] == expect, (k, got, expect)
        for (r0, r1, c0, c1, a0, a1, load) in expect:
            if load:
                inside = (cell[:, 0] >= r0) & (cell[:, 0] <= r1) & (cell[:, 1] >= c0) & (cell[:, 1] <= c1) & (cell[:, 2] >= a0) & (cell[:, 2] <= a1)
                loaded |= inside
            else:
                x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
                f = np.float32
                inside = ((x >= f(r0) * 20) & (x < (f(r1) + 1) * 20) & (y >= f(c0) * 20) & (y < (f(c1) + 1) * 20)
                          & (z >= f(a0) * 20) & (z < (f(a1) + 1) * 20))
                loaded &= ~inside
        n_slabs += len(expect)
        st = m.stats()
        assert st.n_local == loaded.sum() and st.n_global == len(pts), (k, st.n_local, loaded.sum())
    assert n_slabs > 20  # the window really slid
    m.close()


def test_input_filter_chain_matches_oracle():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import modules_oracle as mo
    from norlab_icp_mapper_b200.mapper import Mapper, bounding_box, distance_limit
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(), minimizer="identity", max_iteration_count=1)
    boxes = [((-1.5, -1, -1), (0.5, 1, 0.5)), ((-6, -2.5, -1), (-1.5, 2.5, 1))]  # examples/config.yaml:1-17
    m = Mapper(cfg, True, False, True, False, sensorMaxRange=40.0, inputFilters=[bounding_box(lo, hi, True) for lo, hi in boxes] + [distance_limit(0.5, 2, False)])
    rng = np.random.default_rng(4)
    scan = synth.homog(rng.normal(0, 15, (50_000, 3)) * [1, 1, 0.1])
    got = m.applyInputFilters(scan)
    keep = mo.distance_limit_keep(scan, 40.0)
    for lo, hi in boxes:
        keep &= mo.bounding_box_keep(scan, lo, hi, True)
    keep &= mo.distance_limit_keep(scan, 0.5, dim_index=2)
    assert 0 < keep.sum() < len(scan) and np.array_equal(got, scan[keep])
    m.close()


def test_example_config_pipeline_matches_reference_sequence(oracle):
    """DynamicPoints + Octree + SurfaceNormal + CutAtDescriptorThreshold + input boxes, i.e. the
    reference's examples/config.yaml (with deterministic octree sampling), over a short drive."""
    from norlab_icp_mapper_b200 import _abi
    from norlab_icp_mapper_b200.mapper import Mapper, bounding_box
    from reference_mapper import RefExampleMapper
    cfg = make_config(dim=3, knn=6, max_dist=2.0, outliers=(), minimizer="identity", max_iteration_count=10)
    boxes = [((-1.5, -1, -1), (0.5, 1, 0.5)), ((-6, -2.5, -1), (-1.5, 2.5, 1))]
    dyn = _abi.DynamicParams(thresholdDynamic=0.9, alpha=0.8, beta=0.99, beamHalfAngle=0.01, epsilonA=0.01, epsilonD=0.01)
    gpu = Mapper(cfg, True, False, True, False, updateCondition=("delay", 0.05), sensorMaxRange=200.0, surfaceNormalKnn=10,
                 dynamicPoints=dyn, octree=(0.15, 0), cutAtThreshold=0.65, inputFilters=[bounding_box(lo, hi, True) for lo, hi in boxes],
                 addProbabilityDynamic=0.6)
    ref = RefExampleMapper(max_size=0.15, knn=10, cut=0.65, delay=0.05, dyn=dict(thresholdDynamic=0.9), boxes=boxes)
    for i, (scan, T_true, T_est) in enumerate(_scans(n_scans=4, n_pts=15_000)):
        f_gpu = gpu.applyInputFilters(scan)
        f_ref = ref.apply_input_filters(scan)
        assert np.array_equal(f_gpu, f_ref)
        gpu.processInput(f_gpu, T_true.astype(np.float32), 0.1 * i)   # exact poses: the chain uses the Identity minimiser
        ref.process_input(f_ref, T_true.astype(np.float32), 0.1 * i)
        assert np.array_equal(gpu.getPose(), T_true.astype(np.float32))
        st = gpu.stats()
        assert bool(st.map_updated) == ref.updated
        # the reference's sensor-frame round trip moves map coordinates by ulps per update, which can move a
        # borderline point across an octree face: sizes agree to a fraction of a percent, not exactly
        assert abs(st.n_local - len(ref.map)) <= 0.005 * len(ref.map) + 5, (i, st.n_local, len(ref.map))
    feat, nrm = gpu.getMap()
    assert nrm is not None
    from scipy.spatial import cKDTree
    d, _ = cKDTree(ref.map[:, :3]).query(feat[:, :3])
    assert (d < 1e-4).mean() > 0.995  # same surviving points
    gpu.close()


def test_mapper_with_reading_normals_and_surface_normal_outlier_filter():
    """A configuration in norlab's usual style: SurfaceNormalDataPointsFilter on the reading (`input:`), TrimmedDist +
    SurfaceNormalOutlierFilter in `icp.outlierFilters`, PointDistance module, SurfaceNormal post filter.  The reading's
    normals travel through processInput (transform, icp(input), insert); the filter drops pairs; poses stay on the truth."""
    from norlab_icp_mapper_b200.mapper import Mapper
    world = synth.World3D(seed=21, size=(80.0, 80.0), n_boxes=10)
    outs = {}
    for name, chain in (("with", (("trimmed", 0.9), ("surface_normal", 0.3))), ("without", (("trimmed", 0.9),))):
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=chain, minimizer="point_to_plane", max_iteration_count=25, differential=(1e-3, 1e-3, 3))
        m = Mapper(cfg, True, False, True, False, updateCondition=("distance", 0.5), sensorMaxRange=60.0, minDistNewPoint=0.1, surfaceNormalKnn=10,
                   inputSurfaceNormalKnn=10)
        rng = np.random.default_rng(3)
        errs, overlaps = [], []
        for i in range(5):
            T_true = synth.make_T((1.0 * i, 0.3 * i, 1.5), (0, 0, 2.0 * i))
            S, _ = world.sample(30_000, np.random.default_rng(100 + i), noise=0.01, center=T_true[:3, 3], radius=50.0)
            scan = synth.homog(synth.apply_T(np.linalg.inv(T_true), S))
            T_est = T_true @ synth.make_T(rng.normal(0, 0.03, 3), rng.normal(0, 0.3, 3)) if i else T_true
            m.processInput(scan, T_est.astype(np.float32), 0.1 * i)
            errs.append(synth.pose_error(m.getPose(), T_true))
            overlaps.append(m.stats().overlap)
        feat, nrm = m.getMap()
        outs[name] = (errs, overlaps, len(feat), nrm)
        m.close()
    for name in outs:
        errs, overlaps, n, nrm = outs[name]
        assert max(e[0] for e in errs) < 2e-3 and max(e[1] for e in errs) < 0.03, (name, errs)
        assert n > 30_000 and np.isfinite(nrm).all()
    # the overlap (weighted ratio of used pairs) is lower with the normal filter: it really rejects pairs
    assert np.mean(outs["with"][1][1:]) < np.mean(outs["without"][1][1:]) - 0.01, (outs["with"][1], outs["without"][1])


def _drive(n_scans, n_pts, seed=40):
    world = synth.World3D(seed=seed, size=(120.0, 120.0), n_boxes=14)
    rng = np.random.default_rng(seed)
    for i in range(n_scans):
        T_true = synth.make_T((1.2 * i, 0.3 * i, 1.5), (0, 0, 1.5 * i))
        S, _ = world.sample(n_pts, np.random.default_rng(1000 + seed + i), noise=0.01, center=T_true[:3, 3], radius=55.0)
        scan = synth.homog(synth.apply_T(np.linalg.inv(T_true), S))
        T_est = T_true @ synth.make_T(rng.normal(0, 0.03, 3), rng.normal(0, 0.3, 3)) if i else T_true
        yield scan, T_true, T_est.astype(np.float32)


def test_online_mapper_runs_the_map_update_asynchronously():
    """isOnline (Mapper.cpp:225-228,248-255,280-283; Map.cpp:29-57): processInput dispatches Map::updateLocalPointCloud to a worker
    and returns; registrations go on against the old map until the worker's final setMap; a new update is refused while one is
    running.  (1) Waiting for each update before the next scan reproduces the offline run bit for bit.  (2) Free running with scans
    arriving at sensor rate: the processInput latency no longer contains the update (the worker does it between two scans; the device
    work of one context is serialised, so back-to-back scans would queue behind it), every pose stays on the truth."""
    import time
    from norlab_icp_mapper_b200.mapper import Mapper
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=20)
    kw = dict(updateCondition=("distance", 0.5), sensorMaxRange=80.0, minDistNewPoint=0.05, surfaceNormalKnn=10, reservePoints=1_500_000)
    scans = list(_drive(14, 60_000))
    runs = {}
    for mode in ("offline", "online_stepwise", "online_free"):
        m = Mapper(cfg, True, mode != "offline", True, False, **kw)
        lat, errs, updated, in_flight_seen = [], [], 0, 0
        for i, (scan, T_true, T_est) in enumerate(scans):
            t0 = time.perf_counter()
            m.processRawInput(scan, T_est, 0.1 * i)
            lat.append((time.perf_counter() - t0) * 1e3)
            updated += int(m.stats().map_updated)
            in_flight_seen += int(m.mapUpdateInFlight())
            if mode == "online_stepwise":
                m.waitForMapUpdate()
            elif mode == "online_free":
                time.sleep(0.004)  # scans arrive at sensor rate: the worker updates the map between two scans
            errs.append(synth.pose_error(m.getPose(), T_true))
        m.waitForMapUpdate()
        feat, nrm = m.getMap()
        runs[mode] = dict(lat=np.array(lat), errs=errs, updated=updated, seen=in_flight_seen, feat=feat, nrm=nrm, traj=m.getTrajectory()[0])
        m.close()
    off, step, free = runs["offline"], runs["online_stepwise"], runs["online_free"]
    # (1) same sequence of maps and poses when every update is awaited
    assert step["updated"] == off["updated"] == len(scans)
    assert np.array_equal(step["traj"], off["traj"])
    assert np.array_equal(step["feat"], off["feat"]) and np.array_equal(step["nrm"], off["nrm"])
    # (2) free running: poses on the truth; the update is off the caller's path
    assert max(e[0] for e in free["errs"]) < 3e-3 and max(e[1] for e in free["errs"]) < 0.05, free["errs"]
    assert free["updated"] <= len(scans) and len(free["feat"]) > 0.5 * len(off["feat"])
    steady = slice(3, None)  # (the first scans create the map synchronously and warm the allocator)
    assert np.median(free["lat"][steady]) < 0.8 * np.median(off["lat"][steady]), (np.median(free["lat"][steady]), np.median(off["lat"][steady]))
    assert free["seen"] > 0  # an update really was in flight when processInput returned


def test_cell_manager_spill_tiers_hold_the_same_map(tmp_path):
    """The CellManager seam (CellManager.h:15-18): with a RAMCellManager or the HardDriveCellManager the cells the window leaves
    move out of device memory (b200icp_map_evict_parked) and come back when the window returns (b200icp_map_append_cloud); the local
    map after every scan and the global map at the end are the same point sets as with the default, where they stay in HBM."""
    from norlab_icp_mapper_b200.mapper import Mapper
    cfg = make_config(dim=3, knn=1, max_dist=2.0, outliers=(("trimmed", 0.85),), minimizer="identity", max_iteration_count=1)
    rng = np.random.default_rng(3)
    pts = synth.homog(np.c_[rng.uniform(-250, 250, 120_000), rng.uniform(-250, 250, 120_000), rng.uniform(-30, 30, 120_000)])
    nrm = rng.normal(size=(len(pts), 3)).astype(np.float32)
    scan = synth.homog(rng.normal(0, 5, (500, 3)))
    # out and back: cells are spilled, then loaded again
    path = [(-200 + 26.0 * k, -150 + 18.0 * k, 0.0) for k in range(16)] + [(190 - 26.0 * k, 120 - 18.0 * k, 0.0) for k in range(16)]
    ms = {"device": Mapper(cfg, True, False, False, False, sensorMaxRange=30.0),
          "ram": Mapper(cfg, True, False, False, False, sensorMaxRange=30.0, cellSpill="ram"),
          "disk": Mapper(cfg, True, False, False, True, sensorMaxRange=30.0, cellFolder=str(tmp_path))}
    for m in ms.values():
        m.setMap(pts, nrm)

    def canon(feat, normals):
        order = np.lexsort(feat[:, :3].T)
        return feat[order], normals[order]
    files_seen = 0
    for k, pos in enumerate(path):
        T = synth.make_T(pos, (0, 0, 5.0 * k)).astype(np.float32)
        st = {}
        for name, m in ms.items():
            m.processInput(scan, T, 0.1 * k)
            st[name] = m.stats()
        files_seen = max(files_seen, len(list(tmp_path.glob("cell_*.vtk"))))
        assert st["ram"].n_local == st["disk"].n_local == st["device"].n_local
        assert st["device"].n_global == len(pts)                      # parked cells stay in HBM
        assert st["ram"].n_global == st["ram"].n_local                 # ... or leave it
        assert st["disk"].n_global == st["disk"].n_local
        if k % 5 == 0:
            ref = canon(*ms["device"].getLocalMap())
            for name in ("ram", "disk"):
                got = canon(*ms[name].getLocalMap())
                assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), (k, name)
    assert files_seen > 10  # the hard-drive manager really wrote cell_<row>_<col>_<aisle>.vtk files
    ref = canon(*ms["device"].getMap())
    assert len(ref[0]) == len(pts)
    for name in ("ram", "disk"):
        got = canon(*ms[name].getMap())
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), name
    for m in ms.values():
        m.close()
    assert not list(tmp_path.glob("cell_*.vtk"))  # ~HardDriveCellManager removes its files (HardDriveCellManager.cpp:4-7)
