"""The device-resident scan slot (b200icp_scan_*) and generic descriptors: every step of Mapper::applyInputFilters +
processInput on ONE upload, bit-identical to the host-pointer entry points and to numpy restatements; DataPoints::concatenate's
common-descriptor rule on the device map; the reference-signature MapperModule adapter's `localPointCloud = cloud`."""
import numpy as np
import pytest

from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200._abi import make_config

pytestmark = pytest.mark.gpu


def _ctx(**kw):
    from norlab_icp_mapper_b200.icp import ICP
    args = dict(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=15)
    args.update(kw)
    return ICP(make_config(**args))


@pytest.fixture(scope="module")
def pair():
    return synth.make_pair_3d(n_map=120_000, n_scan=12_000)


def _descriptors(n, seed=5):
    rng = np.random.default_rng(seed)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    prob = rng.uniform(0, 1, n).astype(np.float32)
    # extra block: intensity (1) | observationDirections (3) | t (1) | ring (1)
    extra = np.concatenate([rng.uniform(0, 255, (n, 1)), rng.normal(size=(n, 3)), rng.uniform(0, 0.1, (n, 1)), rng.integers(0, 64, (n, 1))],
                           axis=1).astype(np.float32)
    return nrm, prob, extra


def test_scan_filter_carries_descriptors_and_matches_host_filter(pair):
    from norlab_icp_mapper_b200.mapper import bounding_box, distance_limit, random_sampling
    g = _ctx()
    scan = pair["reading"]
    nrm, prob, extra = _descriptors(len(scan))
    chain = [distance_limit(30.0), bounding_box((-5, -5, -2), (5, 5, 2), True), random_sampling(0.7, seed=11)]
    host = g.filter_cloud(scan, chain)
    g.scan_upload(scan, nrm, prob, extra, rotating_rows=[1])
    n = g.scan_filter(chain)
    f, fn, fp, fx = g.scan_download()
    assert n == len(host) == len(f) and 0 < n < len(scan)
    assert np.array_equal(f, host)
    # which input points survived: features are unique, so match them back
    order = {tuple(p): i for i, p in enumerate(map(tuple, scan))}
    idx = np.array([order[tuple(p)] for p in map(tuple, f)])
    assert np.all(np.diff(idx) > 0)  # ordered compaction
    assert np.array_equal(fn, nrm[idx]) and np.array_equal(fp, prob[idx]) and np.array_equal(fx, extra[idx])
    # an empty chain leaves everything in place; AddDescriptor fills a constant
    g.scan_upload(scan)
    assert g.scan_filter([]) == len(scan)
    g.scan_add_prob(0.6)
    f2, n2, p2, x2 = g.scan_download()
    assert np.array_equal(f2, scan) and n2 is None and x2 is None and np.all(p2 == np.float32(0.6))
    g.close()


def test_scan_transform_rotates_normals_and_observation_directions(pair):
    g = _ctx()
    scan = pair["reading"][:5000]
    nrm, prob, extra = _descriptors(len(scan))
    T = synth.make_T((1.0, -2.0, 0.5), (3.0, -2.0, 40.0)).astype(np.float32)
    g.scan_upload(scan, nrm, prob, extra, rotating_rows=[1])
    g.scan_transform(T)
    f, fn, fp, fx = g.scan_download()
    hf, hn = g.transform(scan, T, nrm)
    assert np.array_equal(f, hf) and np.array_equal(fn, hn)  # same kernel arithmetic as the host-pointer entry point
    # observationDirections rotate exactly like normals do; the other rows are untouched
    _, hod = g.transform(scan, T, extra[:, 1:4])
    assert np.array_equal(fx[:, 1:4], hod)
    assert np.array_equal(fx[:, 0], extra[:, 0]) and np.array_equal(fx[:, 4:], extra[:, 4:]) and np.array_equal(fp, prob)
    # a non-orthogonal matrix is refused like RigidTransformation::compute does
    from norlab_icp_mapper_b200._lib import B200ICPError
    bad = T.copy()
    bad[:3, :3] *= 1.01
    with pytest.raises(B200ICPError) as e:
        g.scan_transform(bad)
    assert e.value.status == 7
    g.close()


def test_scan_register_equals_host_register(pair):
    g = _ctx()
    g.set_map(pair["map"], pair["normals"])
    T_host = g(pair["reading"])
    it_host = g.last_result.iterations
    g.scan_upload(pair["reading"])
    T_dev = g.scan_register()
    assert np.array_equal(T_host, T_dev) and g.last_result.iterations == it_host
    g.close()


def test_scan_register_with_normals_feeds_surface_normal_outlier_filter(pair):
    g = _ctx(outliers=(("trimmed", 0.9), ("surface_normal", 0.3)))
    g.set_map(pair["map"], pair["normals"])
    rn = g.cloud_surface_normals(pair["reading"], 10)
    T_host = g(pair["reading"], reading_normals=rn)
    g.scan_upload(pair["reading"])
    g.scan_surface_normals(10)
    _, dn, _, _ = g.scan_download()
    assert np.array_equal(dn, rn)  # SurfaceNormalDataPointsFilter on the slot == on a host cloud
    T_dev = g.scan_register()
    assert np.array_equal(T_host, T_dev)
    g.close()


def test_point_distance_insert_from_the_slot_concatenates_common_descriptors(pair, oracle):
    g = _ctx()
    n_map = len(pair["map"])
    mn, mp, mx = _descriptors(n_map, seed=9)
    inp = pair["reading"].copy()
    inp[:, :3] = synth.apply_T(pair["correction_true"], pair["reading"])
    inp = inp.astype(np.float32)
    sn, sp, sx = _descriptors(len(inp), seed=10)
    kept, okeep = oracle.point_distance_keep(pair["map"], inp, 0.1)
    okeep = okeep.astype(bool)

    # (a) both clouds carry normals + probabilityDynamic + the same extra layout: everything is concatenated
    g.set_map(pair["map"], pair["normals"])
    g.map_set_prob(mp)
    g.map_set_extra(mx)
    g.scan_upload(inp, sn, sp, sx)
    assert g.scan_insert_point_distance(0.1) == kept
    feat, nrm = g.map_download()
    assert np.array_equal(feat[n_map:], inp[okeep]) and np.array_equal(nrm[n_map:], sn[okeep]) and np.array_equal(nrm[:n_map], pair["normals"])
    assert np.array_equal(g.map_download_prob(), np.concatenate([mp, sp[okeep]]))
    assert np.array_equal(g.map_download_extra(), np.concatenate([mx, sx[okeep]]))

    # (b) the scan lacks `t` and `ring` and has no probabilityDynamic: the caller selects the common rows on both sides
    #     (intensity + observationDirections = rows 0..3), concatenate drops probabilityDynamic
    g.set_map(pair["map"], pair["normals"])
    g.map_set_prob(mp)
    g.map_set_extra(mx)
    g.scan_upload(inp, sn, None, sx[:, :5])
    g.map_select_extra([0, 1, 2, 3])
    g.scan_select_extra([0, 1, 2, 3])
    assert g.scan_insert_point_distance(0.1) == kept
    assert g.map_extra_rows() == 4
    assert np.array_equal(g.map_download_extra(), np.concatenate([mx[:, :4], sx[okeep][:, :4]]))
    from norlab_icp_mapper_b200._lib import B200ICPError
    with pytest.raises(B200ICPError) as e:
        g.map_download_prob()
    assert e.value.status == 8  # InvalidField: the descriptor did not survive the concatenation

    # (c) host-pointer entry point with probabilityDynamic == the slot path (DynamicPoints + PointDistance chain)
    g.set_map(pair["map"], pair["normals"])
    g.map_set_prob(mp)
    assert g.map_insert_point_distance_prob(inp, 0.1, sn, sp) == kept
    assert np.array_equal(g.map_download_prob(), np.concatenate([mp, sp[okeep]]))

    # (d) no points kept still intersects nothing away: the map's descriptors stay
    g.set_map(pair["map"], pair["normals"])
    g.map_set_prob(mp)
    g.scan_upload(pair["map"][:100], pair["normals"][:100], mp[:100])
    assert g.scan_insert_point_distance(0.1) == 0
    assert np.array_equal(g.map_download_prob(), mp)
    g.close()


def test_scan_octree_and_dynamic_points_equal_the_host_pointer_entry_points(pair):
    from norlab_icp_mapper_b200 import _abi
    a, b = _ctx(), _ctx()
    sub, subn = pair["map"][:40_000], pair["normals"][:40_000]
    inp = pair["reading"].copy()
    inp[:, :3] = synth.apply_T(pair["correction_true"], pair["reading"])
    inp = inp.astype(np.float32)
    prob_in = np.full(len(inp), 0.6, np.float32)
    pose = pair["T_true"].astype(np.float32)
    dyn = _abi.DynamicParams(thresholdDynamic=0.9, alpha=0.8, beta=0.99, beamHalfAngle=0.01, epsilonA=0.01, epsilonD=0.01)
    for g in (a, b):
        g.set_map(sub, subn)
        g.map_set_prob(None, 0.6)
    a.map_dynamic_points(inp, prob_in, pose, dyn)
    b.scan_upload(inp, None, prob_in)
    b.scan_dynamic_points(pose, dyn)
    pa, pb = a.map_download_prob(), b.map_download_prob()
    assert np.array_equal(pa, pb) and (pa != np.float32(0.6)).sum() > 100
    # octree on top (concatenate + OctreeGrid, centroid sampler averages every descriptor)
    rn = np.zeros((len(inp), 3), np.float32)
    rn[:, 2] = 1
    na = a.map_octree(inp, 0.3, 2, rn, prob_in)
    b.scan_upload(inp, rn, prob_in)
    nb = b.scan_octree(0.3, 2)
    assert na == nb and na < len(sub) + len(inp)
    fa, nra = a.map_download()
    fb, nrb = b.map_download()
    assert np.array_equal(fa, fb) and np.array_equal(nra, nrb) and np.array_equal(a.map_download_prob(), b.map_download_prob())
    a.close()
    b.close()


def test_replace_local_keeps_parked_cells(pair):
    """`localPointCloud = cloud` for a host-signature module: loaded points replaced, parked ones untouched."""
    g = _ctx()
    g.set_map(pair["map"], pair["normals"])
    # park everything with x >= 0 (cells of 20 m: rows >= 0)
    big = 10 ** 6
    changed = g.map_window(0, [0, big, -big, big, -big, big])
    g.map_commit()
    loaded = pair["map"][:, 0] < 0
    assert changed == (~loaded).sum()
    local, lnrm = g.map_download()
    assert np.array_equal(local, pair["map"][loaded])
    # the "module" keeps every second local point and adds a descriptor-less twist: shift z by 1 cm
    new = local[::2].copy()
    new[:, 2] += np.float32(0.01)
    g.map_replace_local(new, lnrm[::2])
    g.map_commit()
    n_local, n_global = g.map_counts()
    assert n_local == len(new) and n_global == len(new) + (~loaded).sum()
    gf, gn = g.map_download(global_map=True)
    assert np.array_equal(gf[:(~loaded).sum()], pair["map"][~loaded])  # parked points first (order preserved), then the new local cloud
    assert np.array_equal(gf[(~loaded).sum():], new) and np.array_equal(gn[(~loaded).sum():], lnrm[::2])
    # bring the parked cells back: everything is local again
    g.map_window(1, [0, big, -big, big, -big, big])
    g.map_commit()
    assert g.map_counts() == (n_global, n_global)
    g.close()


def test_mapper_dynamic_points_then_point_distance_then_cut(pair):
    """DynamicPointsMapperModule + PointDistanceMapperModule + CutAtDescriptorThreshold -- the chain the
    Python Mapper builds by default with dynamicPoints -- must keep probabilityDynamic on the map across inserts."""
    from norlab_icp_mapper_b200 import _abi
    from norlab_icp_mapper_b200.mapper import Mapper
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=15)
    dyn = _abi.DynamicParams(thresholdDynamic=0.9, alpha=0.8, beta=0.99, beamHalfAngle=0.01, epsilonA=0.01, epsilonD=0.01)
    world = synth.World3D(seed=31, size=(80.0, 80.0), n_boxes=10)
    for raw in (False, True):
        m = Mapper(cfg, True, False, True, False, updateCondition=("delay", 0.05), sensorMaxRange=60.0, minDistNewPoint=0.1, surfaceNormalKnn=10,
                   dynamicPoints=dyn, cutAtThreshold=0.65, addProbabilityDynamic=0.6)
        sizes = []
        for i in range(4):
            T_true = synth.make_T((1.0 * i, 0.3 * i, 1.5), (0, 0, 2.0 * i))
            S, _ = world.sample(20_000, np.random.default_rng(200 + i), noise=0.01, center=T_true[:3, 3], radius=50.0)
            scan = synth.homog(synth.apply_T(np.linalg.inv(T_true), S))
            if raw:
                m.processRawInput(scan, T_true.astype(np.float32), 0.1 * i)
            else:
                m.processInput(m.applyInputFilters(scan), T_true.astype(np.float32), 0.1 * i)
            assert m.stats().map_updated
            sizes.append(m.stats().n_local)
            assert synth.pose_error(m.getPose(), T_true)[1] < 0.03
        prob = m.getMapProbabilityDynamic()
        feat, nrm = m.getMap()
        assert prob is not None and len(prob) == len(feat) and nrm is not None
        assert np.all(prob <= np.float32(0.65))  # what the cut left
        assert sizes[-1] > sizes[0]
        if raw:
            assert sizes == sizes_first  # one upload or two: the same maps
        sizes_first = sizes
        m.close()


def test_mapper_set_map_keeps_probability_dynamic(pair):
    """Mapper::setMap assigns the whole cloud; a reloaded map keeps probabilityDynamic and the
    CutAtDescriptorThreshold post filter / DynamicPoints module keep working on it."""
    from norlab_icp_mapper_b200 import _abi
    from norlab_icp_mapper_b200.mapper import Mapper
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=15)
    dyn = _abi.DynamicParams(thresholdDynamic=0.9, alpha=0.8, beta=0.99, beamHalfAngle=0.01, epsilonA=0.01, epsilonD=0.01)
    m = Mapper(cfg, True, False, True, False, updateCondition=("delay", 0.05), sensorMaxRange=200.0, minDistNewPoint=0.1, surfaceNormalKnn=10,
               dynamicPoints=dyn, cutAtThreshold=0.65, addProbabilityDynamic=0.6)
    sub = pair["map"][:50_000]
    prob = np.random.default_rng(1).uniform(0.1, 0.6, len(sub)).astype(np.float32)
    m.setMap(sub, pair["normals"][:50_000], prob)
    got = m.getMapProbabilityDynamic()
    assert got is not None and np.array_equal(got, prob)
    m.processInput(m.applyInputFilters(pair["scan"]), pair["T_est"].astype(np.float32), 1.0)  # raised InvalidField before the fix
    assert m.stats().map_updated and m.getMapProbabilityDynamic() is not None
    m.close()
