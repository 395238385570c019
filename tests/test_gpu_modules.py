"""Octree / DynamicPoints mapper modules and the CutAtDescriptorThreshold post filter on the device
map, against the numpy restatements in oracle/modules_oracle.py."""
import os
import sys

import numpy as np
import pytest

from norlab_icp_mapper_b200 import _abi, synth
from norlab_icp_mapper_b200._abi import make_config

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import modules_oracle as mo  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu():
    from norlab_icp_mapper_b200.icp import ICP
    g = ICP(make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=10))
    yield g
    g.close()


@pytest.fixture(scope="module")
def pair():
    return synth.make_pair_3d(n_map=120_000, n_scan=20_000)


def _rows(a):
    return {tuple(r) for r in np.ascontiguousarray(a).view(np.uint32).reshape(len(a), -1).tolist()}


@pytest.mark.parametrize("max_size", [0.15, 0.5, 2.0])
def test_octree_first_point_sampling(gpu, pair, max_size):
    m, inp = pair["map"], pair["reading"]
    gpu.set_map(m, None)
    n_after = gpu.map_octree(inp, max_size, sampling_method=0)
    feat, nrm = gpu.map_download()
    allpts = np.r_[m, inp]
    order, ofeat, _ = mo.octree_grid_filter(allpts, max_size, 0)
    assert n_after == len(feat) == len(order)
    assert np.array_equal(feat, ofeat)  # same survivors, same (input) order, bit for bit
    # leaf edge in (maxSize / 2, maxSize]: no two survivors share a voxel of that size
    keys, depth = mo.octree_leaf_keys(feat[:, :3], max_size)
    assert depth > 0


def test_octree_centroid_sampling_with_descriptors(gpu, pair):
    m, inp = pair["map"][:50_000], pair["reading"]
    rng = np.random.default_rng(0)
    gpu.set_map(m, pair["normals"][:50_000])
    pm = rng.uniform(0, 1, len(m)).astype(np.float32)
    gpu.map_set_prob(pm)
    pin = np.full(len(inp), 0.6, np.float32)
    n_after = gpu.map_octree(inp, 0.4, sampling_method=2, input_prob=pin)  # the scan has no normals: the map loses them
    feat, nrm = gpu.map_download()
    prob = gpu.map_download_prob()
    assert nrm is None and len(prob) == len(feat) == n_after
    allpts = np.r_[m, inp]
    order, ofeat, odesc = mo.octree_grid_filter(allpts, 0.4, 2, descriptors=np.r_[pm, pin][:, None])
    assert len(order) == n_after
    np.testing.assert_allclose(feat[:, :3], ofeat[:, :3], rtol=0, atol=2e-5)  # centroid sums: same order, fp32
    np.testing.assert_allclose(prob, odesc[:, 0], rtol=0, atol=2e-6)


@pytest.mark.parametrize("method", [1, 3])
def test_octree_random_and_medoid_sampling(gpu, pair, method):
    """samplingMethod 1 (the one examples/config.yaml:48-51 uses) and 3: one survivor per leaf, a member of that leaf;
    the random pick follows the documented counter-based generator, the medoid is the member closest to the centroid."""
    m, inp = pair["map"][:60_000], pair["reading"]
    gpu.set_map(m, None)
    n_after = gpu.map_octree(inp, 0.5, sampling_method=method)
    feat, _ = gpu.map_download()
    allpts = np.r_[m, inp]
    order, ofeat, _ = mo.octree_grid_filter(allpts, 0.5, method)
    assert n_after == len(feat) == len(order)
    keys_all, _ = mo.octree_leaf_keys(allpts[:, :3], 0.5)
    assert len(np.unique(keys_all)) == n_after
    assert _rows(feat) <= _rows(allpts)  # survivors are input points, untouched
    if method == 3:  # fp32 ties between equidistant members may resolve differently under FMA contraction: allow a few
        assert len(_rows(feat) ^ _rows(ofeat)) <= 2 * max(1, n_after // 2000)
    else:
        assert np.array_equal(feat, ofeat)
    # a second call draws with the next seed: still one member per leaf
    if method == 1:
        n2 = gpu.map_octree(inp[:0], 0.5, sampling_method=1)
        assert n2 == n_after


def test_octree_rejects_what_is_not_implemented(gpu, pair):
    from norlab_icp_mapper_b200._lib import B200ICPError
    gpu.set_map(pair["map"][:1000], None)
    with pytest.raises(B200ICPError) as e:
        gpu.map_octree(pair["reading"][:100], 0.15, max_point_by_node=4)
    assert e.value.status == _abi.ERR_NOT_IMPLEMENTED
    with pytest.raises(B200ICPError) as e:
        gpu.map_octree(pair["reading"][:100], 0.15, sampling_method=7)
    assert e.value.status == _abi.ERR_INVALID_ARG


def test_cut_at_descriptor_threshold(gpu, pair):
    m = pair["map"][:40_000]
    rng = np.random.default_rng(1)
    pm = rng.uniform(0, 1, len(m)).astype(np.float32)
    gpu.set_map(m, pair["normals"][:40_000])
    gpu.map_set_prob(pm)
    removed = gpu.map_cut_at_threshold(0.65, True)
    keep = mo.cut_at_descriptor_threshold(pm, 0.65, True)
    feat, nrm = gpu.map_download()
    assert removed == (~keep).sum() and np.array_equal(feat, m[keep]) and np.array_equal(nrm, pair["normals"][:40_000][keep])
    assert np.array_equal(gpu.map_download_prob(), pm[keep])
    gpu.map_commit()
    ids, d2 = gpu.match(m[keep][:500])
    assert np.all(d2[:, 0] == 0)


def test_dynamic_points_update(gpu, pair):
    m, nrm = pair["map"][:80_000], pair["normals"][:80_000]
    rng = np.random.default_rng(2)
    prob0 = rng.uniform(0.05, 0.7, len(m)).astype(np.float32)  # some already above thresholdDynamic
    pose = pair["T_true"].astype(np.float32)
    inp = synth.homog(synth.apply_T(pose, pair["scan"]))  # the scan in the map frame
    # add a dynamic object: a blob of scan points 3 m in front of a wall, i.e. map points seen *through*
    gpu.set_map(m, nrm)
    gpu.map_set_prob(prob0)
    params = _abi.DynamicParams(sensorMaxRange=80.0)
    gpu.map_dynamic_points(inp, np.full(len(inp), 0.6, np.float32), pose, params)
    got = gpu.map_download_prob()
    want, matched = mo.dynamic_points_update(inp, m, nrm, prob0, pose, sensorMaxRange=80.0)
    changed = want != prob0
    assert matched.sum() > 1000 and changed.sum() > 1000
    # asin/atan2 differ by ulps between libm and CUDA, which can flip an angular nearest neighbour or the
    # radius test for a handful of points; everything else must agree to fp32 rounding
    close = np.abs(got - want) <= 1e-4
    assert close.mean() > 0.998, close.mean()
    assert np.array_equal(got[~matched & ~changed], prob0[~matched & ~changed])
    assert np.all((got >= 0) & (got <= 1))


def test_dynamic_points_missing_fields(gpu, pair):
    from norlab_icp_mapper_b200._lib import B200ICPError
    m = pair["map"][:5000]
    pose = np.eye(4, dtype=np.float32)
    gpu.set_map(m, None)
    gpu.map_set_prob(None, 0.6)
    with pytest.raises(B200ICPError) as e:  # map without normals
        gpu.map_dynamic_points(pair["reading"][:100], np.full(100, 0.6, np.float32), pose)
    assert e.value.status == _abi.ERR_INVALID_FIELD and "normals" in str(e.value)
    gpu.set_map(m, pair["normals"][:5000])
    gpu.map_set_prob(None, 0.6)
    with pytest.raises(B200ICPError) as e:  # input without probabilityDynamic
        gpu.map_dynamic_points(pair["reading"][:100], None, pose)
    assert e.value.status == _abi.ERR_INVALID_FIELD and "probabilityDynamic" in str(e.value)


def test_input_filter_chain_with_random_sampling(gpu, pair):
    """BoundingBox + DistanceLimit + RandomSampling (docs/MapperConfiguration.md input chain) in one predicate pass:
    same survivors, same order as the oracle predicates; the sampling rate is what was asked for."""
    from norlab_icp_mapper_b200.mapper import bounding_box, distance_limit, random_sampling
    pts = pair["reading"]
    chain = [bounding_box((-1.5, -1.0, -1.0), (0.5, 1.0, 0.5), True), random_sampling(0.3, seed=7), distance_limit(60.0, -1, False)]
    out = gpu.filter_cloud(pts, chain)
    keep = mo.bounding_box_keep(pts, (-1.5, -1.0, -1.0), (0.5, 1.0, 0.5), True)
    keep &= mo.random_sampling_keep(len(pts), 0.3, seed=7, slot=1)
    keep &= mo.distance_limit_keep(pts, 60.0, -1, False)
    assert np.array_equal(out, pts[keep])
    rate = mo.random_sampling_keep(len(pts), 0.3, seed=7, slot=1).mean()
    assert abs(rate - 0.3) < 0.02
    assert len(gpu.filter_cloud(pts, [random_sampling(1.0)])) == len(pts) and len(gpu.filter_cloud(pts, [random_sampling(0.0)])) == 0
