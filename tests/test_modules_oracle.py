"""Sanity checks of the numpy module restatements (oracle/modules_oracle.py) themselves."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import modules_oracle as mo  # noqa: E402


def test_octree_leaf_size_and_uniqueness():
    rng = np.random.default_rng(0)
    p = np.c_[rng.uniform(-50, 50, (20000, 3)) * [1, 1, 0.05], np.ones(20000)].astype(np.float32)
    for max_size in (0.15, 1.0, 7.0):
        keys, depth = mo.octree_leaf_keys(p[:, :3], max_size)
        edge = (p[:, :3].max(0) - p[:, :3].min(0)).max() / 2 ** depth
        assert max_size / 2 < edge <= max_size
        order, feat, _ = mo.octree_grid_filter(p, max_size, 0)
        assert len(order) == len(np.unique(keys)) and np.all(np.diff(order) > 0)
        # points sharing a leaf are within one leaf diagonal of their representative
        rep = {int(k): i for i, k in zip(order, keys[order])}
        far = np.linalg.norm(p[:, :3] - p[[rep[int(k)] for k in keys], :3], axis=1)
        assert far.max() <= edge * np.sqrt(3) + 1e-4
    _, cfeat, _ = mo.octree_grid_filter(p, 1.0, 2)
    assert len(cfeat) == len(mo.octree_grid_filter(p, 1.0, 0)[0])


def test_dynamic_points_moves_probabilities_the_right_way():
    """A wall seen exactly where the map has it becomes more static; map points the beams pass through become more dynamic."""
    rng = np.random.default_rng(1)
    y, z = rng.uniform(-2, 2, 4000), rng.uniform(0, 2, 4000)
    wall = np.c_[np.full(4000, 10.0), y, z, np.ones(4000)].astype(np.float32)
    ghost = np.c_[np.full(4000, 5.0), y * 0.5, z * 0.5 + 0.2, np.ones(4000)].astype(np.float32)  # in front of the wall
    m = np.r_[wall, ghost]
    nrm = np.tile(np.array([[1, 0, 0]], np.float32), (len(m), 1))
    prob0 = np.full(len(m), 0.4, np.float32)
    scan = wall.copy()
    scan[:, :3] += rng.normal(0, 0.002, (4000, 3)).astype(np.float32)
    prob, matched = mo.dynamic_points_update(scan, m, nrm, prob0, np.eye(4))
    assert matched[:4000].mean() > 0.9
    assert np.median(prob[:4000][matched[:4000]]) < 0.4   # confirmed static
    gm = matched[4000:]
    assert gm.sum() > 100 and np.median(prob[4000:][gm]) > 0.4  # seen through: more dynamic


def test_filters():
    p = np.array([[0, 0, 0, 1], [1, 1, 1, 1], [3, 0, 0, 1], [0, -4, 0, 1]], np.float32)
    assert mo.bounding_box_keep(p, [-1.5, -1, -1], [0.5, 1, 0.5], True).tolist() == [False, True, True, True]
    assert mo.distance_limit_keep(p, 3.5).tolist() == [True, True, True, False]
    assert mo.cut_at_descriptor_threshold([0.1, 0.65, 0.7], 0.65).tolist() == [True, True, False]


def test_octree_random_and_medoid_oracle():
    rng = np.random.default_rng(5)
    p = np.c_[rng.uniform(-5, 5, (4000, 3)), np.ones(4000)].astype(np.float32)
    keys, _ = mo.octree_leaf_keys(p[:, :3], 1.0)
    for method in (1, 3):
        order, feat, _ = mo.octree_grid_filter(p, 1.0, method)
        assert len(order) == len(np.unique(keys)) and len(np.unique(keys[order])) == len(order)  # one member per leaf
    a = mo.octree_grid_filter(p, 1.0, 1, seed=1)[0]
    b = mo.octree_grid_filter(p, 1.0, 1, seed=2)[0]
    assert not np.array_equal(a, b) and np.array_equal(a, mo.octree_grid_filter(p, 1.0, 1, seed=1)[0])
    # medoid of a leaf with a clear centre: the centre point wins
    q = np.array([[0.1, 0.1, 0.1, 1], [0.2, 0.2, 0.2, 1], [0.3, 0.3, 0.3, 1], [5, 5, 5, 1]], np.float32)
    order, _, _ = mo.octree_grid_filter(q, 3.0, 3)
    assert 1 in order.tolist()


def test_random_sampling_oracle_rate():
    k = mo.random_sampling_keep(20000, 0.25, seed=3)
    assert abs(k.mean() - 0.25) < 0.01 and not np.array_equal(k, mo.random_sampling_keep(20000, 0.25, seed=4))
