"""The Python surface around the hot path that needs no GPU (norlab_icp_mapper_b200/io.py over the C++ mirror's VTK reader /
writer): DataPoints save / load round trips, the RAM and hard-drive cell managers of python/src/*cell_manager.cpp, Trajectory."""
import os

import numpy as np
import pytest


def _cloud(n, seed, with_normals=True, with_prob=True):
    rng = np.random.default_rng(seed)
    f = np.ones((n, 4), np.float32)
    f[:, :3] = rng.normal(size=(n, 3)).astype(np.float32) * 10
    nrm = rng.normal(size=(n, 3)).astype(np.float32) if with_normals else None
    prob = rng.random(n).astype(np.float32) if with_prob else None
    return dict(features=f, normals=nrm, probabilityDynamic=prob)


@pytest.mark.parametrize("binary", [False, True])
def test_vtk_round_trip(tmp_path, binary):
    from norlab_icp_mapper_b200 import io
    c = _cloud(257, 1)
    p = tmp_path / "cloud.vtk"
    io.save_vtk(p, c["features"], c["normals"], c["probabilityDynamic"], binary=binary)
    head = open(p, "rb").read(200)
    assert head.startswith(b"# vtk DataFile") and (b"BINARY" if binary else b"ASCII") in head
    r = io.load_vtk(p)
    tol = 0 if binary else 1e-6  # (ASCII: 9 significant digits)
    for k in ("features", "normals", "probabilityDynamic"):
        assert r[k].shape == c[k].shape and np.allclose(r[k], c[k], rtol=tol, atol=0), k
    c2 = _cloud(5, 2, with_normals=False, with_prob=False)
    io.save_vtk(p, c2["features"], binary=binary)
    r2 = io.load_vtk(p)
    assert r2["normals"] is None and r2["probabilityDynamic"] is None and np.allclose(r2["features"], c2["features"], rtol=tol)
    with pytest.raises(Exception):
        io.load_vtk(tmp_path / "missing.vtk")


def test_cell_managers(tmp_path):
    from norlab_icp_mapper_b200 import io
    for cm in (io.RamCellManager(), io.HardDriveCellManager(folder=tmp_path)):
        assert cm.getAllCellIds() == [] and len(cm.retrieveCell("0_0_0")["features"]) == 0
        a, b = _cloud(40, 3), _cloud(7, 4)
        cm.saveCell("3_-2_0", a)
        cm.saveCell("-1_7_1", b)
        assert sorted(cm.getAllCellIds()) == ["-1_7_1", "3_-2_0"]
        r = cm.retrieveCell("3_-2_0")
        assert np.array_equal(r["features"], a["features"]) and np.array_equal(r["normals"], a["normals"])
        assert np.array_equal(r["probabilityDynamic"], a["probabilityDynamic"])
        cm.saveCell("3_-2_0", b)  # overwrite
        assert len(cm.retrieveCell("3_-2_0")["features"]) == 7 and len(cm.getAllCellIds()) == 2
        if isinstance(cm, io.HardDriveCellManager):
            assert os.path.exists(os.path.join(tmp_path, "cell_-1_7_1.vtk"))
        cm.clearAllCells()
        assert cm.getAllCellIds() == [] and len(cm.retrieveCell("3_-2_0")["features"]) == 0
        if isinstance(cm, io.HardDriveCellManager):
            assert not os.path.exists(os.path.join(tmp_path, "cell_-1_7_1.vtk"))


def test_trajectory(tmp_path):
    from norlab_icp_mapper_b200 import io
    t = io.Trajectory(3)
    for i in range(5):
        P = np.eye(4, dtype=np.float32)
        P[:3, 3] = (i, 2 * i, 0.5)
        c, s = np.cos(0.1 * i), np.sin(0.1 * i)
        P[:2, :2] = [[c, -s], [s, c]]
        t.addPose(P, 0.1 * i)
    p = tmp_path / "traj.vtk"
    t.save(p)
    r = io.load_vtk(p)
    assert np.allclose(r["features"][:, :3], [[i, 2 * i, 0.5] for i in range(5)])
    assert np.allclose(r["normals"][3], [np.cos(0.3), np.sin(0.3), 0], atol=1e-6)
    t.clear()
    assert t.poses == [] and t.timeStamps == []
    with pytest.raises(ValueError):
        t.addPose(np.eye(3), 0.0)
