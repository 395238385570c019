"""Disk formats of the example pipeline: legacy VTK clouds as libpointmatcher writes them and the
example's trajectory CSV, through the C++ example program `build_map_from_scans_and_trajectory` (the mirror of
/root/reference/examples/build_map_from_scans_and_trajectory.cpp).  The --io-only leg needs no GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "norlab_icp_mapper_b200", "build_map_from_scans_and_trajectory")


def _build():
    if not os.path.exists(EXE):
        subprocess.run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "norlab_icp_mapper_b200", "csrc")], check=True)
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "norlab_icp_mapper_b200", "host")], check=True)


def _write_scan(path, pts, rng):
    """the layout of the reference's examples/data/scans/*.vtk: POINTS / VERTICES / POINT_DATA with two scalar descriptors"""
    n = len(pts)
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 3.0\nFile created by libpointmatcher\nASCII\nDATASET POLYDATA\n")
        f.write(f"POINTS {n} float\n")
        for p in pts:
            f.write(f"{p[0]:12.6g} {p[1]:12.6g} {p[2]:12.6g}\n")
        f.write(f"VERTICES {n} {2 * n}\n")
        for i in range(n):
            f.write(f"1 {i}\n")
        f.write(f"POINT_DATA {n}\nSCALARS intensity float\nLOOKUP_TABLE default\n")
        f.write("\n".join(f"{v:.4g}" for v in rng.uniform(0, 255, n)) + "\n")
        f.write("SCALARS t float\nLOOKUP_TABLE default\n")
        f.write("\n".join(f"{v:.6g}" for v in np.linspace(0, 0.1, n)) + "\n")


def _quat_to_R(x, y, z, w):
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _make_dataset(tmp, n_scans=3, n_pts=400, world=None, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(tmp, "scans"))
    header = ("header.stamp.sec,header.stamp.nanosec,header.frame_id,child_frame_id,pose.pose.position.x,pose.pose.position.y,"
              "pose.pose.position.z,pose.pose.orientation.x,pose.pose.orientation.y,pose.pose.orientation.z,pose.pose.orientation.w,"
              "pose.covariance,twist.twist.linear.x")
    lines, clouds, poses = [header], [], []
    for i in range(n_scans):
        yaw = 0.05 * i
        q = (0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2))
        t = np.array([0.5 * i, -0.1 * i, 0.02 * i])
        R = _quat_to_R(*q)
        if world is None:
            pts = rng.uniform(-20, 20, (n_pts, 3)).astype(np.float32)
        else:
            pts = ((world[rng.choice(len(world), n_pts, replace=False)] - t) @ R).astype(np.float32)  # sensor frame
        sec, nsec = 1690309709 + i // 10, (285305600 + 100000000 * i) % 1000000000
        _write_scan(os.path.join(tmp, "scans", f"cloud_{sec}_{nsec:09d}.vtk"), pts, rng)
        vals = ",".join(repr(float(v)) for v in (*t, *q))
        lines.append(f"{sec},{nsec},map,base_link,{vals},[0. 0. 0.],0.0")
        clouds.append(pts)
        poses.append((R, t))
    with open(os.path.join(tmp, "trajectory.csv"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return clouds, poses


def _read_map(path):
    b = open(path, "rb").read()
    binary = b.split(b"\n")[2].startswith(b"BINARY")
    i = b.index(b"POINTS")
    j = b.index(b"\n", i)
    n = int(b[i:j].split()[1])
    if binary:
        return np.frombuffer(b[j + 1:j + 1 + 12 * n], dtype=">f4").reshape(n, 3).astype(np.float32)
    vals = b[j + 1:].split()
    return np.array(vals[:3 * n], dtype=np.float64).reshape(n, 3).astype(np.float32)


@pytest.mark.parametrize("binary", [False, True])
def test_vtk_and_trajectory_round_trip(tmp_path, binary):
    _build()
    clouds, poses = _make_dataset(str(tmp_path))
    args = [EXE, str(tmp_path), "--io-only"] + (["--binary"] if binary else [])
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _read_map(str(tmp_path / "map.vtk"))
    want = np.concatenate([(c.astype(np.float64) @ R.T + t) for c, (R, t) in zip(clouds, poses)]).astype(np.float32)
    # the scans were written with 6 significant digits: compare against what the file holds, re-read in Python
    reread = []
    for f, (R, t) in zip(sorted(os.listdir(tmp_path / "scans")), poses):
        L = open(tmp_path / "scans" / f).read().split("\n")
        n = int(L[4].split()[1])
        P = np.array([l.split() for l in L[5:5 + n]], dtype=np.float64).astype(np.float32)
        reread.append(P @ R.T.astype(np.float32) + t.astype(np.float32))
    want = np.concatenate(reread)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)


def test_missing_columns_and_files_fail_loudly(tmp_path):
    _build()
    os.makedirs(tmp_path / "scans")
    open(tmp_path / "trajectory.csv", "w").write("a,b,c\n1,2,3\n")
    r = subprocess.run([EXE, str(tmp_path), "--io-only"], capture_output=True, text=True)
    assert r.returncode != 0 and "Required columns not found" in r.stderr


@pytest.mark.gpu
def test_example_pipeline_on_the_device(tmp_path):
    """The whole example (input filters, DynamicPoints + Octree modules with random sampling, SurfaceNormal +
    CutAtDescriptorThreshold post filters, delay update condition) on synthetic scans of one world: it runs, the map has
    one point per occupied 15 cm octree leaf at most, normals and probabilityDynamic travel to the VTK file."""
    _build()
    from norlab_icp_mapper_b200 import synth
    world = synth.World3D(seed=3, size=(60.0, 60.0), n_boxes=8)
    W, _ = world.sample(120_000, np.random.default_rng(1), noise=0.01)
    _make_dataset(str(tmp_path), n_scans=4, n_pts=20_000, world=W.astype(np.float64), seed=2)
    r = subprocess.run([EXE, str(tmp_path), "--icp", "point_to_plane"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    txt = open(tmp_path / "map.vtk").read()
    assert "NORMALS normals float" in txt and "SCALARS probabilityDynamic float" in txt
    m = _read_map(str(tmp_path / "map.vtk"))
    assert 5_000 < len(m) <= 80_000
    keys = np.floor(m / 0.075).astype(np.int64)
    assert len(np.unique(keys, axis=0)) > 0.5 * len(m)  # thinned by the octree: few points share a 7.5 cm voxel


@pytest.mark.gpu
def test_example_with_a_reference_signature_module_and_scan_descriptors(tmp_path):
    """--host-module: a module written against the reference's plugin signature (host DataPoints in and out) runs through
    HostMapperModuleAdapter on the device-resident map; the scans' own descriptors (`intensity`, `t`, as in the reference's
    bundled scans) travel through the input filters, both modules and the map download into map.vtk."""
    _build()
    from norlab_icp_mapper_b200 import synth
    world = synth.World3D(seed=3, size=(60.0, 60.0), n_boxes=8)
    W, _ = world.sample(120_000, np.random.default_rng(1), noise=0.01)
    _make_dataset(str(tmp_path), n_scans=4, n_pts=20_000, world=W.astype(np.float64), seed=2)
    r = subprocess.run([EXE, str(tmp_path), "--icp", "point_to_plane", "--host-module"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    txt = open(tmp_path / "map.vtk").read()
    for tag in ("NORMALS normals float", "SCALARS probabilityDynamic float", "SCALARS intensity float", "SCALARS t float"):
        assert tag in txt, tag
    m = _read_map(str(tmp_path / "map.vtk"))
    assert 5_000 < len(m) <= 80_000
    keys = np.floor(m.astype(np.float64) / 0.15).astype(np.int64)
    # the voxel module leaves one point per 15 cm voxel; the CutAtDescriptorThreshold post filter may remove some afterwards
    assert len(np.unique(keys, axis=0)) >= len(m) - 5  # (fp32 vs fp64 voxel keys may disagree for a point on a voxel face)
    # descriptor values are the scans' own: intensity in [0, 255], t in [0, 0.1]
    L = txt.split("\n")
    i0 = L.index("SCALARS intensity float") + 2
    inten = np.array(L[i0:i0 + len(m)], dtype=np.float64)
    assert inten.min() >= 0 and inten.max() <= 255 and inten.std() > 10
