"""Generates tests/golden/example_pair.npz from the reference's bundled example data
(/root/reference/examples/data/scans: the only real-world input the reference ships; it holds no
expected outputs).  Run in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs  : scans 0 and 1 (lexicographic order, examples/build_map_from_scans_and_trajectory.cpp:191),
          sub-sampled with a fixed seed, the reference's BoundingBox input filters applied
          (examples/config.yaml:1-17), placed at their trajectory.csv poses.
Expected: produced by tests/numpy_icp.py (float64 numpy + scipy.spatial.cKDTree) -- an implementation
          independent of oracle/icp_oracle.c and of the CUDA path: exact k-NN ids / squared distances,
          surface normals (knn 10), and the point-to-plane correction of the documented ICP chain
          (docs/MapperConfiguration.md:172-189 with epsilon 0), for a known perturbation.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy_icp  # noqa: E402
from norlab_icp_mapper_b200 import synth  # noqa: E402

REF = "/root/reference/examples/data"


def read_vtk_points(path):
    with open(path) as f:
        lines = f.read().split("\n")
    for i, ln in enumerate(lines):
        if ln.startswith("POINTS"):
            n = int(ln.split()[1])
            vals = np.array(" ".join(lines[i + 1:i + 1 + n]).split(), np.float64)
            return vals.reshape(n, 3)
    raise ValueError("no POINTS section")


def read_poses(path):
    rows = open(path).read().strip().split("\n")[1:]
    poses = []
    for r in rows:
        c = r.split(",")
        t = np.array([float(c[4]), float(c[5]), float(c[6])])
        x, y, z, w = float(c[7]), float(c[8]), float(c[9]), float(c[10])
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = t
        poses.append(T)
    return poses


def bounding_box_remove_inside(p, lo, hi):
    inside = np.all((p >= lo) & (p <= hi), axis=1)
    return p[~inside]


def main():
    files = sorted(os.listdir(os.path.join(REF, "scans")))
    poses = read_poses(os.path.join(REF, "trajectory.csv"))
    rng = np.random.default_rng(20231017)
    clouds = []
    for k in (0, 1):
        p = read_vtk_points(os.path.join(REF, "scans", files[k]))
        p = bounding_box_remove_inside(p, np.array([-1.5, -1, -1]), np.array([0.5, 1, 0.5]))
        p = bounding_box_remove_inside(p, np.array([-6, -2.5, -1]), np.array([-1.5, 2.5, 1]))
        p = p[np.linalg.norm(p, axis=1) < 60.0]
        clouds.append(p)
    map_pts = synth.apply_T(poses[0], clouds[0][rng.choice(len(clouds[0]), 12000, replace=False)])
    scan = clouds[1][rng.choice(len(clouds[1]), 4000, replace=False)]
    map_pts = map_pts.astype(np.float32).astype(np.float64)
    scan = scan.astype(np.float32).astype(np.float64)
    normals = numpy_icp.surface_normals(map_pts, 10)
    T_est = poses[1] @ synth.make_T((0.10, -0.05, 0.02), (0.0, 0.0, 1.0))
    reading = synth.apply_T(T_est, scan).astype(np.float32).astype(np.float64)

    ids6, d6 = numpy_icp.knn(map_pts, reading, 6, max_dist=2.0)
    ids1, d1 = numpy_icp.knn(map_pts, reading, 1)
    T_k6 = numpy_icp.icp(map_pts, normals, reading, knn_k=6, max_dist=2.0, outliers=(), minimizer="point_to_plane", iterations=10)
    T_k1 = numpy_icp.icp(map_pts, normals, reading, knn_k=1, max_dist=2.0, outliers=(("trimmed", 0.85),),
                         minimizer="point_to_plane", iterations=30)
    T_p2p = numpy_icp.icp(map_pts, normals, reading, knn_k=1, max_dist=2.0, outliers=(("trimmed", 0.85),),
                          minimizer="point_to_point", iterations=30)
    # (the clouds go through float32 in the file: the robust correction is computed from what the file holds)
    m32, r32 = synth.homog(map_pts)[:, :3].astype(np.float64), synth.homog(reading)[:, :3].astype(np.float64)
    T_rob = numpy_icp.icp(m32, normals.astype(np.float32), r32, knn_k=1, max_dist=2.0,
                          outliers=(("robust", dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),),
                          minimizer="point_to_plane", iterations=10)
    out = os.path.join(HERE, "example_pair.npz")
    np.savez_compressed(out, map=synth.homog(map_pts), normals=normals.astype(np.float32), reading=synth.homog(reading),
                        knn6_ids=ids6.astype(np.int32), knn6_d2=d6, knn1_ids=ids1.astype(np.int32), knn1_d2=d1,
                        T_plane_k6_it10=T_k6, T_plane_trim_it30=T_k1, T_point_trim_it30=T_p2p, T_plane_robust_cauchy_mad_it10=T_rob,
                        scan_files=np.array(files[:2]))
    print("wrote", out, os.path.getsize(out), "bytes")
    print("correction k6:\n", T_k6)


if __name__ == "__main__":
    main()
