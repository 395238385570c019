"""Development aid: small registrations through the loop kernel's newer paths (Robust scale selects with and without the predicted
window, Robust next to a quantile filter, k > 1, a reading smaller than the grid), for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d = synth.make_pair_3d(n_map=30_000, n_scan=int(os.environ.get("NQ", "6000")), seed=5)
small = synth.make_pair_3d(n_map=30_000, n_scan=90, seed=6)
mad = ("robust", dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad"))
cases = [
    ("mad", d, dict(knn=1, outliers=(mad,), max_iteration_count=12)),
    ("berg", d, dict(knn=1, outliers=(("robust", dict(robustFct="cauchy", tuning=0.05, scaleEstimator="berg")),), max_iteration_count=4)),
    ("trimmed+mad", d, dict(knn=1, outliers=(("trimmed", 0.8), mad), max_iteration_count=12)),
    ("knn3 trimmed", d, dict(knn=3, outliers=(("trimmed", 0.8),), max_iteration_count=6)),
    ("knn3 mad", d, dict(knn=3, outliers=(mad,), max_iteration_count=6)),
    ("tiny mad", small, dict(knn=1, outliers=(mad,), max_iteration_count=6)),
]
for name, data, kw in cases:
    g = ICP(make_config(dim=3, max_dist=1.0, minimizer="point_to_plane", **kw))
    g.set_map(data["map"], data["normals"])
    T = g(data["reading"])
    print(name, "iterations", g.last_result.iterations, "pose error", synth.pose_error(T, data["correction_true"]), flush=True)
    g.close()
