"""Development aid: a small RobustOutlierFilter registration through the loop kernel, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d = synth.make_pair_3d(n_map=30_000, n_scan=int(os.environ.get("NQ", "6000")), seed=5)
for est in ("mad", "berg"):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("robust", dict(robustFct="cauchy", tuning=1.0 if est == "mad" else 0.05, scaleEstimator=est)),),
                      minimizer="point_to_plane", max_iteration_count=4)
    g = ICP(cfg)
    g.set_map(d["map"], d["normals"])
    T = g(d["reading"])
    print(est, "iterations", g.last_result.iterations, "pose error", synth.pose_error(T, d["correction_true"]), flush=True)
    g.close()
