"""Development aid: RobustOutlierFilter chains, loop kernel (scale by in-kernel exact selects) against the kernel-per-step path
(nn_variant bit 26: two device sorts per iteration)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d = synth.make_pair_3d()
for est in ("mad", "berg", "none"):
    for name, variant in (("loop", 0), ("loop, no windows", 0x8000000), ("steps", 0x4000000)):
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("robust", dict(robustFct="cauchy", tuning=1.0 if est != "berg" else 0.05, scaleEstimator=est)),),
                          minimizer="point_to_plane", max_iteration_count=30, nn_variant=variant)
        g = ICP(cfg)
        g.set_map(d["map"], d["normals"])
        ts = []
        for rep in range(8):
            T = g(d["reading"])
            ts.append(g.timing().total_ms)
        er, et = synth.pose_error(T, d["correction_true"])
        print(f"robust cauchy/{est:5s} {name:16s}: total {np.median(ts[2:]):.3f} ms  iterations {g.last_result.iterations}  multi-barrier selects {g.timing().loop_two_barrier_iterations}  pose error {er:.2e} rad {et:.2e} m", flush=True)
        g.close()
