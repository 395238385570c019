"""Development aid: device time of one cfg-2 registration for several quantile-window policies
(B200ICP_WINDOW="gain,floor,max", read at context creation) and lanes-per-query variants of the loop kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config

d = synth.make_pair_3d()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
cases = [("two barriers (variant 16)", 16, None, None), ("three barriers (variant 80)", 80, None, None), ("always search (variant 32)", 32, None, None), ("default", 0, None, None)]
for w in ("2,0.02,0.12", "2,0.005,0.12", "2,0.0005,0.12", "3,0.0015,0.2", "1.5,0.001,0.08"):
    cases.append(("window " + w, 0, w, None))
for mg in ("2,0.001,0.25", "3,0.005,0.25", "4,0.002,0.5", "3,0.002,0.1", "1.5,0.002,0.25", "3,0.02,0.5"):
    cases.append(("margin " + mg, 0, None, mg))
for name, variant, win, mg in cases:
    for key, val in (("B200ICP_WINDOW", win), ("B200ICP_MARGIN", mg)):
        if val is None:
            os.environ.pop(key, None)
        else:
            os.environ[key] = val
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30,
                      nn_variant=variant)
    g = ICP(cfg)
    g.set_map(d["map"], d["normals"])
    for _ in range(3):
        T = g(d["reading"])
    ts, ls, ss = [], [], []
    for _ in range(10):
        flush.fill_(1)
        torch.cuda.synchronize()
        T = g(d["reading"])
        tm = g.timing()
        ts.append(tm.total_ms); ls.append(tm.loop_total_ms); ss.append(tm.loop_search_ms_sum / max(tm.loop_iterations, 1))
    if ref is None:
        ref = T
    er, et = synth.pose_error(T, ref)
    print(f"{name:28s} total {np.median(ts):.3f} ms  loop {np.median(ls):.3f} ms  search/iter {1e3*np.median(ss):.2f} us  "
          f"fast iters {tm.loop_fast_iterations:2d}/{g.last_result.iterations}  searched {tm.loop_searched_queries:8d}  pairs {g.last_result.pairs_last_iter}  d_pose {er:.1e} rad {et:.1e} m", flush=True)
    g.close()
