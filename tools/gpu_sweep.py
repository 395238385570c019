"""Development sweep on the GPU box: k-NN kernel variants x cell sizes on BASELINE config 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config

d = synth.make_pair_3d()
cpps = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["4"])]
variants = [int(x, 0) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
ref_T = None
for cpp in cpps:
    os.environ["B200ICP_CELLS_PER_POINT"] = str(cpp)
    for variant in variants:
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane",
                          max_iteration_count=30, nn_variant=variant & 0xffff, sort_reading=0 if (variant & 0x100000) else 1)
        g = ICP(cfg); g.set_profiling("--prof" in sys.argv); g.set_map(d["map"], d["normals"])
        ts = []
        for rep in range(4):
            T = g(d["reading"]); tm = g.timing(); ts.append((tm.total_ms, 1e3 * tm.nn_ms_sum / max(tm.nn_launches, 1), 1e3 * tm.select_ms_sum / max(tm.nn_launches, 1), 1e3 * tm.acc_ms_sum / max(tm.nn_launches, 1)))
        if ref_T is None:
            ref_T = T
        er = synth.pose_error(T, ref_T)
        print(f"cpp={cpp:5.1f} h={g.grid_info()[0]:.3f} variant={variant:#05x}: total_ms {min(t[0] for t in ts):.3f} nn_us/launch {min(t[1] for t in ts):.2f} sel_us {min(t[2] for t in ts):.2f} acc_us {min(t[3] for t in ts):.2f} "
              f"setmap_ms {g.timing().setmap_ms:.2f} pose-diff-vs-first {er[0]:.1e} {er[1]:.1e}", flush=True)
        g.close()
