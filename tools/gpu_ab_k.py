"""Development aid: k > 1 configurations (cfg 1: 3-D knn 6, cfg 4: 2-D knn 8) through the persistent loop kernel
(nn_variant 0) and the kernel-per-step path (nn_variant 4), same box, same call.  usage: gpu_ab_k.py [variants...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config

variants = [int(a) for a in sys.argv[1:]] or [0, 4]
d1 = synth.make_pair_3d(n_map=41_400, n_scan=41_339, world_size=(120.0, 120.0), n_boxes=14, scan_radius=60.0, dt=(0.10, -0.05, 0.02), drpy_deg=(0, 0, 1.0))
d4 = synth.make_pair_2d()
cases = (("cfg1_knn6_41k", d1, dict(dim=3, knn=6, max_dist=2.0, outliers=(), minimizer="point_to_plane", max_iteration_count=10)),
         ("cfg4_2d_knn8", d4, dict(dim=2, knn=8, max_dist=0.5, outliers=(), minimizer="point_to_point", max_iteration_count=30)),
         ("3d_knn3_trimmed", d1, dict(dim=3, knn=3, max_dist=1.5, outliers=(("trimmed", 0.9),), minimizer="point_to_plane", max_iteration_count=15)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for label, d, kw in cases:
    poses = {}
    for rep in range(2):
        for v in variants:
            g = ICP(make_config(nn_variant=v, **kw))
            g.set_map(d["map"], d.get("normals"))
            for _ in range(3):
                T = g(d["reading"])
            poses[v] = T
            ts, ls, ss = [], [], []
            for _ in range(15):
                flush.fill_(1)
                torch.cuda.synchronize()
                g(d["reading"])
                tm = g.timing()
                ts.append(tm.total_ms)
                ls.append(tm.loop_kernel_ms)
                ss.append(tm.loop_search_ms_sum)
            print(f"{label:16s} variant {v:3d}: total {np.median(ts):.3f} ms  loop {np.median(ls):.3f} ms (V+S {np.median(ss):.3f})  iters {tm.loop_iterations}"
                  f"  fast {tm.loop_fast_iterations} two-barrier {tm.loop_two_barrier_iterations} searched {tm.loop_searched_queries}", flush=True)
            g.close()
    vs = list(poses)
    for v in vs[1:]:
        er, et = synth.pose_error(poses[vs[0]], poses[v])
        print(f"{label:16s} pose diff variant {vs[0]} vs {v}: {er:.2e} rad {et:.2e} m", flush=True)
