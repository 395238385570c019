"""Development aid: MedianDistOutlierFilter chain on cfg 2, windowed iterations (default) vs the general three-barrier path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d = synth.make_pair_3d()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for variant in (0, 16 | 64):
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("median", 3.0),), minimizer="point_to_plane", max_iteration_count=30, nn_variant=variant)
    g = ICP(cfg); g.set_map(d["map"], d["normals"])
    for _ in range(3):
        T = g(d["reading"])
    ts, ls = [], []
    for _ in range(10):
        flush.fill_(1); torch.cuda.synchronize()
        g(d["reading"]); tm = g.timing()
        ts.append(tm.total_ms); ls.append(tm.loop_kernel_ms)
    print(f"variant {variant:3d}: total {np.median(ts):.3f} ms loop {np.median(ls):.3f} ms one-barrier {tm.loop_fast_iterations} two-barrier {tm.loop_two_barrier_iterations} "
          f"pairs {g.last_result.pairs_last_iter} err {synth.pose_error(T, d['correction_true'])}", flush=True)
    g.close()
