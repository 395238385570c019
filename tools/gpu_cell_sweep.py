"""Development aid: cfg-2 registration time vs the index's cell size (B200ICP_CELLS_PER_POINT, read by set_map)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d = synth.make_pair_3d()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for rep in range(2):
    for cpp in ("4", "2", "1", "0.5", "0.25", "8"):
        os.environ["B200ICP_CELLS_PER_POINT"] = cpp
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
        g = ICP(cfg); g.set_map(d["map"], d["normals"])
        edge, dims = g.grid_info()
        for _ in range(3):
            T = g(d["reading"])
        ts, ls = [], []
        for _ in range(10):
            flush.fill_(1); torch.cuda.synchronize()
            T = g(d["reading"]); tm = g.timing()
            ts.append(tm.total_ms); ls.append(tm.loop_kernel_ms)
        if ref is None:
            ref = T
        print(f"cells/pt {cpp:5s} edge {edge:.3f} m dims {dims}: total {np.median(ts):.3f} ms loop {np.median(ls):.3f} ms (cold part {np.median(ts) - np.median(ls):.3f}) "
              f"searched {tm.loop_searched_queries} same pose {np.array_equal(T, ref)}", flush=True)
        g.close()
