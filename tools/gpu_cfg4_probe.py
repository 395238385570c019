"""Development aid: cfg 4 (2-D, 10k-pt scan, knn 8, point-to-point, 30 iterations): per-kernel split and cell-size sweep."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d2 = synth.make_pair_2d()
print("map", len(d2["map"]), "scan", len(d2["reading"]))
for cpp in (None, "1", "16", "64"):
    if cpp is None:
        os.environ.pop("B200ICP_CELLS_PER_POINT", None)
    else:
        os.environ["B200ICP_CELLS_PER_POINT"] = cpp
    for knn in (8, 1):
        cfg = make_config(dim=2, knn=knn, max_dist=0.5, outliers=(), minimizer="point_to_point", max_iteration_count=30)
        g = ICP(cfg); g.set_map(d2["map"], d2["normals"])
        edge, dims = g.grid_info() if hasattr(g, "grid_info") else (0, 0)
        for _ in range(3):
            g(d2["reading"])
        t0 = time.perf_counter()
        for _ in range(10):
            g(d2["reading"])
        ms = 1e2 * (time.perf_counter() - t0)
        g.set_profiling(True)
        g(d2["reading"])
        tm = g.timing()
        g.set_profiling(False)
        print(f"cells/pt {cpp} knn {knn}: {ms:.3f} ms/scan  edge {edge} dims {dims} setmap {tm.setmap_ms:.2f} ms | profiled: nn {tm.nn_ms_sum:.3f} ms over {tm.nn_launches} "
              f"launches, select {tm.select_ms_sum:.3f}, acc {tm.acc_ms_sum:.3f}, total {tm.total_ms:.3f}", flush=True)
        g.close()
