"""Development aid: SurfaceNormal full pass (self k-NN + covariance) on a cfg-2-size map, staged tiles vs shell walk.
usage: gpu_normals_ab.py [n_map] [knn]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
knn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
d = synth.make_pair_3d(n_map=n_map, n_scan=1000)
for variant in (0x200000, 0):  # bit 20: staged tiles; bit 21: shell walk without the speculative bound; 0: default
    g = ICP(make_config(dim=3, knn=1, max_dist=1.0, outliers=(), minimizer="point_to_point", max_iteration_count=5, nn_variant=variant))
    g.set_map(d["map"], None)
    os.environ["B200ICP_FULL_NORMALS"] = "1"
    ts = []
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.map_surface_normals(knn)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"variant {variant:#x}: full normals pass {np.median(ts[1:]):.3f} ms (first {ts[0]:.3f}), redone by the shell walk {g.debug_selfknn_redone()}", flush=True)
    g.close()
