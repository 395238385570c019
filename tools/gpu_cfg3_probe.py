"""Development probe: per-step timings on a dense online-mapping map (BASELINE config 3 style)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
world = synth.World3D(seed=2000, size=(1000.0, 200.0), n_boxes=200)
cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30, differential=(1e-3, 1e-3, 3))
g = ICP(cfg)
def T(f, *a, **k):
    t0 = time.perf_counter(); r = f(*a, **k); return 1e3 * (time.perf_counter() - t0), r
n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for i in range(n_scans):
    x = -450.0 + 2.0 * i
    T_true = synth.make_T((x, 30.0 * np.sin(x / 80.0), 1.5), (0, 0, 0))
    S, _ = world.sample(100_000, np.random.default_rng(2000 + i), noise=0.01, center=T_true[:3, 3], radius=80.0)
    inp = synth.homog(S)
    t_reg = 0.0
    if i > 0:
        t_reg, _ = T(g, inp)
    t_ins, (added, _) = T(g.map_insert_point_distance, inp, 0.05)
    t_com, _ = T(g.map_commit)
    t_nrm, _ = T(g.map_surface_normals, 10)
    if i % 5 == 4 or i == n_scans - 1:
        print(f"scan {i}: local {g.map_counts()[0]} grid {g.grid_info()} | register {t_reg:.2f} ms ({g.last_result.iterations} it) | insert {t_ins:.2f} (+{added}) | commit {t_com:.2f} | normals {t_nrm:.2f}", flush=True)
