"""Development probe: time the steps of Map::updateLocalPointCloud on the device map."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = synth.make_pair_3d(n_map=n_map, n_scan=100_000)
cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
g = ICP(cfg)
inp = synth.homog(synth.apply_T(d["correction_true"], d["reading"]))
def T(f, *a, **k):
    t0 = time.perf_counter(); r = f(*a, **k); return 1e3 * (time.perf_counter() - t0), r
for rep in range(3):
    t_set, _ = T(g.set_map, d["map"], None)
    t_ins, (added, _) = T(g.map_insert_point_distance, inp, 0.05)
    t_com, _ = T(g.map_commit)
    t_nrm, _ = T(g.map_surface_normals, 10)
    t_reg, _ = T(g, d["reading"])
    t_tr, _ = T(g.transform, d["reading"], d["T_est"].astype(np.float32))
    print(f"n_map={n_map} set_map {t_set:.2f} ms (device build {g.timing().setmap_ms:.2f}) | insert {t_ins:.2f} (added {added}) | commit {t_com:.2f} | normals(knn 10) {t_nrm:.2f} | register {t_reg:.2f} | transform(100k, host round trip) {t_tr:.2f}", flush=True)
