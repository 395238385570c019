"""Development aid: per-iteration record of the persistent loop kernel on cfg 2 (path taken, quantile limit,
candidates, queries that needed a search, iteration time on CTA 0).  Env B200ICP_WINDOW / B200ICP_MARGIN apply."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
args = [a for a in sys.argv[1:] if not a.startswith("--")]
variant = (int(args[0]) if args else 0) | 0x80000  # bit 19: write the per-iteration record
if "--cfg1" in sys.argv:
    d = synth.make_pair_3d(n_map=41_400, n_scan=41_339, world_size=(120.0, 120.0), n_boxes=14, scan_radius=60.0, dt=(0.10, -0.05, 0.02), drpy_deg=(0, 0, 1.0))
    cfg = make_config(dim=3, knn=6, max_dist=2.0, outliers=(), minimizer="point_to_plane", max_iteration_count=10, nn_variant=variant)
elif "--cfg4" in sys.argv:
    d = synth.make_pair_2d()
    cfg = make_config(dim=2, knn=8, max_dist=0.5, outliers=(), minimizer="point_to_point", max_iteration_count=30, nn_variant=variant)
else:
    d = synth.make_pair_3d()
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30, nn_variant=variant)
g = ICP(cfg); g.set_map(d["map"], d.get("normals"))
for _ in range(3):
    g(d["reading"])
rec = np.zeros((30, 8), np.uint32)
g._L.b200icp_debug_loop_record.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
g._L.b200icp_debug_loop_record(g._h, rec.ctypes.data, 30)
tm = g.timing()
print(f"total {tm.total_ms:.3f} ms loop {tm.loop_total_ms:.3f} ms one-barrier {tm.loop_fast_iterations} two-barrier {tm.loop_two_barrier_iterations} searched {tm.loop_searched_queries}")
prev = 0
for i, r in enumerate(rec):
    lim = np.array([r[1]], np.uint32).view(np.float32)[0]
    lo, hi = np.array([r[5], r[6]], np.uint32).view(np.float32)
    print(f"it {i:2d} path {r[0]:2d} limit {lim:.6f} cand {r[2]:6d} below {r[3]:6d} searched {int(r[4]) - prev:7d}  next window [{lo:.6f}, {hi:.6f}]  {r[7] / 1e3:7.2f} us")
    prev = int(r[4])
