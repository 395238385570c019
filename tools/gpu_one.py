"""Development aid: N cfg-2 registrations (ITERS iterations each) -- a short command line for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d = synth.make_pair_3d()
cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane",
                  max_iteration_count=int(os.environ.get("ITERS", "30")), nn_variant=int(os.environ.get("VARIANT", "0")))
g = ICP(cfg); g.set_map(d["map"], d["normals"])
for _ in range(int(os.environ.get("REPS", "4"))):
    g(d["reading"])
print("total_ms", g.timing().total_ms, "loop_ms", g.timing().loop_total_ms)
