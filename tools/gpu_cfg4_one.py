"""Development aid: a few cfg-4 registrations (2-D, knn 8, point-to-point) for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d2 = synth.make_pair_2d()
cfg = make_config(dim=2, knn=8, max_dist=0.5, outliers=(), minimizer="point_to_point", max_iteration_count=30)
g = ICP(cfg); g.set_map(d2["map"], d2["normals"])
for _ in range(3):
    g(d2["reading"])
print("total_ms", g.timing().total_ms)
