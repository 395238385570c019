"""Development probe: host-pointer registration (pinned / pageable) vs device-pointer registration."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
from norlab_icp_mapper_b200._abi import Result
d = synth.make_pair_3d()
cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
g = ICP(cfg); g.set_map(d["map"], d["normals"])
pinned = torch.from_numpy(d["reading"]).pin_memory()
dev = pinned.cuda()
T_out = np.zeros(16, np.float32); res = Result()
def run(ptr, device):
    f = g._L.b200icp_register_device if device else g._L.b200icp_register
    t0 = time.perf_counter(); rc = f(g._h, ptr, 4, len(d["reading"]), None, T_out.ctypes.data, ctypes.byref(res)); dt = time.perf_counter() - t0
    assert rc == 0
    return 1e3 * dt, g.timing().total_ms, g.timing().loop_total_ms
for name, ptr, device in (("device", dev.data_ptr(), True), ("pinned", pinned.data_ptr(), False), ("pageable", d["reading"].ctypes.data, False)):
    for _ in range(3): run(ptr, device)
    r = [run(ptr, device) for _ in range(10)]
    print(name, "wall ms", round(min(x[0] for x in r), 3), "event total ms", round(min(x[1] for x in r), 3), "loop kernel ms", round(min(x[2] for x in r), 3))
