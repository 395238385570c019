"""Development aid: A/B of nn_variant values on cfg 2 with the bench's L2-flush protocol.  usage: gpu_ab.py v1 v2 ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
_seed = int(os.environ.get("SEED", "1234"))
d = synth.make_pair_3d(seed=_seed, drpy_deg=(1.0, -1.0, 3.0)) if os.environ.get("HARD") else synth.make_pair_3d(seed=_seed)  # HARD=1: SURVEY 8d's initial error
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
variants = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [0]
sort_modes = (1, 0) if "--sort" in sys.argv else (1,)
for rep in range(2):
    for v, srt in [(v, srt) for v in variants for srt in sort_modes]:
        cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30, nn_variant=v,
                          sort_reading=srt)
        g = ICP(cfg); g.set_map(d["map"], d["normals"])
        for _ in range(3):
            g(d["reading"])
        for mode in ("flushed", "warm"):
            ts, ls = [], []
            for _ in range(15):
                if mode == "flushed":
                    flush.fill_(1)
                torch.cuda.synchronize()
                g(d["reading"]); tm = g.timing()
                ts.append(tm.total_ms); ls.append(tm.loop_kernel_ms)
            print(f"variant {v:4d} sort {srt} {mode:8s}: total {np.median(ts):.3f} ms  loop {np.median(ls):.3f} ms  rest {np.median(ts) - np.median(ls):.3f} ms", flush=True)
        g.close()
