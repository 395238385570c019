"""Development aid: per-phase globaltimer stamps of the last ICP iteration (build with EXTRA=-DB200ICP_STAMPS)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d = synth.make_pair_3d()
cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=int(os.environ.get("ITERS", "30")))
g = ICP(cfg); g.set_map(d["map"], d["normals"])
names = {0: "sel start", 1: "sel p0 hist", 2: "sel p0 sync", 3: "sel p0 pick", 4: "sel p1 hist", 5: "sel p1 sync", 6: "sel p1 pick",
         7: "sel p2 hist", 8: "sel p2 sync", 9: "sel p2 pick", 10: "sel end", 11: "acc start(b0)", 12: "acc loop done(b0)",
         13: "acc last-block ticket", 14: "acc reduced", 15: "acc finished", 16: "nn start(b0)", 17: "nn end(last block)"}
if "--loop" in sys.argv:
    names = {20: "iter start", 21: "nn done", 22: "published / level-0 hist flushed", 23: "barrier 1", 24: "pick 0", 25: "candidates gathered", 26: "barrier 2",
             27: "select done", 31: "acc partial written", 12: "acc barrier", 13: "reduced", 14: "finished", 28: "classified", 29: "counts read", 15: "solved", 16: "delta built", 17: "V + list done", 18: "S done (warp 0)", 19: "S done (CTA 0)", 30: "new matches fetched (thread 0)",
             1: "finish: keys shared", 2: "finish: ranked", 3: "finish: candidates summed (REDUX)", 4: "finish: warp totals shared", 6: "finish: totals", 5: "solve: entered", 7: "solve: sums loaded", 8: "solve: factorised", 9: "solve: substituted"}
for rep in range(3):
    g(d["reading"])
    st = np.zeros(32, np.uint64)
    g._L.b200icp_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    g._L.b200icp_debug_stamps(g._h, st.ctypes.data)
    order = sorted([i for i in names if st[i] > 0], key=lambda i: st[i])
    t0 = st[order[0]]
    print("rep", rep, "total_ms", g.timing().total_ms)
    prev = t0
    for i in order:
        print(f"   {names[i]:24s} +{(int(st[i]) - int(t0)) / 1e3:8.2f} us  (d {(int(st[i]) - int(prev)) / 1e3:6.2f})")
        prev = st[i]
