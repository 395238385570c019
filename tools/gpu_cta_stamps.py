"""Development aid: per-CTA %globaltimer stamps of the loop kernel's last iteration (stamped build:
make -C norlab_icp_mapper_b200/csrc stamps; B200ICP_LIB=libb200icp_stamps.so).  Shows how far apart the CTAs reach the
barriers: iteration start, V/S done, stage-1 barrier in / out, final barrier in / out, iteration end."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config
d = synth.make_pair_3d()
names = ["iter start", "V/S done", "stage-1 barrier in", "stage-1 barrier out", "final barrier in", "final barrier out", "iter end"]
for iters in [int(a) for a in sys.argv[1:]] or [3, 6, 10, 30]:
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=iters)
    g = ICP(cfg); g.set_map(d["map"], d["normals"])
    for _ in range(3):
        g(d["reading"])
    st = np.zeros((148, 32), np.uint64)
    g._L.b200icp_debug_cta_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    g._L.b200icp_debug_cta_stamps(g._h, st.ctypes.data, 148)
    t = st[:, :7].astype(np.int64)
    t0 = t[:, 0].min()
    print(f"== last of {iters} iterations (us relative to the earliest CTA's iteration start): min / median / p90 / max over the 148 CTAs")
    for k, nm in enumerate(names):
        col = (t[:, k] - t0) / 1e3
        if (t[:, k] < t0).any():
            print(f"   {nm:22s} (not reached in this iteration)")
            continue
        print(f"   {nm:22s} {col.min():7.2f} {np.median(col):7.2f} {np.percentile(col, 90):7.2f} {col.max():7.2f}   slowest CTA {int(col.argmax())}")
    dur = (t[:, 1] - t[:, 0]) / 1e3
    order = np.argsort(dur)
    print("   V/S duration: fastest CTAs", [(int(i), round(float(dur[i]), 1)) for i in order[:4]], "slowest", [(int(i), round(float(dur[i]), 1)) for i in order[-6:]])
    g.close()
