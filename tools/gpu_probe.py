"""Development probe run on the GPU box: parity of the CUDA path against the oracle plus timings.
Not part of the product; the real checks live in tests/ (-m gpu) and bench.py."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_binding as ob
from norlab_icp_mapper_b200 import synth
from norlab_icp_mapper_b200.icp import ICP, make_config

out = {}
def log(*a):
    print(*a, flush=True)

rng = np.random.default_rng(0)
# ---- 1. raw knn parity ------------------------------------------------------------------------
for (nref, nq, k, r, dim) in [(50000, 5000, 1, np.inf, 3), (50000, 5000, 6, 2.0, 3), (50000, 5000, 10, np.inf, 3),
                              (20000, 3000, 8, 0.5, 2), (20000, 3000, 32, np.inf, 3), (1000, 500, 1, 0.05, 3)]:
    ref = np.c_[rng.uniform(-20, 20, (nref, dim)), np.ones(nref)].astype(np.float32)
    if dim == 3:
        ref[:, 2] *= 0.1
    q = np.c_[rng.uniform(-25, 25, (nq, dim)), np.ones(nq)].astype(np.float32)
    if dim == 3:
        q[:, 2] *= 0.1
    cfg = make_config(dim=dim, knn=k, max_dist=float(r), outliers=(), minimizer="point_to_point")
    icp = ICP(cfg)
    t = time.time(); ids, d2 = icp.knn(ref, q, k, dim=dim, max_radius=float(r)); t = time.time() - t
    oi, od = ob.knn(ref, q, k, dim=dim, max_radius=float(r))
    same_d = np.array_equal(d2, od)
    same_i = (ids == oi).mean()
    log(f"knn nref={nref} nq={nq} k={k} r={r} dim={dim}: d2 bit-equal={same_d} ids equal={same_i:.5f} t={t*1e3:.1f}ms")
    if not same_d:
        bad = np.argwhere(d2 != od)
        log("  first mismatches", bad[:5], d2[bad[0][0]], od[bad[0][0]], ids[bad[0][0]], oi[bad[0][0]])
    icp.close()

# ---- 2. ICP parity, mid size ----------------------------------------------------------------------
def run_pair(d, cfg, label, trace=True):
    o = ob.OracleICP(cfg)
    o.set_map(d["map"], d["normals"])
    rc_o, T_o, res_o, tr_o, secs = o.register(d["reading"], want_trace=True)
    g = ICP(cfg)
    g.set_trace(True); g.set_profiling(True)
    g.set_map(d["map"], d["normals"])
    log(f"[{label}] grid", g.grid_info(), "mean gpu", g.map_mean(), "mean oracle", o.mean(), "setmap_ms", g.timing().setmap_ms)
    mi, md = g.match(d["reading"]); rc, oi, od = o.match(d["reading"])
    log(f"[{label}] match d2 bit-equal={np.array_equal(md, od)} ids equal={(mi == oi).mean():.5f}")
    try:
        T_g = g(d["reading"])
    except Exception as e:
        log(f"[{label}] GPU register failed: {e}")
        return
    res_g = g.last_result
    tr_g = g.trace()
    er, et = synth.pose_error(T_g, T_o)
    log(f"[{label}] oracle rc={rc_o} it={res_o.iterations} ov={res_o.overlap:.5f} pairs={res_o.pairs_last_iter} | gpu it={res_g.iterations} ov={res_g.overlap:.5f} pairs={res_g.pairs_last_iter}")
    log(f"[{label}] pose diff gpu-vs-oracle: {er:.3e} rad {et:.3e} m ; vs truth gpu {synth.pose_error(T_g, d['correction_true'])} oracle {synth.pose_error(T_o, d['correction_true'])}")
    n = min(len(tr_g), len(tr_o))
    for i in [0, 1, 2, n // 2, n - 1]:
        if 0 <= i < n:
            log(f"    it{i}: trace diff {synth.pose_error(tr_g[i], tr_o[i])}")
    tm = g.timing()
    log(f"[{label}] gpu total_ms={tm.total_ms:.3f} nn_ms_sum={tm.nn_ms_sum:.3f} nn_launches={tm.nn_launches} launches={tm.kernel_launches}; oracle secs={secs}")
    out[label] = dict(total_ms=tm.total_ms, nn_ms=tm.nn_ms_sum, nn_launches=tm.nn_launches, err_rad=er, err_m=et, oracle_s=float(secs[3]))
    return g

d = synth.make_pair_3d(n_map=200_000, n_scan=20_000)
cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
run_pair(d, cfg, "3d-200k")
cfg = make_config(dim=3, knn=6, max_dist=2.0, outliers=(("max_dist", 1.0),), minimizer="point_to_plane", max_iteration_count=10)
run_pair(d, cfg, "3d-200k-k6")
cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_point", max_iteration_count=40, differential=(1e-3, 1e-3, 3))
run_pair(d, cfg, "3d-200k-p2p-diff")
d2 = synth.make_pair_2d()
cfg = make_config(dim=2, knn=8, max_dist=0.5, outliers=(), minimizer="point_to_point", max_iteration_count=30)
run_pair(d2, cfg, "2d-p2p-k8")
cfg = make_config(dim=2, knn=1, max_dist=0.5, outliers=(("median", 3.0),), minimizer="point_to_plane", max_iteration_count=30)
run_pair(d2, cfg, "2d-p2plane-median")

# ---- 3. cfg 2 full size ---------------------------------------------------------------------------
if "--full" in sys.argv:
    d = synth.make_pair_3d()
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
    g = run_pair(d, cfg, "cfg2")
    for variant in [0, 1, 0x40, 0x10, 0x20]:
        for sort in [1, 0]:
            cfgv = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane",
                               max_iteration_count=30, nn_variant=variant, sort_reading=sort)
            gv = ICP(cfgv); gv.set_profiling(True); gv.set_map(d["map"], d["normals"])
            ts = []
            for rep in range(5):
                gv(d["reading"]); tm = gv.timing(); ts.append((tm.total_ms, tm.nn_ms_sum / max(tm.nn_launches, 1)))
            log(f"cfg2 variant={variant:#x} sort={sort}: total_ms {[round(t[0], 3) for t in ts]} nn_ms/launch {[round(t[1], 4) for t in ts]}")
            gv.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
