"""Secondary measurements for the other BASELINE.json configs (1, 3, 4, 5): GPU path vs the CPU oracle
on the same host.  Not the headline bench (bench.py = config 2); writes gpurun_out/configs.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_binding as ob
from norlab_icp_mapper_b200 import synth, batched
from norlab_icp_mapper_b200.icp import ICP, make_config
from norlab_icp_mapper_b200.mapper import Mapper

out = {}
quick = "--quick" in sys.argv
small = "--only-small" in sys.argv  # configs 1 and 4 only


def time_pair(label, d, cfg, reps=10, oracle_reps=2):
    g = ICP(cfg); g.set_map(d["map"], d["normals"])
    for _ in range(3):
        T = g(d["reading"])
    t0 = time.perf_counter()
    for _ in range(reps):
        T = g(d["reading"])
    gpu_ms = 1e3 * (time.perf_counter() - t0) / reps
    dev_ms = g.timing().total_ms
    setmap_ms = g.timing().setmap_ms
    g.close()
    o = ob.OracleICP(cfg)
    t0 = time.perf_counter(); o.set_map(d["map"], d["normals"]); build_s = time.perf_counter() - t0
    o.register(d["reading"])
    t0 = time.perf_counter()
    for _ in range(oracle_reps):
        rc, To, res, _, _ = o.register(d["reading"])
    cpu_ms = 1e3 * (time.perf_counter() - t0) / oracle_reps
    e = synth.pose_error(T, To)
    out[label] = dict(gpu_ms=gpu_ms, gpu_device_ms=dev_ms, gpu_setmap_ms=setmap_ms, cpu_ms=cpu_ms, cpu_build_s=build_s, speedup=cpu_ms / gpu_ms,
                      pose_diff_rad=e[0], pose_diff_m=e[1], cpu_threads=ob.lib().orc_num_threads(), iterations=res.iterations)
    print(label, json.dumps(out[label]), flush=True)


only3 = "--only-cfg3" in sys.argv
# config 1-like: 41k scan vs 41k map, knn 6, maxDist 2, point-to-plane, 10 iterations (docs/MapperConfiguration.md:172-189)
d = None if only3 else synth.make_pair_3d(n_map=41_400, n_scan=41_339, world_size=(120.0, 120.0), n_boxes=14, scan_radius=60.0, dt=(0.10, -0.05, 0.02), drpy_deg=(0, 0, 1.0))
if not only3: time_pair("cfg1_knn6_41k", d, make_config(dim=3, knn=6, max_dist=2.0, outliers=(), minimizer="point_to_plane", max_iteration_count=10))
# config 4: 2-D, 10k-pt scans, point-to-point, dense k = 8
d2 = None if only3 else synth.make_pair_2d()
if not only3: time_pair("cfg4_2d_knn8", d2, make_config(dim=2, knn=8, max_dist=0.5, outliers=(), minimizer="point_to_point", max_iteration_count=30))
# config 2 for reference
if not quick and not only3 and not small:
    time_pair("cfg2", synth.make_pair_3d(), make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30))

# config 5: batched pairs (200k scan vs 1M submap) on this GPU; 8 pairs = one GPU's share of the 64
n_pairs = 0 if (only3 or small) else (4 if quick else 8)
cfg5 = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30)
pairs = [synth.make_pair_3d(n_map=1_000_000, n_scan=200_000, seed=4000 + j) for j in range(n_pairs)]
if n_pairs:
  fn = batched.gpu_register_fn(cfg5, 0)
  batched.register_batch(lambda j: pairs[j], 1, fn)
  t0 = time.perf_counter()
  poses, ov, it = batched.register_batch(lambda j: pairs[j], n_pairs, fn)
  t_gpu = time.perf_counter() - t0
  fn.close()
  o = ob.OracleICP(cfg5)
  t0 = time.perf_counter(); o.set_map(pairs[0]["map"], pairs[0]["normals"]); rc, To, res, _, _ = o.register(pairs[0]["reading"]); t_cpu = time.perf_counter() - t0
  e = synth.pose_error(poses[0], To)
  out["cfg5_batched"] = dict(pairs=n_pairs, gpu_pairs_per_s=n_pairs / t_gpu, gpu_ms_per_pair=1e3 * t_gpu / n_pairs, cpu_ms_per_pair=1e3 * t_cpu,
                           speedup=t_cpu / (t_gpu / n_pairs), pose_diff_rad=e[0], pose_diff_m=e[1], note="per pair: set_map (index build) + 30-iteration ICP; CPU: kd-tree build + ICP")
if n_pairs: print("cfg5", json.dumps(out["cfg5_batched"]), flush=True)

# config 3: online mapping through the host mirror of Mapper::processInput
if small:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_small.json"), "w"), indent=1)
    sys.exit(0)
n_scans = 30 if quick else 120
for a in sys.argv:
    if a.startswith("--cfg3-scans="):
        n_scans = int(a.split("=")[1])
world = synth.World3D(seed=2000, size=(1000.0, 200.0), n_boxes=200)
cfg3 = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30, differential=(1e-3, 1e-3, 3))
m = Mapper(cfg3, True, False, True, False, updateCondition=("distance", 1.0), sensorMaxRange=80.0, minDistNewPoint=0.05, surfaceNormalKnn=10, reservePoints=int(os.environ.get("CFG3_RESERVE", "32000000")))
rng = np.random.default_rng(5)
times, sizes, upd = [], [], []
T_prev_true = None
pose_est = None
for i in range(n_scans):
    x = -450.0 + 2.0 * i
    T_true = synth.make_T((x, 30.0 * np.sin(x / 80.0), 1.5), (0, 0, np.degrees(np.arctan2(30.0 / 80.0 * np.cos(x / 80.0), 1.0))))
    S, _ = world.sample(100_000, np.random.default_rng(2000 + i), noise=0.01, center=T_true[:3, 3], radius=80.0)
    scan = synth.homog(synth.apply_T(np.linalg.inv(T_true), S))
    if pose_est is None:
        T_est = T_true
    else:  # odometry increment with noise composed on the last corrected pose
        inc = np.linalg.inv(T_prev_true) @ T_true @ synth.make_T(rng.normal(0, 0.05, 3), rng.normal(0, 0.5, 3))
        T_est = pose_est @ inc
    t0 = time.perf_counter()
    m.processInput(scan, T_est.astype(np.float32), 0.1 * i)
    times.append(time.perf_counter() - t0)
    if os.environ.get("B200ICP_TRACE_ALLOC"):
        print(f"[scan {i}] {1e3 * times[-1]:.1f} ms", file=sys.stderr, flush=True)
    pose_est = m.getPose().astype(np.float64)
    T_prev_true = T_true
    st = m.stats()
    sizes.append((st.n_local, st.n_global)); upd.append(st.map_updated)
    if (i + 1) % 50 == 0:
        print(f"cfg3 scan {i + 1}: local {st.n_local} global {st.n_global} last {1e3 * times[-1]:.1f} ms median {1e3 * float(np.median(times)):.1f} ms", flush=True)
err = synth.pose_error(pose_est, T_true)
out["cfg3_online"] = dict(scans=n_scans, scans_per_s=n_scans / sum(times), ms_per_scan_median=1e3 * float(np.median(times)),
                          ms_per_update_scan=1e3 * float(np.mean([t for t, u in zip(times, upd) if u])), updates=int(sum(upd)),
                          ms_per_scan_p95=1e3 * float(np.percentile(times, 95)), ms_per_scan_mean=1e3 * float(np.mean(times)),
                          ms_slowest_scans=[round(1e3 * t, 1) for t in sorted(times)[-5:]], slowest_scan_indices=[int(i) for i in np.argsort(times)[-5:]], ms_first_scans=[round(1e3 * t, 1) for t in times[:12]], ms_per_scan_last50=1e3 * float(np.mean(times[-50:])),
                          final_local=sizes[-1][0], final_global=sizes[-1][1], drift_rad=err[0], drift_m=err[1],
                          note="whole Mapper::processInput per scan through the host mirror (upload, input filters, ICP with Counter{30} + "
                               "Differential, PointDistance insert, SurfaceNormal knn 10 over the local map, index rebuild); every scan updates the map")
print("cfg3", json.dumps(out["cfg3_online"]), flush=True)
m.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
