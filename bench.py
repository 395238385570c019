#!/usr/bin/env python
"""bench.py -- scans/sec of the ICP registration hot path on BASELINE.json config 2
("single-GPU ICP: synthetic 100k-pt scan vs 2M-pt map, point-to-plane, 30 iters").

One "step" = one `icp(input)` (Mapper.cpp:213): 30 ICP iterations of a 100 000-point reading against
the 2 000 000-point map already installed by setMap.  Prints ONE JSON line (rank 0).

  value     scans/s with the reading already resident in HBM (b200icp_register_device)
  e2e       scans/s through the host-pointer C-ABI call (b200icp_register): pinned host reading in,
            pose + result out, copies inside the timed region; `e2e.pageable` = the same from ordinary
            (pageable) host memory, which is what the reference's Eigen storage is
  roofline  the dominant kernel (the persistent loop kernel): algorithmic bytes per launch
            / its mean launch duration (CUDA events on the library's stream) against
            MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (restated libpointmatcher/libnabo path, OpenMP over queries like
            libnabo) on the same host: all cores and one thread, per-phase times, and a
            scipy cKDTree(workers=-1) cross-check of the oracle's nearest-neighbour speed
  batched   BASELINE.json config 5 through the product's batched entry point
            (b200icp_register_batch, four contexts per GPU, two when a rank holds fewer than 32 pairs): 64 pairs (200k-pt scan vs 1M-pt submap),
            pair j on rank j mod N -- STRONG scaling, the pose all_gather inside the timed region
  extra     (N = 1) the other configurations: cfg2_hard (SURVEY 8d's 1/-1/3 degree perturbation),
            cfg2_robust (RobustOutlierFilter cauchy / mad instead of TrimmedDist), cfg1-like, cfg4 (2-D,
            knn 8), cfg3 (500-scan online mapping), the search kernel on a 10 M-point map (> L2)

N > 1 (torchrun): the headline stays config 2: every rank registers its own scan against its own map
through the same entry points (independent pairs, no data-path collective) and the K poses of the
timed steps are all_gathered (NCCL) once, INSIDE the timed region; value = all scans / max-over-ranks
time ("weak").

--impl reference times the CPU oracle alone (the reference's own libpointmatcher build cannot be
compiled here: its dependencies are absent), rank 0 only, with ALL host cores
requested explicitly (torchrun exports OMP_NUM_THREADS=1).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "scans/sec (100k-pt scan vs 2M-pt map, point-to-plane, 30 iters)"
UNIT = "scans/s"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg2_hard", "cfg2_robust", "cfg1", "cfg4"],
                    help="the pair workload of the main line (default: BASELINE.json config 2)")
    ap.add_argument("--n-map", type=int, default=None)
    ap.add_argument("--n-scan", type=int, default=None)
    ap.add_argument("--iters", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` and `batched` sub-results")
    ap.add_argument("--pairs", type=int, default=64, help="pairs of the batched (config 5) measurement")
    ap.add_argument("--cfg3-scans", type=int, default=500)
    ap.add_argument("--nn-variant", type=int, default=None)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
WORKLOADS = {
    "cfg2": dict(label="cfg2: single-GPU ICP, synthetic 100k-pt scan vs 2M-pt map, point-to-plane, 30 iters",
                 dim=3, n_map=2_000_000, n_scan=100_000, iters=30, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),),
                 minimizer="point_to_plane", gen={},
                 icp="KDTreeMatcher{knn 1, maxDist 1.0, eps 0} + TrimmedDist{0.85} + PointToPlane + Counter{%d}",
                 world="world and map seed 1234, scan seed 1235+10*rank: 200x200 m ground + 40 boxes + 4 walls, 1 cm noise; scan within 80 m; "
                       "initial error (0.30,-0.20,0.10) m / (0.3,-0.3,1.0) deg = 0.37 m / 1.1 deg"),
    "cfg2_hard": dict(label="cfg2_hard: as cfg2 with SURVEY 8d's initial error (0.30,-0.20,0.10) m / (1,-1,3) deg",
                      dim=3, n_map=2_000_000, n_scan=100_000, iters=30, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),),
                      minimizer="point_to_plane", gen=dict(drpy_deg=(1.0, -1.0, 3.0)),
                      icp="KDTreeMatcher{knn 1, maxDist 1.0, eps 0} + TrimmedDist{0.85} + PointToPlane + Counter{%d}",
                      world="as cfg2; initial error 0.37 m / 3.3 deg (4 m at 80 m range against maxDist 1 m: the reference "
                            "algorithm itself does not converge in 30 iterations)"),
    "cfg2_robust": dict(label="cfg2_robust: as cfg2 with libpointmatcher's RobustOutlierFilter defaults instead of TrimmedDist",
                        dim=3, n_map=2_000_000, n_scan=100_000, iters=30, knn=1, max_dist=1.0,
                        outliers=(("robust", dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad")),),
                        minimizer="point_to_plane", gen={},
                        icp="KDTreeMatcher{knn 1, maxDist 1.0, eps 0} + RobustOutlierFilter{cauchy, tuning 1, scaleEstimator mad} + PointToPlane + Counter{%d}",
                        world="as cfg2 (two exact medians per iteration for the mad scale, inside the loop kernel)"),
    "cfg1": dict(label="cfg1-like: 41k-pt scan vs 41k-pt map, knn 6, maxDist 2, point-to-plane, 10 iters "
                       "(docs/MapperConfiguration.md:172-189 on synthetic clouds of the bundled scans' size)",
                 dim=3, n_map=41_400, n_scan=41_339, iters=10, knn=6, max_dist=2.0, outliers=(), minimizer="point_to_plane",
                 gen=dict(world_size=(120.0, 120.0), n_boxes=14, scan_radius=60.0, dt=(0.10, -0.05, 0.02), drpy_deg=(0, 0, 1.0)),
                 icp="KDTreeMatcher{knn 6, maxDist 2.0, eps 0} + PointToPlane + Counter{%d}",
                 world="120x120 m, 14 boxes; initial error (0.10,-0.05,0.02) m / 1 deg yaw"),
    "cfg4": dict(label="cfg4: 2-D, 10k-pt scan vs 200k-pt map, knn 8, maxDist 0.5, point-to-point, 30 iters",
                 dim=2, n_map=200_000, n_scan=10_000, iters=30, knn=8, max_dist=0.5, outliers=(), minimizer="point_to_point",
                 gen={}, icp="KDTreeMatcher{knn 8, maxDist 0.5, eps 0} + PointToPoint + Counter{%d}",
                 world="60x40 m polygonal room, seed 3000, 1 cm noise; initial error (0.15,-0.10) m / 2 deg"),
}


def resolve(args, name=None):
    w = dict(WORKLOADS[name or args.workload])
    if name is None or name == args.workload:
        for k, a in (("n_map", args.n_map), ("n_scan", args.n_scan), ("iters", args.iters)):
            if a is not None:
                w[k] = a
    return w


def workload_config(w):
    return {"workload": w["label"], "n_map": w["n_map"], "n_scan": w["n_scan"], "iterations": w["iters"],
            "icp": w["icp"] % w["iters"], "world": w["world"]}


def make_cfg(w, nn_variant=None):
    from norlab_icp_mapper_b200._abi import make_config
    kw = {}
    if nn_variant is not None:
        kw["nn_variant"] = nn_variant
    return make_config(dim=w["dim"], knn=w["knn"], max_dist=w["max_dist"], outliers=w["outliers"], minimizer=w["minimizer"],
                       max_iteration_count=w["iters"], **kw)


def make_data(w, rank=0):
    from norlab_icp_mapper_b200 import synth
    if w["dim"] == 2:
        return synth.make_pair_2d(n_map=w["n_map"], n_scan=w["n_scan"], seed=3000 + 10 * rank)
    # every rank: the same world and map, its own scan of it (equal work per GPU: the weak-scaling figure measures the GPUs, not the scenes)
    return synth.make_pair_3d(n_map=w["n_map"], n_scan=w["n_scan"], seed=1234, scan_seed=1235 + 10 * rank, **w["gen"])


def _gen_cfg5_pair(j):
    from norlab_icp_mapper_b200 import synth
    d = synth.make_pair_3d(n_map=1_000_000, n_scan=200_000, seed=4000 + j)
    return {k: d[k] for k in ("map", "normals", "reading", "correction_true")}


def _gen_cfg3_scan(i):
    from norlab_icp_mapper_b200 import synth
    world = _gen_cfg3_scan.world
    x = -450.0 + 2.0 * i
    T_true = synth.make_T((x, 30.0 * np.sin(x / 80.0), 1.5), (0, 0, np.degrees(np.arctan2(30.0 / 80.0 * np.cos(x / 80.0), 1.0))))
    S, _ = world.sample(100_000, np.random.default_rng(2000 + i), noise=0.01, center=T_true[:3, 3], radius=80.0)
    return synth.homog(synth.apply_T(np.linalg.inv(T_true), S)), T_true


def _gen_map_chunk(args):
    from norlab_icp_mapper_b200 import synth
    seed, n = args
    world = synth.World3D(seed=1234)
    P, N = world.sample(n, np.random.default_rng(seed), noise=0.01)
    return synth.homog(P), np.ascontiguousarray(N, np.float32)


def pool_map(fn, items, init=None):
    """Generate independent synthetic clouds on the host cores (fork pool; call BEFORE CUDA is initialised)."""
    import multiprocessing as mp
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nproc = max(1, min(len(items), host_cores() // max(world, 1), 32))
    if nproc == 1:
        if init:
            init()
        return [fn(x) for x in items]
    with mp.get_context("fork").Pool(nproc, initializer=init) as p:
        return p.map(fn, items)


def _init_cfg3_world():
    from norlab_icp_mapper_b200 import synth
    _gen_cfg3_scan.world = synth.World3D(seed=2000, size=(1000.0, 200.0), n_boxes=200)


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle (test infrastructure) timed as the reference arm / cpu_baseline
# ---------------------------------------------------------------------------------------------
def time_oracle(cfg, data, steps, warmup, nthreads):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    o = ob.OracleICP(cfg)
    t0 = time.perf_counter()
    o.set_map(data["map"], data["normals"])
    t_setmap = time.perf_counter() - t0
    for _ in range(warmup):
        o.register(data["reading"], nthreads=nthreads)
    times, phases = [], np.zeros(4)
    T, res = None, None
    for _ in range(steps):
        t0 = time.perf_counter()
        rc, T, res, _, secs = o.register(data["reading"], nthreads=nthreads)
        times.append(time.perf_counter() - t0)
        phases += secs
    total = sum(times)
    phases /= max(steps, 1)
    return dict(value=steps / total, ms_per_step=1e3 * total / steps, cores=nthreads, setmap_s=t_setmap, T=T,
                iterations=res.iterations,
                phases_ms={"kdtree_build_once": 1e3 * t_setmap, "nn_search": 1e3 * phases[0], "outlier_filters": 1e3 * phases[1],
                           "error_minimizer": 1e3 * phases[2], "register_total": 1e3 * phases[3]})


def scipy_crosscheck(data, dim, k, max_dist):
    """scipy.spatial.cKDTree(workers=-1) timing of ONE correspondence search of the workload (SURVEY 8d sanity check of
    the oracle's nearest-neighbour speed)."""
    try:
        from scipy.spatial import cKDTree
        t0 = time.perf_counter()
        tree = cKDTree(data["map"][:, :dim], leafsize=8)
        tb = time.perf_counter() - t0
        q = data["reading"][:, :dim]
        tree.query(q[:1000], k=k, distance_upper_bound=max_dist, workers=-1)
        t0 = time.perf_counter()
        tree.query(q, k=k, distance_upper_bound=max_dist, workers=-1)
        tq = time.perf_counter() - t0
        return {"build_ms": 1e3 * tb, "one_search_ms": 1e3 * tq, "what": "scipy.spatial.cKDTree(leafsize 8).query(workers=-1), one pass "
                "over the reading at its initial pose"}
    except Exception as e:  # scipy absent: not an error of the bench
        return {"error": repr(e)}


def oracle_one_search_ms(cfg, data, nthreads):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    o = ob.OracleICP(cfg)
    o.set_map(data["map"], data["normals"])
    o.match(data["reading"][:1000], nthreads=nthreads)
    t0 = time.perf_counter()
    o.match(data["reading"], nthreads=nthreads)
    return 1e3 * (time.perf_counter() - t0)


def cpu_baseline_block(w, cfg, data, n_all, n_one):
    cores = host_cores()
    r = time_oracle(cfg, data, n_all, 2, cores)
    one = time_oracle(cfg, data, n_one, 0, 1)
    blk = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d full scans of the same workload on %d threads after 2 warm-ups (kd-tree build %.2f s outside), and %d scan(s) "
                     "on 1 thread" % (n_all, cores, r["setmap_s"], n_one),
           "phases_ms": r["phases_ms"],
           "one_thread": {"value": one["value"], "cores": 1, "phases_ms": one["phases_ms"]},
           "what_is_parallel": "OpenMP over queries in the k-NN (what libnabo parallelises); kd-tree build, outlier filters and "
                               "error minimiser single-threaded as upstream",
           "oracle_one_search_ms": oracle_one_search_ms(cfg, data, cores),
           "scipy_ckdtree": scipy_crosscheck(data, w["dim"], w["knn"], w["max_dist"])}
    return blk, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = resolve(args)
    cfg = make_cfg(w)
    data = make_data(w, 0)
    cores = host_cores()  # explicit: torchrun exports OMP_NUM_THREADS=1, which made round 1's N >= 2 reference arm single-threaded
    r = time_oracle(cfg, data, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(w),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d full scans (each %d ICP iterations of %d points vs %d-point map) on %d threads; kd-tree build "
                                   "(%.2f s) outside the timed region" % (args.steps, w["iters"], w["n_scan"], w["n_map"], cores, r["setmap_s"]),
                         "phases_ms": r["phases_ms"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle restating libpointmatcher/libnabo (the reference's own build needs libpointmatcher, "
                "libnabo, Eigen, yaml-cpp, Boost: absent here); OpenMP over queries as libnabo does, rest single-threaded; "
                "one host process whatever --gpus is (the reference has no multi-GPU or batched mode)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for nm, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


class Env:
    """torch / distributed plumbing shared by the measurements."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_poses(self, recs):
        """The only collective of the replicated / batched modes: ONE all_gather of this rank's pose records (K x 18 floats) at the
        end of the K timed steps, inside the timed region -- north_star: "NCCL only to gather final poses"."""
        from norlab_icp_mapper_b200 import batched
        return batched.gather_records(np.ascontiguousarray(recs, np.float32), self.world, self.dist, self.device)

    def timed(self, fn, steps, ext, records=None):
        """Per-step CUDA events (start on the library's stream, end on torch's current stream); L2 flushed between steps, outside
        the events.  With several ranks the K pose records are all_gathered once after the last step; that collective is timed
        (host wall clock around the blocking call: NCCL launch + transfer + read-back) and added.  Returns (device ms, wall ms)
        summed over the steps (+ the gather)."""
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        wall = 0.0
        for i, (a, b) in enumerate(evs):
            self.flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.record(ext)
            rec = fn()
            b.record()
            b.synchronize()
            wall += time.perf_counter() - t0
            if records is not None and rec is not None:
                records[i] = rec
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        if self.world > 1 and records is not None:
            t0 = time.perf_counter()
            self.last_gather = self.gather_poses(records)
            torch.cuda.synchronize()
            g = (time.perf_counter() - t0) * 1e3
            dev_ms += g
            wall += g * 1e-3
        return dev_ms, wall * 1e3


def bench_pair(env, w, data, args, steps, warmup, with_roofline=True):
    """value / e2e / loop-kernel figures of one pair workload on this rank's GPU."""
    torch = env.torch
    from norlab_icp_mapper_b200.icp import ICP
    from norlab_icp_mapper_b200._abi import Result
    cfg = make_cfg(w, args.nn_variant)
    icp = ICP(cfg, device=env.local_rank)
    icp.set_map(data["map"], data["normals"])
    setmap_ms = icp.timing().setmap_ms
    ext = torch.cuda.ExternalStream(icp.stream(), device=env.device)
    rows = w["dim"] + 1
    nq = len(data["reading"])
    reading_pinned = torch.from_numpy(data["reading"]).pin_memory()
    reading_pageable = np.array(data["reading"], copy=True)
    d_reading = reading_pinned.to("cuda", non_blocking=False)
    res = Result()
    T_out = np.zeros(16, np.float32)
    nn = rows * rows
    state = {}

    def finish_step():  # this step's pose record (gathered over the ranks once, after the last timed step)
        rec = np.zeros(18, np.float32)
        rec[:nn] = T_out[:nn]
        rec[16], rec[17] = res.overlap, res.iterations
        return rec

    def step_device():
        rc = icp._L.b200icp_register_device(icp._h, d_reading.data_ptr(), rows, nq, None, T_out.ctypes.data, ctypes.byref(res))
        if rc != 0:
            raise RuntimeError(icp._L.b200icp_last_error(icp._h).decode())
        return finish_step()

    def make_e2e(ptr):
        def step():
            rc = icp._L.b200icp_register(icp._h, ptr, rows, nq, None, T_out.ctypes.data, ctypes.byref(res))
            if rc != 0:
                raise RuntimeError(icp._L.b200icp_last_error(icp._h).decode())
            return finish_step()
        return step

    for _ in range(max(warmup, 3)):
        step_device()
    records = np.zeros((steps, 18), np.float32)
    if env.world > 1:
        env.gather_poses(records)  # (warm-up of the collective too: NCCL sets its connections up on first use)
    torch.cuda.synchronize()
    env.barrier()
    dev_ms, wall_ms = env.timed(step_device, steps, ext, records)
    env.barrier()
    launches_per_step = icp.timing().kernel_launches
    T = T_out[:nn].reshape(rows, rows).T.copy()
    total_ms = env.max_over_ranks(max(dev_ms, wall_ms))  # the call is synchronous: wall time includes the launch overhead

    out = {"ms_per_step": total_ms / steps, "device_ms_per_step": dev_ms / steps, "launches_per_step": int(launches_per_step),
           "setmap_ms": setmap_ms, "T": T, "iterations": int(res.iterations)}
    for key, ptr in (("pinned", reading_pinned.data_ptr()), ("pageable", reading_pageable.ctypes.data)):
        step = make_e2e(ptr)
        for _ in range(3):
            step()
        env.barrier()
        d, wl = env.timed(step, steps, ext, records)
        env.barrier()
        out["e2e_%s_ms" % key] = env.max_over_ranks(max(d, wl)) / steps

    # ---- the persistent loop kernel (what the timed steps run): CUDA events around its launch, averaged over a
    #      separate pass of L2-flushed steps; the in-kernel %globaltimer figures come from the last of them ---
    loop_ms_sum, loop_n, tm = 0.0, 0, None
    for _ in range(steps):
        env.flush.fill_(1)
        torch.cuda.synchronize()
        icp.register_device(d_reading.data_ptr(), nq)
        tm = icp.timing()
        loop_ms_sum += tm.loop_kernel_ms
        loop_n += 1
    out.update(loop_kernel_ms=loop_ms_sum / max(loop_n, 1), loop_n=loop_n, loop_iters=tm.loop_iterations,
               loop_search_ms=tm.loop_search_ms_sum, loop_total_ms=tm.loop_total_ms, loop_fast_iters=tm.loop_fast_iterations,
               loop_two_iters=tm.loop_two_barrier_iterations, loop_searched=tm.loop_searched_queries)
    if with_roofline and w["knn"] == 1:
        # stand-alone search kernels through the kernel-per-step path: per-launch events
        icp.set_profiling(True)
        nn_ms, nn_n = 0.0, 0
        for _ in range(max(steps // 2, 2)):
            env.flush.fill_(1)
            torch.cuda.synchronize()
            icp.register_device(d_reading.data_ptr(), nq)
            tmm = icp.timing()
            nn_ms += tmm.nn_ms_sum
            nn_n += tmm.nn_launches
        icp.set_profiling(False)
        out.update(nn_ms=nn_ms, nn_n=nn_n)
    icp.close()
    return out


def pair_summary(w, r, cpu=None):
    """One `extra` entry: GPU figures of a pair workload (+ the CPU oracle beside it)."""
    from norlab_icp_mapper_b200 import synth
    s = {"workload": w["label"], "value": 1e3 / r["ms_per_step"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
         "e2e_pinned": 1e3 / r["e2e_pinned_ms"], "e2e_pageable": 1e3 / r["e2e_pageable_ms"], "loop_kernel_us": 1e3 * r["loop_kernel_ms"],
         "iterations": r["iterations"], "queries_searched": r["loop_searched"],
         "queries_total": (r["iterations"] - 1) * w["n_scan"],
         "searched_fraction": r["loop_searched"] / max((r["iterations"] - 1) * w["n_scan"], 1),
         "one_barrier_iterations": r["loop_fast_iters"], "two_barrier_iterations": r["loop_two_iters"]}
    if cpu is not None:
        s["cpu"] = {"value": cpu["value"], "cores": cpu["cores"], "phases_ms": cpu["phases_ms"]}
        s["speedup_e2e_pageable_vs_cpu"] = s["e2e_pageable"] / cpu["value"]
        e = synth.pose_error(r["T"], cpu["T"])
        s["pose_diff_vs_oracle"] = {"rad": e[0], "m": e[1]}
    return s


def bench_batched(env, pairs_host, n_pairs_total, passes=3):
    """BASELINE config 5 through b200icp_register_batch: this rank's share of the pairs (pair j -> rank j mod N), two
    contexts on the GPU, the pose all_gather inside the timed region.  Returns pairs/s (whole job) and a breakdown."""
    torch = env.torch
    from norlab_icp_mapper_b200 import batched
    w = WORKLOADS["cfg2"]
    cfg = make_cfg(dict(w, iters=30))
    pinned_keep = []

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pinned_keep.append(t)
        return t.numpy()

    mine = [dict(map=pin(p["map"]), normals=pin(p["normals"]), reading=pin(p["reading"])) for p in pairs_host]
    idx = batched.shard_pairs(n_pairs_total, env.rank, env.world)
    assert len(idx) == len(mine)
    by_j = dict(zip(idx, mine))
    results = {}
    # four contexts per GPU: uploads, index builds, cold searches and loop kernels (each on a quarter of the SMs) of different pairs
    # overlap -- 1 / 2 / 3 / 4 contexts: 562 / 800 / 950 / 990 pairs/s on one GPU (the sweep is part of the N = 1 line)
    # (with fewer than 8 pairs per context the fill and drain of the pipeline outweigh the overlap: 16 pairs per rank at N = 4 ran
    #  3118 pairs/s on two contexts and 2560-3240 on four)
    default_ctx = int(os.environ.get("B200ICP_BATCH_CONTEXTS", "4" if len(mine) >= 32 else "2"))
    for n_ctx in (default_ctx, 1, 2, 3):
        if n_ctx in results:
            continue
        eng = batched.BatchEngine(cfg, devices=(env.local_rank,), contexts_per_device=n_ctx)
        batched.register_batch(lambda j: by_j[j], n_pairs_total, eng, rank=env.rank, world=env.world, dist=env.dist, device=env.device)
        env.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(passes):
            poses, ov, it = batched.register_batch(lambda j: by_j[j], n_pairs_total, eng, rank=env.rank, world=env.world,
                                                   dist=env.dist, device=env.device)
        b.record()
        b.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        env.barrier()
        ms = env.max_over_ranks(max(a.elapsed_time(b), wall_ms))
        lr = eng.last_results
        results[n_ctx] = dict(ms=ms, poses=poses, setmap_ms=float(np.mean([lr[i].setmap_ms for i in range(len(mine))])),
                              register_ms=float(np.mean([lr[i].register_ms for i in range(len(mine))])))
        eng.close()
        if env.world > 1:
            break  # the other context counts are an N = 1 explanation of the overlap, not part of the scaling series
    # pinned H2D rate of this GPU (explains the upload share)
    buf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    dst.copy_(buf)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        dst.copy_(buf, non_blocking=True)
    b.record()
    b.synchronize()
    h2d_gbs = 4 * (64 << 20) / (a.elapsed_time(b) * 1e-3) / 1e9
    bytes_per_pair = 1_000_000 * (16 + 12) + 200_000 * 16
    r2 = results[default_ctx]
    per_pair_ms = r2["ms"] / passes / max(len(mine), 1)
    upload_ms = bytes_per_pair / (h2d_gbs * 1e9) * 1e3
    gpu_ms = r2["setmap_ms"] + r2["register_ms"]
    from norlab_icp_mapper_b200 import synth
    errs = [synth.pose_error(r2["poses"][j], p["correction_true"]) for j, p in zip(idx, pairs_host)]
    out = {"metric": "pairs/sec (200k-pt scan vs 1M-pt submap: setMap + 30-iteration ICP per pair), %d pairs sharded over the GPUs" % n_pairs_total,
           "value": n_pairs_total * passes / (r2["ms"] * 1e-3), "unit": "pairs/s", "n_gpus": env.world, "scaling": "strong",
           "pairs": n_pairs_total, "pairs_this_rank": len(mine), "passes_timed": passes, "ms_per_batch": r2["ms"] / passes,
           "entry_point": "b200icp_register_batch (%d contexts per GPU, each registration loop on its share of the SMs; pair j -> rank j "
                          "mod N; NCCL all_gather of the poses inside the timed region; pinned host submaps)" % default_ctx,
           "per_pair_ms_this_rank": per_pair_ms,
           "breakdown_ms_per_pair": {"h2d_upload_at_measured_rate": upload_ms, "setmap_device": r2["setmap_ms"],
                                     "register_device": r2["register_ms"], "h2d_bytes": bytes_per_pair, "h2d_gbs_pinned": h2d_gbs},
           "limiter": ("host-to-device upload of the submap (PCIe)" if upload_ms > gpu_ms else
                       "GPU work per pair (index build + cold search + loop kernel), uploads hidden behind it")
                      if per_pair_ms < 1.25 * max(upload_ms, gpu_ms) else
                      "host-side serialisation (synchronous set_map / register calls per context): per-pair time exceeds both the "
                      "upload and the GPU work",
           "max_pose_error_vs_truth": {"rad": max(e[0] for e in errs), "m": max(e[1] for e in errs)}}
    out["contexts_per_gpu"] = default_ctx
    if len(results) > 1:
        out["contexts_per_gpu_sweep"] = {str(k): {"value": n_pairs_total * passes / (v["ms"] * 1e-3),
                                                  "per_pair_ms": v["ms"] / passes / max(len(mine), 1),
                                                  "setmap_device_ms": v["setmap_ms"], "register_device_ms": v["register_ms"]}
                                         for k, v in sorted(results.items())}
    return out


def bench_cfg3(env, scans):
    """BASELINE config 3: online mapping through the host mirror of Mapper::processInput (500 scans along an S-curve)."""
    from norlab_icp_mapper_b200 import synth
    from norlab_icp_mapper_b200._abi import make_config
    from norlab_icp_mapper_b200.mapper import Mapper
    cfg3 = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=30,
                       differential=(1e-3, 1e-3, 3))
    m = Mapper(cfg3, True, False, True, False, updateCondition=("distance", 1.0), sensorMaxRange=80.0, minDistNewPoint=0.05,
               surfaceNormalKnn=10, reservePoints=40_000_000, device=env.local_rank)
    rng = np.random.default_rng(5)
    times, upd = [], []
    T_prev_true, pose_est, T_true = None, None, None
    for i, (scan, T_true) in enumerate(scans):
        if pose_est is None:
            T_est = T_true
        else:  # odometry increment with noise composed on the last corrected pose
            inc = np.linalg.inv(T_prev_true) @ T_true @ synth.make_T(rng.normal(0, 0.05, 3), rng.normal(0, 0.5, 3))
            T_est = pose_est @ inc
        t0 = time.perf_counter()
        m.processInput(scan, T_est.astype(np.float32), 0.1 * i)
        times.append(time.perf_counter() - t0)
        pose_est = m.getPose().astype(np.float64)
        T_prev_true = T_true
        upd.append(m.stats().map_updated)
    st = m.stats()
    err = synth.pose_error(pose_est, T_true)
    out = {"workload": "cfg3: online mapping, %d scans x 100k points along a 900 m S-curve at 2 m spacing, odometry noise 5 cm / 0.5 deg, "
                       "PointDistance{0.05} + SurfaceNormal{knn 10}, update condition distance 1.0, sensorMaxRange 80" % len(scans),
           "scans": len(scans), "scans_per_s": len(scans) / sum(times), "ms_per_scan_median": 1e3 * float(np.median(times)),
           "ms_per_scan_p95": 1e3 * float(np.percentile(times, 95)), "ms_per_scan_mean": 1e3 * float(np.mean(times)),
           "ms_first_scan": 1e3 * times[0], "ms_slowest_after_first": 1e3 * float(max(times[1:])) if len(times) > 1 else None,
           "map_updates": int(sum(upd)), "final_local_points": int(st.n_local), "final_global_points": int(st.n_global),
           "drift_vs_truth": {"rad": err[0], "m": err[1]},
           "what": "whole Mapper::processInput per scan through libb200mapper.so (upload, ICP with Counter{30} + Differential, PointDistance "
                   "insert, incremental SurfaceNormal knn 10, cell window, index rebuild), host wall clock"}
    m.close()
    return out


def bench_nn_large(env, chunks, queries):
    """The stand-alone search kernel on a map larger than L2: 10 M points (160 MB of float4 + 128 MB cell table), 100 k queries."""
    torch = env.torch
    from norlab_icp_mapper_b200.icp import ICP
    from norlab_icp_mapper_b200._abi import make_config
    P = np.concatenate([c[0] for c in chunks])
    N = np.concatenate([c[1] for c in chunks])
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=2)
    icp = ICP(cfg, device=env.local_rank)
    icp.set_map(P, N)
    d_reading = torch.from_numpy(queries).to("cuda")
    nq = len(queries)
    icp.set_profiling(True)  # kernel-per-step path: per-launch events around every search kernel
    for _ in range(2):
        icp.register_device(d_reading.data_ptr(), nq)
    ms, n = 0.0, 0
    for _ in range(10):
        env.flush.fill_(1)
        torch.cuda.synchronize()
        icp.register_device(d_reading.data_ptr(), nq)
        tm = icp.timing()
        ms += tm.nn_ms_sum
        n += tm.nn_launches
    icp.close()
    nm = len(P)
    b_nn = 16 * nq + 16 * nm + 8 * nq
    avg_ms = ms / max(n, 1)
    return {"workload": "search kernels on a %d-point map (%.0f MB of points > 126 MB L2), %d queries, L2 flushed per registration" % (nm, 16e-6 * nm, nq),
            "launches_timed": n, "avg_launch_us": 1e3 * avg_ms, "algorithmic_bytes": b_nn, "achieved_gbs": b_nn / (avg_ms * 1e-3) / 1e9,
            "note": "mean over the cold and the warm search of 2-iteration registrations (kernel-per-step path, cudaEvents around each "
                    "launch); B_nn = 16 Nq + 16 Nm + 8 k Nq counts the whole map once although a query only touches ~10 cells"}


def run_b200(args):
    w = resolve(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    extras = not args.no_extras
    # ---- synthetic inputs first (fork pools must not follow CUDA initialisation) -----------------------------
    data = make_data(w, rank)
    from norlab_icp_mapper_b200 import batched
    my_pairs = pool_map(_gen_cfg5_pair, batched.shard_pairs(args.pairs, rank, world)) if extras and args.pairs > 0 else None
    full_extras = extras and world == 1
    cfg3_scans = pool_map(_gen_cfg3_scan, list(range(args.cfg3_scans)), _init_cfg3_world) if full_extras and args.cfg3_scans > 0 else None
    big_chunks = pool_map(_gen_map_chunk, [(9000 + i, 1_250_000) for i in range(8)]) if full_extras else None
    extra_data = {name: make_data(resolve(args, name), rank) for name in ("cfg2_hard", "cfg1", "cfg4")
                  if full_extras and name != args.workload}
    if full_extras and args.workload == "cfg2":
        extra_data["cfg2_robust"] = data  # (the same clouds, another outlier filter)

    env = Env()
    torch = env.torch
    sampler = ClockSampler(env.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    r = bench_pair(env, w, data, args, args.steps, args.warmup)
    clocks = sampler.finish() if sampler else None

    batched_out = bench_batched(env, my_pairs, args.pairs) if my_pairs is not None else None
    extra = {}
    if full_extras:
        cores = host_cores()
        for name, d in extra_data.items():
            wx = resolve(args, name)
            rx = bench_pair(env, wx, d, args, max(args.steps // 2, 5), 3, with_roofline=False)
            cpu = None if args.no_cpu_baseline else time_oracle(make_cfg(wx), d, 3 if wx["n_map"] > 500_000 else 10, 1, cores)
            extra[name] = pair_summary(wx, rx, cpu)
        if cfg3_scans is not None:
            extra["cfg3"] = bench_cfg3(env, cfg3_scans)
        if big_chunks is not None:
            extra["nn_large_map"] = bench_nn_large(env, big_chunks, data["reading"] if w["dim"] == 3 else make_data(resolve(args, "cfg2"))["reading"])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:  # DRAM bytes of one loop-kernel launch from the committed ncu capture (not measured by this run)
            t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["icp_loop_kernel"]
            if args.workload == "cfg2" and w["n_map"] == 2_000_000 and w["n_scan"] == 100_000 and w["iters"] == 30:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        k = w["knn"]
        b_nn = 16 * w["n_scan"] + 16 * w["n_map"] + 8 * k * w["n_scan"]   # one correspondence search
        b_acc = 48 * k * w["n_scan"]                                      # one pair accumulation
        iterations_run = r["iterations"]
        # the dominant kernel is the loop kernel: one launch = (iterations - 1) searches + `iterations` accumulations
        b_loop = (iterations_run - 1) * b_nn + iterations_run * b_acc
        achieved = b_loop / (r["loop_kernel_ms"] * 1e-3) / 1e9 if r["loop_kernel_ms"] > 0 else None
        nn_avg_ms = r.get("nn_ms", 0.0) / max(r.get("nn_n", 0), 1)
        warm_achieved = b_nn / (nn_avg_ms * 1e-3) / 1e9 if r.get("nn_n") else None
        from norlab_icp_mapper_b200 import synth
        err = synth.pose_error(r["T"], data["correction_true"])
        nq = w["n_scan"]
        steps = args.steps
        if extra.get("nn_large_map"):
            extra["nn_large_map"]["frac_of_peak"] = extra["nn_large_map"]["achieved_gbs"] / peak
        line = {
            "metric": METRIC, "value": world * 1e3 / r["ms_per_step"], "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w),
            "run": {"l2": "flushed between steps (256 MiB fill, outside the timed events); within a step the 96 MB index stays "
                          "L2-resident across the 30 iterations by design",
                    "parallelism": "one process per GPU; every rank registers its own scan against its own map (independent pairs), the K "
                                   "poses of the timed steps are all_gathered over NCCL once, inside the timed region" if world > 1 else "single GPU",
                    "host_cores": host_cores()},
            "e2e": {"value": world * 1e3 / r["e2e_pinned_ms"], "unit": UNIT, "h2d_bytes_per_step": 4 * (w["dim"] + 1) * nq,
                    "d2h_bytes_per_step": 512, "ms_per_step": r["e2e_pinned_ms"], "host_memory": "pinned",
                    "pageable": {"value": world * 1e3 / r["e2e_pageable_ms"], "ms_per_step": r["e2e_pageable_ms"],
                                 "note": "the same call from ordinary malloc'ed host memory (what the reference's Eigen matrices are)"}},
            "gpu_launches": int(r["launches_per_step"]) * steps,
            "device_ms_per_step": r["device_ms_per_step"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": "icp_loop_kernel (persistent cooperative kernel, one launch per registration: iterations 1..%d "
                                   "of search + outlier quantile + error sums + solve)" % (iterations_run - 1),
                         "algorithmic_bytes": b_loop, "avg_launch_us": r["loop_kernel_ms"] * 1e3, "launches_timed": r["loop_n"],
                         "peak_source": peak_src,
                         "how": "cudaEvents around the loop kernel's launch on the library's stream, mean over a separate pass of %d "
                                "L2-flushed steps; algorithmic bytes = (iterations - 1) x B_nn + iterations x B_acc with B_nn = 16 Nq + "
                                "16 Nm + 8 k Nq, B_acc = 48 k Nq (SURVEY 8d).  The kernel is latency/issue-bound, not HBM-bound: the index "
                                "stays L2-resident across iterations and most searches are skipped by proof (loop_phase)" % steps,
                         "loop_phase": {"what": "in-kernel %globaltimer record of CTA 0, last step of that pass (not a roofline figure: "
                                                "most queries are proven unchanged and never searched)",
                                        "iterations": r["loop_iters"],
                                        "avg_verify_search_phase_us": (1e3 * r["loop_search_ms"] / r["loop_iters"]) if r["loop_iters"] else None,
                                        "queries_searched": r["loop_searched"], "queries_total": (iterations_run - 1) * nq,
                                        "one_barrier_iterations": r["loop_fast_iters"], "two_barrier_iterations": r["loop_two_iters"],
                                        "loop_kernel_ms_globaltimer": r["loop_total_ms"]},
                         "standalone_search_kernel": {"kernel": "nn1_cold_kernel / nn1_warm_kernel<4> (kernel-per-step path, exhaustive ball search)",
                                                      "avg_launch_us": nn_avg_ms * 1e3, "launches_timed": r.get("nn_n", 0),
                                                      "achieved": warm_achieved,
                                                      "frac": (warm_achieved / peak) if warm_achieved else None,
                                                      "how": "separate pass through the kernel-per-step path, cudaEvents around every "
                                                             "k-NN launch (event-to-event, includes the launch gap)"}},
            "clocks": clocks,
            "setmap_ms": r["setmap_ms"],
            "pose_error_vs_truth": {"rad": err[0], "m": err[1]},
        }
        if batched_out is not None:
            line["batched"] = batched_out
        if world > 1:  # the CPU baseline is a single-GPU-run figure (rank 0 would keep the other ranks waiting for it)
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "timed at N=1 only"}
        elif not args.no_cpu_baseline:
            big = w["n_map"] > 500_000
            blk, rr = cpu_baseline_block(w, make_cfg(w), data, 30 if big else 60, 2 if big else 10)
            line["cpu_baseline"] = blk
            eo = synth.pose_error(r["T"], rr["T"])
            line["pose_diff_vs_oracle"] = {"rad": eo[0], "m": eo[1]}
            if batched_out is not None and my_pairs:
                cfg5 = make_cfg(dict(WORKLOADS["cfg2"], iters=30))
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import oracle_binding as ob
                t0 = time.perf_counter()
                for p in my_pairs[:2]:
                    o = ob.OracleICP(cfg5)
                    o.set_map(p["map"], p["normals"])
                    o.register(p["reading"], nthreads=host_cores())
                dt = (time.perf_counter() - t0) / min(len(my_pairs), 2)
                line["batched"]["cpu"] = {"value": 1.0 / dt, "unit": "pairs/s", "cores": host_cores(),
                                          "sample": "2 pairs (kd-tree build + 30-iteration ICP each) on the CPU oracle"}
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
