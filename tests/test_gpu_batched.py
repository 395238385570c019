"""Batched mode on the GPU: the product register_fn (libb200icp.so) under register_batch equals the
oracle pair by pair; with more than one visible GPU the pairs are really sharded over NCCL."""
import os
import sys

import numpy as np
import pytest

from norlab_icp_mapper_b200 import batched, synth
from norlab_icp_mapper_b200._abi import make_config

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pair(j):
    return synth.make_pair_3d(n_map=100_000, n_scan=10_000, seed=4000 + j, world_size=(100.0, 100.0), n_boxes=12, scan_radius=45.0)


CFG = dict(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=15)


def test_batched_single_gpu_matches_oracle(oracle):
    cfg = make_config(**CFG)
    fn = batched.gpu_register_fn(cfg, 0)
    poses, overlaps, iters = batched.register_batch(_pair, 4, fn)
    fn.close()
    for j in range(4):
        p = _pair(j)
        o = oracle.OracleICP(cfg)
        o.set_map(p["map"], p["normals"])
        rc, T, res, _, _ = o.register(p["reading"])
        er, et = synth.pose_error(poses[j], T)
        assert rc == 0 and er <= 1e-4 and et <= 1e-3, (j, er, et)
        assert iters[j] == res.iterations and abs(overlaps[j] - res.overlap) < 2e-3


def _worker(rank, world, port, n_pairs, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = make_config(**CFG)
    fn = batched.gpu_register_fn(cfg, rank)
    poses, overlaps, iters = batched.register_batch(_pair, n_pairs, fn, rank=rank, world=world, dist=dist, device=f"cuda:{rank}")
    fn.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), poses=poses, overlaps=overlaps, iters=iters)
    dist.barrier()
    dist.destroy_process_group()


def test_batched_multi_gpu_nccl(tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    n_pairs = 6
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 1000, n_pairs, str(tmp_path)), nprocs=world, join=True)
    cfg = make_config(**CFG)
    fn = batched.gpu_register_fn(cfg, 0)
    single = batched.register_batch(_pair, n_pairs, fn)
    fn.close()
    first = np.load(tmp_path / "rank0.npz")
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["poses"], first["poses"]) and np.array_equal(got["iters"], single[2])  # every rank holds the same gathered result
    for j in range(n_pairs):  # (a rank's contexts split the SMs only when it has several pairs: sums sliced differently -> rounding)
        er, et = synth.pose_error(first["poses"][j], single[0][j])
        assert er <= 1e-6 and et <= 1e-5, (j, er, et)


def test_register_batch_c_abi_two_contexts_and_map_reuse(oracle):
    """b200icp_register_batch: pairs dealt over two contexts of one GPU give the poses of one-at-a-time registration; a pair
    without a map registers against the map its context already holds."""
    cfg = make_config(**CFG)
    pairs = [_pair(j) for j in range(5)]
    one = batched.BatchEngine(cfg, devices=(0,), contexts_per_device=1)
    ref = one.register_many(pairs)
    one.close()
    two = batched.BatchEngine(cfg, devices=(0,), contexts_per_device=2)
    got = two.register_many(pairs)
    for (Ta, oa, ia, sa), (Tb, ob_, ib, sb) in zip(ref, got):
        # two contexts on one GPU split its SMs: the error sums are sliced differently, poses agree to rounding
        er, et = synth.pose_error(Ta, Tb)
        assert sa == sb == 0 and ia == ib and abs(oa - ob_) < 1e-5 and er <= 1e-6 and et <= 1e-5
    again = two.register_many(pairs)
    assert all(np.array_equal(a[0], b[0]) for a, b in zip(got, again))  # run-to-run deterministic
    # map reuse: contexts 0 / 1 hold the maps of pairs 4 / 3; a map-less pair j goes to context j % 2
    reuse = two.register_many([dict(reading=pairs[4]["reading"]), dict(reading=pairs[3]["reading"])])
    assert np.array_equal(reuse[0][0], got[4][0]) and np.array_equal(reuse[1][0], got[3][0])
    # a failing pair (no normals for point-to-plane) reports its own status and does not stop the others
    bad = dict(pairs[0], normals=None)
    mixed = two.register_many([pairs[1], bad, pairs[2]])
    assert mixed[0][3] == 0 and mixed[2][3] == 0 and mixed[1][3] == 8  # B200ICP_ERR_INVALID_FIELD
    assert np.array_equal(mixed[0][0], got[1][0]) and np.array_equal(mixed[2][0], got[2][0])
    two.close()
    o = oracle.OracleICP(cfg)
    o.set_map(pairs[2]["map"], pairs[2]["normals"])
    rc, T, res, _, _ = o.register(pairs[2]["reading"])
    er, et = synth.pose_error(got[2][0], T)
    assert rc == 0 and er <= 1e-4 and et <= 1e-3
