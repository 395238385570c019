"""Host logic of the batched (multi-GPU) mode on CPU: pair partition and the single all_gather, with
world_size 2 over gloo and the CPU oracle standing in for the per-rank registration."""
import os
import sys

import numpy as np
import pytest

from norlab_icp_mapper_b200 import batched, synth
from norlab_icp_mapper_b200._abi import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_pairs_partition():
    for n, w in ((64, 8), (7, 2), (3, 4), (0, 2)):
        parts = [batched.shard_pairs(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert all(j % w == r for r, p in enumerate(parts) for j in p)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _pair(j):
    return synth.make_pair_3d(n_map=20_000, n_scan=2_000, seed=4000 + j, world_size=(60.0, 60.0), n_boxes=8, scan_radius=25.0)


def _oracle_fn(cfg):
    import oracle_binding as ob

    def run(pair):
        o = ob.OracleICP(cfg)
        o.set_map(pair["map"], pair["normals"])
        rc, T, res, _, _ = o.register(pair["reading"], nthreads=1)
        assert rc == 0
        return T, res.overlap, res.iterations
    return run


def _worker(rank, world, port, n_pairs, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=8)
    poses, overlaps, iters = batched.register_batch(_pair, n_pairs, _oracle_fn(cfg), rank=rank, world=world, dist=dist)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), poses=poses, overlaps=overlaps, iters=iters)
    dist.barrier()
    dist.destroy_process_group()


def test_register_batch_world2_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp
    n_pairs, world = 5, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_pairs, str(tmp_path)), nprocs=world, join=True)
    cfg = make_config(dim=3, knn=1, max_dist=1.0, outliers=(("trimmed", 0.85),), minimizer="point_to_plane", max_iteration_count=8)
    single = batched.register_batch(_pair, n_pairs, _oracle_fn(cfg))
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["poses"], single[0])  # every rank holds every pose, identical to a 1-rank run
        assert np.array_equal(got["overlaps"], single[1]) and np.array_equal(got["iters"], single[2])
    for j in range(n_pairs):
        e = synth.pose_error(single[0][j], _pair(j)["correction_true"])
        assert e[1] < 0.1
