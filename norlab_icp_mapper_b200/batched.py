"""Batched registration: independent scan<->submap alignments sharded over the GPUs of one node
(BASELINE.json north_star / SURVEY.md 8e).  Pairs are independent units, so there is NO data-path
collective: pair j goes to rank j mod G, every rank builds its own indices and runs its own loops,
and one all_gather of the final poses (+ overlap, iterations) closes the batch.  With backend "nccl"
the gather runs over NVLink; the CPU tests exercise the same code with "gloo".

The registration itself is injected (`register_fn`) so that the host logic can be tested without a
GPU; the product binding is `gpu_register_fn`, which goes through libb200icp.so.
"""
import numpy as np


def shard_pairs(n_pairs, rank, world):
    """Indices of the pairs rank `rank` owns: j = rank (mod world)."""
    return list(range(rank, n_pairs, world))


def gpu_register_fn(cfg, device):
    """register_fn for the product path: one ICP context per rank, map re-installed per pair."""
    from .icp import ICP
    icp = ICP(cfg, device=device)

    def run(pair):
        icp.set_map(pair["map"], pair.get("normals"))
        T = icp(pair["reading"])
        r = icp.last_result
        return T, float(r.overlap), int(r.iterations)

    run.close = icp.close
    return run


def register_batch(get_pair, n_pairs, register_fn, rank=0, world=1, dist=None, device="cpu", dim=3):
    """Registers pairs [0, n_pairs) across `world` ranks.

    get_pair(j) -> dict(map, normals, reading) is called only for the pairs this rank owns.
    Returns (poses [n_pairs, dim+1, dim+1], overlaps [n_pairs], iterations [n_pairs]) on every rank.
    """
    n = dim + 1
    mine = shard_pairs(n_pairs, rank, world)
    per_rank = (n_pairs + world - 1) // world
    rec = np.full((per_rank, n * n + 2), np.nan, np.float32)
    for slot, j in enumerate(mine):
        T, overlap, iters = register_fn(get_pair(j))
        rec[slot, :n * n] = np.asarray(T, np.float32).ravel()
        rec[slot, n * n] = overlap
        rec[slot, n * n + 1] = iters
    if world > 1:
        import torch
        local = torch.from_numpy(rec).to(device)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)  # the only collective of the batched mode
        allrec = np.stack([g.cpu().numpy() for g in gathered])  # [world, per_rank, n*n+2]
    else:
        allrec = rec[None]
    poses = np.zeros((n_pairs, n, n), np.float32)
    overlaps = np.zeros(n_pairs, np.float32)
    iters = np.zeros(n_pairs, np.int32)
    for j in range(n_pairs):
        r, slot = j % world, j // world
        poses[j] = allrec[r, slot, :n * n].reshape(n, n)
        overlaps[j] = allrec[r, slot, n * n]
        iters[j] = int(allrec[r, slot, n * n + 1])
    return poses, overlaps, iters
