"""Batched registration: independent scan<->submap alignments sharded over the GPUs of one node
(BASELINE.json north_star / SURVEY.md 8e, config 5).  Pairs are independent units, so there is NO
data-path collective: pair j goes to rank j mod G, every rank runs its share through
`b200icp_register_batch` (include/b200icp.h) on its own GPU, and one all_gather of the final poses
(+ overlap, iterations) closes the batch.  With backend "nccl" the gather runs over NVLink; the CPU
tests exercise the same host logic with "gloo" and an injected registration function.

On a rank, `BatchEngine` holds `contexts_per_device` ICP contexts per GPU: while one context runs the
ICP loop of pair j (the SMs), the other uploads and indexes the submap of pair j+1 (copy engine, then
a short build) -- per pair the cost goes from upload + build + loop to about max(upload, build + loop).
"""
import ctypes as C

import numpy as np


def shard_pairs(n_pairs, rank, world):
    """Indices of the pairs rank `rank` owns: j = rank (mod world)."""
    return list(range(rank, n_pairs, world))


class BatchEngine:
    """The product path of the batched mode on one rank: ICP contexts + b200icp_register_batch.

    devices: CUDA ordinals this process drives (normally one: its rank's GPU; several = one process
    driving several GPUs).  Raises if libb200icp.so or the GPU is missing -- there is no fallback."""

    def __init__(self, cfg, devices=(0,), contexts_per_device=2):
        from . import _abi
        from .icp import ICP
        self._abi = _abi
        self.cfg = cfg
        self.n = cfg.dim + 1
        self.dim = cfg.dim
        # context order: device-major round robin, so consecutive pairs land on different devices first
        self.ctxs = [ICP(cfg, device=d) for _ in range(max(1, contexts_per_device)) for d in devices]
        self._L = self.ctxs[0]._L
        self._handles = (C.c_void_p * len(self.ctxs))(*[c._h for c in self.ctxs])
        self.last_results = None

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def register_many(self, pairs):
        """pairs: list of dict(map, normals, reading[, T_init]); `map` None = keep the context's map.
        Returns a list of (T (n x n), overlap, iterations, status)."""
        abi = self._abi
        n = len(pairs)
        if n == 0:
            return []
        arr = (abi.Pair * n)()
        keep = []  # the host arrays must outlive the call
        for j, p in enumerate(pairs):
            rd = np.ascontiguousarray(p["reading"], np.float32)
            keep.append(rd)
            arr[j].reading, arr[j].n_reading = rd.ctypes.data, len(rd)
            if p.get("map") is not None:
                m = np.ascontiguousarray(p["map"], np.float32)
                keep.append(m)
                arr[j].map_features, arr[j].n_map = m.ctypes.data, len(m)
                if p.get("normals") is not None:
                    nr = np.ascontiguousarray(p["normals"], np.float32)
                    keep.append(nr)
                    arr[j].map_normals = nr.ctypes.data
            if p.get("T_init") is not None:
                T = np.ascontiguousarray(np.asarray(p["T_init"], np.float32).T).ravel()
                keep.append(T)
                arr[j].T_init = T.ctypes.data
        out = (abi.PairResult * n)()
        self._L.b200icp_register_batch(self._handles, len(self.ctxs), arr, n, out)
        self.last_results = out
        res = []
        nn = self.n * self.n
        for j in range(n):
            T = np.frombuffer(out[j].T, np.float32, nn).reshape(self.n, self.n).T.copy()
            res.append((T, float(out[j].result.overlap), int(out[j].result.iterations), int(out[j].status)))
        return res

    def __call__(self, pair):  # register_fn protocol of register_batch (one pair at a time)
        T, overlap, iters, status = self.register_many([pair])[0]
        if status != 0:
            from ._lib import B200ICPError
            raise B200ICPError(status, "pair failed")
        return T, overlap, iters


def gpu_register_fn(cfg, device, contexts_per_device=2):
    """register_fn for the product path (kept name): a BatchEngine on `device`."""
    return BatchEngine(cfg, devices=(device,), contexts_per_device=contexts_per_device)


def register_batch(get_pair, n_pairs, register_fn, rank=0, world=1, dist=None, device="cpu", dim=3):
    """Registers pairs [0, n_pairs) across `world` ranks.

    get_pair(j) -> dict(map, normals, reading) is called only for the pairs this rank owns.
    register_fn: a BatchEngine (all of the rank's pairs go through ONE b200icp_register_batch call) or any
    callable pair -> (T, overlap, iterations) (the CPU tests inject the oracle here).
    Returns (poses [n_pairs, dim+1, dim+1], overlaps [n_pairs], iterations [n_pairs]) on every rank.
    """
    n = dim + 1
    mine = shard_pairs(n_pairs, rank, world)
    per_rank = (n_pairs + world - 1) // world
    rec = np.full((per_rank, n * n + 2), np.nan, np.float32)
    if hasattr(register_fn, "register_many"):
        results = [r[:3] for r in register_fn.register_many([get_pair(j) for j in mine])]
    else:
        results = [register_fn(get_pair(j)) for j in mine]
    for slot, (T, overlap, iters) in enumerate(results):
        rec[slot, :n * n] = np.asarray(T, np.float32).ravel()
        rec[slot, n * n] = overlap
        rec[slot, n * n + 1] = iters
    allrec = gather_records(rec, world, dist, device)
    poses = np.zeros((n_pairs, n, n), np.float32)
    overlaps = np.zeros(n_pairs, np.float32)
    iters = np.zeros(n_pairs, np.int32)
    for j in range(n_pairs):
        r, slot = j % world, j // world
        poses[j] = allrec[r, slot, :n * n].reshape(n, n)
        overlaps[j] = allrec[r, slot, n * n]
        iters[j] = int(allrec[r, slot, n * n + 1])
    return poses, overlaps, iters


def gather_records(rec, world, dist, device):
    """The only collective of the batched mode: all_gather of the per-rank [per_rank, (dim+1)^2 + 2] records."""
    if world <= 1:
        return rec[None]
    import torch
    local = torch.from_numpy(rec).to(device)
    if local.is_cuda:  # NCCL: one flat output, one kernel, one read-back
        out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out.cpu().numpy()
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    return np.stack([g.cpu().numpy() for g in gathered])  # [world, per_rank, n*n+2]
