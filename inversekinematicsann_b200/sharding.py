"""Multi-GPU plumbing: one process per GPU (torchrun), trajectory rows split into contiguous ranges.

Every target is independent in both solvers (reference inverse.py:120 loop, ann.py:73 row-wise), so
there is no collective inside the compute.  This replaces the reference's only scaling mechanism --
several brokers competing on one RabbitMQ queue with the user splitting the data set (README.md:11,
rpc_broker.py:65-68).  torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests) carries only:
  * the final gather of the (n, 4) angle rows to one rank, and
  * the reduction of the per-call diagnostics (sum of iterations, first offending row, error sums),
    which is what turns a per-shard out-of-limits row into the reference's whole-batch exception.
"""
from dataclasses import fields

import torch
import torch.distributed as dist

from .engine import IkStats

_NO_ROW = 2 ** 62


def shard_range(n, rank, world):
    """Contiguous rows [lo, hi) of rank `rank`: sizes differ by at most one, order preserved."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist_ready():
    return dist.is_available() and dist.is_initialized()


def reduce_stats(stats, row_offset=0, device=None):
    """Combine per-rank IkStats: counters and sums add up, first_* rows become global minima
    (local row + this rank's `row_offset`), exactly what a single-process solve would report."""
    if not _dist_ready() or dist.get_world_size() == 1:
        return stats
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    sum_names = ["n_solved", "sum_iterations", "n_iter_capped", "n_fk_error"]
    min_names = ["first_out_of_limits", "first_zero_division", "first_domain_error", "first_fk_angle_range"]
    sums = torch.tensor([getattr(stats, n) for n in sum_names], dtype=torch.int64, device=dev)
    mins = torch.tensor([getattr(stats, n) + row_offset if getattr(stats, n) >= 0 else _NO_ROW
                         for n in min_names], dtype=torch.int64, device=dev)
    err = torch.tensor([stats.sum_fk_error], dtype=torch.float64, device=dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    dist.all_reduce(err, op=dist.ReduceOp.SUM)
    out = IkStats(**{f.name: getattr(stats, f.name) for f in fields(IkStats)})
    for n, v in zip(sum_names, sums.tolist()):
        setattr(out, n, int(v))
    for n, v in zip(min_names, mins.tolist()):
        setattr(out, n, -1 if v >= _NO_ROW else int(v))
    out.sum_fk_error = float(err.item())
    return out


def gather_rows(local_rows, n_total, dst=0):
    """Gather contiguous row shards (torch tensors, (n_local, c)) to rank `dst` in trajectory order.
    Returns the (n_total, c) tensor on `dst` and None elsewhere.  Shards may differ by one row, so
    they are padded to the largest shard for all_gather_into_tensor."""
    if not _dist_ready() or dist.get_world_size() == 1:
        return local_rows
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    padded = local_rows
    if local_rows.shape[0] < biggest:
        pad = torch.zeros((biggest - local_rows.shape[0],) + tuple(local_rows.shape[1:]),
                          dtype=local_rows.dtype, device=local_rows.device)
        padded = torch.cat([local_rows, pad], dim=0)
    out = torch.empty((world * biggest,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype,
                      device=local_rows.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    if rank != dst:
        return None
    parts = [out[r * biggest: r * biggest + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


class _ShardedIkine:
    """Every rank passes the SAME full trajectory (as the reference's caller would), solves only its contiguous
    range on its own GPU, and rank 0 gets the full (n, 4) result; the reference's exceptions are raised on every
    rank from the reduced diagnostics.  With `fk_error=True` the per-rank error sums are all-reduced as well
    (BASELINE config 4: the FK round trip of every prediction), available as `ik.last_stats.mean_fk_error`."""

    def __init__(self, ik):
        self.ik = ik

    def _solve_local(self, eng, local, fk_error):
        raise NotImplementedError

    def ikine(self, points, gather=True, fk_error=False):
        import numpy as np
        from .kinematics._shared import points_to_array
        arr = points_to_array(points)
        world = dist.get_world_size() if _dist_ready() else 1
        rank = dist.get_rank() if _dist_ready() else 0
        lo, hi = shard_range(arr.shape[0], rank, world)
        local = arr[lo:hi]
        eng = self.ik._engine()
        angles, stats = self._solve_local(eng, local, fk_error) if hi > lo else \
            (np.zeros((0, 4), dtype=self._out_dtype), IkStats())
        total = reduce_stats(stats, row_offset=lo)
        self.ik.last_stats = total
        self.ik._raise_from_stats(points, total)
        if not gather:
            return angles
        dev = f"cuda:{eng.device}" if dist.get_backend() == "nccl" else "cpu"
        full = gather_rows(torch.from_numpy(angles).to(dev), arr.shape[0])
        return None if full is None else full.cpu().numpy()


class ShardedFabrik(_ShardedIkine):
    """Drop-in for FabrikInverseKinematics.ikine over all ranks of a torchrun job."""
    _out_dtype = "float64"

    def _solve_local(self, eng, local, fk_error):
        return eng.fabrik_solve(local, precision=self.ik.precision, fk_stats=fk_error)[:2]


class ShardedAnn(_ShardedIkine):
    """Drop-in for AnnInverseKinematics.ikine over all ranks (weights replicated per GPU at load time)."""
    _out_dtype = "float32"

    def _solve_local(self, eng, local, fk_error):
        eng = self.ik.ann._ensure_uploaded()
        return eng.ann_solve(local, mode=self.ik.ann.mode, fk_stats=fk_error)[:2]
