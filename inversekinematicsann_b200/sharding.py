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


def gather_rows(local_rows, n_total, dst=0, out=None):
    """Gather contiguous row shards (torch tensors, (n_local, c)) to rank `dst` in trajectory order.
    Returns the (n_total, c) tensor on `dst` and None elsewhere.  Point-to-point: every other rank sends its shard
    once and `dst` receives each shard straight into its slice of the result (shards may differ by one row; nobody
    but `dst` allocates or receives anything -- the final gather SURVEY 8e describes, not an all-gather).
    `out` (on `dst`): a preallocated (n_total, c) tensor to receive into."""
    if not _dist_ready() or dist.get_world_size() == 1:
        return local_rows
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    local_rows = local_rows.contiguous()
    ops, full = [], None
    if rank == dst:
        shape = (int(n_total),) + tuple(local_rows.shape[1:])
        full = out if out is not None else torch.empty(shape, dtype=local_rows.dtype, device=local_rows.device)
        if tuple(full.shape) != shape or full.dtype != local_rows.dtype or not full.is_contiguous():
            raise ValueError(f"out must be a contiguous {shape} tensor of {local_rows.dtype}")
        lo, hi = sizes[dst]
        full[lo:hi].copy_(local_rows)
        ops = [dist.P2POp(dist.irecv, full[lo:hi], r) for r, (lo, hi) in enumerate(sizes) if r != dst and hi > lo]
    elif local_rows.shape[0] > 0:
        ops = [dist.P2POp(dist.isend, local_rows, dst)]
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()
    return full


class PeerGather:
    """The final gather without NCCL kernels: every rank maps rank `dst`'s result buffer into its own address space
    (torch symmetric memory over NVLink / NVSwitch peer access) and pushes each finished chunk of its shard there with a
    device-to-device copy on a side stream -- copy-engine DMA, no SMs, so it overlaps the solve of the next chunk
    without competing with it.  `finish()` makes the launching stream wait for the pushes and runs a device-side
    barrier across the ranks; after it `result` (on `dst`) holds all n_total rows in trajectory order.
    Needs an NCCL process group on one NVLink-connected node; construction raises where symmetric memory is
    unavailable and the caller falls back to send/recv (`gather_rows`)."""

    def __init__(self, n_total, cols, dtype, device, dst=0):
        import torch.distributed._symmetric_memory as symm
        self.dst, self.n_total = dst, int(n_total)
        group = dist.group.WORLD
        self.buf = symm.empty((self.n_total, cols), dtype=dtype, device=device)   # same size on every rank
        self.hdl = symm.rendezvous(self.buf, group)
        self.dst_view = self.hdl.get_buffer(dst, (self.n_total, cols), dtype)
        self.stream = torch.cuda.Stream(device)
        self.rank = dist.get_rank()

    def push(self, rows, first_row):
        """Copy `rows` (this rank's finished chunk, produced on the current stream) to rows
        [first_row, first_row + len) of dst's buffer, asynchronously on the side stream."""
        done = torch.cuda.Event()
        done.record()
        self.stream.wait_event(done)
        with torch.cuda.stream(self.stream):
            self.dst_view[first_row: first_row + rows.shape[0]].copy_(rows, non_blocking=True)

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.stream)
        self.hdl.barrier()                      # every rank's pushes have been issued and completed on its streams
        return self.buf if self.rank == self.dst else None


def chunk_ranges(n, chunk_rows):
    """[lo, hi) pieces of a shard of n rows, the grid both ends of the chunked gather agree on."""
    return [(lo, min(lo + chunk_rows, n)) for lo in range(0, n, chunk_rows)]


class _ShardedIkine:
    """The reference scales by running several brokers and letting the user split the data set (README.md:11,
    rpc_broker.py:65-68); here one torchrun job owns the whole trajectory, one contiguous row range per GPU.

    `ikine(points)`: every rank passes the SAME full host trajectory (as the reference's caller would), solves only
    its range on its own GPU, and rank `dst` gets the full (n, 4) result; the reference's exceptions are raised on
    every rank from the reduced diagnostics.  With `fk_error=True` the per-rank error sums are all-reduced as well
    (BASELINE config 4: the FK round trip of every prediction), available as `ik.last_stats.mean_fk_error`.

    `ikine_device(xyz_shard, ...)`: the device-resident form -- each rank holds only ITS rows in HBM, nothing
    touches the host, and the optional gather ships finished chunks to `gather_dst` while the next chunk is being
    solved (copy-engine pushes over NVLink where torch symmetric memory works, else NCCL send/recv).

    `ikine_from_root(points)`: one request that arrives on one rank, served by all GPUs."""

    def __init__(self, ik):
        self.ik = ik

    def _solve_local(self, eng, local, fk_error):
        raise NotImplementedError

    def _solve_device(self, eng, xyz, out, fk_error):
        raise NotImplementedError

    @staticmethod
    def _world_rank():
        return (dist.get_world_size(), dist.get_rank()) if _dist_ready() else (1, 0)

    @staticmethod
    def _collective_device(eng):
        """Where the tensors of the collectives live: the ENGINE's GPU under NCCL (not torch's current device, which
        the caller may never have set), the host under gloo."""
        return f"cuda:{eng.device}" if _dist_ready() and dist.get_backend() == "nccl" else "cpu"

    def ikine(self, points, gather=True, fk_error=False, dst=0):
        import numpy as np
        from .kinematics._shared import points_to_array
        arr = points_to_array(points)
        world, rank = self._world_rank()
        lo, hi = shard_range(arr.shape[0], rank, world)
        local = arr[lo:hi]
        eng = self._engine()
        angles, stats = self._solve_local(eng, local, fk_error) if hi > lo else \
            (np.zeros((0, 4), dtype=self._out_dtype), IkStats())
        dev = self._collective_device(eng)
        total = reduce_stats(stats, row_offset=lo, device=dev)
        self.ik.last_stats = total
        self.ik._raise_from_stats(points, total)
        if not gather or world == 1:
            return angles
        full = gather_rows(torch.from_numpy(angles).to(dev), arr.shape[0], dst=dst)
        return None if full is None else full.cpu().numpy()

    def ikine_device(self, xyz_shard, out_shard=None, n_total=None, gather_dst=None, gather_out=None,
                     chunk_rows=1 << 24, fk_error=False, check=True, peer_gather=None, gather_mode="auto"):
        """Solve this rank's rows where they are.  xyz_shard: (n_local, 3) CUDA tensor = rows
        shard_range(n_total, rank, world) of the trajectory; out_shard: optional (n_local, 4) CUDA tensor.
        Returns out_shard, or -- with gather_dst -- the full (n_total, 4) tensor on that rank and None elsewhere.
        `check=False` skips the diagnostics round trip (one device->host read + one all-reduce): the caller vouches
        for the inputs and reads `eng.stats_fetch_torch()` itself.
        The gather ships finished chunks while the next one is solved: with copy-engine pushes into dst's mapped buffer
        (PeerGather; `gather_mode` "auto" uses it when torch symmetric memory works on every rank, "p2p" insists,
        "nccl" never does; a PeerGather may also be passed in as `peer_gather`) or with NCCL send/recv.  Every rank must
        pass the same `gather_mode`.  With PeerGather the result on dst is its symmetric buffer, valid until ANY rank
        starts the next gather of the same size (its pushes land in the same rows; copy the result out, or synchronise
        the ranks, before that); `gather_out` (a preallocated result on dst) applies to the NCCL transport only."""
        world, rank = self._world_rank()
        eng = self._engine()
        n_local = int(xyz_shard.shape[0])
        if n_total is None:
            if world > 1 and gather_dst is not None:
                raise ValueError("n_total is required to gather (shards may differ by one row)")
            n_total = n_local
        lo, hi = shard_range(n_total, rank, world)
        if hi - lo != n_local:
            raise ValueError(f"rank {rank} of {world} owns rows [{lo}, {hi}) of {n_total}, got {n_local} rows")
        if out_shard is None:
            out_shard = torch.empty((n_local, 4), dtype=self._device_out_dtype(xyz_shard), device=xyz_shard.device)
        if check:
            eng.stats_reset_torch()
        gathering = gather_dst is not None and world > 1
        # NOTE: which transport is used must not depend on anything rank-local (e.g. `gather_out`, which only dst
        # passes): every rank has to take the same branch or the job deadlocks.
        if gathering and peer_gather is None and gather_mode in ("auto", "p2p"):
            peer_gather = self._peer_gather(n_total, out_shard, gather_dst, insist=gather_mode == "p2p")
        if gathering and peer_gather is not None:
            for a, b in chunk_ranges(n_local, chunk_rows):
                self._solve_device(eng, xyz_shard[a:b], out_shard[a:b], fk_error)
                peer_gather.push(out_shard[a:b], lo + a)
            full = peer_gather.finish()
            if check:
                total = reduce_stats(eng.stats_fetch_torch(), row_offset=lo, device=xyz_shard.device)
                self.ik.last_stats = total
                self.ik._raise_from_stats(_RowPrinter(xyz_shard, lo), total)
            return full
        full, pending = None, []
        if gathering and rank == gather_dst:
            full = gather_out if gather_out is not None else \
                torch.empty((n_total, 4), dtype=out_shard.dtype, device=out_shard.device)
        sizes = [shard_range(n_total, r, world) for r in range(world)]
        n_chunks = max(len(chunk_ranges(h - l, chunk_rows)) for l, h in sizes) if gathering else 0
        mine = chunk_ranges(n_local, chunk_rows if gathering else max(n_local, 1))
        for c in range(max(len(mine), n_chunks)):
            if c < len(mine):
                a, b = mine[c]
                self._solve_device(eng, xyz_shard[a:b], out_shard[a:b], fk_error)
            if not gathering:
                continue
            if rank == gather_dst:
                ops = []
                for r, (l, h) in enumerate(sizes):
                    pieces = chunk_ranges(h - l, chunk_rows)
                    if r != rank and c < len(pieces):
                        ops.append(dist.P2POp(dist.irecv, full[l + pieces[c][0]: l + pieces[c][1]], r))
                if ops:
                    pending += dist.batch_isend_irecv(ops)
            elif c < len(mine):
                pending.append(dist.isend(out_shard[mine[c][0]: mine[c][1]], gather_dst))
        if gathering and rank == gather_dst:
            full[lo:hi].copy_(out_shard)
        for work in pending:
            work.wait()
        if check:
            total = reduce_stats(eng.stats_fetch_torch(), row_offset=lo, device=xyz_shard.device)
            self.ik.last_stats = total
            self.ik._raise_from_stats(_RowPrinter(xyz_shard, lo), total)
        if gathering:
            return full
        return out_shard

    _PEER_GATHERS = {}

    @classmethod
    def _peer_gather(cls, n_total, out_shard, dst, insist=False):
        """One cached PeerGather per (rows, dtype, device, dst); None where symmetric memory is not available (decided by
        all ranks together, so that nobody is left waiting in a rendezvous)."""
        if not out_shard.is_cuda or dist.get_backend() != "nccl":
            if insist:
                raise RuntimeError("gather_mode='p2p' needs CUDA tensors and an NCCL process group")
            return None
        key = (int(n_total), str(out_shard.dtype), out_shard.device.index, int(dst))
        if key not in cls._PEER_GATHERS:
            try:
                made = PeerGather(n_total, out_shard.shape[1], out_shard.dtype, out_shard.device, dst=dst)
            except Exception:
                if insist:
                    raise
                made = None
            ok = torch.tensor([1 if made is not None else 0], device=out_shard.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            cls._PEER_GATHERS[key] = made if int(ok.item()) == 1 else None
        return cls._PEER_GATHERS[key]

    def ikine_from_root(self, points=None, root=0, out=None):
        """One request that arrives on ONE rank (the broker of rpc_broker.py:76-100), served by all GPUs: `root` pushes
        the (n, 3) host array to its GPU, the shards travel to their GPUs over NCCL send/recv, every rank solves its
        rows in HBM, the angles are gathered back to `root` and copied into `out` (e.g. a pinned reply buffer).  Every
        rank calls this; `points` / `out` are only read on `root`.  Returns the (n, 4) host array on `root`.
        The request crosses PCIe once, on `root`'s link: that link, not the solve, bounds this path."""
        import numpy as np
        world, rank = self._world_rank()
        eng = self._engine()
        dev = torch.device("cuda", eng.device) if (world == 1 or dist.get_backend() == "nccl") else torch.device("cpu")
        meta = [None]
        if rank == root:
            from .kinematics._shared import points_to_array
            arr = points_to_array(points)
            meta = [(arr.shape[0], str(arr.dtype))]
        if world > 1:
            dist.broadcast_object_list(meta, src=root)
        n_total, dtype_name = meta[0]
        tdtype = torch.float32 if dtype_name == "float32" else torch.float64
        full_in = None
        if rank == root:
            import warnings
            with warnings.catch_warnings():  # a request body decoded in place is a read-only view: it is only read
                warnings.simplefilter("ignore", UserWarning)
                full_in = torch.from_numpy(arr).to(dev, non_blocking=True)
        shard = scatter_rows(full_in, n_total, 3, tdtype, dev, src=root)
        full = self.ikine_device(shard, n_total=n_total, gather_dst=root if world > 1 else None)
        if rank != root:
            return None
        if out is None:
            out = np.empty((n_total, 4), dtype=np.dtype(str(full.dtype).replace("torch.", "")))
        torch.from_numpy(out).copy_(full)      # device -> host (pinned `out`: one DMA)
        return out


def scatter_rows(full_rows, n_total, cols, dtype, device, src=0):
    """The inverse of gather_rows: rank `src` holds (n_total, cols) rows on its device, every rank gets its
    contiguous shard (point-to-point sends of unequal shards; `full_rows` is None elsewhere)."""
    world, rank = (dist.get_world_size(), dist.get_rank()) if _dist_ready() else (1, 0)
    if world == 1:
        return full_rows
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    lo, hi = sizes[rank]
    if rank == src:
        ops = [dist.P2POp(dist.isend, full_rows[l:h], r) for r, (l, h) in enumerate(sizes) if r != src and h > l]
        mine = full_rows[lo:hi]
    else:
        mine = torch.empty((hi - lo, cols), dtype=dtype, device=device)
        ops = [dist.P2POp(dist.irecv, mine, src)] if hi > lo else []
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()
    return mine


class _RowPrinter:
    """`points[i]` for the exception text when the points live in HBM on several ranks: the offending row is
    fetched from this rank if it owns it, else described by its index."""

    def __init__(self, shard, lo):
        self.shard, self.lo = shard, lo

    def __getitem__(self, i):
        j = i - self.lo
        if 0 <= j < self.shard.shape[0]:
            return self.shard[j].tolist()
        return f"#{i}"


class ShardedFabrik(_ShardedIkine):
    """Drop-in for FabrikInverseKinematics.ikine over all ranks of a torchrun job."""
    _out_dtype = "float64"

    def _engine(self):
        return self.ik._engine()

    @staticmethod
    def _device_out_dtype(xyz):
        return xyz.dtype

    def _solve_local(self, eng, local, fk_error):
        return eng.fabrik_solve(local, precision=self.ik.precision, fk_stats=fk_error)[:2]

    def _solve_device(self, eng, xyz, out, fk_error):
        eng.fabrik_solve_device(xyz, out, precision=self.ik.precision, fk_stats=fk_error)


class ShardedAnn(_ShardedIkine):
    """Drop-in for AnnInverseKinematics.ikine over all ranks (weights replicated per GPU at load time)."""
    _out_dtype = "float32"

    def _engine(self):
        return self.ik.ann._ensure_uploaded()

    @staticmethod
    def _device_out_dtype(xyz):
        return torch.float32

    def _solve_local(self, eng, local, fk_error):
        return eng.ann_solve(local, mode=self.ik.ann.mode, fk_stats=fk_error)[:2]

    def _solve_device(self, eng, xyz, out, fk_error):
        eng.ann_solve_device(xyz, out, mode=self.ik.ann.mode, fk_stats=fk_error)
