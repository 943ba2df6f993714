"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the contiguous-range split, the
diagnostics reduction (what turns a per-shard offending row into the reference's whole-batch
exception) and the ordered result gather.  No GPU compute is involved."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_in_order():
    from inversekinematicsann_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 100_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inversekinematicsann_b200.engine import IkStats
    from inversekinematicsann_b200.sharding import gather_rows, reduce_stats, shard_range
    lo, hi = shard_range(n_total, rank, world)
    # a fake "solve": row i of the result encodes i, so ordering mistakes are visible
    local = torch.arange(lo, hi, dtype=torch.float64).unsqueeze(1).repeat(1, 4) + torch.tensor([0.0, 0.25, 0.5, 0.75])
    full = gather_rows(local, n_total)
    # rank 1 saw an out-of-limits row (local index 3) and a later zero division; rank 0 saw neither
    stats = IkStats(n_solved=hi - lo, sum_iterations=10 * (hi - lo), n_iter_capped=rank,
                    first_out_of_limits=3 if rank == 1 else -1, first_zero_division=5 if rank == 1 else -1,
                    sum_fk_error=0.5 * (rank + 1), n_fk_error=hi - lo)
    total = reduce_stats(stats, row_offset=lo)
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full.numpy())
    np.save(os.path.join(out_dir, f"stats{rank}.npy"),
            np.array([total.n_solved, total.sum_iterations, total.n_iter_capped, total.first_out_of_limits,
                      total.first_zero_division, total.first_domain_error, total.sum_fk_error, total.n_fk_error]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_stats(tmp_path):
    world, n_total = 2, 1001  # odd: shards differ by one row
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npy")
    assert full.shape == (n_total, 4)
    np.testing.assert_array_equal(full[:, 0], np.arange(n_total))
    np.testing.assert_array_equal(full[:, 3], np.arange(n_total) + 0.75)
    lo1 = 501  # rank 1 starts after rank 0's 501 rows
    for r in range(world):
        s = np.load(tmp_path / f"stats{r}.npy")
        assert s[0] == n_total and s[1] == 10 * n_total and s[2] == 1
        assert s[3] == lo1 + 3 and s[4] == lo1 + 5 and s[5] == -1   # global row numbers, -1 stays -1
        assert s[6] == pytest.approx(1.5) and s[7] == n_total


def _device_worker(rank, world, port, n_total, chunk_rows, dst, out_dir):
    """The device-resident sharded path with a fake per-row 'solve' on CPU tensors: what is under test is the
    chunk grid both ends of the send/recv gather have to agree on, the row order and the diagnostics reduction."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inversekinematicsann_b200.engine import IkStats
    from inversekinematicsann_b200.sharding import _ShardedIkine, shard_range

    class FakeEngine:
        device = 0

        def __init__(self):
            self.rows = 0

        def stats_reset_torch(self):
            self.rows = 0

        def stats_fetch_torch(self):
            return IkStats(n_solved=self.rows, sum_iterations=3 * self.rows,
                           first_out_of_limits=2 if rank == world - 1 else -1)

    class FakeIk:
        last_stats = None
        raised = None

        def _raise_from_stats(self, points, stats):
            if stats.first_out_of_limits >= 0:
                FakeIk.raised = (stats.first_out_of_limits, points[stats.first_out_of_limits])

    class Fake(_ShardedIkine):
        eng = FakeEngine()

        def _engine(self):
            return self.eng

        @staticmethod
        def _device_out_dtype(xyz):
            return xyz.dtype

        def _solve_device(self, eng, xyz, out, fk_error):
            eng.rows += xyz.shape[0]
            out.copy_(xyz[:, :1] * torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=xyz.dtype))

    lo, hi = shard_range(n_total, rank, world)
    xyz = torch.arange(lo, hi, dtype=torch.float64).unsqueeze(1).repeat(1, 3)
    sh = Fake(FakeIk())
    full = sh.ikine_device(xyz, n_total=n_total, gather_dst=dst, chunk_rows=chunk_rows)
    assert (full is not None) == (rank == dst)
    local_only = sh.ikine_device(xyz, n_total=n_total)          # no gather: the shard's own rows
    assert local_only.shape == (hi - lo, 4)
    if rank == dst:
        np.save(os.path.join(out_dir, "full.npy"), full.numpy())
    # one request that arrives on ONE rank, served by all: scatter -> solve -> gather -> host array on the root
    pts = np.arange(n_total, dtype=np.float64)[:, None].repeat(3, axis=1) if rank == dst else None
    served = sh.ikine_from_root(pts, root=dst)
    assert (served is not None) == (rank == dst)
    if rank == dst:
        np.testing.assert_array_equal(served, np.arange(n_total)[:, None] * np.array([1.0, 2.0, 3.0, 4.0]))
    lo_last = shard_range(n_total, world - 1, world)[0]
    assert sh.ik.last_stats.n_solved == n_total and sh.ik.last_stats.first_out_of_limits == lo_last + 2
    row, printed = FakeIk.raised
    assert row == lo_last + 2
    assert printed == ([float(row)] * 3 if rank == world - 1 else f"#{row}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,chunk_rows,dst", [(2, 1001, 128, 0), (3, 1000, 333, 1), (2, 7, 100, 1)])
def test_device_resident_sharded_gather(tmp_path, world, n_total, chunk_rows, dst):
    mp.spawn(_device_worker, args=(world, _free_port(), n_total, chunk_rows, dst, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npy")
    want = np.arange(n_total)[:, None] * np.array([1.0, 2.0, 3.0, 4.0])
    np.testing.assert_array_equal(full, want)


def test_sharded_api_without_a_process_group():
    """ShardedFabrik / gather_rows in a plain single process (no init_process_group) behave as world = 1."""
    from inversekinematicsann_b200.sharding import gather_rows, _ShardedIkine
    t = torch.arange(12.0).reshape(3, 4)
    assert gather_rows(t, 3) is t
    assert _ShardedIkine._world_rank() == (1, 0)
