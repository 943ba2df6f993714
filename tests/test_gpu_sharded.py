"""Sharded FABRIK ikine over two ranks (gloo rendezvous, both ranks drive cuda:0 with independent
kernels -- no kernel waits on another) equals the single-process result row for row, and the
reference's whole-batch exception is raised on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), IKB_DEVICE="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import OutOfRobotReachException, SixDOFRobot as R
    from inversekinematicsann_b200.sharding import ShardedFabrik
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits, device=0)
    rng = np.random.RandomState(17)
    pts = rng.rand(30_001, 3) * [6, 12, 9] + [0, -6, -3]
    full = ShardedFabrik(ik).ikine(pts)
    if rank == 0:
        np.save(os.path.join(out_dir, "sharded.npy"), full)
        np.save(os.path.join(out_dir, "iters.npy"), np.array([ik.last_stats.sum_iterations, ik.last_stats.n_solved]))
    else:
        assert full is None
    bad = pts.copy()
    bad[29_000, 2] = 6.5  # lives in rank 1's shard: both ranks must raise, with the global row
    try:
        ShardedFabrik(ik).ikine(bad)
        raised = ""
    except OutOfRobotReachException as exc:
        raised = str(exc)
    with open(os.path.join(out_dir, f"raised{rank}.txt"), "w") as f:
        f.write(raised)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_fabrik(tmp_path):
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    rng = np.random.RandomState(17)
    pts = rng.rand(30_001, 3) * [6, 12, 9] + [0, -6, -3]
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    single = ik.ikine(pts, as_array=True)
    sharded = np.load(tmp_path / "sharded.npy")
    assert np.array_equal(single, sharded, equal_nan=True)
    its = np.load(tmp_path / "iters.npy")
    assert its[0] == ik.last_stats.sum_iterations and its[1] == 30_001
    for r in range(2):
        msg = open(tmp_path / f"raised{r}.txt").read()
        assert "is out of manipulator reach area" in msg and "6.5" in msg


def _ann_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), IKB_DEVICE="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inversekinematicsann_b200.kinematics.inverse import AnnInverseKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    from inversekinematicsann_b200.sharding import ShardedAnn
    ann = AnnInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits, device=0)
    ann.load_model(os.path.join(ROOT, "models", "roboarm_b200_r01.h5"))
    rng = np.random.RandomState(23)
    pts = (rng.rand(20_003, 3) * [6, 12, 9] + [0, -6, -3]).astype(np.float32)
    full = ShardedAnn(ann).ikine(pts, fk_error=True)     # config 4's shape: predictions + reduced FK round trip
    if rank == 0:
        np.save(os.path.join(out_dir, "ann.npy"), full)
        np.save(os.path.join(out_dir, "ann_fk.npy"), np.array([ann.last_stats.sum_fk_error, ann.last_stats.n_fk_error]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_ann_with_fk_round_trip(tmp_path):
    from inversekinematicsann_b200.kinematics.inverse import AnnInverseKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    mp.spawn(_ann_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    rng = np.random.RandomState(23)
    pts = (rng.rand(20_003, 3) * [6, 12, 9] + [0, -6, -3]).astype(np.float32)
    ann = AnnInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    ann.load_model(os.path.join(ROOT, "models", "roboarm_b200_r01.h5"))
    single, err = ann.ikine(pts, as_array=True, return_fk_error=True)
    assert np.array_equal(single, np.load(tmp_path / "ann.npy"))
    s, c = np.load(tmp_path / "ann_fk.npy")
    assert c == 20_003 and abs(s / c - float(np.mean(err, dtype=np.float64))) <= 1e-6


def test_device_resident_form_in_one_process():
    """ShardedFabrik.ikine_device / ikine_from_root without a process group are a world of one: the rows stay in HBM,
    the reference's exceptions are raised from the device-side diagnostics, and the numbers equal ikine()'s.
    (The multi-rank gather is exercised with NCCL by bench.py at N > 1, which verifies the row order of the result, and
    with gloo on CPU tensors in tests/test_sharding_gloo.py.)"""
    import torch
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import OutOfRobotReachException, SixDOFRobot as R
    from inversekinematicsann_b200.sharding import ShardedFabrik
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    rng = np.random.RandomState(29)
    pts = (rng.rand(50_001, 3) * [6, 12, 9] + [0, -6, -3]).astype(np.float32)
    want = ik.ikine(pts, out=np.empty((len(pts), 4), np.float32))
    sh = ShardedFabrik(ik)
    got = sh.ikine_device(torch.from_numpy(pts).cuda(), n_total=len(pts), gather_dst=0, chunk_rows=16_384)
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), want, equal_nan=True)
    assert ik.last_stats.n_solved == len(pts)
    served = sh.ikine_from_root(pts)
    assert np.array_equal(served, want, equal_nan=True)
    bad = pts.copy()
    bad[40_000, 0] = -0.5
    with pytest.raises(OutOfRobotReachException) as exc:
        sh.ikine_device(torch.from_numpy(bad).cuda())
    assert "-0.5" in str(exc.value)
