"""GPU parity of K3 (DH forward kinematics) against the reference's goldens and fixtures."""
import numpy as np
import pytest

import conftest as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fkine():
    from inversekinematicsann_b200.kinematics.forward import ForwardKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot
    return ForwardKinematics(SixDOFRobot.dh_matrix)


def test_reference_forward_unit(fkine):
    """reference tests/forward_unit.py:26-31."""
    for angle, dest in zip(C.REF_FORWARD_UNIT_ANGLES, C.REF_FORWARD_UNIT_POINTS):
        fka, chain = fkine.fkine(angle)
        assert fka.shape == (4, 4) and len(chain) == 4
        np.testing.assert_array_almost_equal(dest, [fka[0, 3], fka[1, 3], fka[2, 3]], decimal=4)


def test_chain_matrices_match_reference_fixture(fkine, golden_fk):
    fka, chain = fkine.fkine(list(golden_fk["angles"][0]))
    np.testing.assert_allclose(np.array(chain), golden_fk["chain0"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(fka[3], [0, 0, 0, 1], atol=0)


def test_batched_positions_f64_and_f32(fkine, golden_fk):
    pos = fkine.fkine_positions(golden_fk["angles"])
    np.testing.assert_allclose(pos, golden_fk["positions"], rtol=0, atol=1e-12)
    pos32, err32 = fkine.fkine_positions(golden_fk["angles"].astype(np.float32),
                                         golden_fk["positions"].astype(np.float32))
    assert pos32.dtype == np.float32
    np.testing.assert_allclose(pos32, golden_fk["positions"], rtol=0, atol=5e-6)
    assert err32.max() < 1e-5


def test_angle_range_exception(fkine):
    from inversekinematicsann_b200.robot.robot import OutOfRobotReachException
    with pytest.raises(OutOfRobotReachException) as exc:
        fkine.fkine([7.0, 0, 0, 0])
    assert str(exc.value) == "Forward Kinematics exception, robot joints angles limits are (-2pi, 2pi)"
    with pytest.raises(OutOfRobotReachException):
        fkine.fkine_positions(np.array([[0.1, 0.2, 0.3, 0.4], [0.0, -6.3, 0.0, 0.0]]))
    fkine.fkine([2 * np.pi, -2 * np.pi, 0, 0])  # bounds are inclusive (forward.py:23)


def test_fk_error_of_fabrik_output_matches_oracle(golden_fabrik):
    from inversekinematicsann_b200.kinematics._shared import get_engine
    from oracle import c_oracle
    xyz, ang = golden_fabrik["workspace_xyz"], golden_fabrik["workspace_angles"]
    _, _, want = c_oracle.fk_positions(ang, targets=xyz)
    pos, err, stats = get_engine().fk(ang, xyz)
    np.testing.assert_allclose(err, want, rtol=0, atol=1e-12)
    assert abs(stats.mean_fk_error - want.mean()) < 1e-12 and stats.n_fk_error == len(xyz)


def test_general_dh_table_matches_oracle():
    """A DH table with twists on every joint takes the general (non closed-form) chain path."""
    from inversekinematicsann_b200.kinematics.forward import ForwardKinematics
    from oracle import c_oracle
    dh = [[0, np.pi / 2, 0, 0], [2, 0.5, -0.25, 0.1], [0.3, 2, 1.5, 2.5], [np.pi / 2, 0.3, -0.2, 0.1]]
    rng = np.random.RandomState(4)
    ang = rng.uniform(-np.pi, np.pi, size=(5000, 4))
    st, want, _ = c_oracle.fk_positions(ang, dh=np.array(dh, dtype=np.float64))
    fk = ForwardKinematics(dh)
    np.testing.assert_allclose(fk.fkine_positions(ang), want, rtol=0, atol=1e-12)
    np.testing.assert_allclose(fk.fkine_positions(ang.astype(np.float32)), want, rtol=0, atol=1e-5)
    _, chain = fk.fkine(list(ang[0]))
    _, want_chain = c_oracle.fk_chain(ang[0], dh=np.array(dh, dtype=np.float64))
    np.testing.assert_allclose(np.array(chain), want_chain, rtol=0, atol=1e-12)


# ---- FK error fused into the solvers' epilogues ---------------------------------------------------
def _robot():
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    return R


@pytest.mark.parametrize("n", [501, 100_000])  # below / above the row count where K1 fuses the error (capi.cu)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fabrik_fused_fk_error_equals_k3(dtype, n):
    from inversekinematicsann_b200.kinematics.forward import ForwardKinematics
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    R = _robot()
    rng = np.random.default_rng(8)
    pts = rng.uniform([0, -6, -3], [6, 6, 6], size=(n, 3)).astype(dtype)
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    out = np.empty((pts.shape[0], 4), dtype=dtype)
    angles, fused = ik.ikine(pts, out=out, return_fk_error=True)
    assert fused.dtype == dtype and fused.shape == (pts.shape[0],)
    _, separate = ForwardKinematics(R.dh_matrix).fkine_positions(angles, targets=pts)
    tol = 2e-6 if dtype == np.float32 else 1e-12
    assert np.abs(fused - separate).max() <= tol
    assert abs(ik.last_stats.mean_fk_error - float(np.mean(separate, dtype=np.float64))) <= 1e-6
    # the same angles as without the fused error (same buffer precision: a float32 buffer gets the fp32
    # trigonometric tail, see csrc/fabrik.cu acos_f32), and within the float32-buffer bound of the fp64 result
    assert np.array_equal(angles, ik.ikine(pts, out=np.empty_like(out)), equal_nan=True)
    assert np.nanmax(np.abs(angles - ik.ikine(pts, as_array=True))) <= (1e-6 if dtype == np.float32 else 0.0)


def test_fabrik_fused_fk_error_vs_oracle():
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from oracle import c_oracle
    R = _robot()
    rng = np.random.default_rng(9)
    pts = rng.uniform([0, -6, -3], [6, 6, 6], size=(20_000, 3))
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    angles, it, err = ik.ikine(pts, as_array=True, return_iterations=True, return_fk_error=True)
    want = c_oracle.fabrik_ikine(pts)
    _, _, want_err = c_oracle.fk_positions(want["angles"], targets=pts)
    assert np.array_equal(it, want["iters"])
    # the wrong-branch rows of inverse.py:82-85,102-108 reproduce too: errors match row by row
    assert np.abs(err - want_err).max() <= 2e-3 and np.abs(err - want_err).mean() <= 1e-6


@pytest.mark.parametrize("mode", ["fp16x3_ts", "fp16x3", "fp32"])
def test_ann_fused_fk_error_equals_k3(mode):
    from inversekinematicsann_b200.kinematics.forward import ForwardKinematics
    from inversekinematicsann_b200.kinematics.inverse import AnnInverseKinematics
    from oracle import np_oracle
    R = _robot()
    W, b = np_oracle.synthetic_mlp(seed=3, dims=[3, 500, 500, 4])
    ann = AnnInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    ann.ann.mode = mode
    ann.ann.set_model(W, b, np_oracle.SHIPPED_MEAN_X, np_oracle.SHIPPED_SCALE_X, np_oracle.SHIPPED_MEAN_Y,
                      np_oracle.SHIPPED_SCALE_Y)
    rng = np.random.default_rng(10)
    for n in (1, 127, 128, 129, 40_003):
        pts = rng.uniform([0, -6, -3], [6, 6, 6], size=(n, 3)).astype(np.float32)
        angles, fused = ann.ikine(pts, as_array=True, return_fk_error=True)
        _, separate = ForwardKinematics(R.dh_matrix).fkine_positions(angles, targets=pts)
        assert fused.dtype == np.float32 and np.abs(fused - separate).max() <= 2e-6, (mode, n)
        assert abs(ann.last_stats.mean_fk_error - float(np.mean(separate, dtype=np.float64))) <= 1e-6
        assert np.array_equal(angles, ann.ikine(pts, as_array=True))


def test_device_entry_points_fk_stats_only():
    import torch
    from inversekinematicsann_b200.kinematics._shared import get_engine
    eng = get_engine()
    g = torch.Generator(device="cuda").manual_seed(4)
    xyz = torch.rand(1_000, 3, device="cuda", generator=g) * torch.tensor([6.0, 12.0, 9.0], device="cuda") + \
        torch.tensor([0.0, -6.0, -3.0], device="cuda")
    ang = torch.empty(xyz.shape[0], 4, device="cuda")
    err = torch.empty(xyz.shape[0], device="cuda")
    eng.stats_reset_torch()
    eng.fabrik_solve_device(xyz, ang, fk_stats=True)
    only_stats = eng.stats_fetch_torch()
    eng.stats_reset_torch()
    eng.fabrik_solve_device(xyz, ang, fk_err=err)
    with_rows = eng.stats_fetch_torch()
    assert only_stats.n_fk_error == with_rows.n_fk_error == xyz.shape[0]
    assert abs(only_stats.mean_fk_error - with_rows.mean_fk_error) <= 1e-9
    assert abs(with_rows.mean_fk_error - float(err.double().mean())) <= 1e-6
    eng.stats_reset_torch()
    eng.fabrik_solve_device(xyz, ang)
    assert eng.stats_fetch_torch().n_fk_error == 0


def test_three_fk_error_kernels_agree_and_match_the_oracle():
    """K3 has three code paths for ||FK(angles) - target||: the asynchronous-copy ring (large fp32 batches), the pair
    kernel (its tail, small batches, unaligned buffers) and the generic kernel (positions wanted).  One batch large
    enough for the ring plus a ragged tail: all three against each other and against the fp64 oracle, and the guard of
    forward.py:23-25 reports the same first offending row on every path."""
    import torch
    from inversekinematicsann_b200.kinematics._shared import get_engine
    from oracle import c_oracle
    eng = get_engine()
    n = 1_000_037                                   # 15 625 full trips of 64 rows + 37 rows for the pair kernel
    g = torch.Generator(device="cuda").manual_seed(3)
    ang = (torch.rand(n, 4, device="cuda", generator=g) * 2 - 1) * 3.0
    tgt = torch.rand(n, 3, device="cuda", generator=g) * 6 - 3
    err_ring = torch.empty(n, device="cuda")
    eng.stats_reset_torch()
    eng.fk_device(ang, targets=tgt, err=err_ring)                      # ring + pair tail
    st_ring = eng.stats_fetch_torch()
    err_pair = torch.empty(n - 1, device="cuda")
    eng.fk_device(ang[1:], targets=tgt[1:], err=err_pair)              # 4-byte offset: unaligned -> generic / pair path
    pos = torch.empty(n, 3, device="cuda")
    err_gen = torch.empty(n, device="cuda")
    eng.fk_device(ang, targets=tgt, pos=pos, err=err_gen)              # positions wanted -> generic kernel
    torch.cuda.synchronize()
    assert (err_ring[1:] - err_pair).abs().max().item() <= 2e-6
    assert (err_ring - err_gen).abs().max().item() <= 2e-6
    pick = torch.arange(0, n, 97, device="cuda")
    _, _, want = c_oracle.fk_positions(ang[pick].double().cpu().numpy(), targets=tgt[pick].double().cpu().numpy())
    assert np.abs(err_ring[pick].cpu().numpy() - want).max() <= 1e-5
    assert st_ring.n_fk_error == n and abs(st_ring.mean_fk_error - err_ring.double().mean().item()) <= 1e-6
    bad = ang.clone()
    bad[777_777, 2] = 6.5                            # outside [-2 pi, 2 pi]
    bad[900_001, 0] = -7.0
    eng.stats_reset_torch()
    eng.fk_device(bad, targets=tgt, err=err_ring)
    st = eng.stats_fetch_torch()
    assert st.first_fk_angle_range == 777_777 and torch.isnan(err_ring[777_777]).item()
