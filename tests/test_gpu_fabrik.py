"""GPU parity of K1 (FABRIK + angle extraction) and K0 (limits) through the reference-shaped API and
the C ABI.  Tolerance (BASELINE.json north_star): joint angles within 1e-4 rad of the reference.
The default fp64 mode is held to a much tighter bar (1e-9 rad, identical iteration counts)."""
import numpy as np
import pytest

import conftest as C

pytestmark = pytest.mark.gpu

TOL_NORTH_STAR = 1e-4   # rad, BASELINE.json
TOL_F64_MODE = 1e-9     # rad, what the fp64 iterate actually achieves (float64 angle buffers)
TOL_F32_BUFFER = 1e-6   # rad, float32 angle buffers: fp64 iterate and cosines, trigonometric tail in fp32


@pytest.fixture(scope="module")
def robot():
    from inversekinematicsann_b200.robot.robot import SixDOFRobot
    return SixDOFRobot


@pytest.fixture(scope="module")
def ik(robot):
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    return FabrikInverseKinematics(robot.dh_matrix, robot.links_lengths, robot.effector_workspace_limits)


def test_reference_inverse_unit(ik):
    """reference tests/inverse_unit.py:21-34, same points, same decimal=6 assertion."""
    from inversekinematicsann_b200.robot.robot import OutOfRobotReachException
    predicted = ik.ikine(C.REF_INVERSE_UNIT_POINTS)
    assert isinstance(predicted, list) and isinstance(predicted[0][0], float)
    np.testing.assert_almost_equal(predicted, C.REF_INVERSE_UNIT_FABRIK, decimal=6)
    np.testing.assert_allclose(predicted, C.REF_INVERSE_UNIT_FABRIK, rtol=0, atol=TOL_F64_MODE)
    with pytest.raises(OutOfRobotReachException) as exc:
        ik.ikine(C.REF_INVERSE_UNIT_OUT_OF_REACH)
    assert str(exc.value) == ("Inverse Kinematics exception, point [1.567, 2.22, -3.123] is out of manipulator "
                              "reach area! Limits: {'x': [0, 6], 'y': [-6, 6], 'z': [-3, 6]}")


@pytest.mark.parametrize("name", C.FABRIK_SETS)
def test_fixtures_f64_mode(ik, golden_fabrik, name):
    xyz = golden_fabrik[f"{name}_xyz"]
    angles, iters = ik.ikine(xyz, as_array=True, return_iterations=True)
    assert np.array_equal(iters, golden_fabrik[f"{name}_iters"])
    diff = np.abs(angles - golden_fabrik[f"{name}_angles"])
    assert diff.max() <= TOL_NORTH_STAR
    assert diff.max() <= TOL_F64_MODE, f"{name}: max |dtheta| {diff.max():.3e}"
    assert ik.last_stats.sum_iterations == int(iters.sum())
    assert ik.last_stats.n_solved == len(xyz)


def test_fixtures_f32_mode(robot, golden_fabrik):
    """fp32 iterate (opt-in fast mode): most rows inside 1e-4 rad, the documented minority outside
    (convergence-test flips); iteration counts may differ by one on those rows."""
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    ik32 = FabrikInverseKinematics(robot.dh_matrix, robot.links_lengths, robot.effector_workspace_limits,
                                   precision="f32")
    xyz = golden_fabrik["workspace_xyz"]
    angles, iters = ik32.ikine(xyz, as_array=True, return_iterations=True)
    diff = np.abs(angles - golden_fabrik["workspace_angles"]).max(axis=1)
    frac_out = float((diff > TOL_NORTH_STAR).mean())
    frac_k = float((iters != golden_fabrik["workspace_iters"]).mean())
    print(f"f32 mode: {frac_out:.4%} rows > 1e-4 rad, {frac_k:.4%} iteration mismatches, max {diff.max():.3e}")
    assert frac_out < 0.02 and frac_k < 0.02


@pytest.mark.parametrize("box", ["workspace", "interior"])
def test_large_random_vs_oracle(ik, box):
    from oracle import c_oracle
    rng = np.random.RandomState(20260101)
    n = 200_000
    if box == "workspace":
        xyz = rng.rand(n, 3) * [6, 12, 9] + [0, -6, -3]
    else:
        xyz = rng.rand(n, 3) * [2, 4, 3] + [1, -2, 1]
    want = c_oracle.fabrik_ikine(xyz)
    angles, iters = ik.ikine(xyz, as_array=True, return_iterations=True)
    assert np.array_equal(iters, want["iters"])
    diff = np.abs(angles - want["angles"])
    assert diff.max() <= TOL_F64_MODE, f"max |dtheta| {diff.max():.3e}"
    assert ik.last_stats.n_iter_capped == int(((iters == 100)).sum()) or box == "workspace"


def test_input_forms_and_order(ik, golden_fabrik):
    from inversekinematicsann_b200.kinematics.point import Point
    xyz = golden_fabrik["spring50_xyz"]
    want = golden_fabrik["spring50_angles"]
    for form in (xyz.tolist(), [Point(p) for p in xyz.tolist()], xyz, xyz.astype(np.float32)):
        got = np.asarray(ik.ikine(form))
        tol = 2e-6 if getattr(form, "dtype", None) == np.float32 else TOL_F64_MODE
        np.testing.assert_allclose(got, want, rtol=0, atol=tol)
    assert ik.ikine([]) == []
    # fp32 output buffer through the engine API
    out, _ = ik._engine().fabrik_solve(xyz, out_dtype=np.float32)
    assert out.dtype == np.float32
    np.testing.assert_allclose(out, want, rtol=0, atol=TOL_F32_BUFFER)


def test_float32_buffer_tail_bound(ik):
    """A float32 angle buffer gets the trigonometric tail of the angle extraction (acos / atan2 after the 8-decimal
    rounding of the cosine, inverse.py:79-108) evaluated on the fp32 pipe; a float64 buffer keeps fp64 throughout.
    Both against the oracle on a uniform workspace sample (reachable and out-of-reach rows, both pass loops):
    float32 buffer <= 1e-6 rad (measured 4e-7: two fp32 roundings around pi), float64 buffer <= 1e-9 rad, the
    iteration counts identical either way."""
    from oracle import c_oracle
    rng = np.random.RandomState(77)
    xyz = rng.rand(200_000, 3) * [6, 12, 9] + [0, -6, -3]
    want = c_oracle.fabrik_ikine(xyz)
    ok = np.isfinite(want["angles"]).all(axis=1)
    eng = ik._engine()
    out32, st32, it32 = eng.fabrik_solve(xyz, out_dtype=np.float32, return_iters=True)
    out64, st64, it64 = eng.fabrik_solve(xyz, out_dtype=np.float64, return_iters=True)
    d32 = np.abs(out32.astype(np.float64) - want["angles"])[ok].max()
    d64 = np.abs(out64 - want["angles"])[ok].max()
    print(f"float32 buffer max |dtheta| {d32:.2e}, float64 buffer {d64:.2e}")
    assert d32 <= TOL_F32_BUFFER and d64 <= TOL_F64_MODE
    assert np.array_equal(it32, want["iters"]) and np.array_equal(it64, want["iters"])
    assert st32.sum_iterations == st64.sum_iterations == int(want["iters"].sum())


def test_permutation_invariance(ik):
    """Row i of the output belongs to row i of the input whatever order lanes finish in."""
    rng = np.random.RandomState(3)
    xyz = rng.rand(50_000, 3) * [6, 12, 9] + [0, -6, -3]
    base = ik.ikine(xyz, as_array=True)
    perm = rng.permutation(len(xyz))
    shuffled = ik.ikine(xyz[perm], as_array=True)
    assert np.array_equal(shuffled, base[perm], equal_nan=True)
    again = ik.ikine(xyz, as_array=True)
    assert np.array_equal(again, base, equal_nan=True)  # deterministic


def test_degenerate_and_error_rows(ik):
    from inversekinematicsann_b200.kinematics.point import Point
    from inversekinematicsann_b200.robot.robot import OutOfRobotReachException
    got = ik.ikine([[float("nan"), 1, 1]])          # NaN passes check_limits upstream, result is NaN
    assert np.isnan(got).all()
    with pytest.raises(ZeroDivisionError):           # target == start joint (point.py:40 upstream)
        ik.ikine([[1, 2, 3], [0, 0, 2.0]])
    with pytest.raises(OutOfRobotReachException) as exc:   # first offender wins, Point formatting
        ik.ikine([Point([1, 2, 3]), Point([1, 2, 7]), Point([9, 9, 9])])
    assert "point Point(1, 2, 7) is out of" in str(exc.value)
    with pytest.raises(OutOfRobotReachException):   # limits take precedence over the zero division
        ik.ikine([[0, 0, 2.0], [1, 2, 7]])
    with pytest.raises(ValueError):
        ik.ikine([[1, 2, 3, 4]])
    with pytest.raises(TypeError):
        ik.ikine([["a", "b", "c"]])
    # inclusive bounds
    assert len(ik.ikine([[6, 6, 6], [0, -6, -3], [6, -6, 6]])) == 3
    ik.check_limits([[6, 6, 6], [0, -6, -3]])
    with pytest.raises(OutOfRobotReachException):
        ik.check_limits([[1, 1, 1], [-1e-9, 0, 0]])


def test_host_pipeline_chunks_and_global_row_index(ik):
    """> 2 pipeline stages (4 Mi rows each): results identical to single-chunk solves and the
    out-of-limits row index is global."""
    rng = np.random.RandomState(8)
    n = (1 << 22) * 2 + 12345
    xyz = (rng.rand(n, 3) * [2, 4, 3] + [1, -2, 1]).astype(np.float32)
    eng = ik._engine()
    out, stats = eng.fabrik_solve(xyz, out_dtype=np.float32)
    assert stats.n_solved == n and stats.first_out_of_limits == -1
    for lo in (0, (1 << 22) - 7, (1 << 22) * 2 - 3, n - 100):
        part, _ = eng.fabrik_solve(xyz[lo:lo + 100], out_dtype=np.float32)
        assert np.array_equal(part, out[lo:lo + 100])
    bad = (1 << 22) + 17
    xyz[bad, 2] = 7.0
    xyz[bad + 5_000_000 % (n - bad - 1), 0] = -1.0
    _, stats = eng.fabrik_solve(xyz, out_dtype=np.float32)
    assert stats.first_out_of_limits == bad
    assert eng.check_limits(xyz) == bad


def test_fabrik_calculate_reference_unit(robot):
    """reference tests/fabrik_unit.py:22-35."""
    from inversekinematicsann_b200.kinematics.fabrik import Fabrik
    from inversekinematicsann_b200.kinematics.forward import ForwardKinematics
    from inversekinematicsann_b200.kinematics.point import Point
    fkine = ForwardKinematics(robot.dh_matrix)
    _, fkall = fkine.fkine(robot.dh_matrix[0])
    start = [Point([m[0, 3], m[1, 3], m[2, 3]]) for m in fkall]
    calculated = Fabrik(robot.links_lengths).calculate(start, [1, 2, 3])
    assert len(calculated) == 4 and isinstance(calculated[3], Point)
    np.testing.assert_array_almost_equal(calculated[3], C.REF_FABRIK_UNIT_EFFECTOR)
    with pytest.raises(ValueError):
        Fabrik(robot.links_lengths).calculate(start[:3], [1, 2, 3])
    # bit-faithful 3-D path: equals the C oracle's chain to the last digit or two
    from oracle import c_oracle
    st, chain = c_oracle.fk_chain(c_oracle.DH[0])
    _, want, iters = c_oracle.fabrik_calculate(chain[:, :3, 3], [1, 2, 3])
    np.testing.assert_allclose(np.array(calculated), want, rtol=0, atol=1e-14)


def test_full_size_properties():
    """BASELINE config 3 size (100 M cube_random targets, device resident): FK round trip inside the
    reference's own error for reachable targets, idempotence, and the iteration statistics of
    SURVEY appendix B."""
    import torch
    from inversekinematicsann_b200.kinematics._shared import get_engine
    eng = get_engine()
    n = 100_000_000
    g = torch.Generator(device="cuda").manual_seed(1234)
    xyz = torch.rand(n, 3, device="cuda", generator=g, dtype=torch.float32)
    xyz.mul_(torch.tensor([6.0, 12.0, 9.0], device="cuda")).add_(torch.tensor([0.0, -6.0, -3.0], device="cuda"))
    out = torch.empty(n, 4, device="cuda", dtype=torch.float32)
    eng.stats_reset_torch()
    eng.fabrik_solve_device(xyz, out)
    stats = eng.stats_fetch_torch()
    assert stats.n_solved == n and stats.first_out_of_limits == -1
    mean_iters = stats.sum_iterations / n
    assert 42.0 < mean_iters < 45.0, mean_iters          # 43.6
    err = torch.empty(n, device="cuda", dtype=torch.float32)
    eng.stats_reset_torch()
    eng.fk_device(out, targets=xyz, err=err)
    fk_stats = eng.stats_fetch_torch()
    reach = (xyz - torch.tensor([0.0, 0.0, 2.0], device="cuda")).norm(dim=1)
    inside = reach < 5.9
    med = err[inside][:5_000_000].median().item()
    assert med < 1e-3, med                                  # reference: median 2.2e-4, tol 1e-3
    outside = reach > 6.0
    gap = (err[outside] - (reach[outside] - 6.0)).abs()
    assert gap[:5_000_000].median().item() < 1e-2           # unreachable: error = reach - 6
    assert abs(fk_stats.mean_fk_error - err.double().mean().item()) < 1e-6
    out2 = torch.empty_like(out)
    eng.fabrik_solve_device(xyz, out2)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int32), out2.view(torch.int32))
    # the full-size launch itself against the oracle: 1e5 rows spread over the whole batch (every 1000th row), angles
    # within the float32-buffer bound and identical iteration counts
    from oracle import c_oracle
    it = torch.empty(n, device="cuda", dtype=torch.int32)
    eng.fabrik_solve_device(xyz, out2, iters=it)
    assert torch.equal(out.view(torch.int32), out2.view(torch.int32))      # asking for the counts changes nothing
    pick = torch.arange(0, n, 1000, device="cuda")
    sub = xyz[pick].double().cpu().numpy()
    want = c_oracle.fabrik_ikine(sub)
    got, got_it = out[pick].double().cpu().numpy(), it[pick].cpu().numpy()
    ok = np.isfinite(want["angles"]).all(axis=1)
    assert ok.sum() == len(sub)
    assert np.array_equal(got_it, want["iters"])
    assert np.abs(got - want["angles"]).max() <= TOL_F32_BUFFER


@pytest.mark.parametrize("tol,max_iter", [(1e-2, 100), (1e-3, 10), (1e-5, 30), (1e-3, 1)])
def test_non_default_solver_parameters(robot, tol, max_iter):
    """max_err / max_iterations_num of FabrikInverseKinematics (inverse.py:47-48) reach the kernel."""
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from oracle import c_oracle
    rng = np.random.RandomState(11)
    xyz = rng.rand(20_000, 3) * [6, 12, 9] + [0, -6, -3]
    ik2 = FabrikInverseKinematics(robot.dh_matrix, robot.links_lengths, robot.effector_workspace_limits,
                                  max_err=tol, max_iterations_num=max_iter)
    angles, iters = ik2.ikine(xyz, as_array=True, return_iterations=True)
    want = c_oracle.fabrik_ikine(xyz, tol=tol, max_iter=max_iter)
    assert np.array_equal(iters, want["iters"]) and iters.max() <= max_iter
    assert np.abs(angles - want["angles"]).max() <= TOL_F64_MODE


def test_zero_iteration_corner(robot):
    """tol >= 1 (or max_iter <= 0): the reference's while loop never runs and the seed chain is returned."""
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from oracle import c_oracle
    xyz = np.array([[1.0, 2.0, 3.0], [3.0, -1.0, 0.5]])
    for kw in ({"max_err": 1.5}, {"max_iterations_num": 0}):
        ikz = FabrikInverseKinematics(robot.dh_matrix, robot.links_lengths, robot.effector_workspace_limits, **kw)
        angles, iters = ikz.ikine(xyz, as_array=True, return_iterations=True)
        want = c_oracle.fabrik_ikine(xyz, tol=kw.get("max_err", 1e-3), max_iter=kw.get("max_iterations_num", 100))
        assert (iters == 0).all() and (want["iters"] == 0).all()
        # theta_1 of the vertical seed chain is atan2 of rounding noise upstream; the other angles are pinned
        np.testing.assert_allclose(angles[:, 1:], want["angles"][:, 1:], rtol=0, atol=TOL_F64_MODE)


def test_unequal_link_lengths(robot):
    """joints_distances is independent of the DH table upstream (robot.py:42 vs :40)."""
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from oracle import c_oracle
    links = [2.0, 1.5, 2.5, 1.0]
    rng = np.random.RandomState(12)
    xyz = rng.rand(20_000, 3) * [4, 8, 6] + [0.5, -4, -1]
    ikl = FabrikInverseKinematics(robot.dh_matrix, links, robot.effector_workspace_limits)
    angles, iters = ikl.ikine(xyz, as_array=True, return_iterations=True)
    want = c_oracle.fabrik_ikine(xyz, links=np.array(links))
    ok = want["status"] == 0
    assert np.array_equal(iters[ok], want["iters"][ok])
    finite = ok & np.isfinite(want["angles"]).all(axis=1)
    assert np.abs(angles[finite] - want["angles"][finite]).max() <= TOL_F64_MODE


def test_rotation_about_z_only_moves_theta1(ik):
    """The solve happens in the vertical plane through the z axis: rotating a target about z shifts theta_1 by
    the same angle and leaves theta_2..4 unchanged (a property the 2-D reduction relies on)."""
    rng = np.random.RandomState(13)
    n = 50_000
    r, phi = rng.uniform(0.5, 5.5, n), rng.uniform(-1.2, 1.2, n)
    z = rng.uniform(-3, 6, n)
    dphi = rng.uniform(-0.3, 0.3, n)
    a = ik.ikine(np.stack([r * np.cos(phi), r * np.sin(phi), z], 1), as_array=True)
    b = ik.ikine(np.stack([r * np.cos(phi + dphi), r * np.sin(phi + dphi), z], 1), as_array=True)
    np.testing.assert_allclose(b[:, 0] - a[:, 0], dphi, rtol=0, atol=1e-9)
    # near-straight joints quantise through round(cos, 8): allow one quantum there, exact-ish elsewhere
    d = np.abs(b[:, 1:] - a[:, 1:])
    assert np.percentile(d, 99.9) < 1e-9 and d.max() < 2e-4


def test_float32_and_float64_inputs_agree(ik):
    rng = np.random.RandomState(14)
    xyz = (rng.rand(20_000, 3) * [2, 4, 3] + [1, -2, 1]).astype(np.float32)
    a32 = ik.ikine(xyz, as_array=True)
    a64 = ik.ikine(xyz.astype(np.float64), as_array=True)
    assert np.array_equal(a32, a64)  # the same numbers reach the kernel either way


def test_other_planar_seed_pose():
    """A different seed row / first-link offset is still planar (Rz(theta_1) is the left-most DH factor) and
    takes the 2-D kernel; results follow the reference algorithm for that robot."""
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from oracle import c_oracle
    dh = [[0, np.pi / 3, 0.2, -0.1], [2, 0, 0, 0], [0.5, 2, 2, 2], [np.pi / 2, 0, 0, 0]]
    limits = {'x': [0, 6], 'y': [-6, 6], 'z': [-3, 6]}
    rng = np.random.RandomState(21)
    xyz = rng.rand(20_000, 3) * [5, 10, 8] + [0.5, -5, -2.5]
    ikp = FabrikInverseKinematics(dh, [2, 2, 2, 2], limits)
    angles, iters = ikp.ikine(xyz, as_array=True, return_iterations=True)
    want = c_oracle.fabrik_ikine(xyz, dh=np.array(dh, dtype=np.float64))
    ok = (want["status"] == 0) & np.isfinite(want["angles"]).all(axis=1)
    assert np.array_equal(iters[ok], want["iters"][ok])
    assert np.abs(angles[ok] - want["angles"][ok]).max() <= TOL_F64_MODE


def test_non_planar_robot_takes_the_generic_kernel():
    """Twists on joints 2..4 move the seed chain out of the vertical plane: the 3-D kernel (IEEE fp64, reference
    operation order) serves ikine, with the same flags and statistics."""
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import OutOfRobotReachException
    from oracle import c_oracle
    dh = [[0, np.pi / 2, 0.3, -0.2], [2, 0.1, 0, 0], [0, 2, 2, 2], [np.pi / 2, 0.4, -0.3, 0.2]]
    limits = {'x': [0, 6], 'y': [-6, 6], 'z': [-3, 6]}
    rng = np.random.RandomState(22)
    xyz = rng.rand(10_000, 3) * [5, 10, 8] + [0.5, -5, -2.5]
    ikg = FabrikInverseKinematics(dh, [2, 2, 2, 2], limits)
    angles, iters = ikg.ikine(xyz, as_array=True, return_iterations=True)
    want = c_oracle.fabrik_ikine(xyz, dh=np.array(dh, dtype=np.float64))
    ok = (want["status"] == 0) & np.isfinite(want["angles"]).all(axis=1)
    assert ok.mean() > 0.9
    assert np.array_equal(iters[ok], want["iters"][ok])
    assert np.abs(angles[ok] - want["angles"][ok]).max() <= 1e-8
    assert ikg.last_stats.n_solved == len(xyz) and ikg.last_stats.sum_iterations == int(iters.sum())
    with pytest.raises(OutOfRobotReachException):
        ikg.ikine([[1, 2, 3], [1, 2, 7]])


def test_out_of_reach_split_is_value_identical(ik, tmp_path):
    """Out-of-reach targets run in lockstep batches of 64, everything else through the lane-refill loop (csrc/fabrik.cu
    is_far); IKB_FABRIK_SPLIT=0 sends every row through the lane-refill loop.  The same rows must come out bit-identical
    either way and whatever the batch they arrive in -- in particular in the thin shell around
    |T - S| = d1 + d2 + d3 + tol where the predicate flips."""
    import os
    import subprocess
    import sys
    rng = np.random.RandomState(77)
    n = 90_000
    xyz = rng.rand(n, 3) * [6, 12, 9] + [0, -6, -3]
    # shell: distance from S = (0, 0, 2) within +-2e-3 of the reach limit 6 + tol, directions inside the workspace
    m = 30_000
    d = rng.randn(m, 3)
    d[:, 0] = np.abs(d[:, 0])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    radius = 6.001 + rng.uniform(-2e-3, 2e-3, m)
    radius[:2000] = 6.001 + rng.uniform(-1e-5, 1e-5, 2000)
    shell = np.array([0.0, 0.0, 2.0]) + d * radius[:, None]
    ok = (shell[:, 0] <= 6) & (np.abs(shell[:, 1]) <= 6) & (shell[:, 2] >= -3) & (shell[:, 2] <= 6)
    xyz[:int(ok.sum())] = shell[ok]
    rng.shuffle(xyz)
    whole, it_whole = ik.ikine(xyz, as_array=True, return_iterations=True)
    parts = [ik.ikine(xyz[i:i + 30_000], as_array=True, return_iterations=True) for i in range(0, n, 30_000)]
    pieces = np.concatenate([p[0] for p in parts])
    it_pieces = np.concatenate([p[1] for p in parts])
    assert np.array_equal(it_whole, it_pieces)
    assert np.array_equal(whole, pieces, equal_nan=True)
    # the same rows without the lockstep path, in a fresh process (the switch is read once per process)
    np.save(tmp_path / "xyz.npy", xyz)
    code = ("import sys, numpy as np; sys.path.insert(0, sys.argv[1]);"
            "from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics as F;"
            "from inversekinematicsann_b200.robot.robot import SixDOFRobot as R;"
            "ik = F(R.dh_matrix, R.links_lengths, R.effector_workspace_limits);"
            "a, it = ik.ikine(np.load(sys.argv[2]), as_array=True, return_iterations=True);"
            "a32 = ik.ikine(np.load(sys.argv[2]), out=np.empty((len(a), 4), np.float32));"
            "np.savez(sys.argv[3], a=a, it=it, a32=a32)")
    env = dict(os.environ, IKB_FABRIK_SPLIT="0")
    subprocess.run([sys.executable, "-c", code, C.ROOT, str(tmp_path / "xyz.npy"), str(tmp_path / "nosplit.npz")],
                   env=env, check=True, timeout=600)
    nosplit = np.load(tmp_path / "nosplit.npz")
    assert np.array_equal(it_whole, nosplit["it"])
    assert np.array_equal(whole, nosplit["a"], equal_nan=True)
    whole32 = ik.ikine(xyz, out=np.empty((n, 4), np.float32))
    assert np.array_equal(whole32, nosplit["a32"], equal_nan=True)
    far = np.linalg.norm(xyz - [0, 0, 2], axis=1) > 6.0011
    assert (it_whole[far] == 100).all() and far.sum() > 20_000
    from oracle import c_oracle
    want = c_oracle.fabrik_ikine(xyz)
    assert np.array_equal(it_whole, want["iters"])
    assert np.abs(whole - want["angles"]).max() <= TOL_F64_MODE


@pytest.mark.parametrize("links,tol,max_iter", [([2.0, 1.5, 2.5, 1.0], 1e-3, 100), ([2.0, 2.0, 2.0, 2.0], 1e-2, 37),
                                                ([2.0, 1.0, 1.0, 3.0], 5e-4, 64)])
def test_split_kernel_with_other_links_and_parameters(robot, links, tol, max_iter):
    """The out-of-reach predicate uses d1 + d2 + d3 + tol of the configured arm and max_iter of the configured solver:
    100 000 workspace targets (split kernel) against the oracle, iteration counts exact."""
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from oracle import c_oracle
    rng = np.random.RandomState(31)
    xyz = rng.rand(100_000, 3) * [6, 12, 9] + [0, -6, -3]
    ikl = FabrikInverseKinematics(robot.dh_matrix, links, robot.effector_workspace_limits, max_err=tol,
                                  max_iterations_num=max_iter)
    angles, iters = ikl.ikine(xyz, as_array=True, return_iterations=True)
    want = c_oracle.fabrik_ikine(xyz, links=np.array(links), tol=tol, max_iter=max_iter)
    ok = want["status"] == 0
    assert np.array_equal(iters[ok], want["iters"][ok])
    far = np.linalg.norm(xyz - [0, 0, 2], axis=1) > sum(links[1:]) + tol + 1e-5
    assert far.sum() > 10_000 and (iters[far] == max_iter).all()
    finite = ok & np.isfinite(want["angles"]).all(axis=1)
    assert np.abs(angles[finite] - want["angles"][finite]).max() <= TOL_F64_MODE
    assert ikl.last_stats.sum_iterations == int(iters.sum())
