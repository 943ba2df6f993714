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
