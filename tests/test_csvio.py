"""CSV <-> ndarray helpers against the reference CLI's pandas calls (cli.py:45-49, 74-78, 242)."""
import numpy as np
import pandas as pd
import pytest

from inversekinematicsann_b200 import csvio


def _reference_write(path, rows, columns):
    pd.DataFrame(rows, columns=list(columns)).to_csv(path, index=False)  # what cli.py:45-49 / 74-78 do


def test_reads_what_the_reference_cli_writes(tmp_path):
    rng = np.random.default_rng(3)
    pts = rng.uniform(-6, 6, (5000, 3))
    pts[0] = [0.0, -6.0, 1e-7]
    path = tmp_path / 'points.csv'
    _reference_write(path, pts.tolist(), csvio.POINT_COLUMNS)
    exact = csvio.read_points_csv(path, dtype=np.float64)
    assert np.array_equal(exact, pts)  # shortest round-trip text parses back to the same doubles
    via_pandas = csvio.read_points_csv(path, dtype=np.float64, engine='pandas')
    assert np.array_equal(via_pandas, pd.read_csv(path).values)
    assert np.abs(via_pandas - pts).max() <= 2e-15
    f32 = csvio.read_points_csv(path)
    assert f32.dtype == np.float32 and f32.flags.c_contiguous and np.array_equal(f32, pts.astype(np.float32))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_reference_reader_reads_what_we_write(tmp_path, dtype):
    rng = np.random.default_rng(4)
    angles = rng.uniform(-3.2, 3.2, (4000, 4)).astype(dtype)
    angles[0] = [0.0, 1.0, -2.5, 3.0]
    path = tmp_path / 'angles.csv'
    csvio.write_angles_csv(path, angles)
    assert open(path).readline().strip() == 'theta1,theta2,theta3,theta4'
    frame = pd.read_csv(path, float_precision='round_trip')
    assert list(frame.columns) == list(csvio.ANGLE_COLUMNS)
    assert np.array_equal(frame.values.astype(dtype), angles)
    assert len(pd.read_csv(path).values.tolist()) == 4000  # the call at cli.py:242
    assert np.array_equal(csvio.read_angles_csv(path, dtype=dtype), angles)


def test_shape_errors(tmp_path):
    path = tmp_path / 'angles.csv'
    csvio.write_angles_csv(path, np.zeros((3, 4)))
    with pytest.raises(ValueError):
        csvio.read_points_csv(path)
    with pytest.raises(ValueError):
        csvio.read_points_csv(path, engine='pandas')
    with pytest.raises(ValueError):
        csvio.write_points_csv(path, np.zeros((3, 4)))
    with pytest.raises(ValueError):
        csvio.read_points_csv(path, engine='nope')


def test_empty_file_round_trip(tmp_path):
    path = tmp_path / 'empty.csv'
    csvio.write_points_csv(path, np.zeros((0, 3), dtype=np.float32))
    assert csvio.read_points_csv(path, engine='pandas').shape == (0, 3)


@pytest.mark.gpu
def test_solve_csv_matches_oracle(tmp_path):
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    from inversekinematicsann_b200.robot.position_generator import TrainingDataGenerator
    from oracle import c_oracle
    pts = np.asarray(TrainingDataGenerator.spring(500, 2, 3, 6))
    src, dst = tmp_path / 'spring.csv', tmp_path / 'angles.csv'
    _reference_write(src, pts.tolist(), csvio.POINT_COLUMNS)
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    angles = csvio.solve_csv(ik, src, dst)
    want = c_oracle.fabrik_ikine(pts.astype(np.float32).astype(np.float64))["angles"]
    assert np.abs(angles - want).max() <= 1e-4
    assert np.abs(pd.read_csv(dst).values - want).max() <= 1e-4
