"""CPU-only tests: the C ABI library loads and exports every symbol include/ikb200.h declares, the
host-side mirror keeps the reference's value types and error behaviour, and there is no CPU
fallback for the compute path."""
import ctypes
import os
import re

import numpy as np
import pytest

import conftest as C

ROOT = C.ROOT


def test_library_exports_every_declared_symbol():
    from inversekinematicsann_b200 import _native
    header = open(os.path.join(ROOT, "include", "ikb200.h")).read()
    declared = sorted(set(re.findall(r"\b(ikb_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_native.EXPORTED_SYMBOLS)
    lib = _native.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ikb_version().decode().startswith("ikb200")
    assert ctypes.sizeof(_native.IkbConfig) == 16 * 8 + 4 * 8 + 6 * 8 + 8 + 4 + 4
    assert ctypes.sizeof(_native.IkbStats) == 9 * 8


def test_point_reference_unit():
    """reference tests/point_unit.py:24-43."""
    from inversekinematicsann_b200.kinematics.point import Point, get_distance_between, get_point_between
    p0, p1 = Point([0, 0, 0]), Point([-2.22, 3.123, 0.002])
    with pytest.raises(ValueError) as exc:
        Point([0, 0, 0, 1])
    assert str(exc.value) == "3D Point input shape should be (3,) not (4,)"
    assert [0, 0, 0] == p0 and [-2.22, 3.123, 0.002] == p1
    np.testing.assert_almost_equal(C.REF_POINT_UNIT_DISTANCE, get_distance_between(p0, p1))
    np.testing.assert_almost_equal((np.array(p0) + np.array(p1)) / 2, get_point_between(p0, p1))
    assert str(Point([1, 2, 7])) == "Point(1, 2, 7)" and (p1.x, p1.y, p1.z) == (-2.22, 3.123, 0.002)
    with pytest.raises(ZeroDivisionError):
        get_point_between(p0, p0, 2.0)


def test_robot_constants():
    from inversekinematicsann_b200.robot.robot import OutOfRobotReachException, SixDOFRobot
    assert SixDOFRobot.dh_matrix == [[0, np.pi / 2, 0, 0], [2, 0, 0, 0], [0, 2, 2, 2], [np.pi / 2, 0, 0, 0]]
    assert SixDOFRobot.effector_workspace_limits == {'x': [0, 6], 'y': [-6, 6], 'z': [-3, 6]}
    assert SixDOFRobot.links_lengths == [2, 2, 2, 2] and issubclass(OutOfRobotReachException, Exception)


def test_input_marshalling():
    from inversekinematicsann_b200.kinematics._shared import points_to_array
    from inversekinematicsann_b200.kinematics.point import Point
    a = points_to_array([[1, 2, 3], Point([4, 5, 6])])
    assert a.dtype == np.float64 and a.shape == (2, 3)
    f32 = np.ones((5, 3), dtype=np.float32)
    assert points_to_array(f32) is f32
    assert points_to_array([]).shape == (0, 3)
    with pytest.raises(ValueError):
        points_to_array([[1, 2, 3, 4]])
    with pytest.raises(ValueError):
        points_to_array([[1, 2, 3], [1, 2]])
    with pytest.raises(TypeError):
        points_to_array([["x", "y", "z"]])


def test_constructors_need_no_gpu_but_solving_does():
    import torch
    from inversekinematicsann_b200.kinematics.inverse import AnnInverseKinematics, FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    ann = AnnInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    assert ik.ikine([]) == [] and ann.ikine([]) == []
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ik.ikine([[1, 2, 3]])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "inversekinematicsann_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)


def test_product_and_oracle_synthetic_weights_agree():
    """bench.py's GPU arm takes its synthetic network from the product package, the checker from oracle/; they are
    independent restatements of the same seeded construction."""
    from inversekinematicsann_b200 import models
    from oracle import np_oracle
    Wp, bp = models.synthetic_weights(seed=5)
    Wo, bo = np_oracle.synthetic_mlp(seed=5)
    assert all(np.array_equal(a, b) for a, b in zip(Wp, Wo)) and all(np.array_equal(a, b) for a, b in zip(bp, bo))
    for name in ("SHIPPED_MEAN_X", "SHIPPED_SCALE_X", "SHIPPED_MEAN_Y", "SHIPPED_SCALE_Y"):
        assert np.array_equal(getattr(models, name), getattr(np_oracle, name))
    assert models.REFERENCE_LAYER_DIMS == np_oracle.LAYER_DIMS


def test_out_buffer_validation_happens_before_the_c_abi():
    """`out=` reaches libikb200 as a raw pointer: wrong shape / dtype / strides must be refused in Python."""
    from inversekinematicsann_b200.engine import IkEngine
    ok = np.empty((5, 4), dtype=np.float32)
    assert IkEngine._check_out(ok, 5, 4, (np.float32, np.float64), "out") is ok
    with pytest.raises(ValueError):
        IkEngine._check_out(np.empty((5, 3), np.float32), 5, 4, (np.float32,), "out")      # too small
    with pytest.raises(ValueError):
        IkEngine._check_out(np.empty((4, 4), np.float32), 5, 4, (np.float32,), "out")      # too few rows
    with pytest.raises(ValueError):
        IkEngine._check_out(np.empty((5, 8), np.float32)[:, ::2], 5, 4, (np.float32,), "out")  # strided view
    with pytest.raises(TypeError):
        IkEngine._check_out(np.empty((5, 4), np.float64), 5, 4, (np.float32,), "out")      # ANN writes float32
    with pytest.raises(TypeError):
        IkEngine._check_out([[0.0] * 4] * 5, 5, 4, (np.float32,), "out")
    ro = np.empty((5, 4), np.float32)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        IkEngine._check_out(ro, 5, 4, (np.float32,), "out")


def test_unknown_mode_or_precision_is_a_value_error():
    from inversekinematicsann_b200 import engine
    assert engine._choice(engine._MLP_MODES, "fp16x3_ts", "mode") == engine._native.IKB_MLP_FP16X3_TS
    for table, key, what in ((engine._MLP_MODES, "bf16", "mode"), (engine._FABRIK_PRECISIONS, "f16", "precision"),
                             (engine._MLP_MODES, ["fp32"], "mode")):
        with pytest.raises(ValueError, match=what):
            engine._choice(table, key, what)


def test_raise_from_stats_same_row_in_two_error_classes():
    """A zero-division row and a domain-error row with the same index must not break the exception mapping."""
    from inversekinematicsann_b200.engine import IkStats
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    with pytest.raises(ZeroDivisionError):
        ik._raise_from_stats([[0, 0, 2]] * 4, IkStats(first_zero_division=3, first_domain_error=3))
    with pytest.raises(ValueError):
        ik._raise_from_stats([[0, 0, 2]] * 4, IkStats(first_zero_division=3, first_domain_error=1))


def test_save_model_keeps_the_reference_file_names(tmp_path):
    """ann.py:87-95: `<prefix>_<stamp>.h5` + `_scaler_x.bin` + `_scaler_y.bin`; load_model finds them again.
    (No GPU: nothing is predicted.)"""
    import glob
    from sklearn.preprocessing import StandardScaler
    from inversekinematicsann_b200.kinematics.ann import ANN, DenseStack
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    rng = np.random.default_rng(0)
    ann = ANN(R.effector_workspace_limits, R.dh_matrix)
    ann.model = DenseStack([rng.normal(size=(3, 8)), rng.normal(size=(8, 4))], [rng.normal(size=8), rng.normal(size=4)])
    ann.x_data_skaler = StandardScaler().fit(rng.normal(size=(50, 3)))
    ann.y_data_skaler = StandardScaler().fit(rng.normal(size=(50, 4)))
    prefix = ann.save_model(str(tmp_path / "saved_model"))
    assert glob.glob(str(tmp_path / "saved_model*.h5")) and glob.glob(str(tmp_path / "saved_model*_scaler_x.bin"))
    fresh = ANN(R.effector_workspace_limits, R.dh_matrix)
    assert fresh.load_model(prefix + ".h5") is not None
    assert fresh.model.layer_dims == [3, 8, 4]
    np.testing.assert_array_equal(fresh.model.kernels[1], ann.model.kernels[1])
    np.testing.assert_allclose(fresh.x_data_skaler.mean_, ann.x_data_skaler.mean_)


def test_keras_h5_layout_round_trip(tmp_path):
    """DenseStack.load_h5 on a file in Keras' legacy layout (what tools/h5_to_npz.py converts); needs h5py."""
    pytest.importorskip("h5py")
    from inversekinematicsann_b200.kinematics.ann import DenseStack
    rng = np.random.default_rng(1)
    stack = DenseStack([rng.normal(size=(3, 16)), rng.normal(size=(16, 16)), rng.normal(size=(16, 4))],
                       [rng.normal(size=16), rng.normal(size=16), rng.normal(size=4)])
    path = str(tmp_path / "m.h5")
    stack.save_h5(path)
    assert DenseStack.is_hdf5(path)
    back = DenseStack.load_h5(path)
    assert back.layer_dims == [3, 16, 16, 4]
    for a, b in zip(stack.kernels + stack.biases, back.kernels + back.biases):
        np.testing.assert_array_equal(a, b)


def test_host_side_generators_of_the_reference():
    """position_generator.py:33-37,57-70,92-95: the shapes that stay on the host."""
    from inversekinematicsann_b200.robot.position_generator import TrainingDataGenerator as G
    lim = {'x': [0, 3], 'y': [0, 4], 'z': [-1, 5]}
    np.random.seed(3)
    pts = np.array(G.random(200, lim))
    assert pts.shape == (200, 3)
    for ax, key in enumerate('xyz'):
        assert pts[:, ax].min() == pytest.approx(lim[key][0]) and pts[:, ax].max() == pytest.approx(lim[key][1])
    np.random.seed(3)
    v = np.random.randn(200)
    np.testing.assert_allclose(pts[:, 0], (v - v.min()) / (v.max() - v.min()) * 3.0, atol=1e-12)
    shuffled = np.array(G.random_distribution(50, lim, 'random'))
    assert shuffled.shape == (50, 3)
    np.testing.assert_allclose(np.sort(shuffled[:, 2]), np.linspace(-1, 5, 50))
    circ = list(G.circle_gen(2, 5, (1, 3, 2)))
    assert circ[1] == [1, 3 * np.sin(1), 2 + 2 * np.cos(1)]
    np.random.seed(4)
    cloud = list(G.cube_random_gen(1.0, 2, 3, 1, start=(1, 1, 1)))
    np.random.seed(4)
    assert len(cloud) == 6 and cloud[0] == [2 * np.random.rand() + 1, 3 * np.random.rand() + 1, np.random.rand() + 1]
