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
