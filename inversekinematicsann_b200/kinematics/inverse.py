"""Inverse kinematics front classes -- API of reference kinematics/inverse.py.

``FabrikInverseKinematics.ikine`` and ``AnnInverseKinematics.ikine`` keep the reference's
signatures, return types (list of [theta1..theta4] python floats) and exceptions, so cli.py
(--inverse-kine) and rpc_broker.py run on top of them unchanged.  The per-target python loop of the
reference (inverse.py:120-137) becomes ONE call into libikb200.so for the whole trajectory; pass
``as_array=True`` (and float32/float64 ndarrays in) to skip the python list round trip.

Differences from upstream, all documented in DESIGN.md:
  * ``ikine`` does not write theta_1 of the last target into the caller's ``dh_matrix[0][0]``
    (upstream side effect at inverse.py:125, asserted by no test).
  * degenerate targets (on the z axis / coincident joints) give defined results (NaN angles +
    ZeroDivisionError for a zero-length segment) instead of rounding-noise-dependent ones.
"""
from abc import ABC, abstractmethod

import numpy as np

from ..robot.robot import OutOfRobotReachException
from ._shared import get_engine, points_to_array
from .ann import ANN
from .forward import ForwardKinematics


class InverseKinematics(ABC):
    """Inverse kinematics base (reference inverse.py:18-40)."""

    def __init__(self, dh_matrix, joints_distances, workspace_limits, device=None):
        self.dh_matrix = dh_matrix
        self.joints_distances = joints_distances
        self.workspace_limits = workspace_limits
        self.fkine = ForwardKinematics(self.dh_matrix, device=device)
        self._device = device
        self._max_err = 0.001
        self._max_iter = 100

    def _engine(self):
        return get_engine(self.dh_matrix, self.joints_distances, self.workspace_limits,
                          self._max_err, self._max_iter, self._device)

    def _limits_error(self, dest_points, index):
        # message text of reference inverse.py:32-35; `dest_points[index]` prints as upstream does
        # ([1, 2, 7] for a list row, Point(1, 2, 7) for a Point)
        return OutOfRobotReachException(
            f'Inverse Kinematics exception, point {dest_points[index]} '
            'is out of manipulator reach area! '
            f'Limits: {self.workspace_limits}')

    def check_limits(self, dest_points):
        """Raise OutOfRobotReachException for the first point outside the workspace box."""
        arr = points_to_array(dest_points)
        if arr.shape[0] == 0:
            return
        first_bad = self._engine().check_limits(arr)
        if first_bad >= 0:
            raise self._limits_error(dest_points, first_bad)

    def _raise_from_stats(self, dest_points, stats):
        if stats.first_out_of_limits >= 0:  # whole-batch semantics: limits first (inverse.py:117)
            raise self._limits_error(dest_points, stats.first_out_of_limits)
        rows = [(r, kind) for r, kind in ((stats.first_zero_division, ZeroDivisionError),
                                          (stats.first_domain_error, ValueError)) if r >= 0]
        if rows:
            _, kind = min(rows, key=lambda t: t[0])  # same row in both classes: the division comes first upstream
            raise kind('float division by zero' if kind is ZeroDivisionError else 'math domain error')

    @abstractmethod
    def ikine(self, dest_points):
        """Calculate inverse kinematics"""


class FabrikInverseKinematics(InverseKinematics):
    """Reaching inverse kinematics using the FABRIK method (reference inverse.py:45-139)."""

    def __init__(self, dh_matrix, joints_distances, workspace_limits,
                 max_err=0.001, max_iterations_num=100, device=None, precision='f64'):
        super().__init__(dh_matrix, joints_distances, workspace_limits, device=device)
        from .fabrik import Fabrik
        self.fabrik = Fabrik(joints_distances, max_err, max_iterations_num, device=device)
        self._max_err = max_err
        self._max_iter = max_iterations_num
        self.precision = precision
        self.last_stats = None

    def ikine(self, dest_points, as_array=False, return_iterations=False, out=None, return_fk_error=False):
        """Joint angles [theta1..theta4] for every destination point.

        `out` (optional, implies as_array): a preallocated (n, 4) float32/float64 array -- e.g. pinned
        host memory -- that receives the angles.  `return_fk_error` appends ||FK(angles) - point|| per point (the
        check the reference CLI plots, cli.py:56-61), computed inside the solver kernel."""
        arr = points_to_array(dest_points)
        if arr.shape[0] == 0:
            empty = np.zeros((0, 4)) if as_array else []
            extras = ((np.zeros(0, np.int32),) if return_iterations else ()) + \
                     ((np.zeros(0),) if return_fk_error else ())
            return (empty,) + extras if extras else empty
        if out is not None:
            as_array = True
        res = self._engine().fabrik_solve(arr, out=out, precision=self.precision, return_iters=return_iterations,
                                          return_fk_error=return_fk_error, fk_stats=return_fk_error)
        angles, stats = res[0], res[1]
        self.last_stats = stats
        self._raise_from_stats(dest_points, stats)
        out = angles if as_array else angles.tolist()
        return (out,) + tuple(res[2:]) if len(res) > 2 else out


class AnnInverseKinematics(InverseKinematics):
    """Reaching inverse kinematics using the neural network (reference inverse.py:142-155)."""

    def __init__(self, dh_matrix, joints_distances, workspace_limits, device=None):
        super().__init__(dh_matrix, joints_distances, workspace_limits, device=device)
        self.ann = ANN(workspace_limits, dh_matrix, device=device)
        self.last_stats = None

    def load_model(self, model_name):
        """Load model weights + scalers (reference inverse.py:148-150)."""
        self.ann.load_model(model_name)

    def ikine(self, dest_points, as_array=False, out=None, return_fk_error=False):
        """Predict thetas using the neural network (limits checked first, inverse.py:154)."""
        arr = points_to_array(dest_points)
        if arr.shape[0] == 0:
            empty = np.zeros((0, 4), dtype=np.float32) if as_array else []
            return (empty, np.zeros(0, np.float32)) if return_fk_error else empty
        if out is not None:
            as_array = True
        res = self.ann.predict_with_stats(arr, out=out, return_fk_error=return_fk_error)
        angles, stats = res[0], res[1]
        self.last_stats = stats
        self._raise_from_stats(dest_points, stats)
        angles = angles if as_array else angles.tolist()
        return (angles, res[2]) if return_fk_error else angles
