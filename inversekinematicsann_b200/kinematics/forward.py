"""DH forward kinematics -- API of reference kinematics/forward.py, computed by csrc/fk.cu.

``ForwardKinematics(dh_matrix).fkine(angles)`` returns ``(T_final, [T_1 .. T_4])`` as 4x4 NumPy
arrays exactly like the reference (forward.py:73-94); callers read the position as T[0:3, 3]
(cli.py:60, inverse.py:130).  ``fkine_positions`` is the batched form used for the
position-error check of both solvers.
"""
import numpy as np

from ..robot.robot import OutOfRobotReachException
from ._shared import get_engine

_ANGLE_MSG = 'Forward Kinematics exception, robot joints angles limits are (-2pi, 2pi)'  # forward.py:24-25


class ForwardKinematics:
    """Planar robotic arm forward kinematics"""

    def __init__(self, dh_matrix, device=None):
        assert all(len(row) == len(dh_matrix[0]) for row in dh_matrix)
        self.dh_matrix = dh_matrix
        self.thetas, self.epsilons, self.ais, self.alphas = self.dh_matrix
        self.no_of_features = len(self.thetas)
        assert self.no_of_features >= 3
        if self.no_of_features != 4:
            # upstream's matrices are no_of_features x no_of_features and only work as homogeneous
            # transforms for 4 joints; the engine is built for exactly that case
            raise ValueError('the DH table must describe 4 joints')
        self._device = device

    def _engine(self):
        return get_engine(dh_matrix=self.dh_matrix, device=self._device)

    def fkine(self, angles):
        """All cumulative DH transforms for one set of joint angles -> (T_4, [T_1, T_2, T_3, T_4])."""
        self.thetas = angles  # upstream side effect (forward.py:77)
        chain, status = self._engine().fk_chain(angles)
        if status != 0:
            raise OutOfRobotReachException(_ANGLE_MSG)
        mats = [np.array(chain[i]) for i in range(4)]
        return mats[-1], mats

    def fkine_positions(self, angles, targets=None):
        """Batched end-effector positions for (n, 4) angles; with `targets` also ||pos - target||."""
        pos, err, stats = self._engine().fk(angles, targets)
        if stats.first_fk_angle_range >= 0:
            raise OutOfRobotReachException(_ANGLE_MSG)
        return (pos, err) if targets is not None else pos
