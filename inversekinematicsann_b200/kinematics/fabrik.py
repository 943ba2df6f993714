"""FABRIK solver -- API of reference kinematics/fabrik.py, computed by csrc/fabrik.cu.

``Fabrik(joints_distances, err_margin, max_iter_num).calculate(init_positions, goal)`` keeps the
reference's signature and return type (fabrik.py:13,44-67: a list of four ``Point``).  The chain is
solved on the GPU by the generic 3-D kernel, which follows the reference's operation order in IEEE
fp64; ``calculate_batch`` solves many goals in one launch.
"""
import numpy as np

from ._shared import get_engine
from .point import Point


class Fabrik:
    """Forward And Backward Reaching Inverse Kinematics on a 4-point chain."""

    def __init__(self, joints_distances, err_margin=0.001, max_iter_num=100, device=None):
        self.joints_distances = joints_distances
        self.err_margin = err_margin
        self.max_iter_num = max_iter_num
        self._device = device

    def _engine(self):
        return get_engine(joints_distances=self.joints_distances, max_err=self.err_margin,
                          max_iterations_num=self.max_iter_num, device=self._device)

    def calculate_batch(self, init_joints_positions, goal_effector_positions):
        """(4,3) or (n,4,3) initial chains, (n,3) goals -> (chains[n,4,3], iterations[n])."""
        init = np.asarray(init_joints_positions, dtype=np.float64)
        if init.shape[-2] != len(self.joints_distances):
            raise ValueError('Input vectors should have equal lengths!')  # fabrik.py:46-48
        chains, iters, stats = self._engine().fabrik_calculate(init, goal_effector_positions)
        if stats.first_zero_division >= 0:
            raise ZeroDivisionError('float division by zero')  # point.py:40 upstream
        return chains, iters

    def calculate(self, init_joints_positions, goal_effector_position):
        """Joint positions after FABRIK from `init_joints_positions` to the goal (fabrik.py:44-67)."""
        if len(init_joints_positions) != len(self.joints_distances):
            raise ValueError('Input vectors should have equal lengths!')
        goal = Point(goal_effector_position)
        chains, _ = self.calculate_batch([list(p) for p in init_joints_positions], [list(goal)])
        return [Point(row.tolist()) for row in chains[0]]
