"""Robot description -- same names and values as reference robot/robot.py:38-46.

The arm the reference calls "6 DOF" has four joint angles: a base yaw and three pitch joints,
four links of length 2.  The DH table is stored the reference's way, one ROW per parameter:
``[thetas, epsilons, a, alphas]`` (reference forward.py:16).
"""
from math import pi


class SixDOFRobot:
    """Constants consumed by the IK/FK classes (reference robot/robot.py:38-42)."""
    dh_matrix = [[0, pi / 2, 0, 0],   # theta_i  (seed pose of FabrikInverseKinematics.ikine)
                 [2, 0, 0, 0],        # epsilon_i: offset along z
                 [0, 2, 2, 2],        # a_i: link length along x
                 [pi / 2, 0, 0, 0]]   # alpha_i: twist about x
    effector_workspace_limits = {'x': [0, 6], 'y': [-6, 6], 'z': [-3, 6]}
    links_lengths = [2, 2, 2, 2]


class OutOfRobotReachException(Exception):
    """Raised for targets outside the workspace box (reference inverse.py:32-35) and for joint
    angles outside [-2pi, 2pi] in forward kinematics (reference forward.py:23-25)."""
