"""3-D point value type and the two scalar geometry helpers of reference kinematics/point.py.

Host-side I/O only: the batched device equivalents are the inlined point-at-distance /
distance steps of the FABRIK kernel (csrc/fabrik.cu).  Behaviour mirrored from the reference:
``Point`` is a ``list`` with ``.x .y .z`` (point.py:10-22), construction validates the shape
(point.py:13-14), ``str`` gives ``Point(x, y, z)`` (point.py:18-19).
"""
from math import sqrt

import numpy as np


class Point(list):
    """3D point: a list [x, y, z] that also exposes .x .y .z"""

    def __init__(self, xyz):
        shape = np.array(xyz).shape
        if shape != (3,):
            raise ValueError(f'3D Point input shape should be (3,) not {shape}')
        super().__init__(xyz)
        self.x, self.y, self.z = xyz

    def __str__(self):
        return f'Point{(self.x, self.y, self.z)}'

    def __repr__(self):
        return f'<Point at 0x{id(self):x}, x={self.x}, y={self.y}, z={self.z}>'


def get_distance_between(point_a, point_b):
    """Euclidean distance (reference point.py:25-29)."""
    return sqrt((point_a.x - point_b.x) ** 2 + (point_a.y - point_b.y) ** 2 + (point_a.z - point_b.z) ** 2)


def get_point_between(start_point, end_point, distance=None):
    """Point on the ray start -> end at `distance` from start; the midpoint when distance is None
    (reference point.py:32-45).  A zero-length segment raises ZeroDivisionError as upstream."""
    span = get_distance_between(start_point, end_point)
    if distance is None:
        distance = span / 2
    ratio = distance / span
    return Point([s + ratio * (e - s) for s, e in zip(start_point, end_point)])
