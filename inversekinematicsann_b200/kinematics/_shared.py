"""Helpers shared by the kinematics mirror: engine cache and input marshalling."""
import os

import numpy as np

from ..engine import IkEngine
from ..robot.robot import SixDOFRobot

_ENGINES = {}


def default_device():
    """CUDA ordinal for this process: IKB_DEVICE, else LOCAL_RANK (torchrun), else 0."""
    for var in ("IKB_DEVICE", "LOCAL_RANK"):
        if os.environ.get(var, "") != "":
            return int(os.environ[var])
    return 0


def _freeze(x):
    if isinstance(x, dict):
        return tuple((k, _freeze(v)) for k, v in x.items())
    return tuple(float(v) for v in np.asarray(x, dtype=np.float64).reshape(-1))


def get_engine(dh_matrix=None, joints_distances=None, workspace_limits=None, max_err=0.001,
               max_iterations_num=100, device=None):
    """One IkEngine per distinct (robot, solver, device) description; created on first use so that
    constructing the IK classes never needs a GPU, only solving does."""
    dh = SixDOFRobot.dh_matrix if dh_matrix is None else dh_matrix
    links = SixDOFRobot.links_lengths if joints_distances is None else joints_distances
    limits = SixDOFRobot.effector_workspace_limits if workspace_limits is None else workspace_limits
    dev = default_device() if device is None else int(device)
    # theta_1 of the seed row is overwritten per target upstream (inverse.py:125): not part of the key
    dh_rows = [list(r) for r in dh]
    key_dh = list(dh_rows)
    key_dh[0] = [0.0] + list(dh_rows[0][1:])
    key = (_freeze(key_dh), _freeze(links), _freeze(limits), float(max_err), int(max_iterations_num), dev)
    eng = _ENGINES.get(key)
    if eng is None:
        eng = IkEngine(key_dh, links, limits, max_err, max_iterations_num, device=dev)
        _ENGINES[key] = eng
    return eng


def release_engines():
    for eng in _ENGINES.values():
        eng.close()
    _ENGINES.clear()


def points_to_array(dest_points):
    """list of [x,y,z] / Point / ndarray rows -> contiguous (n, 3) float array.

    float32 ndarrays are passed through (the kernels read fp32 or fp64); everything else becomes
    float64 like the python floats the reference computes with.  Non-numeric content raises
    TypeError, ragged or wrongly shaped input raises ValueError."""
    if isinstance(dest_points, np.ndarray) and dest_points.dtype in (np.float32, np.float64):
        arr = dest_points
    else:
        try:
            arr = np.asarray(dest_points, dtype=np.float64)
        except (TypeError, ValueError) as exc:
            kind = TypeError if isinstance(exc, TypeError) or _has_non_numeric(dest_points) else ValueError
            raise kind(f'destination points must be numeric [x, y, z] rows: {exc}') from exc
    if arr.size == 0:
        return np.zeros((0, 3), dtype=np.float64)
    if arr.ndim != 2 or arr.shape[1] != 3:
        raise ValueError(f'destination points must have shape (n, 3), not {arr.shape}')
    return np.ascontiguousarray(arr)


def _has_non_numeric(points):
    try:
        return any(isinstance(v, (str, bytes, type(None))) for row in points for v in row)
    except TypeError:
        return True
