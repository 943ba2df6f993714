"""Trajectory / training-data shapes -- API of reference robot/position_generator.py, generated on the GPU.

``TrainingDataGenerator.circle / cube / cube_random / spring / random_distribution`` keep the reference's
signatures and return python lists of [x, y, z] (position_generator.py:26-97).  Every shape also has a
``*_device`` form that leaves the points in HBM as a torch tensor (float32 by default), so 1e8-1e9 point
workloads never touch the host.  The random shapes draw from a counter-based Philox stream keyed by
(seed, row): the same distributions as the reference's ``np.random.rand`` / ``scipy.stats.truncnorm`` calls
and reproducible per seed, but not numpy's Mersenne-Twister numbers.
The shapes upstream only ever builds on the host in small numbers stay host-side NumPy here:
``random`` (position_generator.py:65-70; upstream it fails on current scikit-learn because ``minmax_scale`` rejects
a list ``feature_range`` -- the same min-max mapping is done directly), ``distribution='random'`` (:92-95, a shuffled
linspace) and the python-generator variants ``circle_gen`` / ``cube_random_gen`` (:33-37, 57-63).
"""
import ctypes
import math

import numpy as np

from .. import _native
from ..kinematics._shared import get_engine

GEN_CIRCLE, GEN_SPRING, GEN_CUBE, GEN_CUBE_RANDOM, GEN_NORMAL = range(5)


def transpose(data):
    """Transpose a list of lists (reference position_generator.py:14-16)."""
    return list(map(list, zip(*data)))


def _arange_len(length, step):
    return int(len(np.arange(0, length, step)))


def _generate(kind, params, n, dtype="float32", seed=0, row_offset=0, device=None):
    import torch
    eng = get_engine(device=device)
    tdtype = torch.float32 if dtype in ("float32", np.float32, torch.float32) else torch.float64
    out = torch.empty((int(n), 3), dtype=tdtype, device=f"cuda:{eng.device}")
    if n == 0:
        return out
    p = (ctypes.c_double * len(params))(*[float(v) for v in params])
    code = _native.IKB_F32 if tdtype == torch.float32 else _native.IKB_F64
    stream = ctypes.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
    rc = eng._lib.ikb_generate_device(eng._handle, kind, p, len(params), int(n), int(row_offset), out.data_ptr(),
                                      code, ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), stream)
    eng._check(rc, "ikb_generate_device")
    return out


class TrainingDataGenerator:
    """Shapes used to generate training / test / benchmark trajectories."""

    # ---- device forms -------------------------------------------------------------------------------
    @staticmethod
    def circle_device(radius, no_of_samples, centre, **kw):
        return _generate(GEN_CIRCLE, [radius, centre[0], centre[1], centre[2]], no_of_samples, **kw)

    @staticmethod
    def spring_device(no_of_samples, len_x, len_y, len_z, **kw):
        return _generate(GEN_SPRING, [len_x, len_y, len_z, no_of_samples], no_of_samples, **kw)

    @staticmethod
    def cube_device(step, len_x, len_y, len_z, start=(0, 0, 0), **kw):
        nx, ny, nz = (_arange_len(l, step) for l in (len_x, len_y, len_z))
        return _generate(GEN_CUBE, [step, len_x, len_y, len_z, start[0], start[1], start[2], nx, ny], nx * ny * nz, **kw)

    @staticmethod
    def cube_random_device(step, len_x, len_y, len_z, start=(0, 0, 0), seed=1234, no_of_samples=None, **kw):
        """`no_of_samples` overrides the reference's count len(arange(0, volume, step)) (position_generator.py:50),
        whose floating-point step arithmetic is awkward for an exact 1e8."""
        n = _arange_len(len_x * len_y * len_z, step) if no_of_samples is None else int(no_of_samples)
        return _generate(GEN_CUBE_RANDOM, [len_x, len_y, len_z, start[0], start[1], start[2]], n, seed=seed, **kw)

    @staticmethod
    def random_distribution_device(no_of_samples, limits, distribution='normal', std_dev=0.5, seed=1234, **kw):
        lim = [limits[a][i] for a in ('x', 'y', 'z') for i in (0, 1)]
        if distribution == 'normal':
            return _generate(GEN_NORMAL, lim + [std_dev], no_of_samples, seed=seed, **kw)
        if distribution == 'uniform':
            lens = [lim[1] - lim[0], lim[3] - lim[2], lim[5] - lim[4]]
            return _generate(GEN_CUBE_RANDOM, lens + [lim[0], lim[2], lim[4]], no_of_samples, seed=seed, **kw)
        raise ValueError("distribution must be 'normal' or 'uniform' on the device ('random' = shuffled linspace is host-side)")

    # ---- reference-shaped forms (python lists) ---------------------------------------------------------
    @staticmethod
    def circle(radius, no_of_samples, centre):
        """Circle shape (position_generator.py:26-31)."""
        return TrainingDataGenerator.circle_device(radius, no_of_samples, centre, dtype="float64").cpu().tolist()

    @staticmethod
    def cube(step, len_x, len_y, len_z, start=(0, 0, 0)):
        """Cube shape (position_generator.py:39-46)."""
        return TrainingDataGenerator.cube_device(step, len_x, len_y, len_z, start, dtype="float64").cpu().tolist()

    @staticmethod
    def cube_random(step, len_x, len_y, len_z, start=(0, 0, 0), seed=1234):
        """Cube shaped random point cloud (position_generator.py:48-55)."""
        return TrainingDataGenerator.cube_random_device(step, len_x, len_y, len_z, start, seed=seed,
                                                        dtype="float64").cpu().tolist()

    @staticmethod
    def spring(no_of_samples, len_x, len_y, len_z):
        """Horizontal spring shape (position_generator.py:72-78)."""
        return TrainingDataGenerator.spring_device(no_of_samples, len_x, len_y, len_z, dtype="float64").cpu().tolist()

    @staticmethod
    def random(no_of_samples, limits):
        """Random normal distribution min-max scaled into the limits (position_generator.py:65-70): per axis
        `minmax_scale(np.random.randn(n), limits[axis])`, i.e. lo + (v - min v) / (max v - min v) * (hi - lo)."""
        def apply_limits(axis):
            v = np.random.randn(no_of_samples)
            lo, hi = limits[axis]
            span = v.max() - v.min()
            return lo + (v - v.min()) / (span if span > 0 else 1.0) * (hi - lo)
        return [[x, y, z] for x, y, z in zip(apply_limits('x'), apply_limits('y'), apply_limits('z'))]

    @staticmethod
    def circle_gen(radius, no_of_samples, centre):
        """Circle shape generator, formula as upstream INCLUDING its y term `centre[1] * sin(t)` (no radius, no
        offset: position_generator.py:33-37 differs from `circle` there)."""
        for tstamp in range(no_of_samples):
            yield [centre[0], centre[1] * math.sin(tstamp), centre[2] + radius * math.cos(tstamp)]

    @staticmethod
    def cube_random_gen(step, len_x, len_y, len_z, start=(0, 0, 0)):
        """Random cube as a python generator (position_generator.py:57-63): numpy's global stream, x, y, z order."""
        for _ in np.arange(0, len_x * len_y * len_z, step):
            yield [len_x * np.random.rand() + start[0], len_y * np.random.rand() + start[1],
                   len_z * np.random.rand() + start[2]]

    @staticmethod
    def random_distribution(no_of_samples, limits, distribution='normal', std_dev=0.5, seed=1234):
        """Randomly generated data (position_generator.py:80-97)."""
        if distribution == 'random':   # just randomly shuffled data (:92-95)
            positions = []
            for _, limitv in limits.items():
                arr = np.linspace(limitv[0], limitv[1], no_of_samples)
                np.random.shuffle(arr)
                positions.append(arr.tolist())
            return transpose(positions)
        return TrainingDataGenerator.random_distribution_device(no_of_samples, limits, distribution, std_dev,
                                                                seed=seed, dtype="float64").cpu().tolist()


def get_truncated_normal_distribution(mean=0, std_dev=1, low=0, upp=10):
    """Truncated normal distribution object (position_generator.py:18-20), host side (scipy)."""
    from scipy.stats import truncnorm
    return truncnorm((low - mean) / std_dev, (upp - mean) / std_dev, loc=mean, scale=std_dev)


__all__ = ["TrainingDataGenerator", "transpose", "get_truncated_normal_distribution", "math"]
