"""ctypes binding of oracle/ik_oracle.c (TEST INFRASTRUCTURE -- see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libik_oracle.so")
_LIB = None

# robot constants, reference robot/robot.py:40-42 (rows theta, epsilon, a, alpha -- forward.py:16)
DH = np.array([[0.0, np.pi / 2, 0.0, 0.0],
               [2.0, 0.0, 0.0, 0.0],
               [0.0, 2.0, 2.0, 2.0],
               [np.pi / 2, 0.0, 0.0, 0.0]], dtype=np.float64)
LINKS = np.array([2.0, 2.0, 2.0, 2.0], dtype=np.float64)
LIMITS = np.array([0.0, 6.0, -6.0, 6.0, -3.0, 6.0], dtype=np.float64)

STATUS_OK, STATUS_ZERO_DIVISION, STATUS_MATH_DOMAIN, STATUS_FK_ANGLE_RANGE = 0, 1, 2, 3


def build(force: bool = False) -> str:
    """Compile ik_oracle.c with the system gcc (OpenMP when libgomp is usable)."""
    src = os.path.join(_HERE, "ik_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    base = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", _LIB_PATH, src, "-lm"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for extra in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + extra + base, check=True, capture_output=True)
                return _LIB_PATH
            except (OSError, subprocess.CalledProcessError) as exc:
                last = exc
    raise RuntimeError(f"could not build the CPU oracle: {last}")


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.iko_fabrik_ikine.restype = ctypes.c_int64
        L.iko_fabrik_ikine.argtypes = [dp, dp, dp, ctypes.c_double, ctypes.c_int, dp,
                                       ctypes.c_int64, dp, dp, ip, ip]
        L.iko_check_limits.restype = ctypes.c_int64
        L.iko_check_limits.argtypes = [dp, ctypes.c_int64, dp]
        L.iko_fk_chain.restype = ctypes.c_int
        L.iko_fk_chain.argtypes = [dp, dp, dp]
        L.iko_fk_positions.restype = ctypes.c_int
        L.iko_fk_positions.argtypes = [dp, dp, ctypes.c_int64, dp, dp, dp]
        L.iko_fabrik_calculate_one.restype = ctypes.c_int
        L.iko_fabrik_calculate_one.argtypes = [dp, dp, dp, ctypes.c_double, ctypes.c_int, dp, ip]
        L.iko_distance.restype = ctypes.c_double
        L.iko_distance.argtypes = [dp, dp]
        L.iko_point_between_c.restype = ctypes.c_int
        L.iko_point_between_c.argtypes = [dp, dp, ctypes.c_double, ctypes.c_int, dp]
        L.iko_num_threads.restype = ctypes.c_int
        L.iko_set_num_threads.argtypes = [ctypes.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def set_num_threads(n: int) -> None:
    lib().iko_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().iko_num_threads())


def check_limits(xyz, limits=LIMITS) -> int:
    xyz = _f64(xyz, (-1, 3))
    return int(lib().iko_check_limits(_dp(xyz), xyz.shape[0], _dp(_f64(limits))))


def fabrik_ikine(xyz, dh=DH, links=LINKS, limits=LIMITS, tol=1e-3, max_iter=100, want_chain=False):
    """reference inverse.py:115-139.  Returns dict(first_bad, angles[n,4], iters[n], status[n], chain)."""
    xyz = _f64(xyz, (-1, 3))
    n = xyz.shape[0]
    angles = np.full((n, 4), np.nan)
    chain = np.full((n, 4, 3), np.nan) if want_chain else None
    iters = np.zeros(n, dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    bad = lib().iko_fabrik_ikine(_dp(_f64(dh)), _dp(_f64(links)), _dp(_f64(limits)), tol, max_iter,
                                 _dp(xyz), n, _dp(angles), _dp(chain) if want_chain else None,
                                 _ip(iters), _ip(status))
    return {"first_bad": int(bad), "angles": angles, "iters": iters, "status": status, "chain": chain}


def fk_chain(angles, dh=DH):
    """reference forward.py:73-94 -> (status, list of four 4x4 cumulative matrices)."""
    out = np.zeros((4, 4, 4))
    st = lib().iko_fk_chain(_dp(_f64(dh)), _dp(_f64(angles, (4,))), _dp(out))
    return int(st), out


def fk_positions(angles, dh=DH, targets=None):
    angles = _f64(angles, (-1, 4))
    n = angles.shape[0]
    pos = np.zeros((n, 3))
    err = np.zeros(n) if targets is not None else None
    tg = _f64(targets, (-1, 3)) if targets is not None else None
    st = lib().iko_fk_positions(_dp(_f64(dh)), _dp(angles), n, _dp(pos),
                                _dp(tg) if tg is not None else None,
                                _dp(err) if err is not None else None)
    return int(st), pos, err


def fabrik_calculate(init, goal, links=LINKS, tol=1e-3, max_iter=100):
    """reference fabrik.py:44-67 with an explicit initial pose (tests/fabrik_unit.py)."""
    out = np.zeros((4, 3))
    it = np.zeros(1, dtype=np.int32)
    st = lib().iko_fabrik_calculate_one(_dp(_f64(init, (4, 3))), _dp(_f64(goal, (3,))),
                                        _dp(_f64(links)), tol, max_iter, _dp(out), _ip(it))
    return int(st), out, int(it[0])


def distance(a, b) -> float:
    return float(lib().iko_distance(_dp(_f64(a, (3,))), _dp(_f64(b, (3,)))))


def point_between(a, b, dist=None):
    out = np.zeros(3)
    st = lib().iko_point_between_c(_dp(_f64(a, (3,))), _dp(_f64(b, (3,))),
                                   0.0 if dist is None else float(dist), 1 if dist is None else 0,
                                   _dp(out))
    return int(st), out
