"""Broker wire path: the request/reply handling of reference rpc_broker.py:76-100 without
the AMQP transport, plus a binary payload for large trajectories.

``handle_request(engine, body)`` does what ``IkineRPCBroker.callback`` does between receiving ``body`` and
publishing the reply: decode -> ``engine.ikine(positions)`` -> ``{'status': 'OK', 'angles': ...}`` or
``{'status': 'ERROR', 'reason': str(e), 'correlation_id': ...}`` (rpc_broker.py:70-74,84-91).  Unlike
upstream it also maps ``ZeroDivisionError`` (and a malformed message) to an ERROR reply instead of killing the
consumer.

JSON (``{"positions": [[x, y, z], ...]}``) stays supported for compatibility.  The binary form avoids the
~50 bytes and ~1 us of python object per point that make one 10 M-point JSON message impractical:

    request : b'IKB1' | uint32 dtype (0 = float32, 1 = float64) | uint64 n | n*3 little-endian values
    reply   : b'IKB1' | uint32 status | uint64 n | n*4 little-endian angles
              status bits 0-7: 0 OK, 1 ERROR; bits 8-15: dtype of the angles (0 = float32, 1 = float64)
              (ERROR: n = byte length of the UTF-8 reason that follows)

The reply carries the request's dtype (FABRIK) or float32 (ANN: Keras / sklearn return float32, ann.py:70-76).
It is assembled in place: header and angles share ONE page-locked buffer per engine that the solve's
device-to-host copies write into directly (``ikine(out=...)``), so a 10 M-point reply costs no allocation and
no host copy.  ``zero_copy=True`` returns a memoryview of that buffer, valid until the next request on the same
engine (the reference broker publishes the reply before it takes the next message: prefetch_count=1,
rpc_broker.py:67,96-100); the default returns an independent ``bytes`` copy.
"""
import json
import struct

import numpy as np

from .robot.robot import OutOfRobotReachException

MAGIC = b"IKB1"
_HEADER = struct.Struct("<4sIQ")
_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("<f8")}
_CODES = {np.dtype("float32"): 0, np.dtype("float64"): 1}
STATUS_OK, STATUS_ERROR = 0, 1


def encode_binary_request(points):
    arr = np.ascontiguousarray(points)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    if arr.ndim != 2 or arr.shape[1] != 3:
        raise ValueError(f"points must have shape (n, 3), not {arr.shape}")
    code = _CODES[arr.dtype]
    return _HEADER.pack(MAGIC, code, arr.shape[0]) + arr.astype(_DTYPES[code], copy=False).tobytes()


def _unpack_header(body, what):
    if len(body) < _HEADER.size:
        raise ValueError(f"IKB1 {what} is {len(body)} bytes, shorter than its {_HEADER.size}-byte header")
    magic, word, n = _HEADER.unpack_from(body, 0)
    if magic != MAGIC:
        raise ValueError(f"not an IKB1 binary {what}")
    return word, n


def decode_binary_request(body):
    code, n = _unpack_header(body, "request")
    if code not in _DTYPES:
        raise ValueError(f"IKB1 request has unknown dtype code {code}")
    need = _HEADER.size + n * 3 * _DTYPES[code].itemsize
    if len(body) != need:
        raise ValueError(f"IKB1 request announces {n} points ({need} bytes) but carries {len(body)} bytes")
    return np.frombuffer(body, dtype=_DTYPES[code], count=n * 3, offset=_HEADER.size).reshape(n, 3)


def encode_binary_reply(angles=None, error=None):
    if error is not None:
        reason = str(error).encode("utf-8")
        return _HEADER.pack(MAGIC, STATUS_ERROR, len(reason)) + reason
    arr = np.ascontiguousarray(angles)
    if arr.dtype not in _CODES:
        arr = arr.astype(np.float64)
    if arr.ndim != 2 or arr.shape[1] != 4:
        raise ValueError(f"angles must have shape (n, 4), not {arr.shape}")
    code = _CODES[arr.dtype]
    return _HEADER.pack(MAGIC, STATUS_OK | (code << 8), arr.shape[0]) + arr.astype(_DTYPES[code], copy=False).tobytes()


def decode_binary_reply(body):
    """-> {'status': 'OK', 'angles': (n, 4) ndarray in the dtype the header names} or {'status': 'ERROR', ...}."""
    word, n = _unpack_header(body, "reply")
    status, code = word & 0xFF, (word >> 8) & 0xFF
    if status != STATUS_OK:
        if len(body) != _HEADER.size + n:
            raise ValueError(f"IKB1 error reply announces {n} bytes of text but carries {len(body) - _HEADER.size}")
        return {"status": "ERROR", "reason": bytes(body[_HEADER.size:]).decode("utf-8")}
    if code not in _DTYPES:
        raise ValueError(f"IKB1 reply has unknown dtype code {code}")
    need = _HEADER.size + n * 4 * _DTYPES[code].itemsize
    if len(body) != need:
        raise ValueError(f"IKB1 reply announces {n} rows of {_DTYPES[code].name} ({need} bytes) but carries {len(body)} bytes")
    return {"status": "OK",
            "angles": np.frombuffer(body, dtype=_DTYPES[code], count=n * 4, offset=_HEADER.size).reshape(n, 4)}


_ARENAS = {}   # id(native engine) -> (engine, PinnedBuffer): one reusable reply buffer per engine


def _native_engine(ikine_engine):
    """The IkEngine behind a FabrikInverseKinematics / AnnInverseKinematics (anything else: no pinned arena)."""
    try:
        if hasattr(ikine_engine, "ann"):
            return ikine_engine.ann._ensure_uploaded()
        return ikine_engine._engine()
    except AttributeError:
        return None


def _reply_arena(ikine_engine, nbytes):
    eng = _native_engine(ikine_engine)
    if eng is None:
        return None
    held = _ARENAS.get(id(eng))
    if held is None or held[1].nbytes < nbytes:
        size = 1 << 16
        while size < nbytes:
            size <<= 1
        held = (eng, eng.pinned_buffer(size))
        _ARENAS[id(eng)] = held
    return held[1]


def release_arenas():
    for _, buf in _ARENAS.values():
        buf.close()
    _ARENAS.clear()


def _binary_reply(ikine_engine, points, zero_copy):
    n = points.shape[0]
    is_ann = hasattr(ikine_engine, "ann")
    dtype = np.dtype(np.float32) if is_ann else points.dtype
    total = _HEADER.size + n * 4 * dtype.itemsize
    arena = _reply_arena(ikine_engine, total) if n else None
    if arena is None:                      # empty request, or an engine without the native library behind it
        return encode_binary_reply(np.asarray(ikine_engine.ikine(points, as_array=True)).reshape(n, 4))
    angles = arena.array(dtype, (n, 4), offset=_HEADER.size)
    ikine_engine.ikine(points, out=angles)             # device-to-host copies land behind the header
    _HEADER.pack_into(arena.view(0, _HEADER.size), 0, MAGIC, STATUS_OK | (_CODES[dtype] << 8), n)
    view = arena.view(0, total)
    return view if zero_copy else bytes(view)


def handle_request(ikine_engine, body, correlation_id=None, zero_copy=False):
    """One request -> one reply (bytes in, bytes out); JSON in gives JSON out, IKB1 in gives IKB1 out."""
    binary = bytes(body[:4]) == MAGIC
    try:
        if binary:
            return _binary_reply(ikine_engine, decode_binary_request(body), zero_copy)
        from .kinematics.point import Point
        positions = [Point(p) for p in json.loads(body)["positions"]]  # rpc_broker.py:79-80
        return json.dumps({"status": "OK", "angles": ikine_engine.ikine(positions)}).encode()
    except (OutOfRobotReachException, ValueError, TypeError, ZeroDivisionError, KeyError, struct.error) as exc:
        if binary:
            return encode_binary_reply(error=exc)
        return json.dumps({"status": "ERROR", "reason": str(exc), "correlation_id": correlation_id}).encode()
