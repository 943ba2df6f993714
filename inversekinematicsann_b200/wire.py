"""Broker wire path: the request/reply handling of reference rpc_broker.py:76-100 without
the AMQP transport, plus a binary payload for large trajectories.

``handle_request(engine, body)`` does what ``IkineRPCBroker.callback`` does between receiving ``body`` and
publishing the reply: decode -> ``engine.ikine(positions)`` -> ``{'status': 'OK', 'angles': ...}`` or
``{'status': 'ERROR', 'reason': str(e), 'correlation_id': ...}`` (rpc_broker.py:70-74,84-91).  Unlike
upstream it also maps ``ZeroDivisionError`` to an ERROR reply instead of killing the consumer.

JSON (``{"positions": [[x, y, z], ...]}``) stays supported for compatibility.  The binary form avoids the
~50 bytes and ~1 us of python object per point that make one 10 M-point JSON message impractical:

    request : b'IKB1' | uint32 dtype (0 = float32, 1 = float64) | uint64 n | n*3 little-endian values
    reply   : b'IKB1' | uint32 status (0 OK, 1 ERROR) | uint64 n | n*4 little-endian float32/float64 angles
              (ERROR: n = byte length of the UTF-8 reason that follows)
"""
import json
import struct

import numpy as np

from .robot.robot import OutOfRobotReachException

MAGIC = b"IKB1"
_HEADER = struct.Struct("<4sIQ")
_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("<f8")}


def encode_binary_request(points):
    arr = np.ascontiguousarray(points)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    if arr.ndim != 2 or arr.shape[1] != 3:
        raise ValueError(f"points must have shape (n, 3), not {arr.shape}")
    code = 0 if arr.dtype == np.float32 else 1
    return _HEADER.pack(MAGIC, code, arr.shape[0]) + arr.astype(_DTYPES[code], copy=False).tobytes()


def decode_binary_request(body):
    magic, code, n = _HEADER.unpack_from(body, 0)
    if magic != MAGIC or code not in _DTYPES:
        raise ValueError("not an IKB1 binary request")
    need = _HEADER.size + n * 3 * _DTYPES[code].itemsize
    if len(body) != need:
        raise ValueError(f"IKB1 request announces {n} points ({need} bytes) but carries {len(body)} bytes")
    return np.frombuffer(body, dtype=_DTYPES[code], count=n * 3, offset=_HEADER.size).reshape(n, 3)


def encode_binary_reply(angles=None, error=None):
    if error is not None:
        reason = str(error).encode("utf-8")
        return _HEADER.pack(MAGIC, 1, len(reason)) + reason
    arr = np.ascontiguousarray(angles)
    return _HEADER.pack(MAGIC, 0, arr.shape[0]) + arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()


def decode_binary_reply(body, dtype=np.float32):
    magic, status, n = _HEADER.unpack_from(body, 0)
    if magic != MAGIC:
        raise ValueError("not an IKB1 binary reply")
    if status != 0:
        return {"status": "ERROR", "reason": body[_HEADER.size:_HEADER.size + n].decode("utf-8")}
    return {"status": "OK",
            "angles": np.frombuffer(body, dtype=np.dtype(dtype).newbyteorder("<"), count=n * 4, offset=_HEADER.size).reshape(n, 4)}


def handle_request(ikine_engine, body, correlation_id=None):
    """One request -> one reply (bytes in, bytes out); JSON in gives JSON out, IKB1 in gives IKB1 out."""
    binary = bytes(body[:4]) == MAGIC
    try:
        if binary:
            angles = ikine_engine.ikine(decode_binary_request(body), as_array=True)
            return encode_binary_reply(angles)
        from .kinematics.point import Point
        positions = [Point(p) for p in json.loads(body)["positions"]]  # rpc_broker.py:79-80
        return json.dumps({"status": "OK", "angles": ikine_engine.ikine(positions)}).encode()
    except (OutOfRobotReachException, ValueError, TypeError, ZeroDivisionError, KeyError) as exc:
        if binary:
            return encode_binary_reply(error=exc)
        return json.dumps({"status": "ERROR", "reason": str(exc), "correlation_id": correlation_id}).encode()
