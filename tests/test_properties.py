"""Property tests (hypothesis) of the host logic either side of the hot path: the row split of the multi-GPU
solve, the chunk grid of the overlapped gather, and the IKB1 wire codec.  No GPU, no native library."""
import json

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from inversekinematicsann_b200 import wire
from inversekinematicsann_b200.sharding import chunk_ranges, shard_range


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10**12), world=st.integers(1, 64))
def test_shard_ranges_tile_the_trajectory(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))          # contiguous, in rank order
    sizes = [hi - lo for lo, hi in spans]
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
    assert sizes == sorted(sizes, reverse=True)                            # the longer shards come first


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10**6), chunk=st.integers(1, 10**5))
def test_chunk_grid_covers_a_shard_once(n, chunk):
    pieces = chunk_ranges(n, chunk)
    assert len(pieces) == -(-n // chunk)
    assert all(0 < hi - lo <= chunk for lo, hi in pieces)
    assert [lo for lo, _ in pieces] == list(range(0, n, chunk))
    assert not pieces or pieces[-1][1] == n


_finite = st.floats(-1e6, 1e6, allow_nan=False, width=32)


@settings(max_examples=100, deadline=None)
@given(rows=st.lists(st.tuples(_finite, _finite, _finite), max_size=40), f32=st.booleans())
def test_binary_request_round_trip(rows, f32):
    pts = np.array(rows, dtype=np.float32 if f32 else np.float64).reshape(-1, 3)
    body = wire.encode_binary_request(pts)
    assert len(body) == 16 + pts.nbytes
    back = wire.decode_binary_request(body)
    assert back.dtype == pts.dtype and np.array_equal(back, pts)
    for cut in {0, 3, 15, len(body) - 1} - {len(body)}:                    # truncated anywhere: refused, never misread
        with pytest.raises(ValueError):
            wire.decode_binary_request(body[:cut])
    with pytest.raises(ValueError):
        wire.decode_binary_request(body + b"\0")


@settings(max_examples=100, deadline=None)
@given(rows=st.lists(st.tuples(_finite, _finite, _finite, _finite), max_size=40), f32=st.booleans())
def test_binary_reply_round_trip_names_its_dtype(rows, f32):
    ang = np.array(rows, dtype=np.float32 if f32 else np.float64).reshape(-1, 4)
    reply = wire.decode_binary_reply(wire.encode_binary_reply(ang))
    assert reply["status"] == "OK" and reply["angles"].dtype == ang.dtype and np.array_equal(reply["angles"], ang)
    body = wire.encode_binary_reply(ang)
    if len(ang):
        with pytest.raises(ValueError):
            wire.decode_binary_reply(body[:-1])


@settings(max_examples=50, deadline=None)
@given(text=st.text(max_size=60))
def test_error_reply_round_trip(text):
    reply = wire.decode_binary_reply(wire.encode_binary_reply(error=ValueError(text)))
    assert reply == {"status": "ERROR", "reason": text}


class _Echo:
    """Stands in for an IK front class: angles = [x, y, z, x + y + z], no native engine behind it."""

    def ikine(self, points, as_array=False, out=None):
        p = np.array([[q.x, q.y, q.z] if hasattr(q, "x") else list(q) for q in points], dtype=np.float64).reshape(-1, 3)
        a = np.concatenate([p, p.sum(axis=1, keepdims=True)], axis=1)
        return a if as_array else a.tolist()


@settings(max_examples=100, deadline=None)
@given(junk=st.binary(max_size=64))
def test_any_message_gets_a_reply(junk):
    """rpc_broker.py:76-100 answers every message; a malformed one must produce an ERROR reply, not an exception."""
    for body in (junk, wire.MAGIC + junk):
        reply = wire.handle_request(_Echo(), body, correlation_id="c1")
        if bytes(body[:4]) == wire.MAGIC:
            decoded = wire.decode_binary_reply(reply)
        else:
            decoded = json.loads(reply)
        assert decoded["status"] in ("OK", "ERROR")
        if decoded["status"] == "ERROR" and bytes(body[:4]) != wire.MAGIC:
            assert decoded["correlation_id"] == "c1"


def test_json_and_binary_requests_agree():
    pts = [[1.0, 2.0, 3.0], [0.5, -1.0, 2.0]]
    js = json.loads(wire.handle_request(_Echo(), json.dumps({"positions": pts}).encode()))
    bi = wire.decode_binary_reply(wire.handle_request(_Echo(), wire.encode_binary_request(np.array(pts))))
    assert js["status"] == bi["status"] == "OK"
    assert np.array_equal(np.array(js["angles"]), bi["angles"])
