"""SURVEY 8f 'next' rows: the broker wire path (CPU codecs + GPU round trip) and the on-device generators."""
import json

import numpy as np
import pytest

from inversekinematicsann_b200 import wire


def test_binary_codec_round_trip():
    pts = np.random.default_rng(0).uniform(-1, 1, size=(1000, 3)).astype(np.float32)
    body = wire.encode_binary_request(pts)
    assert len(body) == 16 + pts.nbytes and body[:4] == b"IKB1"
    assert np.array_equal(wire.decode_binary_request(body), pts)
    assert wire.decode_binary_request(wire.encode_binary_request(pts.astype(np.float64))).dtype == np.float64
    with pytest.raises(ValueError):
        wire.decode_binary_request(body[:-4])
    with pytest.raises(ValueError):
        wire.encode_binary_request(np.zeros((3, 4)))
    ang = np.random.default_rng(1).normal(size=(7, 4)).astype(np.float32)
    reply = wire.decode_binary_reply(wire.encode_binary_reply(ang))
    assert reply["status"] == "OK" and np.array_equal(reply["angles"], ang)
    err = wire.decode_binary_reply(wire.encode_binary_reply(error=ValueError("boom")))
    assert err == {"status": "ERROR", "reason": "boom"}
    # the reply names its dtype: a float64 (FABRIK) reply decodes as float64 without being told, and a body whose
    # length disagrees with its header is refused instead of being reinterpreted
    ang64 = np.arange(8, dtype=np.float64).reshape(2, 4)
    back = wire.decode_binary_reply(wire.encode_binary_reply(ang64))["angles"]
    assert back.dtype == np.float64 and np.array_equal(back, ang64)
    assert wire.decode_binary_reply(wire.encode_binary_reply(ang))["angles"].dtype == np.float32
    with pytest.raises(ValueError):
        wire.decode_binary_reply(wire.encode_binary_reply(ang64)[:-8])
    for short in (b"IKB1", b"IKB1\x00\x00", b""):
        with pytest.raises(ValueError):
            wire.decode_binary_reply(short)
        with pytest.raises(ValueError):
            wire.decode_binary_request(short)


def test_malformed_binary_request_gets_an_error_reply():
    """A truncated IKB1 message must not kill the consumer (upstream dies on anything it does not catch)."""
    for body in (b"IKB1", b"IKB1\x00\x00\x00\x00\x05", wire.encode_binary_request(np.zeros((4, 3)))[:-1],
                 b"IKB1" + (9).to_bytes(4, "little") + (0).to_bytes(8, "little")):
        reply = wire.decode_binary_reply(wire.handle_request(_FakeEngine(), body))
        assert reply["status"] == "ERROR" and reply["reason"]


class _FakeEngine:
    def ikine(self, points, as_array=False, out=None):
        if len(points) == 0:
            raise ZeroDivisionError("float division by zero")
        return np.zeros((len(points), 4), dtype=np.float32) if as_array else [[0.0] * 4 for _ in points]


def test_binary_request_without_native_engine_uses_plain_reply():
    pts = np.ones((5, 3), dtype=np.float32)
    reply = wire.decode_binary_reply(wire.handle_request(_FakeEngine(), wire.encode_binary_request(pts)))
    assert reply["status"] == "OK" and reply["angles"].shape == (5, 4)


def test_handle_request_json_schema_and_errors():
    """reference rpc_broker.py:76-100: JSON schema, ERROR mapping with correlation id."""
    ok = json.loads(wire.handle_request(_FakeEngine(), json.dumps({"positions": [[1, 2, 3], [3, 2, 1]]}).encode()))
    assert ok == {"status": "OK", "angles": [[0.0] * 4, [0.0] * 4]}
    bad = json.loads(wire.handle_request(_FakeEngine(), json.dumps({"positions": [[1, 2, 3, 4]]}).encode(), "cid-7"))
    assert bad["status"] == "ERROR" and bad["correlation_id"] == "cid-7" and "shape should be (3,)" in bad["reason"]
    zero = json.loads(wire.handle_request(_FakeEngine(), json.dumps({"positions": []}).encode(), "c"))
    assert zero["status"] == "ERROR" and "division by zero" in zero["reason"]  # upstream would crash here


@pytest.mark.gpu
def test_broker_round_trip_on_gpu():
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    from oracle import c_oracle
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    pts = np.random.default_rng(2).uniform([1, -2, 1], [3, 2, 4], size=(5000, 3))
    want = c_oracle.fabrik_ikine(pts)["angles"]
    reply = wire.decode_binary_reply(wire.handle_request(ik, wire.encode_binary_request(pts)))
    assert reply["angles"].dtype == np.float64          # the reply carries the request's dtype
    np.testing.assert_allclose(reply["angles"], want, rtol=0, atol=1e-9)
    view = wire.handle_request(ik, wire.encode_binary_request(pts.astype(np.float32)), zero_copy=True)
    assert isinstance(view, memoryview) and len(view) == 16 + 5000 * 16
    r32 = wire.decode_binary_reply(view)
    assert r32["angles"].dtype == np.float32
    want32 = c_oracle.fabrik_ikine(pts.astype(np.float32).astype(np.float64))["angles"]
    np.testing.assert_allclose(r32["angles"], want32, rtol=0, atol=2e-6)
    js = json.loads(wire.handle_request(ik, json.dumps({"positions": pts[:20].tolist()}).encode()))
    np.testing.assert_allclose(js["angles"], want[:20], rtol=0, atol=1e-9)
    out = json.loads(wire.handle_request(ik, json.dumps({"positions": [[1, 2, 3], [1, 2, 7]]}).encode(), "id1"))
    assert out["status"] == "ERROR" and "Point(1, 2, 7) is out of manipulator reach area" in out["reason"]
    err = wire.decode_binary_reply(wire.handle_request(ik, wire.encode_binary_request(np.array([[1.0, 2.0, 7.0]]))))
    assert err["status"] == "ERROR" and "out of manipulator reach area" in err["reason"]


@pytest.mark.gpu
def test_deterministic_generators_match_reference_shapes(golden_generators):
    from inversekinematicsann_b200.robot.position_generator import TrainingDataGenerator as G
    g = golden_generators
    np.testing.assert_allclose(G.circle(2, 50, (2, 0, 2)), g["circle"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(G.spring(50, 2, 3, 6), g["spring"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(G.cube(0.5, 2, 3, 1.5, (1, -1, 0)), g["cube"], rtol=0, atol=1e-14)
    dev = G.spring_device(1_000_003, 2, 3, 6)
    assert dev.is_cuda and dev.shape == (1_000_003, 3) and abs(dev[-1, 2].item() - 6.0) < 1e-6
    shard = G._generate if hasattr(G, "_generate") else None  # shards produce disjoint row ranges
    from inversekinematicsann_b200.robot import position_generator as pg
    part = pg._generate(pg.GEN_CIRCLE, [2, 2, 0, 2], 10, dtype="float64", row_offset=40).cpu().numpy()
    np.testing.assert_allclose(part, g["circle"][40:50], rtol=0, atol=1e-14)


@pytest.mark.gpu
def test_random_generators_distributions():
    from scipy import stats
    from inversekinematicsann_b200.robot.position_generator import TrainingDataGenerator as G
    lim = {"x": [0, 6], "y": [-6, 6], "z": [-3, 6]}
    n = 400_000
    u = G.cube_random_device(648.0 / n, 6, 12, 9, start=(0, -6, -3), seed=7, dtype="float64").cpu().numpy()
    assert u.shape == (n, 3) and (u >= [0, -6, -3]).all() and (u < [6, 6, 6]).all()
    for ax, (lo, ln) in enumerate([(0, 6), (-6, 12), (-3, 9)]):
        assert stats.kstest(u[:50_000, ax], stats.uniform(lo, ln).cdf).pvalue > 1e-3
    assert abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 0.01
    again = G.cube_random_device(648.0 / n, 6, 12, 9, start=(0, -6, -3), seed=7, dtype="float64").cpu().numpy()
    other = G.cube_random_device(648.0 / n, 6, 12, 9, start=(0, -6, -3), seed=8, dtype="float64").cpu().numpy()
    assert np.array_equal(u, again) and not np.array_equal(u, other)
    nrm = G.random_distribution_device(n, lim, "normal", 0.5, seed=3, dtype="float64").cpu().numpy()
    for ax, key in enumerate("xyz"):
        lo, hi = lim[key]
        assert nrm[:, ax].min() >= lo and nrm[:, ax].max() <= hi
        ref = stats.truncnorm(lo / 0.5, hi / 0.5, loc=0, scale=0.5)
        assert stats.kstest(nrm[:50_000, ax], ref.cdf).pvalue > 1e-3
    assert len(G.random_distribution(100, lim, "normal", 0.35)) == 100


@pytest.mark.gpu
def test_philox_stream_matches_the_numpy_restatement():
    """The device generator's cube_random stream (Philox4x32-10 keyed by seed and row) against its NumPy restatement,
    bit for bit, float64 and float32, with a row offset: bench.py's CPU arm solves rows of this stream."""
    from inversekinematicsann_b200.robot import position_generator as pg
    from oracle import np_oracle
    lens, start = (6.0, 12.0, 9.0), (0.0, -6.0, -3.0)
    for dtype in ("float64", "float32"):
        dev = pg._generate(pg.GEN_CUBE_RANDOM, list(lens) + list(start), 100_003, dtype=dtype, seed=1234,
                           row_offset=7).cpu().numpy()
        host = np_oracle.philox_cube_random(100_003, lens, start, 1234, row_offset=7, dtype=np.dtype(dtype))
        assert np.array_equal(dev, host), dtype
    other = np_oracle.philox_cube_random(1000, lens, start, 1235)
    assert not np.array_equal(other, np_oracle.philox_cube_random(1000, lens, start, 1234))
