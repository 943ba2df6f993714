"""NumPy restatements (TEST INFRASTRUCTURE -- see oracle/__init__.py).

* ``mlp_predict``      -- reference kinematics/ann.py:70-76 (scaler -> Keras Sequential -> scaler)
* ``synthetic_mlp``    -- seeded stand-in weights with the architecture of ann.py:46-56
* generators           -- reference robot/position_generator.py:26-97
* ``fabrik_ikine_np``  -- vectorised fp64 FABRIK (cross-check of ik_oracle.c)

ANN parity is UNPINNED (no real weights, no keras): ``mlp_predict`` restates the published
semantics of its third-party dependencies -- sklearn ``StandardScaler.transform`` =
``(X - mean_) / scale_`` in float64, Keras ``Dense`` = ``x @ kernel + bias`` with kernel shape
(in, out) in float32, ``StandardScaler.inverse_transform`` = in-place ``X *= scale_; X += mean_``
on the float32 prediction array (scale_/mean_ cast to float32 as current sklearn does) (sklearn 1.0.2 per the pickles; keras/tensorflow unpinned,
reference Dockerfile:6).
"""
import numpy as np

# scalers shipped with the reference: models/roboarm_model_1674153800-982793_scaler_{x,y}.bin
# (sklearn StandardScaler, n_samples_seen_=67000; values read with joblib in the survey session,
#  re-checked by tests/test_oracle_vs_reference.py when /root/reference is present)
SHIPPED_MEAN_X = np.array([2.2073088909641334, 0.19405985835497927, 1.494994275926956])
SHIPPED_SCALE_X = np.array([1.7144761363570307, 2.7973201512836416, 2.140079230865925])
SHIPPED_MEAN_Y = np.array([0.052229169532186454, 0.9236331507819656, -1.3859332319838136,
                           -0.42092474514907724])
SHIPPED_SCALE_Y = np.array([0.8768847052996848, 0.6520665510519178, 1.0214536342625566,
                            0.4481255851377674])

LAYER_DIMS = [3] + [500] * 12 + [4]  # ann.py:46-56: Input(3), 12 x Dense(500, tanh), Dense(4)


def synthetic_mlp(seed=1234, dims=LAYER_DIMS, gain=1.0, bias_range=0.1):
    """Seeded Glorot-uniform kernels (Keras' default initialiser, scaled by `gain` so the tanh units
    are neither dead nor saturated) and small uniform biases.  Returns (weights, biases) float32."""
    rng = np.random.default_rng(seed)
    weights, biases = [], []
    for fan_in, fan_out in zip(dims[:-1], dims[1:]):
        lim = gain * np.sqrt(6.0 / (fan_in + fan_out))
        weights.append(rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32))
        biases.append(rng.uniform(-bias_range, bias_range, size=(fan_out,)).astype(np.float32))
    return weights, biases


def mlp_predict(xyz, weights, biases, mean_x=SHIPPED_MEAN_X, scale_x=SHIPPED_SCALE_X,
                mean_y=SHIPPED_MEAN_Y, scale_y=SHIPPED_SCALE_Y, dtype=np.float32, chunk=65536):
    """reference ann.py:70-76.  dtype=float32 mirrors Keras; dtype=float64 bounds fp32 rounding."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    out = np.empty((xyz.shape[0], weights[-1].shape[1]), dtype=dtype)
    Ws = [np.asarray(w, dtype=dtype) for w in weights]
    bs = [np.asarray(b, dtype=dtype) for b in biases]
    for lo in range(0, xyz.shape[0], chunk):
        xs = (xyz[lo:lo + chunk] - mean_x) / scale_x          # StandardScaler.transform, float64
        h = xs.astype(dtype)                                   # Keras casts inputs to float32
        for W, b in zip(Ws[:-1], bs[:-1]):
            h = np.tanh(h @ W + b)                             # Dense(500, tanh)
        y = h @ Ws[-1] + bs[-1]                                # Dense(4), linear
        y *= scale_y.astype(dtype)                             # inverse_transform, in place on the
        y += mean_y.astype(dtype)                              # float32 array (sklearn >= 1.3 casts
        #   scale_/mean_ to X.dtype first; 1.0.2 multiplied by the float64 vectors and rounded the
        #   result -- a 1-ulp(fp32) difference, the reference's sklearn version is unpinned)
        out[lo:lo + chunk] = y
    return out


# ---- trajectory generators, reference robot/position_generator.py ------------------------------

def circle(radius, n, centre):
    """position_generator.py:26-31 (integer 'timestamps' in radians)."""
    t = np.arange(n, dtype=np.float64)
    return np.stack([np.full(n, float(centre[0])), centre[1] + radius * np.sin(t),
                     centre[2] + radius * np.cos(t)], axis=1)


def cube(step, len_x, len_y, len_z, start=(0, 0, 0)):
    """position_generator.py:39-46: x fastest, then y, then z."""
    xs, ys, zs = (np.arange(0, l, step) for l in (len_x, len_y, len_z))
    z, y, x = np.meshgrid(zs, ys, xs, indexing="ij")
    return np.stack([x.ravel() + start[0], y.ravel() + start[1], z.ravel() + start[2]], axis=1)


def cube_random(step, len_x, len_y, len_z, start=(0, 0, 0)):
    """position_generator.py:48-55.  One np.random.rand() per axis in x,y,z order per point, which
    np.random.rand(n, 3) reproduces bit for bit under the same global seed."""
    n = len(np.arange(0, len_x * len_y * len_z, step))
    u = np.random.rand(n, 3)
    return u * np.array([len_x, len_y, len_z], dtype=np.float64) + np.array(start, dtype=np.float64)


def spring(n, len_x, len_y, len_z):
    """position_generator.py:72-78."""
    z = np.linspace(0, len_z, n)
    x = (np.sin(z) * len_x) + len_x
    y = (np.cos(z) * len_y) + len_y
    return np.stack([x / 2, y / 2, z], axis=1)


def random_distribution_normal(n, limits, std_dev=0.5):
    """position_generator.py:80-97, distribution='normal': truncnorm(mean 0) per axis, x then y then z."""
    from scipy.stats import truncnorm
    cols = [truncnorm(lo / std_dev, hi / std_dev, loc=0, scale=std_dev).rvs(n)
            for lo, hi in (limits[k] for k in ("x", "y", "z"))]
    return np.stack(cols, axis=1)


# ---- vectorised FABRIK (cross-check) -----------------------------------------------------------

def _pb(a, b, L):
    d = b - a
    n = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)
    return a + (L / n)[:, None] * d


def _dist(a, b):
    d = a - b
    return np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)


def fabrik_ikine_np(xyz, links=(2.0, 2.0, 2.0, 2.0), tol=1e-3, max_iter=100,
                    init=((0, 0, 2), (0, 0, 4), (0, 0, 6), (0, 0, 8))):
    """fabrik.py:44-67 + inverse.py:54-112 on arrays, starting from the constant vertical pose the
    reference's seed FK produces up to 1e-16.  round(x, 8) is approximated by
    np.round.  Returns (angles[n,4], iters[n])."""
    T = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    n = T.shape[0]
    P = [np.tile(np.asarray(p, dtype=np.float64), (n, 1)) for p in init]
    S = P[0].copy()
    active = np.ones(n, dtype=bool)
    iters = np.zeros(n, dtype=np.int32)
    d = links
    with np.errstate(all="ignore"):
        for _ in range(max_iter):
            if not active.any():
                break
            idx = np.nonzero(active)[0]
            p0, p1, p2 = P[0][idx], P[1][idx], P[2][idx]
            t, s = T[idx], S[idx]
            b2 = _pb(t, p2, d[2]); b1 = _pb(b2, p1, d[1]); b0 = _pb(b1, p0, d[0])
            se = _dist(b0, s)
            f1 = _pb(s, b1, d[1]); f2 = _pb(f1, b2, d[2]); f3 = _pb(f2, t, d[3])
            ge = _dist(f3, t)
            P[0][idx], P[1][idx], P[2][idx], P[3][idx] = s, f1, f2, f3
            iters[idx] += 1
            active[idx] = (se > tol) | (ge > tol)
        B, C, D, E = P
        A = np.zeros_like(B)
        th1 = np.arctan2(E[:, 1], E[:, 0])
        ab, bc, cd, de = _dist(A, B), _dist(B, C), _dist(C, D), _dist(D, E)
        ac, bd, ce = _dist(A, C), _dist(B, D), _dist(C, E)
        a2 = np.arccos(np.round((ab ** 2 + bc ** 2 - ac ** 2) / (2 * ab * bc), 8))
        th2 = np.where(C[:, 0] * D[:, 0] < 0, (3 * np.pi / 2) - a2, -(np.pi / 2 - a2))
        a3 = np.arccos(np.round((bc ** 2 + cd ** 2 - bd ** 2) / (2 * bc * cd), 8))
        th3 = -(np.pi - a3)
        a4 = np.arccos(np.round((cd ** 2 + de ** 2 - ce ** 2) / (2 * cd * de), 8))
        mid = _pb(C, E, ce / 2)
        th4 = np.where(bd > _dist(B, mid), -(np.pi - a4), np.pi - a4)
    return np.stack([th1, th2, th3, th4], axis=1), iters


# ---- the benchmark's input stream, restated ------------------------------------------------------
# csrc/generators.cu draws cube_random / random_dist targets from a counter-based Philox4x32-10 stream keyed by
# (seed, row): 128 bits for counter 2*row (x, y) and 2*row + 1 (z).  The same stream in NumPy, so that the CPU arm of
# bench.py solves the very rows the GPU arm solves (pinned bit-exact against the device generator by
# tests/test_wire_and_generators.py::test_philox_stream_matches_the_numpy_restatement).

def _philox4x32_10(seed, ctr):
    """seed: python int (64 bit); ctr: uint64 array -> four uint32 arrays."""
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    c0, c1 = ctr & mask, ctr >> np.uint64(32)
    c2 = np.full_like(c0, 0x1BD11BDA)
    c3 = np.full_like(c0, 0x5851F42D)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = m0 * c0, m1 * c2                      # 32 x 32 -> 64 bit products
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u53(a, b):
    return (((a >> np.uint64(5)) << np.uint64(26)) | (b >> np.uint64(6))).astype(np.float64) * (1.0 / 9007199254740992.0)


def philox_cube_random(n, lens, start, seed, row_offset=0, dtype=np.float64, chunk=1 << 22):
    """Rows [row_offset, row_offset + n) of TrainingDataGenerator.cube_random_device(..., seed=seed): start + len * U[0,1)
    per axis (position_generator.py:48-55) with the device generator's Philox stream."""
    out = np.empty((int(n), 3), dtype=dtype)
    for lo in range(0, int(n), chunk):
        hi = min(int(n), lo + chunk)
        g = np.arange(row_offset + lo, row_offset + hi, dtype=np.uint64)
        r0, r1, r2, r3 = _philox4x32_10(seed, np.uint64(2) * g)
        s0, s1, _, _ = _philox4x32_10(seed, np.uint64(2) * g + np.uint64(1))
        u = np.stack([_u53(r0, r1), _u53(r2, r3), _u53(s0, s1)], axis=1)
        out[lo:hi] = (u * np.asarray(lens, dtype=np.float64) + np.asarray(start, dtype=np.float64)).astype(dtype)
    return out
