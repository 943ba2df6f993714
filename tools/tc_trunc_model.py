"""CPU model of the split-fp16 tensor-core accumulation of csrc/mlp_tc2.cu, to compare accumulation ORDERS.

tcgen05.mma (kind::f16, fp32 accumulate) truncates its accumulator toward zero after every K = 16 step (measured on
B200).  This script replays one network through that model in NumPy -- exact products, fp64 sum
of the 16 products of a step, accumulator truncated to 24 significant bits after each step -- for

  interleaved : per k chunk  x_hi w_hi, x_lo w_hi, x_hi w_lo            (round 1 kernel: 96 truncating steps at full size)
  corr_first  : all k chunks x_lo w_hi + x_hi w_lo, then all x_hi w_hi  (32 steps at full size, 64 at 2^-11 of it)

and prints the error of each against the fp64 restatement next to the plain fp32 NumPy one.  Test infrastructure only.

    python tools/tc_trunc_model.py models/roboarm_b200_r01 --rows 20000
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

X_SCALE = 64.0


def trunc24(x):
    """fp64 array -> nearest-toward-zero value with a 24-bit significand (as float64)."""
    m, e = np.frexp(x)
    return np.ldexp(np.trunc(m * 16777216.0) / 16777216.0, e)


def split16(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def comp(steps, scale=1.0):
    return 1.0 + scale * 0.36067376022224085 * 5.9604644775390625e-8 * steps


def layer(act, W, b, order, comp_steps):
    """act: (n, fin) float32 activations (unscaled); returns tanh(act @ W + b) float32 through the model."""
    n, fin = act.shape
    fout = W.shape[1]
    kp = 512
    Wp = np.zeros((kp, fout), np.float32)
    Wp[:fin] = W
    Wp[fin] = b                                   # bias row, multiplied by the constant-1 feature
    A = np.zeros((n, kp), np.float32)
    A[:, :fin] = act
    A[:, fin] = 1.0
    wmax = max(20.0, np.abs(W).max(), np.abs(b).max())
    e = int(np.clip(12 - int(np.ceil(np.log2(wmax))), -8, 24))
    sw = np.float32(2.0 ** e)
    whi, wlo = split16(Wp * sw)
    xhi, xlo = split16(A * np.float32(X_SCALE))
    acc = np.zeros((n, fout), np.float64)

    def step(x, w, k0):
        nonlocal acc
        acc = trunc24(acc + x[:, k0:k0 + 16] @ w[k0:k0 + 16])

    if order == "interleaved":
        for kc in range(0, kp, 64):
            for x, w in ((xhi, whi), (xlo, whi), (xhi, wlo)):
                for k0 in range(kc, kc + 64, 16):
                    step(x, w, k0)
    elif order == "corr_first":
        for kc in range(0, kp, 64):
            for x, w in ((xlo, whi), (xhi, wlo)):
                for k0 in range(kc, kc + 64, 16):
                    step(x, w, k0)
        for k0 in range(0, kp, 16):
            step(xhi, whi, k0)
    elif order == "rn":   # same operands, round-to-nearest accumulation: isolates the truncation
        for kc in range(0, kp, 64):
            for x, w in ((xhi, whi), (xlo, whi), (xhi, wlo)):
                for k0 in range(kc, kc + 64, 16):
                    acc = (acc + x[:, k0:k0 + 16] @ w[k0:k0 + 16]).astype(np.float32).astype(np.float64)
    else:
        raise ValueError(order)
    oscale = np.float32(comp(comp_steps) / (float(sw) * X_SCALE)) if order != "rn" else np.float32(1.0 / (float(sw) * X_SCALE))
    pre = acc.astype(np.float32) * oscale
    return np.tanh(pre)


def run(pts, ann, order, comp_steps):
    xs, ys = ann.x_data_skaler, ann.y_data_skaler
    W, b = ann.model.kernels, ann.model.biases
    h = ((pts.astype(np.float64) - xs.mean_) / xs.scale_).astype(np.float32)
    h = np.tanh(h @ W[0] + b[0])                                    # layer 1 on the CUDA cores (fp32)
    for l in range(1, len(W) - 1):
        h = layer(h, W[l], b[l], order, comp_steps)
    y = h @ W[-1] + b[-1]
    return y * ys.scale_.astype(np.float32) + ys.mean_.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--rows", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=7)
    args = ap.parse_args()
    from inversekinematicsann_b200.kinematics.ann import ANN
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    from oracle import np_oracle
    ann = ANN(R.effector_workspace_limits, R.dh_matrix)
    ann.load_model(args.model + ".h5")
    xs, ys = ann.x_data_skaler, ann.y_data_skaler
    rng = np.random.default_rng(args.seed)
    pts = (rng.random((args.rows, 3)) * [6, 12, 9] + [0, -6, -3]).astype(np.float32)
    o64 = np_oracle.mlp_predict(pts, ann.model.kernels, ann.model.biases, xs.mean_, xs.scale_, ys.mean_, ys.scale_,
                                dtype=np.float64)
    o32 = np_oracle.mlp_predict(pts, ann.model.kernels, ann.model.biases, xs.mean_, xs.scale_, ys.mean_, ys.scale_,
                                dtype=np.float32)

    def stats(a):
        d = np.abs(a.astype(np.float64) - o64).max(axis=1)
        return {"mean": float(np.abs(a.astype(np.float64) - o64).mean()), "p99": float(np.quantile(d, 0.99)),
                "max": float(d.max()), "rows_over_1e-5": int((d > 1e-5).sum()), "frac_over_1e-5": float((d > 1e-5).mean())}

    rep = {"rows": args.rows, "numpy_fp32": stats(o32)}
    for order, steps in (("rn", 0), ("interleaved", 96), ("corr_first", 32), ("corr_first", 36)):
        rep[f"{order}_comp{steps}"] = stats(run(pts, ann, order, steps))
        print(order, steps, json.dumps(rep[f"{order}_comp{steps}"]), flush=True)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
