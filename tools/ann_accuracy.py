"""Accuracy of the three K2 modes on a given weight file, against the fp64 and fp32 NumPy restatements.

    python tools/ann_accuracy.py models/roboarm_b200_r01 [--rows 200000]

(IKB200_LIB selects an alternative build of the library, e.g. an experimental variant.)
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('model', help='path prefix: <prefix>.npz, <prefix>_scaler_x.bin, <prefix>_scaler_y.bin; or '
                    '"synthetic:<seed>:<gain>" for seeded Glorot weights with the shipped scalers')
    ap.add_argument('--rows', type=int, default=200_000)
    ap.add_argument('--modes', default='fp32,fp16x3,fp16x3_ts')
    ap.add_argument('--seed', type=int, default=7)
    ap.add_argument('--comp-sweep', default='', help='comma list of IKB_TC_TRUNC_COMP values to try')
    ap.add_argument('--sweep-modes', default='fp16x3_ts')
    args = ap.parse_args()
    from inversekinematicsann_b200.kinematics.ann import ANN
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
    from oracle import np_oracle
    from threadpoolctl import threadpool_limits

    ann = ANN(R.effector_workspace_limits, R.dh_matrix)
    if args.model.startswith('synthetic:'):
        _, seed, gain = args.model.split(':')
        W, b = np_oracle.synthetic_mlp(seed=int(seed), gain=float(gain))
        ann.set_model(W, b, np_oracle.SHIPPED_MEAN_X, np_oracle.SHIPPED_SCALE_X, np_oracle.SHIPPED_MEAN_Y,
                      np_oracle.SHIPPED_SCALE_Y)
    else:
        ann.load_model(args.model + '.h5')
    xs, ys = ann.x_data_skaler, ann.y_data_skaler
    rng = np.random.default_rng(args.seed)
    pts = (rng.random((args.rows, 3)) * [6, 12, 9] + [0, -6, -3]).astype(np.float32)
    with threadpool_limits(limits=os.cpu_count() or 1):
        o64 = np_oracle.mlp_predict(pts, ann.model.kernels, ann.model.biases, xs.mean_, xs.scale_, ys.mean_,
                                    ys.scale_, dtype=np.float64)
        o32 = np_oracle.mlp_predict(pts, ann.model.kernels, ann.model.biases, xs.mean_, xs.scale_, ys.mean_,
                                    ys.scale_, dtype=np.float32)

    def stats(a, b):
        d = np.abs(a.astype(np.float64) - b).max(axis=1)
        return {'max': float(d.max()), 'p99.9': float(np.quantile(d, 0.999)), 'p99': float(np.quantile(d, 0.99)),
                'mean': float(np.abs(a.astype(np.float64) - b).mean()), 'rows_over_1e-5': int((d > 1e-5).sum()),
                'signed_mean': float((a.astype(np.float64) - b).mean())}

    report = {'rows': args.rows, 'lib': os.environ.get('IKB200_LIB', 'default'),
              'fp32_oracle_vs_fp64': stats(o32, o64)}
    for mode in args.modes.split(','):
        ann.mode = mode
        got = ann.predict(pts)
        report[mode] = {'vs_fp64': stats(got, o64), 'vs_fp32_oracle': stats(got, o32.astype(np.float64))}
    for comp in [c for c in args.comp_sweep.split(',') if c]:
        os.environ['IKB_TC_TRUNC_COMP'] = comp
        for mode in args.sweep_modes.split(','):
            ann._uploaded = False
            ann.mode = mode
            report[f'{mode}_comp_{comp}'] = {'vs_fp64': stats(ann.predict(pts), o64)}
    print(json.dumps(report, indent=1))


if __name__ == '__main__':
    main()
