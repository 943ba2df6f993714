"""Label targets with the FABRIK kernel, train the reference's network on them, report the trained model's accuracy.

SURVEY 8f rank 4: the reference's model file (.h5) is missing from its tree, so every ANN number so far used
synthetic weights.  This produces real ones (FABRIK-labelled, as the shipped model evidently was) and measures, through
the product kernels, what BASELINE's metric asks for the ANN: mean FK position error ||FK(ann(p)) - p||.

    python tools/train_fabrik_model.py --samples 2000000 --epochs 60 --out gpurun_out/roboarm_b200

Recipe = kinematics/training.py (the reference's, ann.py:27-68) with a larger batch and step size than the reference's
32 / 1e-5 so that it finishes in GPU-minutes; both are flags.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--samples', type=int, default=2_000_000)
    ap.add_argument('--epochs', type=int, default=60)
    ap.add_argument('--batch-size', type=int, default=2048)
    ap.add_argument('--learning-rate', type=float, default=1e-3)
    ap.add_argument('--final-learning-rate', type=float, default=1e-5)
    ap.add_argument('--max-label-error', type=float, default=1e-2,
                    help='drop labels whose FK error exceeds this (unreachable targets and the wrong-branch rows '
                         'of inverse.py:82-85,102-108)')
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--patience', type=int, default=12, help='EarlyStopping patience (reference: 12)')
    ap.add_argument('--out', default='gpurun_out/roboarm_b200')
    ap.add_argument('--tf32', action='store_true')
    args = ap.parse_args()

    from inversekinematicsann_b200.kinematics.ann import ANN
    from inversekinematicsann_b200.kinematics.forward import ForwardKinematics
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    from inversekinematicsann_b200.robot.robot import SixDOFRobot as R

    limits = R.effector_workspace_limits
    lo = np.array([limits[a][0] for a in 'xyz'], dtype=np.float64)
    hi = np.array([limits[a][1] for a in 'xyz'], dtype=np.float64)
    rng = np.random.default_rng(args.seed)
    pts = (lo + (hi - lo) * rng.random((args.samples, 3))).astype(np.float32)

    fab = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, limits)
    fk = ForwardKinematics(R.dh_matrix)
    t0 = time.perf_counter()
    labels, iters = fab.ikine(pts, as_array=True, return_iterations=True)
    _, label_err = fk.fkine_positions(labels, targets=pts)
    label_s = time.perf_counter() - t0
    keep = (iters < 100) & (label_err <= args.max_label_error) & np.isfinite(labels).all(axis=1)
    x, y = pts[keep].astype(np.float64), labels[keep]
    print(f'labelled {args.samples} targets in {label_s:.2f} s; kept {keep.sum()} '
          f'({100.0 * keep.mean():.1f} %) with FK error <= {args.max_label_error}', flush=True)

    ann = ANN(limits, R.dh_matrix)
    t0 = time.perf_counter()
    ann.train_model(args.epochs, x, y, batch_size=args.batch_size, learning_rate=args.learning_rate,
                    final_learning_rate=args.final_learning_rate, seed=args.seed, allow_tf32=args.tf32,
                    patience=args.patience)
    train_s = time.perf_counter() - t0

    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    from joblib import dump
    ann.model.save_npz(args.out + '.npz')
    dump(ann.x_data_skaler, args.out + '_scaler_x.bin', compress=True)
    dump(ann.y_data_skaler, args.out + '_scaler_y.bin', compress=True)

    # held-out accuracy through the product kernels
    test = (lo + (hi - lo) * np.random.default_rng(args.seed + 1).random((1_000_000, 3))).astype(np.float32)
    want, t_iters = fab.ikine(test, as_array=True, return_iterations=True)
    _, fab_err = fk.fkine_positions(want, targets=test)
    reach = (t_iters < 100) & (fab_err <= args.max_label_error)
    report = {'samples': int(args.samples), 'kept': int(keep.sum()), 'epochs_run': len(ann.history['loss']),
              'batch_size': args.batch_size, 'learning_rate': [args.learning_rate, args.final_learning_rate],
              'label_seconds': label_s, 'train_seconds': train_s, 'loss': ann.history['loss'][-1],
              'best_val_loss': ann.history['best_val_loss'], 'held_out_rows': int(reach.sum()), 'modes': {}}
    preds = {}
    for mode in ('fp32', 'fp16x3', 'fp16x3_ts'):
        ann.mode = mode
        pred = ann.predict(test[reach])
        preds[mode] = pred
        _, err = fk.fkine_positions(pred, targets=test[reach])
        report['modes'][mode] = {
            'fk_error_mean': float(err.mean()), 'fk_error_median': float(np.median(err)),
            'fk_error_p99': float(np.quantile(err, 0.99)),
            'angle_abs_diff_vs_fabrik_mean': float(np.abs(pred - want[reach]).mean()),
            'max_abs_diff_vs_fp32_kernel': float(np.abs(pred - preds['fp32']).max())}
    report['fabrik_fk_error_mean_same_rows'] = float(fab_err[reach].mean())
    with open(args.out + '_report.json', 'w') as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))


if __name__ == '__main__':
    main()
