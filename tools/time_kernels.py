"""Quick device-resident timings of the three kernels (CUDA events, mean of `--reps` after warm-up), one JSON line.
For A/B runs of library variants (IKB200_LIB=gpurun_variants/<name>.so) and for ncu captures (`--only`).

    python tools/time_kernels.py [--only k1w,k1r,k2,k3] [--rows N] [--reps R]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inversekinematicsann_b200.kinematics.inverse import AnnInverseKinematics, FabrikInverseKinematics  # noqa: E402
from inversekinematicsann_b200.robot.position_generator import TrainingDataGenerator as Gen  # noqa: E402
from inversekinematicsann_b200.robot.robot import SixDOFRobot as R  # noqa: E402


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="k1w,k1r,k2,k3")
    ap.add_argument("--rows", type=int, default=50_000_000)
    ap.add_argument("--ann-rows", type=int, default=2_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warm", type=int, default=2)
    args = ap.parse_args()
    only = args.only.split(",")
    out = {"lib": os.environ.get("IKB200_LIB", "default")}
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    eng = ik._engine()
    n = args.rows

    def cube(box_len, start, seed):
        return Gen.cube_random_device(1.0, *box_len, start=start, seed=seed, no_of_samples=n)

    for key, box in (("k1w", ((6, 12, 9), (0, -6, -3))), ("k1r", ((2, 4, 3), (1, -2, 1)))):
        if key not in only:
            continue
        xyz = cube(*box, seed=1234)
        ang = torch.empty(n, 4, device="cuda")
        eng.stats_reset_torch()
        eng.fabrik_solve_device(xyz, ang)
        st = eng.stats_fetch_torch()
        ms = timed(lambda: eng.fabrik_solve_device(xyz, ang), args.reps, args.warm)
        out[key] = {"rows": n, "ms": ms, "solves_per_s": n / ms * 1e3, "mean_iterations": st.sum_iterations / n,
                    "algorithmic_tflops": (114 * st.sum_iterations + 126 * n) / ms / 1e9}
        if key == "k1w" and "k3" in only:
            err = torch.empty(n, device="cuda")
            ms = timed(lambda: eng.fk_device(ang, targets=xyz, err=err), args.reps, args.warm)
            out["k3"] = {"rows": n, "ms": ms, "rows_per_s": n / ms * 1e3, "GBps": n * 32 / ms / 1e6}
            del err
        del xyz, ang
    if "k2" in only:
        ann = AnnInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
        ann.load_model(os.path.join(ROOT, "models", "roboarm_b200_r01.h5"))
        aeng = ann.ann._ensure_uploaded()
        m = args.ann_rows
        xyz = Gen.random_distribution_device(m, R.effector_workspace_limits, "normal", 0.5, seed=4321)
        ang = torch.empty(m, 4, device="cuda")
        ms = timed(lambda: aeng.ann_solve_device(xyz, ang, mode="fp16x3_ts"), args.reps, args.warm)
        out["k2"] = {"rows": m, "ms": ms, "solves_per_s": m / ms * 1e3,
                     "algorithmic_tflops": 2.0 * aeng.mlp_macs_per_row * m / ms / 1e9}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
