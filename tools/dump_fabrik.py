"""Solve a seeded workspace sample with whatever library IKB200_LIB selects and save angles + iterations (.npz);
used to compare two builds bit for bit (e.g. a kernel change that must not alter results)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics  # noqa: E402
from inversekinematicsann_b200.robot.robot import SixDOFRobot as R  # noqa: E402

out, rows = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
rng = np.random.default_rng(99)
pts = (rng.random((rows, 3)) * [6, 12, 9] + [0, -6, -3]).astype(np.float32)
pts[:1000] = (rng.random((1000, 3)) * [2, 4, 3] + [1, -2, 1]).astype(np.float32)
results = {}
for prec in ("f64", "f32"):
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits, precision=prec)
    res = np.empty((rows, 4), dtype=np.float32)
    ik.ikine(pts[:1000], as_array=True)
    t0 = time.perf_counter()
    _, iters = ik.ikine(pts, out=res, return_iterations=True)
    print(prec, "seconds", time.perf_counter() - t0, "mean iterations", iters.mean(), flush=True)
    results[f"angles_{prec}"] = res.copy()
    results[f"iters_{prec}"] = iters
np.savez(out, **results)
