"""Smoke-sized launches of every hot kernel for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py [k1] [k2] [k3]
    compute-sanitizer --tool racecheck python tools/sanitize_run.py k1

K1 runs both pass loops (lockstep batches for out-of-reach targets, lane refill for the rest) with float32 and float64
buffers and the fused FK error; K2 the default tensor-core mode and the two cross-check modes; K3 the pair kernel and
the generic one.  Results are checked against the oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inversekinematicsann_b200.kinematics.inverse import AnnInverseKinematics, FabrikInverseKinematics  # noqa: E402
from inversekinematicsann_b200.robot.robot import SixDOFRobot as R  # noqa: E402
from oracle import c_oracle, np_oracle  # noqa: E402

which = sys.argv[1:] or ["k1", "k2", "k3"]
rng = np.random.RandomState(5)
if "k1" in which or "k3" in which:
    n = int(os.environ.get("IKB_SANITIZE_ROWS", 20_000))
    pts = rng.rand(n, 3) * [6, 12, 9] + [0, -6, -3]
    ik = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    want = c_oracle.fabrik_ikine(pts)
if "k1" in which:
    a64, it = ik.ikine(pts, as_array=True, return_iterations=True)
    a32 = ik.ikine(pts.astype(np.float32), out=np.empty((n, 4), np.float32))
    small, err = ik.ikine(pts[:700], as_array=True, return_fk_error=True)          # fused FK epilogue variant
    assert np.array_equal(it, want["iters"]) and np.nanmax(np.abs(a64 - want["angles"])) <= 1e-9
    w32 = c_oracle.fabrik_ikine(pts.astype(np.float32).astype(np.float64))["angles"]
    assert np.nanmax(np.abs(a32 - w32)) <= 1e-6
    print("k1 ok", n, "rows, far rows", int((want["iters"] == 100).sum()), flush=True)
if "k3" in which:
    ang32 = want["angles"].astype(np.float32)
    pos, err = ik.fkine.fkine_positions(ang32, pts.astype(np.float32))           # generic kernel (positions wanted)
    eng = ik._engine()
    _, err_only, _ = eng.fk(ang32, pts.astype(np.float32), want_pos=False)       # pair kernel
    assert np.nanmax(np.abs(err - err_only)) <= 2e-6
    _, _, want_err = c_oracle.fk_positions(ang32.astype(np.float64), targets=pts.astype(np.float32).astype(np.float64))
    assert np.nanmax(np.abs(err_only - want_err)) <= 1e-5
    print("k3 ok", flush=True)
if "k2" in which:
    m = int(os.environ.get("IKB_SANITIZE_ANN_ROWS", 700))
    W, b = np_oracle.synthetic_mlp(seed=11)
    ann = AnnInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    ann.ann.set_model(W, b, np_oracle.SHIPPED_MEAN_X, np_oracle.SHIPPED_SCALE_X, np_oracle.SHIPPED_MEAN_Y,
                      np_oracle.SHIPPED_SCALE_Y)
    xyz = rng.rand(m, 3) * [6, 12, 9] + [0, -6, -3]
    ref = np_oracle.mlp_predict(xyz, W, b)
    for mode in os.environ.get("IKB_SANITIZE_ANN_MODES", "fp16x3_ts,fp16x3,fp32").split(","):
        ann.ann.mode = mode
        got, fk = ann.ikine(xyz, as_array=True, return_fk_error=True)
        assert np.abs(got - ref).max() <= 1e-5, mode
        print("k2 ok", mode, flush=True)
print("sanitize_run done", flush=True)
