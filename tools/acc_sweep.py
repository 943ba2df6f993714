import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inversekinematicsann_b200.kinematics._shared import get_engine
from oracle import np_oracle
eng = get_engine()
rng = np.random.default_rng(0)
xyz = rng.uniform([0, -6, -3], [6, 6, 6], size=(100_000, 3))
for seed, gain in [(1, 1.0), (2, 1.0), (3, 1.0), (4, 0.8), (5, 1.2)]:
    W, b = np_oracle.synthetic_mlp(seed=seed, gain=gain)
    eng.mlp_load(W, b, np_oracle.SHIPPED_MEAN_X, np_oracle.SHIPPED_SCALE_X, np_oracle.SHIPPED_MEAN_Y, np_oracle.SHIPPED_SCALE_Y)
    w64 = np_oracle.mlp_predict(xyz, W, b, dtype=np.float64)
    w32 = np_oracle.mlp_predict(xyz, W, b)
    row = [f"seed {seed} gain {gain}: fp32-oracle vs fp64 {np.abs(w32-w64).max():.2e}"]
    for mode in ("fp32", "fp16x3", "fp16x3_ts"):
        got, _ = eng.ann_solve(xyz, mode=mode)
        row.append(f"{mode} vs fp32-oracle {np.abs(got-w32).max():.2e} vs fp64 {np.abs(got-w64).max():.2e}")
    print(" | ".join(row), flush=True)
