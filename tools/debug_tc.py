import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inversekinematicsann_b200.kinematics._shared import get_engine
from oracle import np_oracle
eng = get_engine()
cases = [([3, 128, 128, 4], 64), ([3, 128, 128, 4], 1000), ([3, 100, 50, 70, 4], 333), ([3, 256, 256, 256, 4], 500),
         ([3, 500, 500, 500, 4], 4096), (np_oracle.LAYER_DIMS, 20000)]
only = os.environ.get("TC_CASE")
for ci, (dims, n) in enumerate(cases):
    if only is not None and int(only) != ci:
        continue
    W, b = np_oracle.synthetic_mlp(seed=7, dims=dims)
    eng.mlp_load(W, b, np_oracle.SHIPPED_MEAN_X, np_oracle.SHIPPED_SCALE_X, np_oracle.SHIPPED_MEAN_Y, np_oracle.SHIPPED_SCALE_Y)
    rng = np.random.default_rng(n)
    xyz = rng.uniform([0, -6, -3], [6, 6, 6], size=(n, 3))
    want = np_oracle.mlp_predict(xyz, W, b)
    want64 = np_oracle.mlp_predict(xyz, W, b, dtype=np.float64)
    t0 = time.time()
    got, st = eng.ann_solve(xyz, mode=os.environ.get("TC_MODE", "fp16x3"))
    ref32, _ = eng.ann_solve(xyz, mode="fp32")
    print(f"case {ci} dims {dims[:3]}..x{len(dims)-2} n {n}: tc vs fp32-oracle {np.abs(got-want).max():.3e} vs fp64 {np.abs(got-want64).max():.3e} "
          f"| simt vs fp64 {np.abs(ref32-want64).max():.3e} | nan {np.isnan(got).sum()} ({time.time()-t0:.2f}s)", flush=True)
