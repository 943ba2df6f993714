import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inversekinematicsann_b200.kinematics._shared import get_engine
from oracle import c_oracle
eng = get_engine()
rng = np.random.RandomState(3)
xyz = rng.rand(50_000, 3) * [6, 12, 9] + [0, -6, -3]
base, st, it0 = eng.fabrik_solve(xyz, return_iters=True)
again, st, it1 = eng.fabrik_solve(xyz, return_iters=True)
perm = rng.permutation(len(xyz))
sh, st, it2 = eng.fabrik_solve(xyz[perm], return_iters=True)
want = c_oracle.fabrik_ikine(xyz)
print("again differs rows:", np.nonzero((again != base).any(axis=1))[0][:20])
d = np.nonzero((sh != base[perm]).any(axis=1))[0]
print("shuffled differs rows:", len(d), d[:20])
for j in d[:10]:
    i = perm[j]
    print(i, xyz[i], "base", base[i], it0[i], "shuf", sh[j], it2[j], "oracle", want["angles"][i], want["iters"][i])
print("base vs oracle max", np.abs(base - want["angles"]).max(), "shuf vs oracle", np.abs(sh - want["angles"][perm]).max())
