"""One device-resident FABRIK solve of N uniform workspace targets (for ncu launch lists of the solver alone)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inversekinematicsann_b200.kinematics._shared import get_engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = get_engine()
g = torch.Generator(device="cuda").manual_seed(1234)
xyz = torch.rand(n, 3, device="cuda", generator=g) * torch.tensor([6.0, 12.0, 9.0], device="cuda") + \
    torch.tensor([0.0, -6.0, -3.0], device="cuda")
out = torch.empty(n, 4, device="cuda")
for _ in range(reps):
    eng.fabrik_solve_device(xyz, out)
torch.cuda.synchronize()
print("done", eng.stats_fetch_torch())
