"""A few device-resident K3 launches (FK position error of N rows) for ncu captures of fk_kernel alone."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inversekinematicsann_b200.kinematics._shared import get_engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
eng = get_engine()
g = torch.Generator(device="cuda").manual_seed(5)
xyz = torch.rand(n, 3, device="cuda", generator=g) * 4.0
ang = (torch.rand(n, 4, device="cuda", generator=g) - 0.5) * 6.0
err = torch.empty(n, device="cuda")
for _ in range(3):
    eng.fk_device(ang, targets=xyz, err=err)
torch.cuda.synchronize()
print("done", float(err[:1000].mean()))
