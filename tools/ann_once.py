"""A few device-resident K2 launches (default tensor-core mode, trained model) for ncu captures of the MLP kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inversekinematicsann_b200.kinematics.inverse import AnnInverseKinematics  # noqa: E402
from inversekinematicsann_b200.robot.robot import SixDOFRobot as R  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "fp16x3_ts"
ann = AnnInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
ann.load_model(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "models", "roboarm_b200_r01.h5"))
eng = ann.ann._ensure_uploaded()
g = torch.Generator(device="cuda").manual_seed(5)
xyz = torch.rand(n, 3, device="cuda", generator=g) * torch.tensor([6.0, 12.0, 9.0], device="cuda") + \
    torch.tensor([0.0, -6.0, -3.0], device="cuda")
out = torch.empty(n, 4, device="cuda")
err = torch.empty(n, device="cuda")
for _ in range(3):
    eng.ann_solve_device(xyz, out, mode=mode, fk_err=err)
torch.cuda.synchronize()
print("done", float(err.mean()))
