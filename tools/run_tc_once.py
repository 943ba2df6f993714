import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inversekinematicsann_b200.kinematics._shared import get_engine
from oracle import np_oracle
eng = get_engine()
W, b = np_oracle.synthetic_mlp()
eng.mlp_load(W, b, np_oracle.SHIPPED_MEAN_X, np_oracle.SHIPPED_SCALE_X, np_oracle.SHIPPED_MEAN_Y, np_oracle.SHIPPED_SCALE_Y)
m = int(os.environ.get("TC_ROWS", 1_000_000))
xyz = torch.rand(m, 3, device="cuda") * torch.tensor([6., 12., 9.], device="cuda") + torch.tensor([0., -6., -3.], device="cuda")
out = torch.empty(m, 4, device="cuda")
for _ in range(3):
    eng.ann_solve_device(xyz, out, mode=os.environ.get("TC_MODE", "fp16x3"))
torch.cuda.synchronize()
print("ok", out[:2].tolist())
