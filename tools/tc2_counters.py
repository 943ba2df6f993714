import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inversekinematicsann_b200.kinematics._shared import get_engine
from inversekinematicsann_b200 import _native
from oracle import np_oracle
eng = get_engine()
W, b = np_oracle.synthetic_mlp()
eng.mlp_load(W, b, np_oracle.SHIPPED_MEAN_X, np_oracle.SHIPPED_SCALE_X, np_oracle.SHIPPED_MEAN_Y, np_oracle.SHIPPED_SCALE_Y)
m = 148 * 128 * 20
xyz = torch.rand(m, 3, device="cuda") * torch.tensor([6., 12., 9.], device="cuda") + torch.tensor([0., -6., -3.], device="cuda")
out = torch.empty(m, 4, device="cuda")
lib = _native.load()
buf = (ctypes.c_ulonglong * 16)()
eng.ann_solve_device(xyz, out, mode="fp16x3_ts"); torch.cuda.synchronize()
lib.ikbdbg_tc2_counters(buf, 1)
eng.ann_solve_device(xyz, out, mode="fp16x3_ts"); torch.cuda.synchronize()
lib.ikbdbg_tc2_counters(buf, 1)
names = ["mma_total", "mma_wait_w_full", "mma_wait_act_ready", "mma_wait_d_empty", "epi_wait_d_full", "epi_wait_a_free", "epi_half_total", "epi_compute", "epi_waitfree+store", "epi_drain"]
tiles = 20
for n, v in zip(names, buf):
    print(f"{n:22s} {v:12d} cycles  per layer {v / tiles / 11:10.0f}")
