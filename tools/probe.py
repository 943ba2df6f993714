"""Quick device-resident timing probe (development aid; bench.py is the judged harness)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from inversekinematicsann_b200.kinematics._shared import get_engine
from inversekinematicsann_b200.engine import fabrik_algorithmic_flops


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


def main():
    eng = get_engine()
    out = {}
    for dt in ("f32", "f64"):
        out[f"fma_peak_{dt}_tflops"] = eng.microbench_fma(dt)
    n = int(os.environ.get("PROBE_N", 20_000_000))
    g = torch.Generator(device="cuda").manual_seed(1234)
    u = torch.rand(n, 3, device="cuda", generator=g)
    boxes = {"W": ([6., 12., 9.], [0., -6., -3.]), "R": ([2., 4., 3.], [1., -2., 1.])}
    for name, (ln, st) in boxes.items():
        xyz = u * torch.tensor(ln, device="cuda") + torch.tensor(st, device="cuda")
        ang = torch.empty(n, 4, device="cuda")
        for prec in ("f64", "f32"):
            eng.stats_reset_torch()
            eng.fabrik_solve_device(xyz, ang, precision=prec)
            st_ = eng.stats_fetch_torch()
            best, med = timed(lambda: eng.fabrik_solve_device(xyz, ang, precision=prec))
            flops = fabrik_algorithmic_flops(st_.sum_iterations, n)
            out[f"fabrik_{name}_{prec}"] = {"solves_per_s": n / best, "ms": best * 1e3, "mean_iters": st_.sum_iterations / n,
                                           "alg_tflops": flops / best / 1e12}
        err = torch.empty(n, device="cuda")
        best, med = timed(lambda: eng.fk_device(ang, targets=xyz, err=err))
        out[f"fk_{name}"] = {"rows_per_s": n / best, "GBps": n * 32 / best / 1e9}
    # MLP
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import np_oracle
    W, b = np_oracle.synthetic_mlp()
    eng.mlp_load(W, b, np_oracle.SHIPPED_MEAN_X, np_oracle.SHIPPED_SCALE_X, np_oracle.SHIPPED_MEAN_Y, np_oracle.SHIPPED_SCALE_Y)
    m = int(os.environ.get("PROBE_M", 1_000_000))
    xyz = u[:m] * torch.tensor(boxes["W"][0], device="cuda") + torch.tensor(boxes["W"][1], device="cuda")
    ang = torch.empty(m, 4, device="cuda")
    for mode in [x for x in os.environ.get("PROBE_MODES", "fp32").split(",") if x != "none"]:
        best, med = timed(lambda: eng.ann_solve_device(xyz, ang, mode=mode), reps=3, warm=1)
        out[f"mlp_{mode}"] = {"solves_per_s": m / best, "ms": best * 1e3, "tflops": m * 2 * eng.mlp_macs_per_row / best / 1e12}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
