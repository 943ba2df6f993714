"""Generate tests/golden/*.npz by running the UNMODIFIED reference (this container only).

    python tools/make_golden.py

Imports /root/reference through oracle/ref_import.py (keras stubbed; FABRIK/FK/generators only)
and stores inputs + reference outputs so that the GPU box, which has no /root/reference, can run
the same parity checks.  Re-running is deterministic (fixed seeds)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def main():
    ref = ref_import.load()
    gen = ref.position_generator.TrainingDataGenerator
    limits = ref.robot.SixDOFRobot.effector_workspace_limits
    sets = {}
    # (W) full workspace box, cube_random semantics (position_generator.py:48-55), seed 1234
    np.random.seed(1234)
    n_w = 4000
    sets["workspace"] = np.array(gen.cube_random(648.0 / n_w, 6, 12, 9, start=(0, -6, -3)))[:n_w]
    # (R) reachable interior box
    np.random.seed(1234)
    n_r = 2000
    sets["interior"] = np.array(gen.cube_random(24.0 / n_r, 2, 4, 3, start=(1, -2, 1)))[:n_r]
    sets["spring50"] = np.array(gen.spring(50, 2, 3, 6))          # cli.py:195 example
    sets["spring500"] = np.array(gen.spring(500, 2, 3, 6))
    sets["circle200"] = np.array(gen.circle(2, 200, (2, 0, 2)))
    np.random.seed(1234)
    sets["normal05"] = np.array(gen.random_distribution(1000, limits, "normal", 0.5))
    # near the reach boundary ||T-(0,0,2)|| ~ 6: slow convergence / iteration cap
    rng = np.random.RandomState(7)
    dirs = rng.randn(600, 3); dirs[:, 0] = np.abs(dirs[:, 0]); dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    pts = np.array([0, 0, 2.0]) + dirs * rng.uniform(5.5, 6.3, size=(600, 1))
    ok = (pts[:, 0] <= 6) & (np.abs(pts[:, 1]) <= 6) & (pts[:, 2] >= -3) & (pts[:, 2] <= 6)
    sets["boundary"] = pts[ok][:400]
    # edge cases that are legal under check_limits and non-degenerate
    sets["edge"] = np.array([[6, 6, 6], [6, -6, -3], [6, 6, -3], [0.5, 0, 8 - 2.5], [3, 0, 2],
                             [1e-3, 1e-3, 5.9], [6, 0, 2], [1, 2, 3], [0.001, 5.999, 2.0],
                             [1.0, 2.1, 3.0], [1.567, 2.22, -2.123], [1.02, 3.33, 4.99]], dtype=float)
    out = {}
    t0 = time.time()
    for name, pts in sets.items():
        angles, iters = ref_import.fabrik_ikine_with_iterations(ref, pts.tolist())
        out[f"{name}_xyz"] = pts
        out[f"{name}_angles"] = np.array(angles)
        out[f"{name}_iters"] = np.array(iters, dtype=np.int32)
        print(f"{name}: n={len(pts)} mean iters {np.mean(iters):.2f} ({time.time() - t0:.1f}s)")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fabrik_reference.npz"), **out)

    # FK: reference ForwardKinematics.fkine on random angle sets within (-2pi, 2pi)
    dh, _, _ = ref_import.fresh_robot_constants(ref)
    fk = ref.forward.ForwardKinematics(dh)
    rng = np.random.RandomState(99)
    ang = rng.uniform(-np.pi, np.pi, size=(2000, 4))
    pos = np.array([[m[0, 3], m[1, 3], m[2, 3]] for m in (fk.fkine(list(a))[0] for a in ang)])
    chain0 = np.array(fk.fkine(list(ang[0]))[1])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fk_reference.npz"),
                        angles=ang, positions=pos, chain0=chain0)
    print("fk:", pos.shape)

    # generators: reference outputs for the shapes the benchmarks use
    np.random.seed(1234)
    cr = np.array(gen.cube_random(648.0 / 1000, 6, 12, 9, start=(0, -6, -3)))
    np.random.seed(1234)
    rd = np.array(gen.random_distribution(500, limits, "normal", 0.35))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "generators_reference.npz"),
                        cube_random=cr, normal035=rd,
                        circle=np.array(gen.circle(2, 50, (2, 0, 2))),
                        spring=np.array(gen.spring(50, 2, 3, 6)),
                        cube=np.array(gen.cube(0.5, 2, 3, 1.5, start=(1, -1, 0))))
    print("done")


if __name__ == "__main__":
    main()
