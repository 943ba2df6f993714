"""Independent fp32 evaluation of reference kinematics/ann.py:70-76 with torch on the CPU (TEST INFRASTRUCTURE --
see oracle/__init__.py): a second, third-party implementation of Dense + tanh next to the NumPy restatement in
np_oracle.mlp_predict, so that the ANN kernels are not only ever compared with the builder's own arithmetic.

``torch.nn.functional.linear`` (oneDNN / MKL GEMM, its own blocking and summation order) and ``torch.tanh`` in
float32; the scalers follow sklearn's semantics exactly as np_oracle does (transform in float64, cast to float32 for
the network as Keras does, inverse_transform in place on the float32 prediction).  ANN parity with Keras itself stays
UNPINNED: neither keras nor the reference's weights exist in this environment.
"""
import contextlib

import numpy as np


def gemm_rel_error(strict=True):
    """Relative error of a 512 x 512 fp32 F.linear against float64: ~1e-7 for IEEE fp32 accumulation.  oneDNN may run
    fp32 matmuls through reduced-precision units on CPUs that have them (seen on a B200 host: 5e-4 rad on the trained
    network with mkldnn enabled), which is not the fp32 arithmetic Keras' CPU kernels use."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    a, w = torch.randn(512, 512, generator=g), torch.randn(512, 512, generator=g)
    with _strict_fp32() if strict else contextlib.nullcontext():
        got = F.linear(a, w).double()
    want = F.linear(a.double(), w.double())
    return float((got - want).abs().max() / want.abs().max())


@contextlib.contextmanager
def _strict_fp32():
    """IEEE fp32 GEMMs: oneDNN off (MKL sgemm is used instead), float32 matmul precision 'highest'."""
    import torch
    old = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        with torch.backends.mkldnn.flags(enabled=False):
            yield
    finally:
        torch.set_float32_matmul_precision(old)


def mlp_predict(xyz, weights, biases, mean_x, scale_x, mean_y, scale_y, chunk=65536):
    import torch
    import torch.nn.functional as F
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    Ws = [torch.from_numpy(np.ascontiguousarray(np.asarray(w, dtype=np.float32).T)) for w in weights]   # (out, in)
    bs = [torch.from_numpy(np.asarray(b, dtype=np.float32)) for b in biases]
    sy, my = torch.from_numpy(np.asarray(scale_y, np.float32)), torch.from_numpy(np.asarray(mean_y, np.float32))
    out = np.empty((xyz.shape[0], Ws[-1].shape[0]), dtype=np.float32)
    with torch.no_grad(), _strict_fp32():
        for lo in range(0, xyz.shape[0], chunk):
            xs = (xyz[lo:lo + chunk] - np.asarray(mean_x)) / np.asarray(scale_x)       # StandardScaler.transform, fp64
            h = torch.from_numpy(xs.astype(np.float32))
            for W, b in zip(Ws[:-1], bs[:-1]):
                h = torch.tanh(F.linear(h, W, b))
            y = F.linear(h, Ws[-1], bs[-1])
            y.mul_(sy).add_(my)                                                         # inverse_transform, in place
            out[lo:lo + chunk] = y.numpy()
    return out


def mlp_predict_fresh_process(xyz, weights, biases, mean_x, scale_x, mean_y, scale_y, timeout=900):
    """mlp_predict in a new interpreter.  In a long-lived process that has already loaded CUDA, OpenMP code of its own
    and several BLAS users, torch's CPU GEMMs were seen to come out 5e-4 rad off on one B200 host while the very same
    call in a fresh process agreed with NumPy to 1.5e-6 (gpurun_out diagnostics, round 2); a checker must not depend on
    that, so the tests and bench.py run it out of process."""
    import os
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        arrays = {"xyz": np.asarray(xyz), "mean_x": np.asarray(mean_x), "scale_x": np.asarray(scale_x),
                  "mean_y": np.asarray(mean_y), "scale_y": np.asarray(scale_y), "n_layers": np.array(len(weights))}
        for i, (w, b) in enumerate(zip(weights, biases)):
            arrays[f"W{i}"], arrays[f"b{i}"] = np.asarray(w), np.asarray(b)
        src, dst = os.path.join(tmp, "in.npz"), os.path.join(tmp, "out.npy")
        np.savez(src, **arrays)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        subprocess.run([sys.executable, os.path.abspath(__file__), src, dst], check=True, timeout=timeout, env=env)
        return np.load(dst)


if __name__ == "__main__":
    import sys
    data = np.load(sys.argv[1])
    nl = int(data["n_layers"])
    out = mlp_predict(data["xyz"], [data[f"W{i}"] for i in range(nl)], [data[f"b{i}"] for i in range(nl)],
                      data["mean_x"], data["scale_x"], data["mean_y"], data["scale_y"])
    np.save(sys.argv[2], out)
