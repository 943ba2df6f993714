import os, sys, numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import np_oracle, torch_oracle
from joblib import load
m = os.path.join(ROOT, 'models/roboarm_b200_r01')
d = np.load(m + '.npz'); nl = len([k for k in d.files if k.startswith('W')])
W = [d[f'W{i}'] for i in range(nl)]; b = [d[f'b{i}'] for i in range(nl)]
sx, sy = load(m + '_scaler_x.bin'), load(m + '_scaler_y.bin')
sc = (sx.mean_, sx.scale_, sy.mean_, sy.scale_)
rng = np.random.default_rng(23)
pts = (rng.random((30000, 3)) * [6, 12, 9] + [0, -6, -3]).astype(np.float32)
n32 = np_oracle.mlp_predict(pts, W, b, *sc); n64 = np_oracle.mlp_predict(pts, W, b, *sc, dtype=np.float64)
t32 = torch_oracle.mlp_predict(pts, W, b, *sc)
print("torch vs numpy", np.abs(t32 - n32).max(), "numpy vs 64", np.abs(n32 - n64).max(), "torch vs 64", np.abs(t32 - n64).max())
# layer by layer
xs = ((pts.astype(np.float64) - sx.mean_) / sx.scale_)
h64 = xs.copy(); hn = xs.astype(np.float32); ht = torch.from_numpy(xs.astype(np.float32))
for l in range(nl - 1):
    h64 = np.tanh(h64 @ W[l].astype(np.float64) + b[l].astype(np.float64))
    hn = np.tanh(hn @ W[l] + b[l])
    with torch.no_grad():
        ht = torch.tanh(F.linear(ht, torch.from_numpy(np.ascontiguousarray(W[l].T)), torch.from_numpy(b[l])))
    print(l, "numpy err", float(np.abs(hn - h64).max()), "torch err", float(np.abs(ht.numpy() - h64).max()), "dtype", ht.dtype)
for shape in ((30000, 500, 500), (30000, 512, 512), (4096, 500, 500), (30000, 3, 500)):
    M, K, N = shape
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) * 0.05
    want = a.double() @ w.double().T
    print(shape, "F.linear err", float((F.linear(a, w).double() - want).abs().max()), "np err", float(np.abs((a.numpy() @ w.numpy().T) - want.numpy()).max()))
