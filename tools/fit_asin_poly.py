"""Fit the polynomial P(t) with asin(x) = x + x*t*P(t), t = x^2 in [0, 0.25], used by the fp64
epilogue of csrc/fabrik.cu (acos / atan2 replacements).  Weighted least squares on Chebyshev nodes
in 60-digit arithmetic (mpmath), then rounded to double; prints the max abs error of asin."""
import sys
import mpmath as mp
import numpy as np

mp.mp.dps = 60
DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 11
N = 400
nodes = [(mp.mpf(0.25) / 2) * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * N))) for k in range(N)]


def target(t):
    x = mp.sqrt(t)
    if t < mp.mpf(10) ** -30:
        return mp.mpf(1) / 6
    return (mp.asin(x) - x) / (x * t)


A = mp.matrix(N, DEG + 1)
b = mp.matrix(N, 1)
for i, t in enumerate(nodes):
    w = mp.sqrt(t) * t  # weight: absolute error of asin, not of P
    for j in range(DEG + 1):
        A[i, j] = w * t ** j
    b[i] = w * target(t)
coef = mp.lu_solve(A.T * A, A.T * b)
c64 = [float(c) for c in coef]
xs = np.linspace(0, 0.5, 20001)
err = 0.0
for x in xs:
    t = x * x
    p = 0.0
    for c in reversed(c64):
        p = p * t + c
    val = x + x * t * p
    err = max(err, abs(val - float(mp.asin(mp.mpf(float(x))))))
print("degree", DEG, "max abs err", err)
for c in c64:
    print(f"    {c!r},")
