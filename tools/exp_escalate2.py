"""Experiment (CPU, numpy): fp32 state until the guard band |r - Re tau_k| < G * eps32 * rms|X| catches a coefficient,
complex128 state from the next iteration on.  Prints the switch iteration and the rel-L2 error vs all-complex128."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pseudo_3d_interpolation_b200 import synth
from oracle import pocs_oracle as orc

cfg = int(sys.argv[1]); sids = [int(a) for a in sys.argv[2].split(",")]
Gs = [float(a) for a in sys.argv[3].split(",")]
noise = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
kw = {}
if len(sys.argv) > 5:
    n = int(sys.argv[5]); kw = dict(n_il=n, n_xl=n)
d, fold, c = synth.sparse_freq_slices(cfg, sids, noise=noise, **kw)
mask = orc.mask_from_fold(fold)
niter = c["niter"]
keep = 1 - mask
EPS32 = 2.0 ** -24

def run(x, G):
    N = x.size
    X0 = np.fft.fft2(x.astype(np.complex128))
    tau = orc.threshold_table(X0, niter, "exponential", 0.99, 1e-5)
    nnz = np.count_nonzero(x)
    u = EPS32 * np.sqrt((np.abs(X0) ** 2).sum() / N * (N / nnz))
    g = G * u
    xp = x.astype(np.complex64)
    K = niter
    wide = True
    for k in range(niter):
        if k == K:
            xp = xp.astype(np.complex128)
        lo = k < K
        X = np.fft.fft2(xp)
        t = np.complex64(tau[k]) if lo else tau[k]
        r = np.abs(X)
        a, b = t.real, t.imag
        kill = (r < a) | ((r == a) & (0 < b))
        if lo and G > 0 and (np.abs(r - a) < g).any():
            K = k + 1
        Y = np.where(kill, 0, X)
        y = np.fft.ifft2(Y)
        y *= keep
        y += (x.astype(np.complex64) if lo else x)
        xp = y
    return xp.astype(np.complex64), K

for i, sid in enumerate(sids):
    x = d[i]
    if not np.count_nonzero(x):
        continue
    ref, _ = run(x, 1e30)   # switches at k=1 -> all complex128
    assert _ == 1
    line = f"cfg {cfg} slice {sid}:"
    for G in Gs:
        y, K = run(x, G)
        line += f"  G={G:g}: K={K} err={np.linalg.norm(y - ref) / np.linalg.norm(ref):.1e}"
    print(line, flush=True)
