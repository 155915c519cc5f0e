"""Experiment (CPU, numpy): fp32 state for the first K iterations, complex128 afterwards.
How late can the switch be before the result leaves the 1e-4 of the float64 reference?"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pseudo_3d_interpolation_b200 import synth
from oracle import pocs_oracle as orc

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sid = int(sys.argv[2]) if len(sys.argv) > 2 else 300
d, fold, c = synth.sparse_freq_slices(cfg, [sid])
mask = orc.mask_from_fold(fold)
x = d[0]
niter = c["niter"]
N = x.size

X0 = np.fft.fft2(x.astype(np.complex128))
tau = orc.threshold_table(X0, niter, "exponential", 0.99, 1e-5)
keep = 1 - mask

def run(K, trace=False):
    xp = x.astype(np.complex64)
    hist = []
    for k in range(niter):
        if k == K:
            xp = xp.astype(np.complex128)
        X = np.fft.fft2(xp)
        t = tau[k] if k >= K else np.complex64(tau[k])
        r = np.abs(X)
        a, b = t.real, t.imag
        kill = (r < a) | ((r == a) & (0 < b))
        if trace:
            nrm = np.linalg.norm(X)
            g = 8 * 6e-8 * nrm / np.sqrt(N) * np.sqrt(np.log2(N))
            hist.append((k, float(a / abs(tau[0])), float(1 - kill.mean()), int((np.abs(r - a) < g).sum())))
        Y = np.where(kill, 0, X)
        y = np.fft.ifft2(Y)
        y *= keep
        y += (x if k >= K else x.astype(np.complex64))
        xp = y
    return xp.astype(np.complex128), hist

t0 = time.time()
ref, hist = run(0, trace=True)
print("ref time", time.time() - t0)
for h in hist[::3]:
    print("k %3d tau/tau0 %.2e kept %.3e guard %d" % h)
for K in [int(a) for a in (sys.argv[3].split(",") if len(sys.argv) > 3 else "100,80,70,60,50,40,30,20".split(","))]:
    y, _ = run(K)
    print("K", K, "err", np.linalg.norm(y.astype(np.complex64) - ref.astype(np.complex64)) / np.linalg.norm(ref))
