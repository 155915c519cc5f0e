"""After a switch at K (complex64 state -> complex128), how does the deviation from the all-complex128 trajectory evolve?"""
import sys
import numpy as np
sys.path.insert(0, ".")
from pseudo_3d_interpolation_b200 import synth
from oracle import pocs_oracle as orc
cfg = int(sys.argv[1]); sid = int(sys.argv[2]); K = int(sys.argv[3])
d, fold, c = synth.sparse_freq_slices(cfg, [sid])
mask = orc.mask_from_fold(fold); keep = 1 - mask; niter = c["niter"]
EPS32 = 2.0 ** -24
x = d[0]; N = x.size
X0 = np.fft.fft2(x.astype(np.complex128))
tau = orc.threshold_table(X0, niter, "exponential", 0.99, 1e-5)
nnz = np.count_nonzero(x)
u_rms = EPS32 * np.sqrt((np.abs(X0) ** 2).sum() / nnz)
u_max = EPS32 * np.abs(X0).max() * N / nnz
print(f"u_max/u_rms {u_max / u_rms:.1f}")
xa = x.astype(np.complex64); xb = x.astype(np.complex128)
for k in range(niter):
    if k == K: xa = xa.astype(np.complex128)
    Xa = np.fft.fft2(xa); Xb = np.fft.fft2(xb)
    ra, rb = np.abs(Xa).astype(np.float64), np.abs(Xb)
    a = tau[k].real
    ka = ra < (np.float32(a) if k < K else a); kb = rb < a
    dev = np.abs(ra - rb)
    near = np.abs(rb - a) < 0.25 * a
    gap = np.sort(np.abs(rb - a).ravel())[:3] / u_rms
    nflip = int((ka != kb).sum())
    if k >= K - 3 or nflip:
        print(f"k {k:3d} tau/z {a / abs(tau[0].real) * .99:.1e} kept {1 - kb.mean():.2e} dev near tau: max {dev[near].max() / u_rms:7.2f} rms {np.sqrt((dev[near]**2).mean()) / u_rms:7.3f} u_rms | all: max {dev.max() / u_rms:8.1f} | closest gaps {gap[0]:.1f} {gap[1]:.1f} {gap[2]:.1f} u_rms | flips {nflip}")
    ya = np.fft.ifft2(np.where(ka, 0, Xa)); ya *= keep; ya += (x.astype(np.complex64) if k < K else x); xa = ya
    yb = np.fft.ifft2(np.where(kb, 0, Xb)); yb *= keep; yb += x; xb = yb
print("final err", np.linalg.norm(xa - xb) / np.linalg.norm(xb))
