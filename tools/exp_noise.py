"""Lockstep complex64 / complex128 trajectories: size of the fp32 deviation of the coefficients and the first decision flip."""
import sys
import numpy as np
sys.path.insert(0, ".")
from pseudo_3d_interpolation_b200 import synth
from oracle import pocs_oracle as orc
cfg = int(sys.argv[1]); sids = [int(a) for a in sys.argv[2].split(",")]
d, fold, c = synth.sparse_freq_slices(cfg, sids)
mask = orc.mask_from_fold(fold); keep = 1 - mask; niter = c["niter"]
EPS32 = 2.0 ** -24
for i, sid in enumerate(sids):
    x = d[i]; N = x.size
    if not np.count_nonzero(x): continue
    X0 = np.fft.fft2(x.astype(np.complex128))
    tau = orc.threshold_table(X0, niter, "exponential", 0.99, 1e-5)
    nnz = np.count_nonzero(x)
    u_rms = EPS32 * np.sqrt((np.abs(X0) ** 2).sum() / nnz)
    u_max = EPS32 * np.abs(X0).max() * N / nnz
    xa = x.astype(np.complex64); xb = x.astype(np.complex128)
    print(f"slice {sid}: u_max/u_rms = {u_max / u_rms:.1f}")
    for k in range(niter):
        Xa = np.fft.fft2(xa); Xb = np.fft.fft2(xb)
        ra, rb = np.abs(Xa), np.abs(Xb)
        a32 = np.complex64(tau[k]).real; a64 = tau[k].real
        ka = ra < a32; kb = rb < a64
        dev = np.abs(ra.astype(np.float64) - rb)
        near = np.abs(rb - a64) < 0.2 * a64
        flips = np.flatnonzero(ka != kb)
        msg = f"  k {k:3d} tau/z {a64 / abs(tau[0].real) * 0.99:.1e} max dev {dev.max() / u_rms:8.1f} u_rms = {dev.max() / u_max:6.2f} u_max; near-tau dev {dev[near].max() / u_rms if near.any() else 0:8.1f} u_rms"
        if flips.size:
            j = flips[0]
            msg += f"  FLIPS {flips.size}: |r-a| = {abs(rb.flat[j] - a64) / u_rms:.1f} u_rms = {abs(rb.flat[j] - a64) / u_max:.2f} u_max"
            print(msg); break
        if k % 4 == 0: print(msg)
        ya = np.fft.ifft2(np.where(ka, 0, Xa)); ya *= keep; ya += x.astype(np.complex64); xa = ya
        yb = np.fft.ifft2(np.where(kb, 0, Xb)); yb *= keep; yb += x; xb = yb
