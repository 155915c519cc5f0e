"""Rader / mixed-radix plans vs generic path vs float64 oracle on one awkward case (development diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pseudo_3d_interpolation_b200 as p3d
from oracle import pocs_oracle as orc
from oracle.golden_cases import make_input

def rel(a, b): return float(np.linalg.norm(a.astype(np.complex128) - b) / np.linalg.norm(b))

import sys as _s
shape = tuple(int(v) for v in _s.argv[1:3]) if len(_s.argv) > 2 else (1201, 48)
for seed in (9, 10, 11):
    for op, model, alpha, version in (("garrote", "exponential", 0.7, "adaptive"), ("garrote", "exponential", 1.0, "regular"), ("soft", "exponential", 0.7, "adaptive")):
        x, mask = make_input(dict(seed=seed, shape=shape, keep=0.3, nwaves=5))
        x = x.astype(np.complex64)[None]
        for niter in (1, 2, 4, 7):
            params = dict(niter=niter, thresh_op=op, thresh_model=model, eps=0.0, alpha=alpha, p_max=0.99, p_min=1e-3)
            y, _ = p3d.PocsPlan(*shape).run(x, mask, version=version, **params)
            g = p3d.PocsPlan(*shape); g.set_option("force_generic", 1)
            yg, _ = g.run(x, mask, version=version, **params)
            ref = orc.pocs_slice(x[0].astype(np.complex128), mask, version=version, **params)
            tau = orc.threshold_table(np.fft.fft2(x[0].astype(np.complex128)), max(niter, 2), model, 0.99, 1e-3)
            print(f"seed {seed} {op:8s} a={alpha} {version:8s} niter={niter}: rader-ref {rel(y[0], ref):.2e}  generic-ref {rel(yg[0], ref):.2e}  rader-generic {rel(y[0], yg[0]):.2e}  Im/Re(tau0) {tau[0].imag/tau[0].real:+.2f}")
