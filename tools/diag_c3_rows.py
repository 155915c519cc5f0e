import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import pseudo_3d_interpolation_b200 as p3d
from oracle import pocs_oracle as orc
from oracle.golden_cases import make_input
def rel(a, b): return float(np.linalg.norm(a.astype(np.complex128) - b) / np.linalg.norm(b))
for shape, which in (((48, 1201), 1), ((847, 1201), 0)):
    x, mask = make_input(dict(seed=9, shape=shape, keep=0.3, nwaves=5))
    xs = np.stack([x, 0.25 * np.conj(x)]).astype(np.complex64)[which:which + 1]
    for niter in (2, 3, 4, 5, 6, 7):
        params = dict(niter=niter, thresh_op="garrote", thresh_model="exponential", eps=0.0, alpha=0.7, p_max=0.99, p_min=1e-3)
        y, _ = p3d.PocsPlan(*shape).run(xs, mask, version="adaptive", **params)
        g = p3d.PocsPlan(*shape); g.set_option("force_generic", 1)
        yg, _ = g.run(xs, mask, version="adaptive", **params)
        y64, _ = p3d.PocsPlan(*shape, precision=64).run(xs, mask, version="adaptive", **params)
        ref = orc.pocs_slice(xs[0].astype(np.complex128), mask, version="adaptive", **params)
        print(shape, "niter", niter, "rader-ref %.2e generic-ref %.2e f64-ref %.2e" % (rel(y[0], ref), rel(yg[0], ref), rel(y64[0], ref)))
