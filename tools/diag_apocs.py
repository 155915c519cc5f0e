import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pseudo_3d_interpolation_b200 as p3d
from oracle.golden_cases import CASES, make_input
from oracle import pocs_oracle as orc
def rel(a, b):
    a = np.asarray(a, np.complex128); b = np.asarray(b, np.complex128)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
case = [c for c in CASES if c["name"] == "apocs_a08"][0]
x, mask = make_input(case)
for ver in ("adaptive", "regular"):
    for op in ("hard", "soft"):
        for alpha in (1.0, 0.8):
            for niter in (2, 3, 5, 20):
                params = dict(niter=niter, thresh_op=op, thresh_model="exponential", eps=0.0, alpha=alpha, p_max=0.99, p_min=1e-4)
                ref = orc.pocs_slice(x.astype(np.complex128), mask, version=ver, **params)
                plan = p3d.pocs.get_plan(*x.shape)
                y, _ = plan.run(x, mask, version=ver, **params)
                other = orc.pocs_slice(x.astype(np.complex128), mask, version=("regular" if ver == "adaptive" else "adaptive"), **params)
                print(f"{ver:9s} {op:5s} alpha={alpha} niter={niter:2d}: vs-same-version {rel(y[0], ref):.3e}   vs-other-version {rel(y[0], other):.3e}")
