import sys; sys.path.insert(0, ".")
import numpy as np
import pseudo_3d_interpolation_b200 as p3d
from oracle.golden_cases import CASES, make_input
g = np.load("tests/golden/reference_pocs.npz")
for c in CASES:
    if not c["params"]["thresh_op"].endswith("-percentile"): continue
    x, mask = make_input(c)
    y = p3d.POCS(x, mask, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT", **c["params"])
    ref = g[c["name"] + "__y"]
    print(c["name"], c["params"], np.linalg.norm(y - ref) / np.linalg.norm(ref))
