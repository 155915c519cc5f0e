"""Randomised cross-check of the numpy oracle against the live reference (build container only).

``python oracle/validate_against_reference.py [n_trials]`` draws random shapes, operators,
schedules and parameters, runs the reference's ``POCS_algorithm`` (imported from
/root/reference) and ``oracle.pocs_oracle.pocs_slice`` on the same input and requires
bit-for-bit equal outputs, iteration counts and cost histories.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import _reference_loader, pocs_oracle as orc     # noqa: E402
from oracle.golden_cases import make_input                   # noqa: E402


def main(n_trials=40):
    ref = _reference_loader.load()
    rng = np.random.default_rng(2024)
    ops = ["hard", "soft", "garrote"]
    models = ["linear", "exponential", "exponential-2", "data-driven", "inverse_proportional"]
    bad = 0
    for t in range(n_trials):
        case = dict(seed=int(rng.integers(1 << 30)), shape=(int(rng.integers(8, 70)), int(rng.integers(8, 70))),
                    noise=float(rng.choice([0.0, 0.02])), keep=float(rng.uniform(0.2, 0.7)),
                    real_input=bool(rng.random() < 0.25))
        params = dict(niter=int(rng.integers(4, 40)), thresh_op=str(rng.choice(ops)),
                      thresh_model=str(rng.choice(models)), eps=float(rng.choice([0.0, 1e-9, 1e-6])),
                      alpha=float(rng.choice([1.0, 0.7])), p_max=0.99,
                      p_min=(("adaptive" if rng.random() < 0.2 else float(rng.choice([1e-5, 1e-3])))),
                      sqrt_decay=bool(rng.random() < 0.15))
        version = str(rng.choice(["regular", "fast", "adaptive"]))
        x, mask = make_input(case)
        xin = x.astype(np.float64 if case["real_input"] else np.complex128)
        fn = {"regular": ref.POCS, "fast": ref.FPOCS, "adaptive": ref.APOCS}[version]
        i_ref, i_orc = {}, {}
        y_ref = fn(xin, mask, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT",
                   results_dict=i_ref, **params)
        y_orc = orc.pocs_slice(xin, mask, version=version, info=i_orc, **params)
        ok = np.array_equal(y_ref, y_orc, equal_nan=True) and i_ref["niterations"] == i_orc["niterations"]
        bad += (not ok)
        print(f"[{t:3d}] {'OK ' if ok else 'BAD'} {case['shape']} {version:8s} {params['thresh_op']:7s} "
              f"{params['thresh_model']:20s} its={i_ref['niterations']}")
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 40) else 0)
