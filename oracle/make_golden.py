"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported from
/root/reference) on the seeded cases of ``golden_cases.py``.

Run in the build container only:  ``python oracle/make_golden.py``.
The stored arrays are outputs of the unmodified reference code with
``transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind='FFT'`` on inputs
upcast to complex128/float64 (the float64 reference), plus
``get_threshold_decay`` tables and ``threshold`` outputs.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import _reference_loader            # noqa: E402
from oracle.golden_cases import CASES, SCHEDULES, make_input   # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref = _reference_loader.load()
    os.makedirs(OUT, exist_ok=True)
    store = {}
    for case in CASES:
        x, mask = make_input(case)
        xin = x.astype(np.float64 if case.get("real_input") else np.complex128)
        fn = {"regular": ref.POCS, "fast": ref.FPOCS, "adaptive": ref.APOCS}[case.get("version", "regular")]
        info = {}
        tmp = os.path.join(OUT, "_costs.tmp")
        if os.path.exists(tmp):
            os.remove(tmp)
        y = fn(xin, mask, None, transform=np.fft.fft2, itransform=np.fft.ifft2,
               transform_kind="FFT", results_dict=info, path_results=tmp, **case["params"])
        with open(tmp) as f:
            fields = f.read().strip().split(";")
        os.remove(tmp)
        costs = np.array([float(v) for v in fields[2:]], dtype=np.float64)
        n = case["name"]
        store[f"{n}__y"] = np.asarray(y)
        store[f"{n}__niterations"] = np.int64(info["niterations"])
        store[f"{n}__costs"] = costs
        # the reference's OWN complex64 path (what numpy >= 2 runs in production) against its
        # float64 result: the fp32 conditioning floor of this case
        if not case.get("all_zero"):
            y32 = fn(x, mask, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT", **case["params"])
            store[f"{n}__c64_drift"] = np.float64(np.linalg.norm(y32 - y) / np.linalg.norm(y))
        print(f"{n:28s} shape={x.shape} niterations={info['niterations']:3d} cost={float(info['cost']):.3e}")

    # schedule tables on the spectrum of case 0
    x, mask = make_input(CASES[0])
    X0 = np.fft.fft2(x.astype(np.complex128))
    for i, sp in enumerate(SCHEDULES):
        tau = ref.get_threshold_decay(sp["thresh_model"], sp["niter"], "FFT", sp["p_max"], sp["p_min"],
                                      x_fwd=X0, kind=sp["kind"])
        store[f"schedule_{i}"] = np.asarray(tau)

    # threshold operator, complex tau, including edge values (|X| == Re(tau), X == 0)
    rng = np.random.default_rng(5)
    X = (rng.standard_normal((6, 7)) + 1j * rng.standard_normal((6, 7)))
    X[0, 0] = 0.0
    X[1, 1] = 0.6 + 0.8j       # |X| == 1.0 exactly
    X[2, 2] = -1.0
    store["thr_X"] = X
    for tname, tau in (("pos", 1.0 + 0.25j), ("neg", 1.0 - 0.25j), ("real", 0.8)):
        for kind in ("hard", "soft", "garrote"):
            store[f"thr_{kind}_{tname}"] = ref.threshold(X, tau, sub=0, kind=kind)
    np.savez_compressed(os.path.join(OUT, "reference_pocs.npz"), **store)
    print("wrote", os.path.join(OUT, "reference_pocs.npz"))


if __name__ == "__main__":
    main()
