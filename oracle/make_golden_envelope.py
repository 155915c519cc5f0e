"""Generate tests/golden/reference_envelope.npz with the REFERENCE's own ``functions.signal.envelope``
(imported from /root/reference; dask / matplotlib are absent here and unused by that function, so stubs are
registered).  Run in the build container only:  ``python oracle/make_golden_envelope.py``.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_envelope.npz")

# (name, nt, n_il, n_xl): even / odd / power-of-two record lengths
CASES = [("nt64", 64, 5, 7), ("nt75_odd", 75, 4, 6), ("nt512", 512, 3, 5), ("nt100", 100, 2, 9)]


def make_cube(name, nt, n_il, n_xl):
    rng = np.random.default_rng(sum(map(ord, name)))
    t = np.arange(nt)[:, None, None]
    f = rng.uniform(0.02, 0.2, (1, n_il, n_xl))
    x = np.cos(2 * np.pi * f * t + rng.uniform(0, 6, (1, n_il, n_xl))) * np.exp(-((t - nt * rng.uniform(0.3, 0.7, (1, n_il, n_xl))) / (0.15 * nt)) ** 2)
    x += 0.05 * rng.standard_normal((nt, n_il, n_xl))
    return x.astype(np.float32)


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "dask", "dask.array", "xarray", "tqdm", "segyio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["dask.array"].Array = type("Array", (), {})          # scipy's array-api helpers probe this name
    sys.path.insert(0, os.environ.get("P3D_REFERENCE_ROOT", "/root/reference"))
    import pseudo_3D_interpolation.functions.signal as sg
    return sg


def main():
    sg = load_reference()
    store = {}
    for name, nt, n_il, n_xl in CASES:
        x = make_cube(name, nt, n_il, n_xl)
        e32 = sg.envelope(x, axis=0)                       # what the pipeline computes (float32 in -> float32 out)
        e64 = sg.envelope(x.astype(np.float64), axis=0)    # the float64 reference
        assert e32.dtype == np.float32
        store[name + "__env32"] = e32
        store[name + "__env64"] = e64
        print(f"{name:10s} shape={x.shape} max={e64.max():.4f} f32-vs-f64 {np.abs(e32 - e64).max():.2e}")
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
