"""Generate tests/golden/reference_postprocessing.npz by running the REFERENCE's own
``remove_acquisition_footprint`` / ``spatial_antialiasing`` (imported from /root/reference; xarray, dask and
matplotlib are absent here and only needed by other functions of that module, so empty stubs are registered).

Run in the build container only:  ``python oracle/make_golden_postprocessing.py``.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_postprocessing.npz")

CASES = [
    dict(name="fp_both", fn="footprint", shape=(96, 80), kw=dict(sigma=3, direction="both", buffer_center=0.25, buffer_filter=2)),
    dict(name="fp_iline", fn="footprint", shape=(75, 112), kw=dict(sigma=4, direction="iline", buffer_center=0.2, buffer_filter=3)),
    dict(name="fp_xline_complex", fn="footprint", shape=(64, 64), complex=True, kw=dict(sigma=2, direction="xline", buffer_center=0.3, buffer_filter=1)),
    dict(name="aa_iline", fn="antialias", shape=(120, 60), kw=dict(direction="iline", factors_upsampling={"iline": 4, "xline": 1}, sigma=3)),
    dict(name="aa_xline", fn="antialias", shape=(50, 128), kw=dict(direction="xline", factors_upsampling={"iline": 1, "xline": 2}, sigma=2)),
]


def make_slice(case):
    rng = np.random.default_rng(sum(map(ord, case["name"])))
    ny, nx = case["shape"]
    i, j = np.mgrid[:ny, :nx]
    d = np.zeros((ny, nx))
    for _ in range(4):
        d += rng.uniform(0.3, 1) * np.cos(2 * np.pi * (rng.uniform(-.2, .2) * i + rng.uniform(-.2, .2) * j) + rng.uniform(0, 6))
    d[::6, :] *= 1.5                                   # an "acquisition footprint"
    d += 0.05 * rng.standard_normal((ny, nx))
    if case.get("complex"):
        d = d + 1j * np.roll(d, 3, axis=1)
    return d


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "xarray", "dask",
                 "dask.diagnostics", "dask.array", "segyio", "tqdm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["dask.diagnostics"].ProgressBar = object
    sys.path.insert(0, os.environ.get("P3D_REFERENCE_ROOT", "/root/reference"))
    import pseudo_3D_interpolation.cube_postprocessing_3D as pp
    return pp


def main():
    pp = load_reference()
    store = {}
    for c in CASES:
        d = make_slice(c)
        fn = pp.remove_acquisition_footprint if c["fn"] == "footprint" else pp.spatial_antialiasing
        y, f = fn(d, return_filter=True, verbose=0, **c["kw"])
        store[c["name"] + "__y"] = y
        store[c["name"] + "__filter"] = f
        print(f"{c['name']:20s} shape={d.shape} |y|={np.linalg.norm(y):.4f} filter range [{f.min():.4f}, {f.max():.4f}]")
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
