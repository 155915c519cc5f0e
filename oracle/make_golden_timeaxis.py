"""Generate tests/golden/reference_timeaxis.npz by running the REFERENCE's own host pieces of steps 12 / 14:
``get_freq_filter_win`` / ``_get_stopband`` / ``get_freq_filter_mask`` (cube_apply_FFT.py:49-181) and ``rescale_dask``
(functions/utils.py:444-473, used by cube_apply_IFFT.py:121-140), imported from /root/reference.

xarray, xrft and dask are absent in this image and are needed by those modules only at import time and for the
``DataArray`` container the window is returned in, so minimal stub modules are registered first (``xr.set_options``, a
``DataArray`` that carries ``data`` / ``dims`` / ``coords`` and compares like its array).  The xrft transform itself is NOT
pinned by this file (the fork is not in the image): only the host pieces are.

Run in the build container only:  ``python oracle/make_golden_timeaxis.py``.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_timeaxis.npz")

DT_MS = 0.05

# (name, filter type, corner frequencies in kHz, nt, rfft axis?)
WINDOW_CASES = [
    ("lowpass_rfft", "lowpass", [1.2, 3.0], 512, True),
    ("highpass_rfft", "highpass", [0.4, 1.1], 512, True),
    ("bandpass_rfft", "bandpass", [0.3, 0.9, 3.1, 4.4], 2048, True),
    ("lowpass_fft", "lowpass", [1.0, 2.5], 256, False),          # fftfreq order (negative frequencies last)
    ("bandpass_odd", "bandpass", [0.5, 1.0, 2.0, 2.6], 375, True),
    ("highpass_narrow", "highpass", [0.9, 0.95], 300, True),      # stopband of one or two samples
]


class _DataArray:
    """The few things the reference touches on an xarray.DataArray in the functions pinned here."""

    def __init__(self, data, dims=None, coords=None, **kw):
        self.data = np.asarray(data)
        self.values = self.data
        self.dims = list(dims) if dims is not None else []
        self.coords = dict(coords) if coords is not None else {}

    def __getitem__(self, key):
        return _DataArray(self.coords[key], dims=[key], coords={key: self.coords[key]})

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def __lt__(self, o): return self.data < o
    def __le__(self, o): return self.data <= o
    def __gt__(self, o): return self.data > o
    def __ge__(self, o): return self.data >= o


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "xarray", "xrft", "dask",
                 "dask.diagnostics", "dask.array", "segyio", "tqdm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["dask.diagnostics"].ProgressBar = object
    xr = sys.modules["xarray"]
    xr.set_options = lambda **kw: None
    xr.DataArray = _DataArray
    xr.Dataset = object
    sys.path.insert(0, os.environ.get("P3D_REFERENCE_ROOT", "/root/reference"))
    import pseudo_3D_interpolation.cube_apply_FFT as fwd
    import pseudo_3D_interpolation.functions.utils as utils
    return fwd, utils


def axis(nt, real):
    return np.fft.rfftfreq(nt, DT_MS) if real else np.fft.fftfreq(nt, DT_MS)


def rescale_inputs():
    rng = np.random.default_rng(77)
    a = rng.standard_normal((40, 6, 5)).astype(np.float32)          # (twt, iline, xline) with negative samples
    b = np.abs(rng.standard_normal((33, 4, 3))).astype(np.float32) + 0.25
    c = np.full((8, 2, 2), 0.5, dtype=np.float32)                     # constant: returned unchanged
    return {"mixed": a, "positive": b, "constant": c}


def main():
    fwd, utils = load_reference()
    store = {}
    for name, kind, freqs, nt, real in WINDOW_CASES:
        f = axis(nt, real)
        da = _DataArray(f, dims=["freq_twt"], coords={"freq_twt": f})
        win = fwd.get_freq_filter_win(list(freqs), da, dim="freq_twt", filter_type=kind)
        keep = fwd.get_freq_filter_mask(_DataArray(np.zeros_like(f), dims=["freq_twt"], coords={"freq_twt": f}), "freq_twt", list(freqs), kind)
        store[f"win__{name}"] = np.asarray(win.data, dtype=np.float64)
        store[f"keep__{name}"] = np.asarray(keep, dtype=bool)
        print(f"{name:18s} n={f.size:5d} sum(win)={win.data.sum():.6f} kept={int(np.count_nonzero(keep))}")
    for n in (0, 1, 2, 3, 7, 8):
        for kind in ("lowpass", "highpass"):
            store[f"stopband__{kind}_{n}"] = np.asarray(fwd._get_stopband(n, kind), dtype=np.float64)
    # --rescale-envelope (cube_apply_IFFT.py:121-140): clip below zero, global min / max, rescale_dask per trace
    for name, x in rescale_inputs().items():
        clipped = np.where(x < 0, 0, x)
        amin, amax = clipped.min(), clipped.max()
        y = utils.rescale_dask(clipped, amin=amin, amax=amax)
        store[f"rescale__{name}"] = np.asarray(y)
        print(f"rescale {name:10s} -> [{np.min(y):.4f}, {np.max(y):.4f}] dtype {np.asarray(y).dtype}")
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
