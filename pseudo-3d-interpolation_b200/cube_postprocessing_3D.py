"""kx-ky domain filters of the reference's post-processing step, on the B200 path.

Mirrors (names, arguments, defaults, return values) the two functions of
``cube_postprocessing_3D.py`` whose per-slice arithmetic is ``ifft2(filter * fft2(slice)).real``:

* ``remove_acquisition_footprint``  (cube_postprocessing_3D.py:179-260)
* ``spatial_antialiasing``          (cube_postprocessing_3D.py:263-347)

plus their helpers ``gaussian_kernel_2d`` (:131-176) and ``rescale`` (functions/utils.py:413-441).
The filter plane is built once on the host (a few small numpy/scipy calls, as in the reference);
the transform -> multiply -> inverse transform of every slice runs on the GPU through
``p3d_kxky_filter_run`` (the fused three-pass kernels of one POCS iteration).  Unlike the reference,
``data`` may also be a stack of slices ``(n, ny, nx)`` sharing one filter - that is how a cube is
processed without one launch per slice.  There is no CPU fallback.

SURVEY.md section 8(f-2).  The remaining options of the reference's script (upsampling, smoothing,
AGC) are not part of the FFT hot path.
"""
from __future__ import annotations

import numpy as np
from scipy import signal as _signal

from .pocs import get_plan

__all__ = ["rescale", "gaussian_kernel_2d", "footprint_filter", "antialiasing_filter",
           "remove_acquisition_footprint", "spatial_antialiasing", "apply_kxky_filter"]


def rescale(a, vmin=0, vmax=1):
    """Linear map of ``a`` onto [vmin, vmax]; constant input is returned unchanged (utils.py:413-441)."""
    a = np.asarray(a)
    lo, hi = np.nanmin(a), np.nanmax(a)
    vmin = lo if vmin is None else vmin
    vmax = hi if vmax is None else vmax
    if lo == hi:
        return a
    return vmin + (a - lo) * ((vmax - vmin) / (hi - lo))


def gaussian_kernel_2d(sigma: int = 7, n=None, normalized: bool = True, orientation: str = "equal"):
    """Outer product of two Gaussian windows (cube_postprocessing_3D.py:131-176)."""
    ny, nx = n if isinstance(n, tuple) else (n, n)
    fy, fx = {"equal": (8, 8), "iline": (2, 8), "xline": (8, 2)}[orientation]
    if ny is None:
        ny = sigma * fy + 1
    if nx is None:
        nx = sigma * fx + 1
    ny += 1 - ny % 2          # odd sizes
    nx += 1 - nx % 2
    kernel = np.outer(_signal.windows.gaussian(ny, sigma), _signal.windows.gaussian(nx, sigma))
    if normalized:
        kernel /= 2 * np.pi * sigma ** 2
    return kernel


def _orientation(direction, dims, shape=None):
    if direction == "iline":
        return "horizontal" if dims[0] == "iline" else "vertical"
    if direction == "xline":
        return "vertical" if dims[1] == "xline" else "horizontal"
    if direction == "twt" and shape is not None:
        return "vertical" if shape[0] > shape[1] else "horizontal"
    return direction


def footprint_filter(shape, sigma=7, direction="both", buffer_center=0.25, buffer_filter=3, dims=("iline", "xline")):
    """Centred (fftshift-ordered) notch filter of ``remove_acquisition_footprint`` for slices of ``shape``
    (cube_postprocessing_3D.py:218-252): 1 everywhere except along the kx / ky axes away from the centre."""
    ny, nx = shape
    npad = sigma * 5
    nyp, nxp = ny + npad, nx + npad
    stencil = np.zeros((nyp, nxp), dtype="int8")
    direction = _orientation(direction, dims, shape)
    if direction in ("both", "horizontal"):
        c = nxp // 2 + 1
        w = round(nyp * (1 - buffer_center) + .5) // 2
        stencil[:w, c - buffer_filter: c + buffer_filter + 1] = 1
        stencil[-w:, c - buffer_filter: c + buffer_filter + 1] = 1
    if direction in ("both", "vertical"):
        c = nyp // 2 + 1
        w = round(nxp * (1 - buffer_center) + .5) // 2
        stencil[c - buffer_filter: c + buffer_filter + 1, :w] = 1
        stencil[c - buffer_filter: c + buffer_filter + 1, -w:] = 1
    smooth = _signal.fftconvolve(stencil, gaussian_kernel_2d(sigma=sigma), mode="same")
    return 1 - rescale(smooth[npad // 2: -npad // 2, npad // 2: -npad // 2])


def antialiasing_filter(shape, direction, factors_upsampling, sigma=7, dims=("iline", "xline")):
    """Centred low-pass of ``spatial_antialiasing`` (cube_postprocessing_3D.py:303-340)."""
    il, xl = dims
    if sorted(dims) != sorted(factors_upsampling.keys()):
        raise ValueError(f"Coordinates {dims} not found in `factors_upsampling` {factors_upsampling.keys()}")
    ny, nx = shape
    npad = sigma * 5
    shrink = 0.98
    stencil = np.zeros((ny + npad, nx + npad), dtype="int8")
    direction = _orientation(direction, dims)
    if direction == "horizontal":
        perc = 1 - factors_upsampling.get(xl, 1) / factors_upsampling.get(il, 1)
        hw = round(ny * perc * shrink) // 2 + npad
        stencil[hw:-hw, :] = 1
    elif direction == "vertical":
        perc = 1 - factors_upsampling.get(il, 1) / factors_upsampling.get(xl, 1)
        hw = round(nx * perc * shrink) // 2 + npad
        stencil[:, hw:-hw] = 1
    smooth = _signal.fftconvolve(stencil, gaussian_kernel_2d(sigma=sigma), mode="same")
    return rescale(smooth[npad // 2: -npad // 2, npad // 2: -npad // 2], vmin=1e-3, vmax=1)


def apply_kxky_filter(data, ffilter, device=0):
    """``ifft2(ifftshift(ffilter) * fft2(data)).real`` for one slice (ny, nx) or a stack (n, ny, nx), on the GPU."""
    data = np.asarray(data)
    ny, nx = data.shape[-2:]
    plan = get_plan(ny, nx, device=device, precision=32)
    y = plan.kxky_filter(data.astype(np.complex64, copy=False), np.fft.ifftshift(np.asarray(ffilter)).astype(np.float32))
    out = y.real
    return out.astype(np.float64) if data.dtype in (np.float64, np.complex128) else np.ascontiguousarray(out)


def remove_acquisition_footprint(data, sigma: int = 7, direction: str = "both", buffer_center: float = 0.25,
                                 buffer_filter: int = 3, return_filter: bool = False, dims: tuple = ("iline", "xline"),
                                 verbose: int = 1, device: int = 0):
    """Remove the acquisition footprint from a slice (or a stack of slices) in the kx-ky domain."""
    data = np.asarray(data)
    ffilter = footprint_filter(data.shape[-2:], sigma, direction, buffer_center, buffer_filter, dims)
    data_filt = apply_kxky_filter(data, ffilter, device=device)
    return (data_filt, ffilter) if return_filter else data_filt


def spatial_antialiasing(data, direction: str, factors_upsampling: dict, sigma: int = 7, dims: tuple = ("iline", "xline"),
                         return_filter: bool = False, verbose: int = 1, device: int = 0):
    """Spatial de-aliasing in the kx-ky domain after iline / xline upsampling."""
    data = np.asarray(data)
    ffilter = antialiasing_filter(data.shape[-2:], direction, factors_upsampling, sigma, dims)
    data_filt = apply_kxky_filter(data, ffilter, device=device)
    return (data_filt, ffilter) if return_filter else data_filt
