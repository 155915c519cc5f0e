"""Steps 12 -> 13 -> 14 chained on one GPU without leaving device memory.

The reference runs three scripts with a netCDF file between each pair
(cube_apply_FFT.py:316-319 -> cube_POCS_interpolation_3D.py:231-233, 370-405 ->
cube_apply_IFFT.py:53-57); the complex spectrum makes two host/disk round trips and is split
into ``.real`` / ``.imag`` and recombined on the way.  Here the time cube is uploaded once, the
slice-major spectrum written by ``p3d_time_fft`` is exactly the ``(n_slices, n_iline, n_xline)``
layout ``p3d_pocs_run`` consumes, and its result feeds ``p3d_time_ifft`` in place; only the
interpolated time cube comes back.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .pocs import get_plan, make_params, mask_from_fold


def interpolate_time_cube(x, twt, fold, compute_real=True, device=0, precision=None, window=None,
                          return_spectrum=False, results=None, **metadata):
    """``x`` (nt, n_il, n_xl) float32 sparse time cube, ``twt`` (nt,), ``fold`` (n_il, n_xl).

    Returns the interpolated time cube (nt_even, n_il, n_xl) float32 (and, optionally, the
    interpolated spectrum in rfftfreq / fftfreq order).  ``metadata`` are the POCS keywords
    (niter, eps, thresh_op, thresh_model, alpha, p_max, p_min, ...).
    """
    lib = _lib.load()
    _lib.require_gpu()
    for k in ("transform", "itransform", "transform_kind", "auxiliary_data", "verbose", "results_dict", "path_results"):
        metadata.pop(k, None)
    params = make_params(**metadata)
    x = np.asarray(x)
    twt = np.asarray(twt, dtype=np.float64)
    if x.shape[0] % 2:
        x, twt = x[:-1], twt[:-1]
    x = np.ascontiguousarray(x, dtype=np.float32)
    nt, n1, n2 = x.shape
    ntr = n1 * n2
    dt, t0 = float(twt[1] - twt[0]), float(twt[0])
    nf = nt // 2 + 1 if compute_real else nt
    mask = np.ascontiguousarray(mask_from_fold(fold), dtype=np.uint8)
    win = None if window is None else np.ascontiguousarray(window, dtype=np.float64)

    d_x = _lib.DeviceBuffer(x.nbytes, device)
    d_f = _lib.DeviceBuffer(nf * ntr * 8, device)
    d_y = _lib.DeviceBuffer(nf * ntr * 8, device)
    d_m = _lib.DeviceBuffer(mask.nbytes, device)
    try:
        d_x.upload(x)
        d_m.upload(mask)
        _lib.check(lib.p3d_time_fft(device, C.c_void_p(d_x.ptr), _lib.MEM_DEVICE, C.c_void_p(d_f.ptr), _lib.MEM_DEVICE,
                                    nt, nt, ntr, dt, t0, 1 if compute_real else 0, _lib.ptr(win)))
        plan = get_plan(n1, n2, device, precision)
        nit = np.zeros(nf, np.int32)
        cost = np.zeros(nf, np.float64)
        plan.run_device(d_f.ptr, d_m.ptr, d_y.ptr, nf, params, nit=nit, cost=cost)
        # the POCS output keeps the order of its input (fftfreq / rfftfreq), so ascending = 0 here
        _lib.check(lib.p3d_time_ifft(device, C.c_void_p(d_y.ptr), _lib.MEM_DEVICE, C.c_void_p(d_x.ptr), _lib.MEM_DEVICE,
                                     nt, nt, ntr, dt, t0, 1 if compute_real else 0, 0))
        out = np.empty_like(x)
        d_x.download(out)
        if isinstance(results, dict):
            results["niterations"] = nit
            results["cost"] = cost
        if return_spectrum:
            spec = np.empty((nf, n1, n2), dtype=np.complex64)
            d_y.download(spec)
            return out, spec
        return out
    finally:
        for b in (d_x, d_f, d_y, d_m):
            b.free()
