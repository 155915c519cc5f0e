"""pseudo-3d-interpolation_b200: B200-native FFT-POCS hot path of fwrnke/pseudo-3D-interpolation.

Importable as ``pseudo_3d_interpolation_b200`` (see the shim package of that name at the
repository root; a directory name with hyphens cannot be imported directly).

Public surface (mirrors the reference for the one accelerated path):
  pocs.POCS_algorithm / POCS / FPOCS / APOCS / get_threshold_decay / threshold / pocs_cube
  timeaxis.time_fft / time_ifft / freq_filter_window
  cube_postprocessing_3D.remove_acquisition_footprint / spatial_antialiasing  (kx-ky domain filters)
  cube_POCS_interpolation_3D.main, cube_apply_FFT.main, cube_apply_IFFT.main  (CLI mirrors)
"""
from . import _lib                                   # noqa: F401
from .pocs import (POCS_algorithm, POCS, FPOCS, APOCS, get_threshold_decay, threshold, pocs_cube,  # noqa: F401
                   PocsPlan, make_params, mask_from_fold, fft2, ifft2, band_bounds, set_default_precision, get_plan)

__version__ = "0.1.0"
