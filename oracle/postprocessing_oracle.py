"""CPU restatement (numpy / scipy, float64) of the reference's kx-ky domain filters.  TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Follows /root/reference/pseudo_3D_interpolation/cube_postprocessing_3D.py:
  kernel            gaussian_kernel_2d                  :131-176
  footprint         remove_acquisition_footprint        :179-260   (arithmetic :254)
  antialias         spatial_antialiasing                :263-347   (arithmetic :342)
and functions/utils.py:413-441 (rescale).  Pinned: tests/golden/reference_postprocessing.npz holds outputs of the
reference's own functions (oracle/make_golden_postprocessing.py imports them from /root/reference).
"""
import numpy as np
from scipy.signal import fftconvolve
from scipy.signal.windows import gaussian


def _rescale(a, vmin=0.0, vmax=1.0):
    lo, hi = np.nanmin(a), np.nanmax(a)
    return a if lo == hi else vmin + (a - lo) * ((vmax - vmin) / (hi - lo))


def kernel(sigma=7):
    n = sigma * 8 + 1
    n += (n % 2 == 0)
    g = gaussian(n, sigma)
    return np.outer(g, g) / (2 * np.pi * sigma ** 2)


def _smoothed(stencil, sigma):
    npad = sigma * 5
    s = fftconvolve(stencil, kernel(sigma), mode="same")
    return s[npad // 2: -npad // 2, npad // 2: -npad // 2]


def footprint_filter(shape, sigma=7, direction="both", buffer_center=0.25, buffer_filter=3):
    """direction: 'both' | 'horizontal' (reference 'iline' with dims=('iline','xline')) | 'vertical' ('xline')."""
    ny, nx = shape
    nyp, nxp = ny + sigma * 5, nx + sigma * 5
    st = np.zeros((nyp, nxp), dtype=np.int8)
    b = buffer_filter
    if direction in ("both", "horizontal"):
        c, w = nxp // 2 + 1, round(nyp * (1 - buffer_center) + .5) // 2
        st[:w, c - b:c + b + 1] = 1
        st[-w:, c - b:c + b + 1] = 1
    if direction in ("both", "vertical"):
        c, w = nyp // 2 + 1, round(nxp * (1 - buffer_center) + .5) // 2
        st[c - b:c + b + 1, :w] = 1
        st[c - b:c + b + 1, -w:] = 1
    return 1 - _rescale(_smoothed(st, sigma))


def antialias_filter(shape, direction, f_il, f_xl, sigma=7):
    """direction: 'horizontal' (reference 'iline') | 'vertical' ('xline'); f_il / f_xl = upsampling factors."""
    ny, nx = shape
    npad = sigma * 5
    st = np.zeros((ny + npad, nx + npad), dtype=np.int8)
    if direction == "horizontal":
        hw = round(ny * (1 - f_xl / f_il) * 0.98) // 2 + npad
        st[hw:-hw, :] = 1
    else:
        hw = round(nx * (1 - f_il / f_xl) * 0.98) // 2 + npad
        st[:, hw:-hw] = 1
    return _rescale(_smoothed(st, sigma), 1e-3, 1.0)


def apply(data, ffilter):
    """ifft2(ifftshift(filter) * fft2(data)).real over the last two axes (cube_postprocessing_3D.py:254,342)."""
    return np.fft.ifft2(np.fft.ifftshift(ffilter) * np.fft.fft2(data)).real
