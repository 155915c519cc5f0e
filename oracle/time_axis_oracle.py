"""CPU oracle for the time-axis transforms of steps 12 and 14.  TEST INFRASTRUCTURE ONLY.

Parity status: **parity unpinned**.  The arithmetic of these two steps lives in a
third-party dependency that is not vendored and not pinned by the reference:
``xrft @ git+https://github.com/fwrnke/xrft.git`` (requirements.txt:14, setup.cfg:42),
called at cube_apply_FFT.py:240-254 and cube_apply_IFFT.py:83-94.  The package is not
installed here and its source is not mounted, and the reference holds no test or golden
vector for this boundary.  What is restated below is xrft's published algorithm for
``fft(..., shift=False, true_phase=True, true_amplitude=True)`` and
``ifft(..., shift=True, true_phase=True, true_amplitude=True)``:

    forward :  F[k] = dt * exp(-2*pi*i*f_k*t0) * sum_n x[n] exp(-2*pi*i*k*n/N)
    inverse :  x[n] = (1/dt) * IDFT( F[k] * exp(+2*pi*i*f_k*t0) )[n]   (real part, float32)

with ``t0 = twt[0]``, ``dt = twt[1]-twt[0]``, ``f_k = fftfreq(N, dt)`` (or ``rfftfreq``
with ``--compute_real``).  The checkable properties are the round trip
``inverse(forward(x)) == x`` to float32 rounding, agreement with ``numpy.fft`` and the
``(twt, iline, xline)`` layout conventions of the two scripts.

Reference lines restated (relative to /root/reference/pseudo_3D_interpolation):
  cube_apply_FFT.py:223-233   odd number of samples -> drop the last one
  cube_apply_FFT.py:236,250   ``shape=`` zero padding to ``upsampling_factor * nt``
  cube_apply_FFT.py:240-254   forward transform, cast to complex64
  cube_apply_FFT.py:49-143    Hann-tapered frequency window (lowpass/highpass/bandpass)
  cube_apply_FFT.py:146-181,281-286   mask of kept slices for ``--drop-filtered-freq``
  cube_apply_IFFT.py:73-79    complex = real + 1j*imag
  cube_apply_IFFT.py:83-94    inverse transform (input ascending in frequency), float32
  cube_apply_IFFT.py:121-140, functions/utils.py:444-473   clip < 0 and global rescale
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "freq_axis", "time_fft", "time_ifft", "freq_filter_window", "freq_filter_keep",
    "rescale_envelope",
]


def freq_axis(nfft: int, dt: float, compute_real: bool, ascending: bool = False):
    if compute_real:
        return np.fft.rfftfreq(nfft, dt)
    f = np.fft.fftfreq(nfft, dt)
    return np.fft.fftshift(f) if ascending else f


def time_fft(x, twt, compute_real=False, upsampling_factor=1, window=None):
    """``x`` (nt, n_il, n_xl) float -> (nf, n_il, n_xl) complex64 in fftfreq/rfftfreq order."""
    x = np.asarray(x)
    twt = np.asarray(twt, dtype=np.float64)
    nt = x.shape[0]
    if nt % 2:                       # cube_apply_FFT.py:223-233
        x, twt, nt = x[:-1], twt[:-1], nt - 1
    dt = float(twt[1] - twt[0])
    t0 = float(twt[0])
    nfft = int(upsampling_factor) * nt
    f = freq_axis(nfft, dt, compute_real)
    xd = x.astype(np.float64)
    F = (np.fft.rfft if compute_real else np.fft.fft)(xd, n=nfft, axis=0)
    phase = np.exp(-2j * np.pi * f * t0) * dt
    F = F * phase.reshape((-1,) + (1,) * (x.ndim - 1))
    F = F.astype(np.complex64)
    if window is not None:           # cube_apply_FFT.py:273-278
        w = np.asarray(window, dtype=np.float64).reshape((-1,) + (1,) * (x.ndim - 1))
        F = (F * w).astype(np.complex64)
    return F, f


def time_ifft(F, dt, t0, compute_real=False, ascending=True, nfft=None):
    """Inverse of :func:`time_fft`.  ``F`` (nf, n_il, n_xl) complex.

    ``ascending=True`` means the frequency axis is sorted ascending (fftshift order),
    which is what step 13's merged output looks like when the full spectrum is used
    (SURVEY 3.1 footnote); with ``compute_real`` the axis is rfftfreq and already ascending.
    """
    F = np.asarray(F).astype(np.complex128)
    if compute_real:
        nfft = 2 * (F.shape[0] - 1) if nfft is None else nfft
        f = np.fft.rfftfreq(nfft, dt)
    else:
        nfft = F.shape[0]
        if ascending:
            F = np.fft.ifftshift(F, axes=0)
        f = np.fft.fftfreq(nfft, dt)
    phase = np.exp(2j * np.pi * f * t0) / dt
    G = F * phase.reshape((-1,) + (1,) * (F.ndim - 1))
    if compute_real:
        x = np.fft.irfft(G, n=nfft, axis=0)
    else:
        x = np.fft.ifft(G, axis=0).real
    return x.astype(np.float32)


# ---- frequency-domain taper (cube_apply_FFT.py:49-143) -------------------------------------
def _stopband(nstop: int, kind: str):
    size = nstop * 2
    size += 1 if size % 2 == 0 else 0
    sl = slice(1, size // 2 + 1) if kind == "highpass" else slice(size // 2, -1)
    return np.hanning(size)[sl]


def freq_filter_window(filter_freqs, frequencies, filter_type="lowpass"):
    frequencies = np.asarray(frequencies)
    if filter_type in ("lowpass", "highpass"):
        fmin, fmax = min(filter_freqs), max(filter_freqs)
        const = (0, 1) if filter_type == "highpass" else (1, 0)
        n_lower = np.count_nonzero(frequencies < fmin)
        n_stop = np.count_nonzero((frequencies >= fmin) & (frequencies <= fmax))
        n_higher = np.count_nonzero(frequencies > fmax)
        stop = _stopband(n_stop, filter_type)
    elif filter_type == "bandpass":
        f1, f2, f3, f4 = sorted(filter_freqs)
        const = (0, 0)
        n_lower = np.count_nonzero(frequencies < f1)
        n_lo = np.count_nonzero((frequencies >= f1) & (frequencies <= f2))
        n_mid = np.count_nonzero((frequencies > f2) & (frequencies < f3))
        n_hi = np.count_nonzero((frequencies >= f3) & (frequencies <= f4))
        n_higher = np.count_nonzero(frequencies > f4)
        stop = np.hstack((_stopband(n_lo, "highpass"), np.ones((n_mid,)), _stopband(n_hi, "lowpass")))
    else:
        raise ValueError(filter_type)
    return np.pad(stop, pad_width=(n_lower, n_higher), mode="constant", constant_values=(const,))


def freq_filter_keep(frequencies, freqs, filter_type="lowpass"):
    frequencies = np.asarray(frequencies)
    ff = sorted(freqs)
    if filter_type == "lowpass":
        return frequencies <= ff[-1]
    if filter_type == "highpass":
        return frequencies >= ff[0]
    return np.logical_and(frequencies >= ff[0], frequencies <= ff[-1])


def rescale_envelope(x):
    """clip < 0 then global min/max rescale to [0, 1] (cube_apply_IFFT.py:121-140)."""
    x = np.where(x < 0, 0, x)
    amin, amax = x.min(), x.max()
    if amin == amax:
        return x
    return 0 + (x - amin) * ((1 - 0) / (amax - amin))


def envelope(x, axis=-1):
    """Amplitude envelope |x + i Hilbert(x)| along ``axis`` in float64: the reference's ``functions/signal.py:672-690``
    (``scipy.signal.hilbert`` = one-sided spectrum weights 1, 2, ..., 2, (1 at Nyquist for even N), 0, ... then
    ``abs``), restated with numpy FFTs.  Pinned by tests/golden/reference_envelope.npz (oracle/make_golden_envelope.py)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[axis]
    h = np.zeros(n)
    if n % 2 == 0:
        h[0] = h[n // 2] = 1.0
        h[1:n // 2] = 2.0
    else:
        h[0] = 1.0
        h[1:(n + 1) // 2] = 2.0
    shape = [1] * x.ndim
    shape[axis] = n
    return np.abs(np.fft.ifft(np.fft.fft(x, axis=axis) * h.reshape(shape), axis=axis))
