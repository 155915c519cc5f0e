"""Time-axis forward / inverse transforms of steps 12 and 14 on the GPU (ndarray level).

``time_fft`` restates the arithmetic of ``xrft.fft(dim, real_dim?, shift=False, true_phase=True,
true_amplitude=True)`` at cube_apply_FFT.py:240-254 and ``time_ifft`` that of
``xrft.ifft(shift=True, true_phase=True, true_amplitude=True)`` at cube_apply_IFFT.py:83-94
(closed forms in SURVEY.md 8c).  Arrays are time-major ``(twt, iline, xline)`` /
``(freq, iline, xline)`` like the reference's netCDF variables.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__all__ = ["freq_axis", "time_fft", "time_ifft", "freq_filter_window", "freq_filter_keep", "rescale_envelope", "envelope"]


def freq_axis(nfft, dt, compute_real, ascending=False):
    """Frequency coordinate written by step 12 (fftfreq / rfftfreq order) or, with
    ``ascending``, the sorted axis step 13's merge produces."""
    if compute_real:
        return np.fft.rfftfreq(nfft, dt)
    f = np.fft.fftfreq(nfft, dt)
    return np.fft.fftshift(f) if ascending else f


def time_fft(x, twt, compute_real=False, upsampling_factor=1, window=None, device=0):
    """(nt, n_il, n_xl) float32 -> ((nf, n_il, n_xl) complex64, freq axis).

    An odd number of samples drops the last one (cube_apply_FFT.py:223-233);
    ``upsampling_factor`` zero-pads to ``factor * nt`` (the fork-only ``shape=`` argument,
    cube_apply_FFT.py:236,250); ``window`` (nf,) multiplies the spectrum (cube_apply_FFT.py:273-278).
    """
    _lib.require_gpu()
    x = np.asarray(x)
    twt = np.asarray(twt, dtype=np.float64)
    if x.shape[0] % 2:
        x, twt = x[:-1], twt[:-1]
    nt = x.shape[0]
    if nt < 2:
        raise ValueError("need at least two time samples")
    dt = float(twt[1] - twt[0])
    t0 = float(twt[0])
    nfft = int(upsampling_factor) * nt
    x = np.ascontiguousarray(x, dtype=np.float32)
    ntr = int(np.prod(x.shape[1:], dtype=np.int64))
    nf = nfft // 2 + 1 if compute_real else nfft
    out = np.empty((nf,) + x.shape[1:], dtype=np.complex64)
    win = None
    if window is not None:
        win = np.ascontiguousarray(window, dtype=np.float64)
        if win.shape != (nf,):
            raise ValueError(f"window must have shape ({nf},)")
    _lib.check(_lib.load().p3d_time_fft(device, _lib.ptr(x), _lib.MEM_HOST, _lib.ptr(out), _lib.MEM_HOST, nt, nfft, ntr,
                                        dt, t0, 1 if compute_real else 0, _lib.ptr(win)))
    return out, freq_axis(nfft, dt, compute_real)


def time_ifft(F, dt, t0, compute_real=False, ascending=True, nt_out=None, device=0):
    """(nf, n_il, n_xl) complex64 -> (nt, n_il, n_xl) float32 (real part, cube_apply_IFFT.py:92)."""
    _lib.require_gpu()
    F = np.ascontiguousarray(F, dtype=np.complex64)
    nf = F.shape[0]
    nfft = 2 * (nf - 1) if compute_real else nf
    if nfft % 2 or nfft < 2:
        raise ValueError("the frequency axis must come from an even-length transform")
    nt_out = nfft if nt_out is None else int(nt_out)
    ntr = int(np.prod(F.shape[1:], dtype=np.int64))
    out = np.empty((nt_out,) + F.shape[1:], dtype=np.float32)
    _lib.check(_lib.load().p3d_time_ifft(device, _lib.ptr(F), _lib.MEM_HOST, _lib.ptr(out), _lib.MEM_HOST, nfft, nt_out, ntr,
                                         float(dt), float(t0), 1 if compute_real else 0, 1 if ascending else 0))
    return out


# ---- frequency-domain taper of step 12 (host, O(nf)) -------------------------------------------------
def _taper(n_samples, rising):
    size = 2 * n_samples
    size += 1 if size % 2 == 0 else 0
    w = np.hanning(size)
    return w[1:size // 2 + 1] if rising else w[size // 2:-1]


def freq_filter_window(filter_freqs, frequencies, filter_type="lowpass"):
    """Hann-tapered pass window over the frequency axis (cube_apply_FFT.py:72-143)."""
    fr = np.asarray(frequencies)
    if filter_type in ("lowpass", "highpass"):
        lo, hi = min(filter_freqs), max(filter_freqs)
        below = np.count_nonzero(fr < lo)
        inside = np.count_nonzero((fr >= lo) & (fr <= hi))
        above = np.count_nonzero(fr > hi)
        ramp = _taper(inside, rising=(filter_type == "highpass"))
        edge = (0, 1) if filter_type == "highpass" else (1, 0)
    elif filter_type == "bandpass":
        f1, f2, f3, f4 = sorted(filter_freqs)
        below = np.count_nonzero(fr < f1)
        n_up = np.count_nonzero((fr >= f1) & (fr <= f2))
        n_flat = np.count_nonzero((fr > f2) & (fr < f3))
        n_down = np.count_nonzero((fr >= f3) & (fr <= f4))
        above = np.count_nonzero(fr > f4)
        ramp = np.concatenate([_taper(n_up, True), np.ones(n_flat), _taper(n_down, False)])
        edge = (0, 0)
    else:
        raise ValueError(f"unknown filter type {filter_type!r}")
    return np.pad(ramp, (below, above), mode="constant", constant_values=(edge,))


def freq_filter_keep(frequencies, freqs, filter_type="lowpass"):
    """Boolean mask of frequency slices kept by ``--drop-filtered-freq`` (cube_apply_FFT.py:146-181)."""
    fr = np.asarray(frequencies)
    ff = sorted(freqs)
    if filter_type == "lowpass":
        return fr <= ff[-1]
    if filter_type == "highpass":
        return fr >= ff[0]
    if filter_type == "bandpass":
        return np.logical_and(fr >= ff[0], fr <= ff[-1])
    raise ValueError(f"unknown filter type {filter_type!r}")


def rescale_envelope(x):
    """clip below zero, then global min/max rescale to [0, 1] (cube_apply_IFFT.py:121-140)."""
    x = np.where(x < 0, 0, x)
    lo, hi = x.min(), x.max()
    if lo == hi:
        return x
    return (x - lo) * (1.0 / (hi - lo))


def envelope(signal, axis=-1, device=0):
    """Amplitude envelope of a trace (1D), section (2D) or cube (3D) along ``axis`` via the Hilbert transform, on
    the GPU: same signature and result dtype as the reference's ``functions/signal.py:672-690`` (step 11's ``env``
    variable, cube_preprocessing_3D.py:341-353).  The kernel works on a time-major ``(nt, n_traces)`` view, so
    ``axis=0`` of a ``(twt, iline, xline)`` cube needs no copy."""
    _lib.require_gpu()
    signal = np.asarray(signal)
    x = np.moveaxis(signal, axis, 0)
    lead = x.shape
    x2 = np.ascontiguousarray(x.reshape(lead[0], -1), dtype=np.float32)
    out = np.empty_like(x2)
    if x2.size:
        _lib.check(_lib.load().p3d_time_envelope(int(device), _lib.ptr(x2), _lib.MEM_HOST, _lib.ptr(out), _lib.MEM_HOST,
                                                 x2.shape[0], x2.shape[1]))
    return np.moveaxis(out.reshape(lead), 0, axis).astype(signal.dtype, copy=False)
