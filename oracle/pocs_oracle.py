"""CPU oracle for the FFT-POCS slice algorithm.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement (complex128 arithmetic unless the caller passes
complex64) of what the reference computes for ``transform_kind='FFT'``.  It is written
in an explicit, branch-per-quirk form instead of the reference's call structure, and it
is the checker that the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; the product package never does.

Parity status: PINNED.  ``oracle/validate_against_reference.py`` runs this file
against the reference's own ``POCS_algorithm`` / ``get_threshold_decay`` / ``threshold``
imported from ``/root/reference`` (bit-for-bit equality is required there), and
``oracle/make_golden.py`` stores outputs *of the reference itself* under
``tests/golden/`` so the pin travels to machines without ``/root/reference``.

Reference lines restated here (paths relative to /root/reference/pseudo_3D_interpolation):
  functions/POCS.py:286-288      complex ``x_fwd.max()`` -> lexicographic maximum
  functions/POCS.py:296-299,326  tau_min (``p_min`` float or 'adaptive')
  functions/POCS.py:327-331      tau_max, decay_kind='factors'
  functions/POCS.py:336-343      iteration multiplier (i-1)/(niter-1)
  functions/POCS.py:348-362      linear / exponential[-q] / data-driven schedules
  functions/POCS.py:251-274      inverse-proportional schedule
  functions/POCS.py:515-521      all-zero slice shortcut
  functions/POCS.py:535-632      the iteration loop, cost, early exit
  functions/POCS.py:653-656      complex in -> complex out, real in -> real part
  functions/threshold_operator.py:20-56,59-95,98-123   soft / garrote / hard
  cube_POCS_interpolation_3D.py:242-244   mask = min(fold, 1)
"""
from __future__ import annotations

import math
import numpy as np

__all__ = [
    "lexmax", "threshold_table", "apply_threshold", "pocs_slice", "pocs_cube",
    "mask_from_fold",
]


# --------------------------------------------------------------------------------------
# complex ordering helpers (numpy orders complex numbers lexicographically: real, then imag)
# --------------------------------------------------------------------------------------
def lexmax(X: np.ndarray) -> complex:
    """Element of ``X`` with the largest real part; ties broken by the largest imag part.

    Restates ``x_fwd.max()`` on a complex array (functions/POCS.py:288).
    """
    X = np.asarray(X).ravel()
    if not np.iscomplexobj(X):
        return X.max()
    re = X.real
    m = re.max()
    cand = X[re == m]
    return cand[np.argmax(cand.imag)]


def _lex_less(a_re, a_im, b_re, b_im):
    """a < b in numpy's complex ordering."""
    return (a_re < b_re) | ((a_re == b_re) & (a_im < b_im))


def _parse_q(thresh_model: str) -> float:
    if "-" in thresh_model:
        try:
            return float(thresh_model.split("-")[-1])
        except ValueError:
            return 1.0
    return 1.0


def threshold_table(X0, niter, thresh_model="exponential", p_max=0.99, p_min=1e-5,
                    decay_kind="values"):
    """Per-slice threshold schedule tau[0..niter-1] from the initial spectrum ``X0``.

    Returns a complex array (linear / exponential / data-driven) or a float array
    (inverse-proportional), exactly like ``get_threshold_decay(..., 'FFT', ...)``.
    """
    niter = int(niter)
    it = np.arange(1, niter + 1)

    if "inverse" in thresh_model and "proportional" in thresh_model:
        r = np.abs(X0)
        vmax, vmin = r.max(), r.min()
        q = _parse_q(thresh_model)
        a = (niter ** q * (vmax - vmin)) / (niter ** q - 1)
        b = (niter ** q * vmin - vmax) / (niter ** q - 1)
        return a / (it ** q) + b

    if decay_kind == "values":
        z = lexmax(X0)
        if isinstance(p_min, str) and p_min == "adaptive":
            # 0.01 * rms(|X0|); real-valued, so ln(tau_min / tau_max) below is complex
            nrm = np.linalg.norm(np.asarray(X0), axis=None)
            tau_min = 0.01 * np.sqrt(nrm ** 2 / np.asarray(X0).size)
        else:
            tau_min = p_min * z
        tau_max = p_max * z
    elif decay_kind == "factors":
        tau_max, tau_min = p_max, p_min
    else:
        raise ValueError('Parameter `kind` only supports arguments "values" or "factors"')

    mu = (it - 1) / (niter - 1)

    if thresh_model == "linear":
        return tau_max - (tau_max - tau_min) * mu
    if "exponential" in thresh_model:
        q = _parse_q(thresh_model)
        c = np.log(tau_min / tau_max)
        return tau_max * np.exp(c * mu ** q)
    if thresh_model == "data-driven":
        X0 = np.asarray(X0)
        tmn, tmx = complex(tau_min), complex(tau_max)
        sel = _lex_less(tmn.real, tmn.imag, X0.real, X0.imag) & \
            _lex_less(X0.real, X0.imag, tmx.real, tmx.imag)
        v = X0[sel]
        order = np.lexsort((v.imag, v.real))[::-1]      # descending lexicographic
        v = v[order]
        nv = v.size
        tau = np.zeros((niter,), dtype=X0.dtype)
        tau[0] = v[0]
        tau[1:] = v[np.ceil((it[1:] - 1) * (nv - 1) / (niter - 1)).astype("int")]
        return tau
    raise NotImplementedError(f"{thresh_model} is not implemented for FFT transform!")


def apply_threshold(X, tau, kind="hard"):
    """Threshold operator with the reference's complex-``tau`` semantics."""
    X = np.asarray(X)
    # ``tau`` is used as the numpy scalar it is in the reference (an element of the schedule
    # array): numpy >= 2 treats numpy scalars as strongly typed, so a complex128 tau promotes a
    # complex64 spectrum to complex128 in the soft / garrote branches exactly as the reference
    # does, and a real tau uses real division (a / r) where a complex tau uses complex division.
    if kind.endswith("-percentile"):
        # '<op>-percentile' wrappers (functions/POCS.py:43-58): the scheduled value is a percentile of |X|
        tau = np.percentile(np.abs(X), tau)
        kind = kind[: -len("-percentile")]
    a, b = np.real(tau), np.imag(tau)
    r = np.abs(X)
    if kind == "hard":
        kill = (r < a) | ((r == a) & (0.0 < b))
        return np.where(kill, 0, X)
    # The shrink factor is computed with the same numpy complex expressions the reference
    # uses (so rounding is identical); the ordering semantics of ``clip(min=0)`` on a
    # complex array are then applied explicitly.
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == "soft":
            f = 1 - tau / r
        elif kind in ("garrote", "garotte"):
            f = 1 - tau ** 2 / r ** 2
        else:
            raise NotImplementedError(kind)
        # lexicographic clip: f := 0 where f < 0, i.e. Re f < 0 or (Re f == 0 and Im f < 0).
        # |X| == 0 gives Re f = -inf -> 0.  (If additionally Im(tau) == 0 exactly while tau is
        # complex, numpy yields (-inf, nan) and the reference propagates nan; that
        # measure-zero corner is not reproduced: the factor is 0 here and on the GPU.)
        zero = (np.real(f) < 0) | ((np.real(f) == 0) & (np.imag(f) < 0)) | (r == 0)
        f = np.where(zero, 0, f)
        return X * f


def mask_from_fold(fold):
    """mask = fold where fold <= 1 else 1 (cube_POCS_interpolation_3D.py:242-244)."""
    fold = np.asarray(fold)
    return np.minimum(fold, 1).astype(fold.dtype)


def pocs_slice(x, mask, niter=50, thresh_op="hard", thresh_model="exponential", eps=1e-9,
               alpha=1.0, p_max=0.99, p_min=1e-5, sqrt_decay=False, decay_kind="values",
               version="regular", info=None):
    """One slice of FFT-POCS.  ``x`` (N1,N2) real or complex, ``mask`` (N1,N2) in {0,1}."""
    x = np.asarray(x)
    mask = np.asarray(mask)
    if mask.max() > 1:
        raise ValueError(f"mask should be quasi-boolean (0 or 1) but has maximum of {mask.max()}")
    niter, eps, p_max, alpha = int(niter), float(eps), float(p_max), float(alpha)
    is_complex = np.iscomplexobj(x)
    costs = []
    if np.count_nonzero(x) == 0:
        if info is not None:
            info.update(niterations=0, cost=0, costs=[0])
        return x

    X0 = np.fft.fft2(x)
    tau = threshold_table(X0, niter, thresh_model, p_max, p_min, decay_kind)
    keep = 1 - alpha * mask
    v_fast = 1
    x_prev = x
    k_done = 0
    for k in range(niter):
        if version == "regular":
            x_in = x_prev
        elif version == "fast":
            # x_old aliases x_inv in the reference (functions/POCS.py:629), so the momentum term
            # is frac * 0; kept literally because frac (float64) promotes the dtype
            v1 = (1 + np.sqrt(1 + 4 * v_fast ** 2)) / 2
            frac = (v_fast - 1) / (v1 + 1)
            v_fast = v1
            x_in = x_prev + frac * (x_prev - x_prev)
        elif version == "adaptive":
            x_tmp = alpha * x + keep * x_prev
            x_in = x_tmp + (1 - alpha) * (x - mask * x_prev)
        else:
            raise ValueError(version)
        X = np.fft.fft2(x_in)
        t = np.sqrt(tau[k]) if sqrt_decay else tau[k]
        Y = apply_threshold(X, t, thresh_op)
        y = np.fft.ifft2(Y)
        # in place, like the reference (functions/POCS.py:616-619): a complex64 iterate stays
        # complex64 under numpy >= 2 (the production path), a complex128 one stays complex128
        y *= keep
        y += x * alpha
        x_new = y
        # cost: reference sums (|x_k| - |x_{k-1}|) element-wise first (functions/POCS.py:622)
        cost = np.sum(np.abs(x_new) - np.abs(x_prev)) ** 2 / np.sum(np.abs(x_new)) ** 2
        costs.append(float(cost))
        x_prev = x_new
        k_done = k + 1
        if k > 2 and cost < eps:
            break
    if info is not None:
        info.update(niterations=k_done, cost=costs[-1], costs=costs)
    return x_prev if is_complex else np.real(x_prev)


def pocs_cube(cube, fold_or_mask, upcast=True, infos=None, **params):
    """Loop ``pocs_slice`` over axis 0 of ``cube`` (slices are independent).

    ``upcast=True`` computes in complex128/float64 (the "float64 reference" of
    BASELINE.json) and casts the result back to the input dtype, which is what
    ``np.vectorize(otypes=[cube.dtype])`` does in the reference driver
    (cube_POCS_interpolation_3D.py:314-336).
    """
    cube = np.asarray(cube)
    mask = mask_from_fold(fold_or_mask)
    out = np.empty_like(cube)
    for s in range(cube.shape[0]):
        xs = cube[s]
        if upcast:
            xs = xs.astype(np.complex128 if np.iscomplexobj(xs) else np.float64)
        info = {} if infos is not None else None
        out[s] = pocs_slice(xs, mask, info=info, **params)
        if infos is not None:
            infos.append(info)
    return out
