"""Host-side mirror of the reference's POCS interface, executing on the B200 through the C ABI.

Mirrors ``pseudo_3D_interpolation/functions/POCS.py`` for ``transform_kind='FFT'``:

* :func:`POCS_algorithm` and the partials :data:`POCS`, :data:`FPOCS`, :data:`APOCS`
  (functions/POCS.py:371-391, 659-661) -- same arguments, defaults, return conventions,
  ``results_dict`` / ``path_results`` side effects and ``ValueError`` / ``NotImplementedError``
  behaviour, but the work is done by ``p3d_pocs_run`` (include/p3d_b200.h);
* :func:`get_threshold_decay` (functions/POCS.py:169-368) and :func:`threshold`
  (functions/POCS.py:61-102, functions/threshold_operator.py) -- small host utilities kept
  for API compatibility; the device path never calls them (the schedule is derived from
  device-side statistics inside ``p3d_pocs_run``);
* :func:`pocs_cube` -- the slice loop of cube_POCS_interpolation_3D.py:303-340 for a whole
  ``(n_slices, n_iline, n_xline)`` array, sharded as contiguous bands over the listed GPUs.

There is no CPU fallback: without the shared library or a CUDA device these functions raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import time
from functools import partial

import numpy as np

from . import _lib

TRANSFORMS = ("FFT", "WAVELET", "SHEARLET", "CURVELET", "DCT")

__all__ = ["POCS_algorithm", "POCS", "FPOCS", "APOCS", "get_threshold_decay", "threshold", "pocs_cube",
           "PocsPlan", "make_params", "mask_from_fold", "fft2", "ifft2", "set_default_precision", "release_plans"]


# --------------------------------------------------------------------------------------------------
# parameter translation
# --------------------------------------------------------------------------------------------------
def _parse_model(thresh_model: str):
    """'exponential-2' -> (MODEL_EXPONENTIAL, 2.0) etc. (functions/POCS.py:251,262-267,351-352)."""
    tm = str(thresh_model)
    q = 1.0
    if all(s in tm for s in ("inverse", "proportional")):
        if "-" in tm:
            try:
                q = float(tm.split("-")[-1])
            except ValueError:
                q = 1.0
        return _lib.MODELS["inverse_proportional"], q
    if tm == "linear":
        return _lib.MODELS["linear"], q
    if "exponential" in tm:
        if "-" in tm:
            q = float(tm.split("-")[-1])
        return _lib.MODELS["exponential"], q
    if tm == "data-driven":
        return _lib.MODELS["data-driven"], q
    raise NotImplementedError(f"{thresh_model} is not implemented for FFT transform!")


def make_params(niter=50, thresh_op="hard", thresh_model="exponential", eps=1e-9, alpha=1.0, p_max=0.99,
                p_min=1e-5, sqrt_decay=False, decay_kind="values", version="regular",
                absmax_threshold=False) -> _lib.PocsParams:
    niter = int(niter)
    eps = float(eps)
    p_max = float(p_max)
    alpha = float(alpha)
    percentile = isinstance(thresh_op, str) and thresh_op.endswith("-percentile")
    if percentile:
        thresh_op = thresh_op[: -len("-percentile")]
    if thresh_op not in _lib.OPS:
        raise NotImplementedError(f"unknown threshold operator {thresh_op!r}")
    model, q = _parse_model(thresh_model)
    if decay_kind not in ("values", "factors"):
        raise ValueError('Parameter `kind` only supports arguments "values" or "factors"')
    if version not in _lib.VERSIONS:
        raise ValueError(f"unknown POCS version {version!r}")
    adaptive = isinstance(p_min, str) and p_min == "adaptive"
    if isinstance(p_min, str) and not adaptive:
        # the reference fails at ``p_min * x_fwd_max`` for a str p_min
        raise TypeError(f"p_min must be a float or 'adaptive' (got {p_min!r})")
    if adaptive and decay_kind == "factors":
        raise TypeError("p_min='adaptive' cannot be combined with decay_kind='factors'")
    p = _lib.PocsParams()
    p.niter, p.thresh_op, p.thresh_model, p.version = niter, _lib.OPS[thresh_op], model, _lib.VERSIONS[version]
    p.q, p.eps, p.alpha, p.p_max = q, eps, alpha, p_max
    p.p_min = 0.0 if adaptive else float(p_min)
    p.p_min_adaptive = 1 if adaptive else 0
    p.sqrt_decay = 1 if sqrt_decay else 0
    p.decay_factors = 1 if decay_kind == "factors" else 0
    p.absmax_threshold = 1 if absmax_threshold else 0
    p.thresh_percentile = 1 if percentile else 0
    if percentile:
        # '<op>-percentile' (functions/POCS.py:43-58): the scheduled value is handed to np.percentile(|X|, .), which
        # only works for real values in [0, 100], i.e. decay_kind='factors' with p_max / p_min given in percent
        if decay_kind != "factors" or adaptive or not (0.0 <= p_max <= 100.0 and 0.0 <= p.p_min <= 100.0):
            raise ValueError("Percentiles must be in the range [0, 100]")
        if model not in (_lib.MODELS["linear"], _lib.MODELS["exponential"]):
            raise NotImplementedError(f"{thresh_model} schedule with percentile operators is not implemented")
    return p


def mask_from_fold(fold):
    """mask = fold clipped to {0,1}, dtype preserved (cube_POCS_interpolation_3D.py:242-244)."""
    fold = np.asarray(fold)
    return np.minimum(fold, 1).astype(fold.dtype)


# --------------------------------------------------------------------------------------------------
# plan wrapper
# --------------------------------------------------------------------------------------------------
class PocsPlan:
    """One ``p3d_plan`` (one GPU, one slice shape).  Not re-entrant; use one per thread/GPU."""

    def __init__(self, n_iline, n_xline, device=0, max_slices=0, band_slices=0, precision=0):
        """``precision``: 0 / "auto" (default) = escalating: every slice iterates in fp32 until a coefficient of its
        spectrum comes within a guard band of the threshold and in complex128 from that iterate on (within 1e-4 of the
        float64 reference); 32 = fp32 only (fastest; threshold decisions can differ from float64 in the late
        iterations); 64 = float64 state mode (complex128 throughout, result rounded once to complex64)."""
        precision = _precision_code(precision)
        lib = _lib.load()
        _lib.require_gpu()
        self.n_iline, self.n_xline, self.device = int(n_iline), int(n_xline), int(device)
        h = C.c_void_p()
        _lib.check(lib.p3d_plan_create(C.byref(h), self.device, self.n_iline, self.n_xline, int(max_slices), int(band_slices)))
        self._h = h
        self._lock = threading.Lock()
        self.precision = int(precision)
        self.set_option("precision", self.precision)

    def escalation(self):
        """(slices that switched to complex128, slice-iterations they ran there) of the last run."""
        a, b = C.c_int64(), C.c_int64()
        _lib.check(_lib.load().p3d_plan_get_escalation(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().p3d_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- options / introspection -----------------------------------------------------------------
    def set_option(self, key, value):
        _lib.check(_lib.load().p3d_plan_set_option(self._h, key.encode(), int(value)))

    def describe(self) -> str:
        buf = C.create_string_buffer(2048)
        _lib.check(_lib.load().p3d_plan_describe(self._h, buf, 2048))
        return buf.value.decode()

    def set_profiling(self, on=True):
        _lib.check(_lib.load().p3d_plan_set_profiling(self._h, 1 if on else 0))

    def get_profile(self, reset=True):
        ms = (C.c_double * len(_lib.PROFILE_KINDS))()
        n = (C.c_int64 * len(_lib.PROFILE_KINDS))()
        _lib.check(_lib.load().p3d_plan_get_profile(self._h, ms, n, 1 if reset else 0))
        return {k: dict(ms=ms[i], launches=n[i]) for i, k in enumerate(_lib.PROFILE_KINDS)}

    def event_record(self, slot):
        _lib.check(_lib.load().p3d_plan_event_record(self._h, int(slot)))

    def event_elapsed_ms(self, a, b) -> float:
        ms = C.c_double()
        _lib.check(_lib.load().p3d_plan_event_elapsed_ms(self._h, int(a), int(b), C.byref(ms)))
        return ms.value

    # -- compute -----------------------------------------------------------------------------------
    def run(self, x, mask, out=None, params=None, slices_per_mask=None, want_costs=False, **kw):
        """POCS over ``x`` (n_slices, n_iline, n_xline) complex64 host array -> (out, info)."""
        if params is None:
            params = make_params(**kw)
        x = np.ascontiguousarray(x, dtype=np.complex64)
        if x.ndim == 2:
            x = x[None]
        ns = x.shape[0]
        if x.shape[1:] != (self.n_iline, self.n_xline):
            raise ValueError(f"slice shape {x.shape[1:]} does not match the plan ({self.n_iline}, {self.n_xline})")
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        if mask.ndim == 2:
            mask = mask[None]
        if mask.shape[1:] != (self.n_iline, self.n_xline):
            raise ValueError(f"mask shape {mask.shape[1:]} does not match the plan")
        if mask.size and mask.max() > 1:
            raise ValueError(f"mask should be quasi-boolean (0 or 1) but has maximum of {mask.max()}")
        spm = int(slices_per_mask) if slices_per_mask else max(ns, 1)
        if mask.shape[0] < (ns + spm - 1) // spm:
            raise ValueError("not enough masks for the given slices_per_mask")
        if out is None:
            out = np.empty_like(x)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.complex64 and out.shape == x.shape and out.flags.c_contiguous
                  and out.flags.writeable):
            raise ValueError("out must be a writeable C-contiguous complex64 array of the shape of x")
        nit = np.zeros(ns, dtype=np.int32)
        cost = np.zeros(ns, dtype=np.float64)
        costs = np.full((ns, params.niter), np.nan, dtype=np.float64) if want_costs else None
        with self._lock:
            _lib.check(_lib.load().p3d_pocs_run(self._h, C.byref(params), _lib.ptr(x), _lib.MEM_HOST, _lib.ptr(mask), spm,
                                                _lib.ptr(out), _lib.MEM_HOST, ns, _lib.ptr(nit), _lib.ptr(cost),
                                                _lib.ptr(costs) if costs is not None else None))
        return out, dict(niterations=nit, cost=cost, costs=costs)

    def run_device(self, x_ptr, mask_ptr, out_ptr, n_slices, params, slices_per_mask=None, nit=None, cost=None):
        """Device-resident variant: raw device pointers (ints), nothing copied."""
        spm = int(slices_per_mask) if slices_per_mask else max(int(n_slices), 1)
        with self._lock:
            _lib.check(_lib.load().p3d_pocs_run(self._h, C.byref(params), C.c_void_p(int(x_ptr)), _lib.MEM_DEVICE,
                                                C.c_void_p(int(mask_ptr)), spm, C.c_void_p(int(out_ptr)), _lib.MEM_DEVICE,
                                                int(n_slices), _lib.ptr(nit), _lib.ptr(cost), None))

    def schedule(self, x, params=None, **kw):
        """Threshold schedule tau (n_slices, niter) complex128 computed like the device path does."""
        if params is None:
            params = make_params(**kw)
        x = np.ascontiguousarray(x, dtype=np.complex64)
        if x.ndim == 2:
            x = x[None]
        tau = np.zeros((x.shape[0], params.niter, 2), dtype=np.float64)
        with self._lock:
            _lib.check(_lib.load().p3d_pocs_schedule(self._h, C.byref(params), _lib.ptr(x), _lib.MEM_HOST, x.shape[0], _lib.ptr(tau)))
        return tau[..., 0] + 1j * tau[..., 1]

    def kxky_filter(self, x, filt, out=None):
        """``ifft2(filt * fft2(x))`` per slice (complex64).  ``filt``: real (n_iline, n_xline) plane in FFT order."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        x3 = x[None] if x.ndim == 2 else x
        if x3.shape[1:] != (self.n_iline, self.n_xline):
            raise ValueError(f"slice shape {x3.shape[1:]} does not match the plan ({self.n_iline}, {self.n_xline})")
        filt = np.ascontiguousarray(filt, dtype=np.float32)
        if filt.shape != (self.n_iline, self.n_xline):
            raise ValueError(f"filter shape {filt.shape} does not match the plan")
        if out is None:
            out = np.empty_like(x3)
        with self._lock:
            _lib.check(_lib.load().p3d_kxky_filter_run(self._h, _lib.ptr(x3), _lib.MEM_HOST, _lib.ptr(filt), _lib.MEM_HOST,
                                                       _lib.ptr(out), _lib.MEM_HOST, x3.shape[0]))
        return out[0] if x.ndim == 2 else out

    def kxky_filter_device(self, x_ptr, filt_ptr, out_ptr, n_slices):
        """Device-resident variant of :meth:`kxky_filter` (raw device pointers)."""
        with self._lock:
            _lib.check(_lib.load().p3d_kxky_filter_run(self._h, C.c_void_p(int(x_ptr)), _lib.MEM_DEVICE, C.c_void_p(int(filt_ptr)),
                                                       _lib.MEM_DEVICE, C.c_void_p(int(out_ptr)), _lib.MEM_DEVICE, int(n_slices)))

    def fft2(self, x, inverse=False):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        x3 = x[None] if x.ndim == 2 else x
        out = np.empty_like(x3)
        with self._lock:
            _lib.check(_lib.load().p3d_fft2(self._h, _lib.ptr(x3), _lib.MEM_HOST, _lib.ptr(out), _lib.MEM_HOST, x3.shape[0], 1 if inverse else 0))
        return out[0] if x.ndim == 2 else out


def _precision_code(p):
    if isinstance(p, str):
        p = {"auto": 0, "escalating": 0, "0": 0, "32": 32, "64": 64}.get(p.strip().lower(), p)
    if p not in (0, 32, 64):
        raise ValueError('precision must be "auto" (0), 32 or 64')
    return int(p)


_PLANS = {}
_PLANS_LOCK = threading.Lock()
_DEFAULT_PRECISION = [_precision_code(os.environ.get("P3D_PRECISION", "auto"))]


def set_default_precision(bits):
    """"auto" / 0 (escalating fp32 -> complex128, the default), 32 (fp32 only) or 64 (float64 state mode) for
    POCS_algorithm / pocs_cube calls that do not say otherwise.  Also settable with the environment variable
    P3D_PRECISION."""
    _DEFAULT_PRECISION[0] = _precision_code(bits)


def release_plans():
    """Destroy the cached plans (and with them their device buffers, which are sized for the largest call so far)."""
    with _PLANS_LOCK:
        for p in _PLANS.values():
            p.close()
        _PLANS.clear()


def get_plan(n_iline, n_xline, device=0, precision=None) -> PocsPlan:
    precision = _DEFAULT_PRECISION[0] if precision is None else _precision_code(precision)
    key = (int(n_iline), int(n_xline), int(device), precision)
    with _PLANS_LOCK:
        p = _PLANS.get(key)
        if p is None:
            p = _PLANS[key] = PocsPlan(key[0], key[1], device=key[2], precision=precision)
        return p


def fft2(x, device=0):
    """numpy.fft.fft2-compatible forward transform on the GPU (complex64)."""
    x = np.asarray(x)
    return get_plan(x.shape[-2], x.shape[-1], device).fft2(x)


def ifft2(x, device=0):
    x = np.asarray(x)
    return get_plan(x.shape[-2], x.shape[-1], device).fft2(x, inverse=True)


# --------------------------------------------------------------------------------------------------
# reference-compatible entry points
# --------------------------------------------------------------------------------------------------
def POCS_algorithm(
    x,
    mask,
    auxiliary_data=None,
    transform=None,
    itransform=None,
    transform_kind: str = None,
    niter: int = 50,
    thresh_op: str = "hard",
    thresh_model: str = "exponential",
    eps: float = 1e-9,
    alpha: float = 1.0,
    p_max: float = 0.99,
    p_min: float = 1e-5,
    sqrt_decay: bool = False,
    decay_kind: str = "values",
    verbose: bool = False,
    version: str = "regular",
    results_dict: dict = None,
    path_results: str = None,
    device: int = 0,
):
    """Interpolate one sparse 2-D slice with FFT-POCS on the GPU.

    Same contract as the reference's ``POCS_algorithm`` (functions/POCS.py:371-656):
    ``x`` real or complex 2-D, ``mask`` in {0, 1}; complex input returns complex, real input
    returns the real part; ``results_dict`` receives ``niterations``/``runtime``/``cost``;
    ``path_results`` gets the ``niterations;runtime;cost_0;...`` line appended.
    ``transform``/``itransform`` must be supplied (the reference raises otherwise) but only
    ``transform_kind='FFT'`` is executed here: the 2-D FFTs are the library's own kernels.
    """
    x = np.asarray(x)
    mask = np.asarray(mask)
    if np.max(mask) > 1:
        raise ValueError(f"mask should be quasi-boolean (0 or 1) but has maximum of {np.max(mask)}")
    if any(v is None for v in [transform, itransform]):
        raise ValueError("Forward and inverse transform function have to be supplied")
    if transform_kind is None or transform_kind.upper() not in TRANSFORMS:
        raise ValueError(f"Unsupported transform. Please select one of: {TRANSFORMS}")
    transform_kind = transform_kind.upper()
    if transform_kind != "FFT":
        raise NotImplementedError(f"transform_kind={transform_kind!r}: only the FFT transform runs on the B200 path")
    if x.ndim != 2:
        raise ValueError("x must be a 2-D slice")
    params = make_params(niter, thresh_op, thresh_model, eps, alpha, p_max, p_min, sqrt_decay, decay_kind, version)

    is_complex_input = np.iscomplexobj(x)
    plan = get_plan(x.shape[0], x.shape[1], device)
    t0 = time.perf_counter()
    out, info = plan.run(x.astype(np.complex64), mask, params=params, want_costs=path_results is not None or verbose)
    runtime = time.perf_counter() - t0
    niterations = int(info["niterations"][0])
    cost = float(info["cost"][0]) if niterations > 0 else 0
    if niterations == 0:
        runtime = 0
    if verbose:
        print("\n" + "-" * 20)
        print(f"# iterations:  {niterations:4d}")
        print(f"cost function: {cost}")
        print(f"runtime:       {runtime:.3f} s")
        print("-" * 20)
    if isinstance(results_dict, dict):
        results_dict["niterations"] = niterations
        results_dict["runtime"] = round(runtime, 3)
        results_dict["cost"] = cost
    if path_results is not None:
        costs = [0] if niterations == 0 else [float(c) for c in info["costs"][0][:niterations]]
        with open(path_results, mode="a", newline="\n") as f:
            f.write(";".join([str(i) for i in [niterations, runtime] + costs]) + "\n")
    if niterations == 0:
        return x                                    # all-zero slice: input returned unchanged
    y = out[0]
    if is_complex_input:
        return y.astype(x.dtype if x.dtype in (np.complex64, np.complex128) else np.complex64, copy=False)
    yr = np.real(y)
    return yr.astype(x.dtype, copy=False) if np.issubdtype(x.dtype, np.floating) else yr


POCS = partial(POCS_algorithm, version="regular")
FPOCS = partial(POCS_algorithm, version="fast")
APOCS = partial(POCS_algorithm, version="adaptive")


def get_threshold_decay(thresh_model, niter: int, transform_kind: str = None, p_max: float = 0.99,
                        p_min: float = 1e-3, x_fwd=None, kind: str = "values"):
    """Iteration-based threshold schedule for the FFT transform (host utility, float64).

    Mirrors functions/POCS.py:169-368 including the complex lexicographic maximum
    (``x_fwd.max()`` on a complex array).  Returns a complex array for linear /
    exponential[-q] / data-driven and a float array for inverse-proportional[-q].
    """
    if transform_kind is not None:
        if transform_kind.upper() not in TRANSFORMS and (kind == "values" or thresh_model == "data-driven"):
            raise ValueError(f"Unsupported transform. Please select one of: {TRANSFORMS}")
        transform_kind = transform_kind.upper()
        if transform_kind not in ("FFT", "DCT", "CURVELET"):
            raise NotImplementedError(f"{transform_kind} schedules are not part of the B200 path")
    if x_fwd is None and (kind == "values" or thresh_model == "data-driven"):
        raise ValueError('`x_fwd` must be specified for thresh_model="data-driven" or kind="values"!')
    niter = int(niter)
    steps = np.arange(1, niter + 1)
    model, q = _parse_model(thresh_model)

    if model == _lib.MODELS["inverse_proportional"]:
        mag = np.abs(x_fwd)
        hi, lo = mag.max(), mag.min()
        a = (niter ** q * (hi - lo)) / (niter ** q - 1)
        b = (niter ** q * lo - hi) / (niter ** q - 1)
        return a / (steps ** q) + b

    if kind == "values":
        if transform_kind is None:
            raise ValueError('`transform_kind` must be specified for thresh_model="data-driven" or kind="values"!')
        xf = np.asarray(x_fwd)
        peak = xf.max()                 # complex: numpy's lexicographic max (largest real part)
        if isinstance(p_min, str) and p_min == "adaptive":
            tau_lo = 0.01 * np.sqrt(np.linalg.norm(xf) ** 2 / xf.size)
        else:
            tau_lo = p_min * peak
        tau_hi = p_max * peak
    elif kind == "factors":
        tau_hi, tau_lo = p_max, p_min
    else:
        raise ValueError('Parameter `kind` only supports arguments "values" or "factors"')

    frac = (steps - 1) / (niter - 1)
    if model == _lib.MODELS["linear"]:
        return tau_hi - (tau_hi - tau_lo) * frac
    if model == _lib.MODELS["exponential"]:
        return tau_hi * np.exp(np.log(tau_lo / tau_hi) * frac ** q)
    # data-driven: order statistics of the coefficients between the two bounds
    xf = np.asarray(x_fwd)
    cand = np.sort(xf[(xf > tau_lo) & (xf < tau_hi)])[::-1]
    picks = np.ceil((steps[1:] - 1) * (cand.size - 1) / (niter - 1)).astype("int")
    tau = np.zeros((niter,), dtype=xf.dtype)
    tau[0] = cand[0]
    tau[1:] = cand[picks]
    return tau


def threshold(data, thresh, sub=0, kind="soft"):
    """Apply a threshold operator to ``data`` (host utility; functions/POCS.py:61-102).

    ``thresh`` may be complex, in which case numpy's lexicographic ordering applies exactly as
    in the reference.  Percentile variants take ``thresh`` as a percentile of |data|.
    """
    data = np.asarray(data)
    if kind.endswith("-percentile"):
        thresh = np.percentile(np.abs(data), thresh)
        kind = kind[: -len("-percentile")]
    mag = np.absolute(data)
    if kind == "hard":
        return np.where(np.less(mag, thresh), sub, data)
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == "soft":
            shrink = 1 - thresh / mag
        elif kind in ("garrote", "garotte"):
            shrink = 1 - thresh ** 2 / mag ** 2
        else:
            return None                               # the reference falls through silently
        shrink = np.maximum(shrink, 0)                # ndarray.clip(min=0): lexicographic for complex
        res = data * shrink
    if sub == 0:
        return res
    return np.where(np.less(mag, thresh), sub, res)


# --------------------------------------------------------------------------------------------------
# whole-cube driver: contiguous frequency bands per GPU, no inter-GPU traffic
# --------------------------------------------------------------------------------------------------
def band_bounds(n_slices: int, n_parts: int):
    """Contiguous band [start, stop) of every part (ceil split)."""
    per = -(-n_slices // max(n_parts, 1))
    return [(min(i * per, n_slices), min((i + 1) * per, n_slices)) for i in range(n_parts)]


def pocs_cube(cube, fold_or_mask, devices=None, out=None, results=None, precision=None, **metadata):
    """POCS over every slice of ``cube`` (n_slices, n_iline, n_xline).

    ``cube`` complex64 (frequency domain) or float32 (time domain, iterated as complex and
    returned as the real part).  ``fold_or_mask`` (n_iline, n_xline) uint8 is clipped
    to {0,1}.  ``devices``: GPU ids; the slice axis is split into one contiguous band per GPU,
    each processed by its own host thread and plan.  ``metadata`` are POCS_algorithm keywords
    (``transform``/``itransform``/``transform_kind`` are accepted and checked like the reference).
    """
    metadata = dict(metadata)
    tk = metadata.pop("transform_kind", "FFT")
    if tk is None or str(tk).upper() not in TRANSFORMS:
        raise ValueError(f"Unsupported transform. Please select one of: {TRANSFORMS}")
    if str(tk).upper() != "FFT":
        raise NotImplementedError(f"transform_kind={tk!r}: only the FFT transform runs on the B200 path")
    for k in ("transform", "itransform", "auxiliary_data", "verbose", "results_dict", "path_results"):
        metadata.pop(k, None)
    params = make_params(**metadata)
    cube = np.asarray(cube)
    if cube.ndim != 3:
        raise ValueError("cube must be (n_slices, n_iline, n_xline)")
    is_complex = np.iscomplexobj(cube)
    mask = mask_from_fold(fold_or_mask).astype(np.uint8)
    ns, n1, n2 = cube.shape
    if mask.shape != (n1, n2):
        raise ValueError(f"mask shape {mask.shape} does not match slices ({n1}, {n2})")
    xin = cube if cube.dtype == np.complex64 else cube.astype(np.complex64)
    # the caller's `out` is written in place only when it is a dense complex64 array of the right shape
    direct = (out is not None and isinstance(out, np.ndarray) and out.dtype == np.complex64 and out.shape == (ns, n1, n2)
              and out.flags.c_contiguous and out.flags.writeable)
    res = out if direct else np.empty((ns, n1, n2), dtype=np.complex64)
    nit = np.zeros(ns, dtype=np.int32)
    cost = np.zeros(ns, dtype=np.float64)
    if devices is None:
        devices = [0]
    _lib.require_gpu()
    bounds = band_bounds(ns, len(devices))
    errors = []

    def work(dev, lo, hi):
        try:
            if hi <= lo:
                return
            plan = get_plan(n1, n2, dev, precision)
            try:
                _, info = plan.run(xin[lo:hi], mask, out=res[lo:hi], params=params)
            except MemoryError:
                # cached plans of other shapes keep their lane buffers (sized for their largest call): drop them and retry
                release_plans()
                plan = get_plan(n1, n2, dev, precision)
                _, info = plan.run(xin[lo:hi], mask, out=res[lo:hi], params=params)
            nit[lo:hi] = info["niterations"]
            cost[lo:hi] = info["cost"]
        except Exception as e:      # noqa: BLE001 - re-raised below
            errors.append(e)

    if len(devices) == 1:
        work(devices[0], *bounds[0])
    else:
        ts = [threading.Thread(target=work, args=(d, lo, hi)) for d, (lo, hi) in zip(devices, bounds)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    if errors:
        raise errors[0]
    if isinstance(results, dict):
        results["niterations"] = nit
        results["cost"] = cost
    if is_complex:
        if out is not None and out is not res:
            out[...] = res
            return out
        return res if cube.dtype == np.complex64 else res.astype(cube.dtype)
    real = res.real.astype(cube.dtype if np.issubdtype(cube.dtype, np.floating) else np.float32)
    if out is not None:
        out[...] = real
        return out
    return real
