"""Deterministic synthetic pseudo-3D cubes for tests and the benchmark.

A cube is a sum of K dipping planar events with Ricker wavelets,
``d(t, il, xl) = sum_k a_k * ricker_k(t - t0 - tau_k - p_k (il - il_c) - q_k (xl - xl_c))``.
Two equivalent generators are offered:

* :func:`time_cube` samples that expression on ``twt = t0 + n dt`` -> ``(nt, n_il, n_xl)``
  float32, the layout step 11 of the reference writes (cube_preprocessing_3D.py:374);
* :func:`freq_slices` evaluates the analytic spectrum of the same events per frequency,
  ``D(f) = sum_k a_k R_k(f) exp(-2 pi i f (t0 + tau_k + p_k (il-il_c) + q_k (xl-xl_c)))``
  with ``R_k(f) = (2/sqrt(pi)) f^2/f_k^3 exp(-f^2/f_k^2)``, i.e. what step 12 would emit,
  without ever materialising the time cube (needed for the 2000x2000x4096 config).

Masks: Bernoulli keep-probability, or the pseudo-3D line pattern of config 4.
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np

T0_MS = 725.0
DT_MS = 0.05


@dataclass
class Events:
    amp: np.ndarray
    tau: np.ndarray
    p: np.ndarray
    q: np.ndarray
    fpk: np.ndarray      # Ricker peak frequency (kHz when dt is in ms)


# BASELINE.json configs (id -> parameters); C3/C4 iteration counts follow SURVEY.md §8d
CONFIGS = {
    1: dict(n_il=200, n_xl=200, nt=512, keep=0.30, niter=50, thresh_op="hard",
            thresh_model="exponential", alpha=1.0),
    2: dict(n_il=1000, n_xl=1000, nt=2048, keep=0.20, niter=100, thresh_op="hard",
            thresh_model="exponential", alpha=1.0),
    3: dict(n_il=1201, n_xl=847, nt=3000, keep=0.20, niter=100, thresh_op="soft",
            thresh_model="linear", alpha=1.0),
    4: dict(n_il=2000, n_xl=2000, nt=4096, keep="lines", niter=100, thresh_op="hard",
            thresh_model="data-driven", alpha=0.7),
    5: dict(n_il=256, n_xl=256, nt=1024, keep=0.20, niter=30, thresh_op="garrote",
            thresh_model="exponential", alpha=1.0, n_cubes=64),
}


def draw_events(rng, nt, n_il, n_xl, dt=DT_MS, n_events=8) -> Events:
    f_nyq = 0.5 / dt
    sign = rng.choice([-1.0, 1.0], size=n_events)
    amp = rng.uniform(0.3, 1.0, size=n_events) * sign
    tau = rng.uniform(0.15, 0.85, size=n_events) * nt * dt
    # total move-out across the cube <= 0.25 * record length
    smax_il = 0.125 * nt * dt / max(n_il, 1)
    smax_xl = 0.125 * nt * dt / max(n_xl, 1)
    p = rng.uniform(-smax_il, smax_il, size=n_events)
    q = rng.uniform(-smax_xl, smax_xl, size=n_events)
    fpk = rng.uniform(0.08, 0.2, size=n_events) * f_nyq
    return Events(amp, tau, p, q, fpk)


def make_fold(rng, n_il, n_xl, keep):
    """uint8 fold map.  ``keep`` = Bernoulli probability or 'lines' (config-4 pattern)."""
    if isinstance(keep, str) and keep == "lines":
        il = np.arange(n_il)[:, None]
        xl = np.arange(n_xl)[None, :]
        a = (il % 12 == 0)
        b = (xl % 60 == 0)
        return (a.astype(np.uint8) + b.astype(np.uint8)).astype(np.uint8)   # 2 at crossings
    return (rng.random((n_il, n_xl)) < keep).astype(np.uint8)


def _delays(ev: Events, n_il, n_xl, t0):
    il = np.arange(n_il, dtype=np.float64)[:, None] - (n_il - 1) / 2.0
    xl = np.arange(n_xl, dtype=np.float64)[None, :] - (n_xl - 1) / 2.0
    return [t0 + ev.tau[k] + ev.p[k] * il + ev.q[k] * xl for k in range(len(ev.amp))]


def time_cube(ev: Events, nt, n_il, n_xl, dt=DT_MS, t0=T0_MS):
    twt = t0 + dt * np.arange(nt, dtype=np.float64)
    d = np.zeros((nt, n_il, n_xl), dtype=np.float64)
    for k, delay in enumerate(_delays(ev, n_il, n_xl, t0)):
        u = twt[:, None, None] - delay[None]
        a = (np.pi * ev.fpk[k] * u) ** 2
        d += ev.amp[k] * (1.0 - 2.0 * a) * np.exp(-a)
    return d.astype(np.float32), twt


def freq_slices(ev: Events, freqs, n_il, n_xl, t0=T0_MS, dtype=np.complex64):
    """Analytic step-12 output for the listed frequencies -> (len(freqs), n_il, n_xl)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    out = np.zeros((freqs.size, n_il, n_xl), dtype=np.complex128)
    il = np.arange(n_il, dtype=np.float64) - (n_il - 1) / 2.0
    xl = np.arange(n_xl, dtype=np.float64) - (n_xl - 1) / 2.0
    for i, f in enumerate(freqs):
        acc = np.zeros((n_il, n_xl), dtype=np.complex128)
        for k in range(len(ev.amp)):
            rk = (2.0 / np.sqrt(np.pi)) * f * f / ev.fpk[k] ** 3 * np.exp(-(f / ev.fpk[k]) ** 2)
            # the plane-wave phase is separable in (il, xl): one outer product per event
            a0 = ev.amp[k] * rk * np.exp(-2j * np.pi * f * (t0 + ev.tau[k]))
            a1 = np.exp(-2j * np.pi * f * ev.p[k] * il)
            a2 = np.exp(-2j * np.pi * f * ev.q[k] * xl)
            acc += np.outer(a0 * a1, a2)
        out[i] = acc
    return out.astype(dtype)


def config_case(config_id, n_il=None, n_xl=None, nt=None, noise=0.0):
    """(events, fold uint8, rng, params) for a BASELINE config, optionally shrunk.

    Draw order is events -> mask -> noise, ``rng = default_rng(1000 + config_id)``.
    """
    c = dict(CONFIGS[config_id])
    n_il = n_il or c["n_il"]
    n_xl = n_xl or c["n_xl"]
    nt = nt or c["nt"]
    rng = np.random.default_rng(1000 + config_id)
    ev = draw_events(rng, nt, n_il, n_xl)
    fold = make_fold(rng, n_il, n_xl, c["keep"])
    c.update(n_il=n_il, n_xl=n_xl, nt=nt, noise=noise)
    return ev, fold, rng, c


def sparse_freq_slices(config_id, slice_ids, n_il=None, n_xl=None, nt=None, noise=0.0):
    """Sparse (masked) complex64 frequency slices of a config + its fold map + parameters."""
    ev, fold, rng, c = config_case(config_id, n_il, n_xl, nt, noise)
    f = np.fft.rfftfreq(c["nt"], DT_MS)[np.asarray(slice_ids)]
    d = freq_slices(ev, f, c["n_il"], c["n_xl"], dtype=np.complex128)
    if noise > 0:
        sig = noise * np.abs(d).max()
        d = d + sig * (rng.standard_normal(d.shape) + 1j * rng.standard_normal(d.shape)) / np.sqrt(2)
    d = d * (fold > 0)[None]
    return d.astype(np.complex64), fold, c
