"""Property tests (hypothesis) of the numpy oracle: the invariants SURVEY 8c lists for the reference algorithm, on random
small slices.  They guard the checker itself; the GPU path is held to the same properties at full size in
tests/test_gpu_parity.py::test_full_size_properties_all_configs."""
import os
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pocs_oracle as orc                  # noqa: E402
from oracle import time_axis_oracle as tor             # noqa: E402

OPS = st.sampled_from(["hard", "soft", "garrote"])
MODELS = st.sampled_from(["linear", "exponential", "exponential-2", "data-driven"])


def _slice(seed, n1, n2, keep):
    rng = np.random.default_rng(seed)
    i, j = np.mgrid[:n1, :n2]
    x = sum(rng.uniform(0.3, 1) * np.exp(2j * np.pi * (rng.uniform(-.3, .3) * i + rng.uniform(-.3, .3) * j + rng.random())) for _ in range(3))
    mask = (rng.random((n1, n2)) < keep).astype(np.uint8)
    mask[0, 0] = 1
    return (x * mask).astype(np.complex128), mask


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), n1=st.integers(6, 20), n2=st.integers(6, 20), op=OPS, model=MODELS, niter=st.integers(2, 8))
def test_observed_traces_exact_and_scale_equivariance(seed, n1, n2, op, model, niter):
    x, mask = _slice(seed, n1, n2, 0.4)
    params = dict(niter=niter, thresh_op=op, thresh_model=model, eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)
    info = {}
    y = orc.pocs_slice(x, mask, info=info, **params)
    assert info["niterations"] == niter and len(info["costs"]) == niter
    # alpha = 1: observed samples come back bit for bit (x_inv * 0 + x)
    assert np.array_equal(y[mask == 1], x[mask == 1])
    # thresholds are p * z(x): a power-of-two scale passes through every operation exactly
    y4 = orc.pocs_slice(4.0 * x, mask, **params)
    assert np.array_equal(y4, 4.0 * y)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(8, 24), niter=st.integers(3, 12), p_min=st.sampled_from([1e-5, 1e-3, 1e-1]))
def test_schedule_end_points(seed, n, niter, p_min):
    x, _ = _slice(seed, n, n + 3, 1.0)
    X0 = np.fft.fft2(x)
    z = orc.lexmax(X0)
    for model in ("linear", "exponential", "exponential-3"):
        tau = orc.threshold_table(X0, niter, model, 0.99, p_min)
        assert tau.shape == (niter,)
        np.testing.assert_allclose(tau[0], 0.99 * z, rtol=1e-12)
        np.testing.assert_allclose(tau[-1], p_min * z, rtol=1e-9)
    t = orc.threshold_table(X0, niter, "inverse_proportional", 0.99, p_min)
    np.testing.assert_allclose([t[0], t[-1]], [np.abs(X0).max(), np.abs(X0).min()], rtol=1e-9)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10_000), nt=st.integers(4, 40), real=st.booleans())
def test_time_axis_round_trip(seed, nt, real):
    rng = np.random.default_rng(seed)
    nt -= nt % 2
    x = rng.standard_normal((max(nt, 2), 3, 2)).astype(np.float32)
    twt = 725.0 + 0.05 * np.arange(x.shape[0])
    F, f = tor.time_fft(x, twt, compute_real=real)
    assert F.dtype == np.complex64 and F.shape[0] == (x.shape[0] // 2 + 1 if real else x.shape[0])
    Fin = F if real else np.fft.fftshift(F, axes=0)               # step 13's merge sorts the frequency axis
    xb = tor.time_ifft(Fin, 0.05, 725.0, compute_real=real, ascending=True)
    np.testing.assert_allclose(xb, x, atol=2e-5 * np.abs(x).max())


def test_fpocs_equals_pocs_and_all_zero_slice():
    x, mask = _slice(3, 12, 10, 0.5)
    params = dict(niter=6, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0)
    assert np.array_equal(orc.pocs_slice(x, mask, version="regular", **params), orc.pocs_slice(x, mask, version="fast", **params))
    info = {}
    z = np.zeros_like(x)
    assert orc.pocs_slice(z, mask, info=info, **params) is z and info["niterations"] == 0
