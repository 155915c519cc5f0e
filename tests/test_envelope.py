"""Amplitude envelope along the time axis: oracle against outputs of the reference's own
``functions.signal.envelope`` (tests/golden/reference_envelope.npz, made by oracle/make_golden_envelope.py),
GPU kernels against both."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import time_axis_oracle as orc                                  # noqa: E402
from oracle.make_golden_envelope import CASES, make_cube                   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "reference_envelope.npz")
ATOL32 = 3e-6        # float32 transform pair on O(1) amplitudes (the reference's own float32 path is 2.5e-7 off its float64 one)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference(case, gold):
    x = make_cube(*case)
    np.testing.assert_allclose(orc.envelope(x, axis=0), gold[case[0] + "__env64"], rtol=0, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gpu_envelope_matches_reference_golden(case, gold):
    from pseudo_3d_interpolation_b200 import timeaxis
    x = make_cube(*case)
    e = timeaxis.envelope(x, axis=0)
    assert e.dtype == np.float32 and e.shape == x.shape
    np.testing.assert_allclose(e, gold[case[0] + "__env64"], rtol=0, atol=ATOL32)
    np.testing.assert_allclose(e, gold[case[0] + "__env32"], rtol=0, atol=ATOL32)


@pytest.mark.gpu
@pytest.mark.parametrize("nt,ntr", [(512, 1001), (1024, 130), (2048, 64), (4096, 7), (300, 33), (1201, 5), (1, 4), (2, 3),
                                    (1000, 37), (2000, 9), (2500, 5), (3000, 11), (4000, 3), (5000, 4)])
def test_gpu_envelope_lengths_and_axes(nt, ntr):
    """register-resident pipeline (512 ... 4096, odd trace counts), generic direct kernel (any length, Bluestein),
    degenerate lengths; other axes go through a time-major view."""
    from pseudo_3d_interpolation_b200 import timeaxis
    rng = np.random.default_rng(nt + ntr)
    x = rng.standard_normal((nt, ntr)).astype(np.float32)
    e = timeaxis.envelope(x, axis=0)
    ref = orc.envelope(x, axis=0)
    np.testing.assert_allclose(e, ref, rtol=0, atol=2e-5 * max(1.0, np.sqrt(np.log2(max(nt, 2)))))
    if nt in (300, 512):
        y = np.ascontiguousarray(x.T)
        np.testing.assert_allclose(timeaxis.envelope(y, axis=-1), e.T, rtol=0, atol=1e-6)
        z = x.reshape(nt, ntr, 1).astype(np.float64)
        ez = timeaxis.envelope(z, axis=0)
        assert ez.dtype == np.float64 and ez.shape == z.shape


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["tma", "pipeline"])
@pytest.mark.parametrize("nt,ntr", [(512, 740), (1024, 260), (2048, 108), (4096, 44), (1000, 332), (2000, 100), (2500, 52), (4000, 36)])
def test_gpu_envelope_one_pass_and_pipeline(nt, ntr, path, monkeypatch):
    """the one-pass TMA-staged envelope kernel (trace counts divisible by 4; three CTAs, so every CTA walks over several tiles,
    the last one partial) and the transposing pipeline on the same input"""
    from pseudo_3d_interpolation_b200 import timeaxis, _lib
    monkeypatch.setenv("P3D_TIME_PATH", path)
    monkeypatch.setenv("P3D_TIME_GRID", "3")
    rng = np.random.default_rng(nt + ntr)
    x = rng.standard_normal((nt, ntr)).astype(np.float32)
    e = timeaxis.envelope(x, axis=0)
    assert _lib.load().p3d_time_last_path().decode() == path
    np.testing.assert_allclose(e, orc.envelope(x, axis=0), rtol=0, atol=2e-5 * np.sqrt(np.log2(nt)))
