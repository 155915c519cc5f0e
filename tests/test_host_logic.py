"""CPU-only checks: host-side mirrors against the reference's golden vectors, parameter
translation, band sharding, the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np
import pytest

import pseudo_3d_interpolation_b200 as p3d
from pseudo_3d_interpolation_b200 import _lib, pocs, timeaxis, synth
from oracle.golden_cases import CASES, SCHEDULES, make_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "p3d_b200.h")).read()
    declared = set(re.findall(r"\b(p3d_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"p3d_plan"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.p3d_abi_version() == 3


def test_params_struct_layout_matches_header():
    # 4 int32, 5 double, 5 int32 (+ 4 bytes tail padding) -> 16 + 40 + 24 = 80 bytes
    assert ctypes.sizeof(_lib.PocsParams) == 80
    p = pocs.make_params(niter=7, thresh_op="garrote", thresh_model="exponential-2", eps=1e-6, alpha=0.7,
                         p_max=0.9, p_min="adaptive", sqrt_decay=True)
    assert (p.niter, p.thresh_op, p.thresh_model, p.q, p.p_min_adaptive, p.sqrt_decay) == (7, 2, 1, 2.0, 1, 1)
    assert pocs.make_params(thresh_model="inverse-proportional-3").q == 3.0
    with pytest.raises(NotImplementedError):
        pocs.make_params(thresh_model="cubic")
    # percentile operators: only with decay_kind='factors' and percentiles in [0, 100] (np.percentile's own error otherwise)
    with pytest.raises(ValueError, match="Percentiles must be in the range"):
        pocs.make_params(thresh_op="hard-percentile")
    with pytest.raises(ValueError, match="Percentiles must be in the range"):
        pocs.make_params(thresh_op="soft-percentile", decay_kind="factors", p_max=120.0, p_min=1.0)
    pp = pocs.make_params(thresh_op="garrote-percentile", decay_kind="factors", p_max=99.0, p_min=1.0)
    assert (pp.thresh_op, pp.thresh_percentile, pp.decay_factors) == (2, 1, 1)
    with pytest.raises(NotImplementedError):
        pocs.make_params(thresh_op="median-percentile", decay_kind="factors", p_max=99.0, p_min=1.0)
    with pytest.raises(ValueError):
        pocs.make_params(decay_kind="nope")
    with pytest.raises(TypeError):
        pocs.make_params(p_min="1e-5")


def test_no_cpu_fallback_without_gpu():
    if _lib.load().p3d_device_count() > 0:
        pytest.skip("GPU present")
    x = np.ones((8, 8), dtype=np.complex64)
    with pytest.raises(_lib.P3dError):
        p3d.POCS(x, np.ones((8, 8), np.uint8), None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT")
    with pytest.raises(_lib.P3dError):
        p3d.pocs_cube(x[None], np.ones((8, 8), np.uint8))
    with pytest.raises(_lib.P3dError):
        timeaxis.time_fft(np.zeros((8, 2, 2), np.float32), np.arange(8.0))


@pytest.mark.parametrize("i", range(len(SCHEDULES)))
def test_get_threshold_decay_matches_reference(i, golden):
    sp = SCHEDULES[i]
    x, _ = make_input(CASES[0])
    X0 = np.fft.fft2(x.astype(np.complex128))
    tau = pocs.get_threshold_decay(sp["thresh_model"], sp["niter"], "FFT", sp["p_max"], sp["p_min"], x_fwd=X0, kind=sp["kind"])
    assert np.array_equal(tau, golden[f"schedule_{i}"])


@pytest.mark.parametrize("kind", ["hard", "soft", "garrote"])
@pytest.mark.parametrize("tname,tau", [("pos", 1.0 + 0.25j), ("neg", 1.0 - 0.25j), ("real", 0.8)])
def test_threshold_matches_reference(kind, tname, tau, golden):
    got = pocs.threshold(golden["thr_X"], tau, sub=0, kind=kind)
    assert np.array_equal(got, golden[f"thr_{kind}_{tname}"], equal_nan=True)


def test_threshold_errors():
    with pytest.raises(ValueError):
        pocs.get_threshold_decay("linear", 5, "FFT", x_fwd=None)
    with pytest.raises(ValueError):
        pocs.get_threshold_decay("linear", 5, "NOPE", x_fwd=np.ones(3, complex))
    with pytest.raises(ValueError):
        pocs.get_threshold_decay("linear", 5, "FFT", x_fwd=np.ones(3, complex), kind="bad")


def test_band_bounds_cover_and_are_contiguous():
    for n in (0, 1, 7, 1025, 2049):
        for parts in (1, 2, 4, 8):
            b = pocs.band_bounds(n, parts)
            assert len(b) == parts and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b if hi > lo or n == 0) <= -(-n // parts)


def test_mask_from_fold_bit_exact():
    fold = np.array([[0, 1, 2, 7], [255, 0, 1, 3]], dtype=np.uint8)
    m = p3d.mask_from_fold(fold)
    assert m.dtype == np.uint8 and np.array_equal(m, [[0, 1, 1, 1], [1, 0, 1, 1]])


def test_synth_time_and_freq_agree():
    """Analytic frequency slices equal the step-12 transform of the sampled time cube."""
    from oracle import time_axis_oracle as tor
    rng = np.random.default_rng(9)
    nt, n1, n2 = 256, 6, 5
    ev = synth.draw_events(rng, nt, n1, n2)
    d, twt = synth.time_cube(ev, nt, n1, n2)
    F, f = tor.time_fft(d, twt, compute_real=True)
    ids = [5, 17, 40]
    A = synth.freq_slices(ev, f[ids], n1, n2)
    err = np.linalg.norm(A - F[ids]) / np.linalg.norm(F[ids])
    assert err < 1e-3, err


def test_freq_filter_window_matches_oracle():
    from oracle import time_axis_oracle as tor
    f = np.fft.rfftfreq(512, 0.05)
    for ftype, freqs in (("lowpass", [3.0, 5.0]), ("highpass", [0.5, 1.5]), ("bandpass", [0.5, 1.5, 4.0, 6.0])):
        np.testing.assert_array_equal(timeaxis.freq_filter_window(list(freqs), f, ftype), tor.freq_filter_window(list(freqs), f, ftype))
        np.testing.assert_array_equal(timeaxis.freq_filter_keep(f, freqs, ftype), tor.freq_filter_keep(f, freqs, ftype))


def test_numa_binding_helpers(tmp_path, monkeypatch):
    """cpulist parsing and the sysfs walk of bind_to_gpu_numa_node on a fake /sys tree (no GPU, no NVML needed)."""
    from pseudo_3d_interpolation_b200 import distributed as pd
    assert pd._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert pd._parse_cpulist("") == []
    # unknown GPU -> not bound, never raises
    assert "not bound" in pd.bind_to_gpu_numa_node(0, sysfs=str(tmp_path))
    # fake topology: GPU on node 1 whose cpus are the ones this process may use
    monkeypatch.setattr(pd, "gpu_numa_node", lambda dev, sysfs="/sys": 1)
    node = tmp_path / "devices" / "system" / "node" / "node1"
    node.mkdir(parents=True)
    allowed = sorted(os.sched_getaffinity(0))
    (node / "cpulist").write_text(",".join(str(c) for c in allowed[:2]) + "\n")
    before = os.sched_getaffinity(0)
    try:
        msg = pd.bind_to_gpu_numa_node(0, sysfs=str(tmp_path))
        assert "node 1" in msg and os.sched_getaffinity(0) == set(allowed[:2])
    finally:
        os.sched_setaffinity(0, before)
    (node / "cpulist").write_text("9999\n")
    assert "no allowed cpus" in pd.bind_to_gpu_numa_node(0, sysfs=str(tmp_path))


def test_postprocessing_direction_aliases():
    from pseudo_3d_interpolation_b200 import cube_postprocessing_3D as pp
    a = pp.footprint_filter((40, 60), sigma=2, direction="iline")
    b = pp.footprint_filter((40, 60), sigma=2, direction="xline", dims=("xline", "iline"))       # swapped dims: same stencil
    assert np.array_equal(a, b)
    t = pp.footprint_filter((40, 60), sigma=2, direction="twt")                                    # ny < nx -> horizontal
    assert np.array_equal(t, a)
    assert not np.array_equal(a, pp.footprint_filter((40, 60), sigma=2, direction="xline"))
    f = pp.antialiasing_filter((64, 48), "xline", {"iline": 1, "xline": 4}, sigma=2)
    assert f.shape == (64, 48) and abs(f.min() - 1e-3) < 1e-12 and abs(f.max() - 1.0) < 1e-12


def test_band_bounds_cover_everything():
    for n, w in ((1025, 8), (7, 3), (1, 2), (0, 4), (5, 5)):
        b = pocs.band_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) == -(-n // w) and min(sizes) >= 0          # ceil split: nobody gets more than ceil(n / w)
