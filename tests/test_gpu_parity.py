"""GPU parity tests: the CUDA path (through the C ABI) against the reference's golden outputs
and against the numpy oracle on seeded inputs.  Tolerance: relative L2 <= 1e-4 (device result vs
the float64 reference, BASELINE.json north_star) in the DEFAULT mode of the library (escalating
precision: fp32 pilot, exact float64 restart); observed traces exact at
alpha = 1; iteration counts equal.  Tests of the opt-in fp32-only mode say so and carry their own
(looser, measured) bounds."""
import multiprocessing as mp

import numpy as np
import pytest

from oracle import pocs_oracle as orc
from oracle.golden_cases import CASES, make_input

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)


@pytest.fixture(scope="module")
def p3d():
    import pseudo_3d_interpolation_b200 as m
    m._lib.require_gpu()
    return m


def _oracle_job(job):
    x, mask, params = job
    return orc.pocs_slice(x.astype(np.complex128), mask, **params).astype(np.complex64)


def oracle_cube(d, fold, **params):
    """float64 oracle over the slices of ``d``, one process per slice (full-size slices take 5 - 90 s each)."""
    mask = orc.mask_from_fold(fold)
    jobs = [(d[i], mask, params) for i in range(d.shape[0])]
    if len(jobs) == 1:
        return np.stack([_oracle_job(jobs[0])])
    with mp.get_context("fork").Pool(min(len(jobs), mp.cpu_count())) as pool:
        return np.stack(pool.map(_oracle_job, jobs, chunksize=1))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_pocs_matches_reference_golden(case, golden, p3d):
    x, mask = make_input(case)
    info = {}
    fn = {"regular": p3d.POCS, "fast": p3d.FPOCS, "adaptive": p3d.APOCS}[case.get("version", "regular")]
    y = fn(x, mask, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT",
           results_dict=info, **case["params"])
    n = case["name"]
    ref = golden[f"{n}__y"]
    assert y.shape == ref.shape
    assert np.iscomplexobj(y) == np.iscomplexobj(ref)
    assert info["niterations"] == int(golden[f"{n}__niterations"])
    assert rel_l2(y, ref) <= RTOL, rel_l2(y, ref)       # (the percentile operators run on complex128 state in the default mode)
    if case["params"]["alpha"] == 1.0 and not case.get("all_zero"):
        obs = mask == 1
        assert np.array_equal(np.asarray(y)[obs], x[obs])     # observed traces reproduced exactly


def test_cost_history_matches(golden, p3d, tmp_path):
    case = CASES[0]
    x, mask = make_input(case)
    path = tmp_path / "costs.out"
    p3d.POCS(x, mask, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT",
             path_results=str(path), **case["params"])
    fields = path.read_text().strip().split(";")
    assert int(fields[0]) == int(golden["hard_exp__niterations"])
    costs = np.array([float(v) for v in fields[2:]])
    ref = golden["hard_exp__costs"]
    assert costs.shape == ref.shape
    big = ref > 1e-10          # costs near the float32 noise floor of sum|x| are not comparable
    np.testing.assert_allclose(costs[big], ref[big], rtol=2e-2)


@pytest.mark.parametrize("shape", [(16, 16), (64, 48), (100, 40), (121, 77), (37, 58), (200, 200), (1, 32), (32, 1), (3, 5)])
def test_fft2_matches_numpy(shape, p3d):
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((3,) + shape) + 1j * rng.standard_normal((3,) + shape)).astype(np.complex64)
    plan = p3d.PocsPlan(shape[0], shape[1])
    X = plan.fft2(x)
    ref = np.fft.fft2(x.astype(np.complex128))
    assert rel_l2(X, ref) < 2e-6
    xb = plan.fft2(X, inverse=True)
    assert rel_l2(xb, x) < 2e-6


@pytest.mark.parametrize("model,kw", [
    ("linear", {}), ("exponential", {}), ("exponential-2", {}), ("data-driven", {}),
    ("inverse_proportional", {}), ("exponential", dict(p_min="adaptive")), ("linear", dict(decay_kind="factors", p_max=3.0, p_min=0.1)),
    ("exponential", dict(sqrt_decay=True)),
])
def test_schedule_matches_oracle(model, kw, p3d):
    case = CASES[4]
    x, mask = make_input(case)
    niter = 13
    plan = p3d.PocsPlan(*x.shape)
    params = dict(niter=niter, thresh_model=model, p_max=0.99, p_min=1e-5)
    params.update(kw)
    tau = plan.schedule(x, **params)[0]
    X0 = np.fft.fft2(x.astype(np.complex128))
    okw = {k: v for k, v in params.items() if k in ("p_max", "p_min", "decay_kind")}
    ref = orc.threshold_table(X0, niter, model, **okw)
    if kw.get("sqrt_decay"):
        ref = np.sqrt(ref)
    np.testing.assert_allclose(tau, ref, rtol=5e-5, atol=1e-6 * np.abs(ref).max())


def test_cube_matches_oracle_config1_shrunk(p3d):
    from pseudo_3d_interpolation_b200 import synth
    d, fold, c = synth.sparse_freq_slices(1, slice_ids=[3, 9, 17, 40], n_il=50, n_xl=40, nt=128)
    params = dict(niter=25, thresh_op=c["thresh_op"], thresh_model=c["thresh_model"], eps=0.0, alpha=c["alpha"],
                  p_max=0.99, p_min=1e-5)
    res = {}
    y = p3d.pocs_cube(d, fold, results=res, transform_kind="FFT", **params)
    ref = orc.pocs_cube(d, fold, **params)
    assert rel_l2(y, ref) <= RTOL
    for s in range(d.shape[0]):
        assert rel_l2(y[s], ref[s]) <= RTOL
    obs = orc.mask_from_fold(fold) == 1
    assert np.array_equal(y[:, obs], d[:, obs])
    assert np.all(res["niterations"] == 25)


@pytest.mark.parametrize("shape,op,model,niter,alpha", [
    ((256, 256), "hard", "exponential", 12, 1.0),
    ((256, 256), "soft", "linear", 10, 0.7),
    ((256, 256), "garrote", "exponential", 10, 1.0),
    ((200, 200), "hard", "exponential", 12, 1.0),
    ((1000, 1000), "hard", "exponential", 8, 1.0),
    ((1000, 256), "garrote", "linear", 6, 1.0),
    ((256, 1000), "soft", "exponential", 6, 0.7),
    ((2000, 200), "hard", "exponential", 5, 1.0),
    ((200, 2000), "hard", "linear", 5, 1.0),
])
def test_specialised_kernels_match_oracle(shape, op, model, niter, alpha, p3d):
    """shapes served by the register-resident kernels (p3d_pocs_spec.cu), vs the float64 oracle
    and vs the generic kernels of the same library.  p_min = 1e-2 keeps the hard-threshold
    cases away from the dense small-coefficient population (well conditioned in fp32)."""
    case = dict(seed=77, shape=shape, keep=0.3, nwaves=5)
    x, mask = make_input(case)
    params = dict(niter=niter, thresh_op=op, thresh_model=model, eps=0.0, alpha=alpha, p_max=0.99,
                  p_min=1e-2 if op == "hard" else 1e-4)
    plan = p3d.PocsPlan(*shape)
    assert "spec<" in plan.describe()
    y, info = plan.run(x, mask, **params)
    ref = orc.pocs_slice(x.astype(np.complex128), mask, **params)
    assert rel_l2(y[0], ref) <= RTOL, rel_l2(y[0], ref)
    if alpha == 1.0:
        assert np.array_equal(y[0][mask == 1], x[mask == 1])
    plan.set_option("force_generic", 1)
    yg, _ = plan.run(x, mask, **params)
    assert rel_l2(y[0], yg[0]) <= 5e-5
    if shape == (1000, 1000):
        plan.set_option("force_generic", 0)
        plan.set_option("spec_variant", 1)
        y1, _ = plan.run(x, mask, **params)
        assert rel_l2(y1[0], ref) <= RTOL


def test_early_exit_many_slices(p3d):
    """per-slice early exit (device-side stop flags) on a spec shape: iteration counts match the oracle."""
    xs, refs, nits = [], [], []
    params = dict(niter=60, thresh_op="soft", thresh_model="exponential", eps=1e-9, alpha=1.0, p_max=0.99, p_min=1e-5)
    mask = None
    for sd in range(5):
        x, m = make_input(dict(seed=11, shape=(256, 256), keep=0.35, nwaves=3 + sd))
        mask = m
        xs.append(x * (1.0 + sd))
    x = np.stack(xs)
    x[2] = 0                                         # an all-zero slice in the middle
    for i in range(x.shape[0]):
        info = {}
        refs.append(orc.pocs_slice(x[i].astype(np.complex128), mask, info=info, **params))
        nits.append(info["niterations"])
    plan = p3d.PocsPlan(256, 256)
    y, info = plan.run(x, mask, **params)
    assert info["niterations"][2] == 0 and np.array_equal(y[2], x[2])
    for i in range(x.shape[0]):
        assert int(info["niterations"][i]) == nits[i], (i, info["niterations"][i], nits[i])
        assert rel_l2(y[i], refs[i]) <= RTOL


def test_fp32_only_mode_flip_statistics(p3d):
    """precision = 32 (opt-in, fastest): fp32 hard thresholding must be as close to the float64 reference as the
    reference algorithm run in complex64 (numpy >= 2 keeps complex64) is -- same typical error,
    and decision flips (error > 1e-4) not more frequent.  The default mode has no flips: next assertions."""
    params = dict(niter=20, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-4)
    plan = p3d.PocsPlan(64, 64, precision=32)
    auto = p3d.PocsPlan(64, 64)
    e_gpu, e_c64 = [], []
    for seed in range(300, 324):
        x, mask = make_input(dict(seed=seed, shape=(64, 64), keep=0.4, nwaves=4))
        ref = orc.pocs_slice(x.astype(np.complex128), mask, **params)
        e_c64.append(rel_l2(orc.pocs_slice(x, mask, **params), ref))
        y, _ = plan.run(x, mask, **params)
        e_gpu.append(rel_l2(y[0], ref))
        ya, _ = auto.run(x, mask, **params)
        assert rel_l2(ya[0], ref) <= 2e-7, (seed, rel_l2(ya[0], ref))        # default mode: one rounding to complex64
    e_gpu, e_c64 = np.array(e_gpu), np.array(e_c64)
    print("gpu   :", np.sort(e_gpu)[[0, 12, -3, -2, -1]], (e_gpu > RTOL).sum())
    print("np c64:", np.sort(e_c64)[[0, 12, -3, -2, -1]], (e_c64 > RTOL).sum())
    assert np.median(e_gpu) <= max(2e-6, 3 * np.median(e_c64))
    assert (e_gpu > RTOL).sum() <= (e_c64 > RTOL).sum() + 4
    assert e_gpu.max() <= max(2e-2, 3 * e_c64.max())


def test_full_size_config_slices(p3d):
    """BASELINE configs 1 and 2 at their real slice sizes, iteration counts and parameters (noise-free synthetic cube,
    hard / exponential, p_min = 1e-5, eps = 0): <= 1e-4 against the float64 oracle in the default mode, observed traces
    exact.  (fp32 alone is 1e-3 .. 3e-3 off here, and so is the reference's own complex64 path: SURVEY 8a-C.)"""
    from pseudo_3d_interpolation_b200 import synth
    for cfg, ids in ((1, [20, 60, 100, 140, 200]), (2, [40, 300, 640, 1000])):
        d, fold, c = synth.sparse_freq_slices(cfg, slice_ids=ids)
        params = dict(niter=c["niter"], thresh_op=c["thresh_op"], thresh_model=c["thresh_model"], eps=0.0,
                      alpha=c["alpha"], p_max=0.99, p_min=1e-5)
        res = {}
        y = p3d.pocs_cube(d, fold, results=res, **params)
        ref = oracle_cube(d, fold, **params)
        e = rel_l2(y, ref)
        print(f"config {cfg}: cube rel-L2 vs float64 oracle {e:.3e}; per slice {[float('%.1e' % rel_l2(y[i], ref[i])) for i in range(len(ids))]}")
        assert e <= RTOL, e
        for i in range(len(ids)):
            assert rel_l2(y[i], ref[i]) <= RTOL, (cfg, ids[i], rel_l2(y[i], ref[i]))
        assert np.all(res["niterations"] == c["niter"])
        obs = orc.mask_from_fold(fold) == 1
        assert np.array_equal(y[:, obs], d[:, obs])


@pytest.mark.parametrize("cfg,ids", [(3, [700]), (4, [1200]), (5, [30, 200, 500])])
def test_full_size_config_slices_c3_c4_c5(cfg, ids, p3d):
    """BASELINE configs 3 (1201 x 847, soft / linear, 100 iterations), 4 (2000 x 2000, hard / data-driven, alpha = 0.7,
    100 iterations, line-pattern fold with fold = 2 crossings) and 5 (256 x 256, garrote / exponential, 30 iterations) at
    their real sizes and parameters (p_min = 1e-5, eps = 0) against the float64 oracle: <= 1e-4."""
    from pseudo_3d_interpolation_b200 import synth
    d, fold, c = synth.sparse_freq_slices(cfg, slice_ids=ids)
    assert (fold.max() == 2) == (cfg == 4)
    params = dict(niter=c["niter"], thresh_op=c["thresh_op"], thresh_model=c["thresh_model"], eps=0.0,
                  alpha=c["alpha"], p_max=0.99, p_min=1e-5)
    res = {}
    y = p3d.pocs_cube(d, fold, results=res, **params)
    ref = oracle_cube(d, fold, **params)
    for i in range(len(ids)):
        e = rel_l2(y[i], ref[i])
        print(f"config {cfg} slice {ids[i]}: rel-L2 vs float64 oracle {e:.3e}")
        assert e <= RTOL, (cfg, ids[i], e)
    assert np.all(res["niterations"] == c["niter"])
    if c["alpha"] == 1.0:
        obs = orc.mask_from_fold(fold) == 1
        assert np.array_equal(y[:, obs], d[:, obs])


@pytest.mark.parametrize("cfg,ids,niter", [(1, [20, 100, 200], 50), (2, [300, 700], 25), (3, [500], 12), (4, [900], 8), (5, [40, 400, 333], 30)])
def test_full_size_properties_all_configs(cfg, ids, niter, p3d):
    """Size-independent properties at the real slice sizes of all five BASELINE configs (each config's own operator,
    schedule and alpha; iteration counts shortened where a slice is 4 M points):
      * scale equivariance, bit for bit: POCS(2 x) == 2 POCS(x) (thresholds are p * z(x); powers of two are exact in fp32);
      * batch independence, bit for bit: a slice's result does not depend on which other slices share the call;
      * observed traces reproduced exactly when alpha = 1, the result is finite, unobserved traces get filled."""
    from pseudo_3d_interpolation_b200 import synth
    d, fold, c = synth.sparse_freq_slices(cfg, slice_ids=ids)
    params = dict(niter=niter, thresh_op=c["thresh_op"], thresh_model=c["thresh_model"], eps=0.0, alpha=c["alpha"], p_max=0.99, p_min=1e-5)
    mask = orc.mask_from_fold(fold)
    plan = p3d.PocsPlan(c["n_il"], c["n_xl"])
    y, info = plan.run(d, mask, **params)
    assert list(info["niterations"]) == [niter] * len(ids) and np.isfinite(y.view(np.float32)).all()
    y2, _ = plan.run((2 * d).astype(np.complex64), mask, **params)
    assert np.array_equal(y2, 2 * y)
    y1, _ = plan.run(d[-1:], mask, **params)
    assert np.array_equal(y1[0], y[-1])
    obs = mask == 1
    if c["alpha"] == 1.0:
        assert np.array_equal(y[:, obs], d[:, obs])
    assert np.abs(y[:, ~obs]).max() > 0


def test_full_size_config_slices_soft(p3d):
    """Soft operator on config-1 slices at full size.  With the reference's complex tau (Q1) even the soft and garrote
    operators jump at a threshold modulus (below it the factor is clipped to 0, at it the factor is -i Im(tau)/|X|), so
    fp32 alone flips single coefficients here too (2e-4 measured); the default mode holds 1e-4."""
    from pseudo_3d_interpolation_b200 import synth
    d, fold, c = synth.sparse_freq_slices(1, slice_ids=[20, 100, 200])
    for op in ("soft", "garrote"):
        params = dict(niter=c["niter"], thresh_op=op, thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-5)
        y = p3d.pocs_cube(d, fold, **params)
        ref = oracle_cube(d, fold, **params)
        print(f"{op} full size: rel-L2 {rel_l2(y, ref):.3e}")
        assert rel_l2(y, ref) <= RTOL


# ------------------------------------------------------------------------------------------------
# float64 state mode: parity with the reference's float64 results at the 1e-4 of BASELINE.json
# (in fact at complex64 rounding level) on every case, including the ill-conditioned ones
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_f64_mode_matches_reference_golden(case, golden, p3d):
    x, mask = make_input(case)
    n = case["name"]
    ref = golden[f"{n}__y"]
    plan = p3d.PocsPlan(x.shape[0], x.shape[1], precision=64)
    y, info = plan.run(x.astype(np.complex64), mask, version=case.get("version", "regular"), want_costs=True, **case["params"])
    nit = int(info["niterations"][0])
    assert nit == int(golden[f"{n}__niterations"])
    if nit == 0:
        assert np.array_equal(y[0], x.astype(np.complex64))
        return
    got = y[0] if np.iscomplexobj(ref) else y[0].real
    assert rel_l2(got, ref) <= 2e-7, rel_l2(got, ref)            # = one rounding to complex64
    costs = info["costs"][0][:nit]
    refc = golden[f"{n}__costs"]
    big = refc > 1e-24
    np.testing.assert_allclose(costs[big], refc[big], rtol=1e-6)


def test_f64_mode_full_size_config_slices(p3d):
    """configs 1 (5 slices, hard and soft) and 2 (1 slice) at full size: <= 1e-4 vs float64 reference."""
    from pseudo_3d_interpolation_b200 import synth
    for cfg, ids, ops in ((1, [20, 60, 100, 140, 200], ("hard", "soft")), (2, [300], ("hard",))):
        d, fold, c = synth.sparse_freq_slices(cfg, slice_ids=ids)
        for op in ops:
            params = dict(niter=c["niter"], thresh_op=op, thresh_model=c["thresh_model"], eps=0.0, alpha=c["alpha"], p_max=0.99, p_min=1e-5)
            y = p3d.pocs_cube(d, fold, precision=64, **params)
            ref = orc.pocs_cube(d, fold, **params)
            e = rel_l2(y, ref)
            print(f"f64 mode config {cfg} {op}: cube rel-L2 {e:.3e}")
            assert e <= RTOL, e
            obs = orc.mask_from_fold(fold) == 1
            assert np.array_equal(y[:, obs], d[:, obs])


def test_f64_mode_ill_conditioned_spec_shape(p3d):
    """256 x 256 hard / exponential, p_min = 1e-4, 12 iterations: fp32 (and the float64 oracle under a
    1e-7 input perturbation) moves by 5e-3 here; the float64 mode reproduces the reference."""
    x, mask = make_input(dict(seed=77, shape=(256, 256), keep=0.3, nwaves=5))
    params = dict(niter=12, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-4)
    ref = orc.pocs_slice(x.astype(np.complex128), mask, **params)
    y, _ = p3d.PocsPlan(256, 256, precision=64).run(x, mask, **params)
    assert rel_l2(y[0], ref) <= 2e-7, rel_l2(y[0], ref)


@pytest.mark.parametrize("shape", [(200, 200), (256, 256), (200, 256), (1000, 48), (48, 2000)])
@pytest.mark.parametrize("op,model,alpha,version", [("hard", "exponential", 1.0, "regular"), ("soft", "linear", 0.7, "regular"),
                                                     ("garrote", "data-driven", 1.0, "adaptive")])
def test_f64_mode_register_kernels_match_generic(shape, op, model, alpha, version, p3d):
    """float64 state mode: the register-resident complex128 iteration kernels (1000 / 2000 / 256 / 200, also
    mixed with a generic axis) against the generic complex128 kernels and the float64 oracle."""
    x, mask = make_input(dict(seed=5, shape=shape, keep=0.3, nwaves=4))
    x = np.stack([x, 0.5 * x[::-1, ::-1] * mask]).astype(np.complex64)
    params = dict(niter=9, thresh_op=op, thresh_model=model, eps=0.0, alpha=alpha, p_max=0.99, p_min=1e-3)
    plan = p3d.PocsPlan(*shape, precision=64)
    assert "spec64<" in plan.describe() or "mix64<" in plan.describe()
    y, info = plan.run(x, mask, version=version, want_costs=True, **params)
    gen = p3d.PocsPlan(*shape, precision=64)
    gen.set_option("force_generic", 1)
    assert "spec64<" not in gen.describe() and "mix64<" not in gen.describe()
    yg, infog = gen.run(x, mask, version=version, want_costs=True, **params)
    assert rel_l2(y, yg) <= 2e-7
    np.testing.assert_allclose(info["costs"], infog["costs"], rtol=1e-3, atol=1e-22)   # tiny costs = cancelling sums
    for i in range(2):
        ref = orc.pocs_slice(x[i].astype(np.complex128), mask, version=version, **params)
        assert rel_l2(y[i], ref) <= 2e-7, rel_l2(y[i], ref)
    if alpha == 1.0 and version == "regular":
        obs = mask == 1
        assert np.array_equal(y[:, obs], x[:, obs])


def test_f64_mode_data_driven_and_schedule(p3d):
    case = CASES[4]
    x, mask = make_input(case)
    plan = p3d.PocsPlan(*x.shape, precision=64)
    X0 = np.fft.fft2(x.astype(np.complex128))
    for model in ("data-driven", "exponential", "linear", "inverse_proportional"):
        tau = plan.schedule(x, niter=11, thresh_model=model, p_max=0.99, p_min=1e-5)[0]
        ref = orc.threshold_table(X0, 11, model, 0.99, 1e-5)
        np.testing.assert_allclose(tau, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("precision", ["auto", 32])
@pytest.mark.parametrize("shape,op", [((256, 256), "soft-percentile"), ((200, 120), "garrote-percentile"), ((64, 1000), "hard-percentile")])
def test_percentile_operators_spec_shapes(shape, op, precision, p3d):
    """'<op>-percentile' (functions/POCS.py:43-58) on register-resident / mixed plans, several slices per call
    (each slice gets its own per-iteration percentile), host-buffer path.  The default mode runs them on complex128
    state (plain 1e-4); precision=32 is the fast fp32 path."""
    x, mask = make_input(dict(seed=31, shape=shape, keep=0.35, nwaves=5, noise=0.01))
    x = np.stack([x, 0.3 * x, np.conj(x)]).astype(np.complex64)
    params = dict(niter=8, thresh_op=op, thresh_model="exponential", eps=0.0, alpha=1.0, p_max=99.5, p_min=30.0, decay_kind="factors")
    y, info = p3d.PocsPlan(*shape, precision=precision).run(x, mask, **params)
    assert list(info["niterations"]) == [8, 8, 8]
    for i in range(3):
        ref = orc.pocs_slice(x[i].astype(np.complex128), mask, **params)
        e = rel_l2(y[i], ref)
        # fp32 only, hard: a coefficient within fp32 rounding of the percentile value may fall on either side
        assert e <= (2e-3 if (precision == 32 and op.startswith("hard")) else RTOL), (i, e)


@pytest.mark.parametrize("shape", [(60, 847), (1201, 48), (1201, 847), (48, 1201), (847, 1201)])
@pytest.mark.parametrize("op,model,alpha,version,eps", [("soft", "linear", 1.0, "regular", 0.0), ("garrote", "exponential", 0.7, "adaptive", 0.0),
                                                         ("hard", "data-driven", 1.0, "regular", 0.0), ("soft", "exponential", 1.0, "regular", 1e-6)])
def test_config3_plans_match_generic_and_oracle(shape, op, model, alpha, version, eps, p3d):
    """config 3's axes: xline 847 = 11 x 7 x 11 (three-pass mixed-radix register plan), iline 1201 (prime: Rader's
    algorithm over two 1200-point register transforms), alone (other axis generic) and together, against the
    generic shared-memory path (Bluestein for 1201) and the float64 oracle."""
    x, mask = make_input(dict(seed=9, shape=shape, keep=0.3, nwaves=5))
    x = np.stack([x, 0.25 * np.conj(x)]).astype(np.complex64)
    params = dict(niter=7, thresh_op=op, thresh_model=model, eps=eps, alpha=alpha, p_max=0.99, p_min=1e-3)
    plan = p3d.PocsPlan(*shape)
    d = plan.describe()
    dc, dr = d.split("cols_iter=")[1].split(";")[0], d.split("rows_iter=")[1].split(";")[0]
    assert ("rader<1201" in dc) == (shape[0] == 1201) and ("mix<847" in dc) == (shape[0] == 847), d
    assert ("rader<1201" in dr) == (shape[1] == 1201) and ("mix<847" in dr) == (shape[1] == 847), d
    y, info = plan.run(x, mask, version=version, want_costs=True, **params)
    gen = p3d.PocsPlan(*shape)
    gen.set_option("force_generic", 1)
    yg, infog = gen.run(x, mask, version=version, want_costs=True, **params)
    assert list(info["niterations"]) == list(infog["niterations"])
    assert rel_l2(y, yg) <= RTOL, rel_l2(y, yg)
    for i in range(2):
        oinfo = {}
        ref = orc.pocs_slice(x[i].astype(np.complex128), mask, version=version, info=oinfo, **params)
        assert oinfo["niterations"] == info["niterations"][i]
        assert rel_l2(y[i], ref) <= RTOL, (i, rel_l2(y[i], ref))
    if alpha == 1.0 and version == "regular":
        obs = mask == 1
        assert np.array_equal(y[:, obs], x[:, obs])


MORE_SIZES = [128, 512, 1024, 2048, 400, 500, 800, 1600, 300, 600, 700, 900, 1100, 1300, 1200, 2400,
              1400, 1500, 1800, 2100, 2200, 3000, 2500, 4096, 768, 1280, 1536, 1792, 2304, 2560, 3072, 847]


@pytest.mark.parametrize("n", MORE_SIZES)
def test_more_register_plans_match_oracle(n, p3d):
    """every further length with a register plan (p3d_pocs_spec_more.cu, p3d_pocs_spec_mix.cu), as the iline axis and
    as the xline axis (the other axis runs the generic kernels), against the float64 oracle."""
    for shape in ((n, 24), (20, n)):
        x, mask = make_input(dict(seed=n, shape=shape, keep=0.3, nwaves=4))
        x = np.stack([x, 0.5 * np.conj(x)]).astype(np.complex64)
        params = dict(niter=5, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=0.8, p_max=0.99, p_min=1e-3)
        plan = p3d.PocsPlan(*shape)
        d = plan.describe()
        key = "cols_iter=" if shape[0] == n else "rows_iter="
        assert f"<{n}," in d.split(key)[1].split(";")[0], d
        y, info = plan.run(x, mask, **params)
        for i in range(2):
            ref = orc.pocs_slice(x[i].astype(np.complex128), mask, **params)
            assert rel_l2(y[i], ref) <= RTOL, (shape, i, rel_l2(y[i], ref))
    # square slice of this size: both axes on register plans, hard threshold, observed traces exact
    if n <= 1024:
        x, mask = make_input(dict(seed=n + 1, shape=(n, n), keep=0.25, nwaves=5))
        params = dict(niter=4, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-2)
        y, _ = p3d.PocsPlan(n, n).run(x.astype(np.complex64), mask, **params)
        ref = orc.pocs_slice(x.astype(np.complex128), mask, **params)
        assert rel_l2(y[0], ref) <= RTOL
        assert np.array_equal(y[0][mask == 1], x.astype(np.complex64)[mask == 1])


def test_per_cube_masks(p3d):
    """config-5 style batch: independent cubes, each with its own mask (slices_per_mask)."""
    rng = np.random.default_rng(3)
    n1, n2, per, ncubes = 32, 32, 3, 4
    xs, masks = [], []
    for cidx in range(ncubes):
        case = dict(seed=100 + cidx, shape=(n1, n2), keep=0.3)
        for s in range(per):
            x, m = make_input(dict(case, seed=100 + cidx))
            xs.append(x * (1 + 0.1 * s))
        masks.append(m)
    x = np.stack(xs).astype(np.complex64)
    mask = np.stack(masks)
    plan = p3d.PocsPlan(n1, n2)
    params = dict(niter=12, thresh_op="garrote", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-4)
    y, info = plan.run(x, mask, slices_per_mask=per, **params)
    for i in range(x.shape[0]):
        ref = orc.pocs_slice(x[i].astype(np.complex128), mask[i // per], **params)
        assert rel_l2(y[i], ref) <= RTOL


def test_chunked_and_banded_equals_single(p3d):
    """band scheduler / chunk pipeline must not change results."""
    from pseudo_3d_interpolation_b200 import synth
    d, fold, c = synth.sparse_freq_slices(2, slice_ids=list(range(2, 21)), n_il=40, n_xl=48, nt=128)
    params = dict(niter=10, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=0.7, p_max=0.99, p_min=1e-4)
    mask = orc.mask_from_fold(fold)
    a = p3d.PocsPlan(40, 48)
    ya, _ = a.run(d, mask, **params)
    b = p3d.PocsPlan(40, 48, max_slices=4, band_slices=3)
    yb, _ = b.run(d, mask, **params)
    assert np.array_equal(ya, yb)


def test_errors(p3d):
    x = np.ones((8, 8), dtype=np.complex64)
    m = np.ones((8, 8), dtype=np.uint8)
    kw = dict(transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT")
    with pytest.raises(ValueError):
        p3d.POCS(x, m * 2, None, **kw)
    with pytest.raises(ValueError):
        p3d.POCS(x, m, None, transform=None, itransform=None, transform_kind="FFT")
    with pytest.raises(ValueError):
        p3d.POCS(x, m, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="nope")
    with pytest.raises(NotImplementedError):
        p3d.POCS(x, m, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="WAVELET")
    with pytest.raises(NotImplementedError):
        p3d.POCS(x, m, None, thresh_model="quadratic", **kw)


def test_time_axis_roundtrip_and_oracle(p3d):
    from oracle import time_axis_oracle as tor
    from pseudo_3d_interpolation_b200 import timeaxis, synth
    rng = np.random.default_rng(1)
    for nt, shape, real, up in [(64, (5, 7), True, 1), (64, (5, 7), False, 1), (100, (3, 4), True, 1),
                                (50, (4, 5), False, 2), (63, (2, 3), True, 1), (74, (3, 3), True, 1)]:
        x = rng.standard_normal((nt,) + shape).astype(np.float32)
        twt = synth.T0_MS + synth.DT_MS * np.arange(nt)
        F, f = timeaxis.time_fft(x, twt, compute_real=real, upsampling_factor=up)
        Fr, fr = tor.time_fft(x, twt, compute_real=real, upsampling_factor=up)
        assert F.shape == Fr.shape
        np.testing.assert_allclose(f, fr)
        assert rel_l2(F, Fr) < 5e-6
        nte = nt - (nt % 2)
        Fin = np.fft.fftshift(F, axes=0) if not real else F       # step 13 leaves the axis ascending
        xb = timeaxis.time_ifft(Fin, synth.DT_MS, synth.T0_MS, compute_real=real, ascending=True, nt_out=nte)
        xr = tor.time_ifft(np.fft.fftshift(Fr, axes=0) if not real else Fr, synth.DT_MS, synth.T0_MS, compute_real=real)
        assert rel_l2(xb, xr[:nte]) < 5e-6
        assert rel_l2(xb, x[:nte]) < 5e-6


def test_time_axis_window(p3d):
    from oracle import time_axis_oracle as tor
    from pseudo_3d_interpolation_b200 import timeaxis, synth
    rng = np.random.default_rng(2)
    nt = 128
    x = rng.standard_normal((nt, 4, 6)).astype(np.float32)
    twt = synth.T0_MS + synth.DT_MS * np.arange(nt)
    f = np.fft.rfftfreq(nt, synth.DT_MS)
    for ftype, freqs in (("lowpass", [4.0, 6.0]), ("highpass", [1.0, 2.0]), ("bandpass", [1.0, 2.0, 5.0, 7.0])):
        w = timeaxis.freq_filter_window(list(freqs), f, ftype)
        wr = tor.freq_filter_window(list(freqs), f, ftype)
        np.testing.assert_array_equal(w, wr)
        F, _ = timeaxis.time_fft(x, twt, compute_real=True, window=w)
        Fr, _ = tor.time_fft(x, twt, compute_real=True, window=wr)
        assert rel_l2(F, Fr) < 5e-6


@pytest.mark.parametrize("nt,shape,real,up", [(512, (9, 7), True, 1), (512, (9, 7), False, 1), (256, (5, 11), True, 2),
                                              (1024, (6, 5), True, 1), (2048, (3, 5), False, 1), (4096, (2, 3), True, 1),
                                              (300, (4, 5), True, 1), (1201, (2, 2), False, 1),
                                              (1000, (7, 9), True, 1), (1000, (3, 5), False, 1), (2000, (5, 3), True, 1), (2500, (3, 3), True, 1),
                                              (3000, (5, 7), True, 1), (3000, (2, 3), False, 1), (4000, (2, 5), True, 1), (5000, (3, 1), True, 1),
                                              (1500, (3, 3), True, 2)])
def test_time_axis_register_pipeline_sizes(nt, shape, real, up, p3d):
    """record lengths served by the transposing register-resident pipeline (512..4096, here with
    odd trace counts and zero padding) and generic / Bluestein lengths, vs the numpy oracle."""
    from oracle import time_axis_oracle as tor
    from pseudo_3d_interpolation_b200 import timeaxis, synth
    rng = np.random.default_rng(nt)
    x = rng.standard_normal((nt,) + shape).astype(np.float32)
    twt = synth.T0_MS + synth.DT_MS * np.arange(nt)
    F, f = timeaxis.time_fft(x, twt, compute_real=real, upsampling_factor=up)
    Fr, fr = tor.time_fft(x, twt, compute_real=real, upsampling_factor=up)
    assert F.shape == Fr.shape and rel_l2(F, Fr) < 5e-6
    nte = nt - (nt % 2)
    Fin = np.fft.fftshift(F, axes=0) if not real else F
    xb = timeaxis.time_ifft(Fin, synth.DT_MS / 1.0, synth.T0_MS, compute_real=real, ascending=True, nt_out=nte)
    if up == 1:
        assert rel_l2(xb, x[:nte]) < 5e-6
    else:
        xr = tor.time_ifft(np.fft.fftshift(Fr, axes=0) if not real else Fr, synth.DT_MS, synth.T0_MS, compute_real=real)
        assert rel_l2(xb, xr[:nte]) < 5e-6


@pytest.mark.parametrize("path", ["tma", "pipeline", "direct"])
@pytest.mark.parametrize("nt,shape,real,up", [(2048, (12, 11), True, 1), (2048, (4, 9), False, 1), (1024, (8, 25), True, 1), (512, (16, 23), False, 1),
                                              (256, (32, 7), True, 2), (4096, (4, 5), True, 1), (4096, (2, 6), False, 1),
                                              (1000, (8, 21), True, 1), (1000, (4, 17), False, 1), (2000, (4, 13), True, 1), (2500, (4, 9), True, 1),
                                              (4000, (4, 7), True, 1), (500, (4, 5), True, 2), (1024, (4, 3), False, 2)])
def test_time_axis_paths(nt, shape, real, up, path, p3d, monkeypatch):
    """The three implementations of the time-axis transforms (one pass with TMA-staged tiles - the default -, the
    transposing pipeline, the direct register kernels) against the numpy oracle, on trace counts divisible by 4 (what the
    one-pass kernels need) with a partial last tile, and with so few CTAs that every CTA walks over several tiles
    (both shared-memory stages, both mbarrier phases).  up = 2: zero padding to the plan length."""
    from oracle import time_axis_oracle as tor
    from pseudo_3d_interpolation_b200 import timeaxis, synth, _lib
    monkeypatch.setenv("P3D_TIME_PATH", path)
    monkeypatch.setenv("P3D_TIME_GRID", "3")
    rng = np.random.default_rng(nt + shape[1])
    x = rng.standard_normal((nt,) + shape).astype(np.float32)
    twt = synth.T0_MS + synth.DT_MS * np.arange(nt)
    kw = dict(upsampling_factor=up)
    F, f = timeaxis.time_fft(x, twt, compute_real=real, **kw)
    used = _lib.load().p3d_time_last_path().decode()
    n_plan = nt * up
    if path == "tma":
        assert used == "tma", used
    elif path == "pipeline":
        assert used == "pipeline", used
    else:
        assert used == ("direct" if n_plan in (512, 1024, 2048, 4096, 2000, 4000) else "generic"), used
    Fr, fr = tor.time_fft(x, twt, compute_real=real, **kw)
    assert F.shape == Fr.shape and rel_l2(F, Fr) < 5e-6
    np.testing.assert_allclose(f, fr)
    nte = nt - (nt % 2)
    Fin = np.fft.fftshift(F, axes=0) if not real else F
    xb = timeaxis.time_ifft(Fin, synth.DT_MS, synth.T0_MS, compute_real=real, ascending=True, nt_out=nte)
    used = _lib.load().p3d_time_last_path().decode()
    # the inverse of a two-sided spectrum needs twice the stage: the pipeline takes it for the lengths whose tiles are two pairs wide
    assert used == path or path == "direct" or (path == "tma" and not real and n_plan >= 2500 and used == "pipeline"), used
    xr = tor.time_ifft(np.fft.fftshift(Fr, axes=0) if not real else Fr, synth.DT_MS, synth.T0_MS, compute_real=real)
    assert rel_l2(xb, xr[:nte]) < 5e-6
    if n_plan == nt:
        assert rel_l2(xb, x[:nte]) < 5e-6


@pytest.mark.parametrize("nt,ntr", [(2048, 20004), (1024, 30000), (2000, 12004)])
def test_time_axis_one_pass_many_tiles_per_cta(nt, ntr, p3d):
    """default path and grid (one persistent CTA per SM): every CTA walks over a dozen tiles; against numpy's rfft / irfft"""
    from pseudo_3d_interpolation_b200 import timeaxis, synth, _lib
    rng = np.random.default_rng(nt)
    x = rng.standard_normal((nt, ntr)).astype(np.float32)
    twt = synth.T0_MS + synth.DT_MS * np.arange(nt)
    F, f = timeaxis.time_fft(x.reshape(nt, ntr, 1), twt, compute_real=True)
    assert _lib.load().p3d_time_last_path().decode() == "tma"
    ph = synth.DT_MS * np.exp(-2j * np.pi * f * synth.T0_MS)
    Fr = np.fft.rfft(x.astype(np.float64), axis=0) * ph[:, None]
    assert rel_l2(F[:, :, 0], Fr) < 5e-6
    xb = timeaxis.time_ifft(F, synth.DT_MS, synth.T0_MS, compute_real=True, nt_out=nt)
    assert _lib.load().p3d_time_last_path().decode() == "tma"
    assert rel_l2(xb[:, :, 0], x) < 5e-6


# ------------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------------
def test_edge_empty_single_and_degenerate_shapes(p3d):
    plan = p3d.PocsPlan(12, 10)
    y, info = plan.run(np.zeros((0, 12, 10), np.complex64), np.ones((12, 10), np.uint8), niter=5)
    assert y.shape == (0, 12, 10) and info["niterations"].shape == (0,)
    for shape in [(1, 32), (32, 1), (2, 3), (7, 1)]:
        x, mask = make_input(dict(seed=3, shape=shape, keep=0.6))
        params = dict(niter=6, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-2)
        ref = orc.pocs_slice(x.astype(np.complex128), mask, **params)
        got = p3d.POCS(x, mask, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT", **params)
        assert rel_l2(got, ref) <= RTOL, (shape, rel_l2(got, ref))


def test_edge_full_and_empty_masks(p3d):
    x, _ = make_input(dict(seed=4, shape=(40, 36), keep=1.0))
    ones = np.ones(x.shape, np.uint8)
    params = dict(niter=8, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)
    y = p3d.POCS(x, ones, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT", **params)
    assert np.array_equal(y, x)                                  # nothing missing: observed data returned exactly
    zeros = np.zeros(x.shape, np.uint8)                          # nothing "observed": x is still re-inserted (alpha * x)
    y0 = p3d.POCS(x, zeros, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT", **params)
    ref0 = orc.pocs_slice(x.astype(np.complex128), zeros, **params)
    assert rel_l2(y0, ref0) <= RTOL


def test_edge_niter_one_and_dtypes(p3d):
    x, mask = make_input(dict(seed=6, shape=(24, 30), keep=0.5))
    kw = dict(transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT")
    # niter = 1: (niter - 1) = 0 makes the reference's multiplier 0/0 = nan -> tau = nan -> nothing thresholded
    with np.errstate(all="ignore"):
        ref1 = orc.pocs_slice(x.astype(np.complex128), mask, niter=1, thresh_op="hard", thresh_model="linear", eps=0.0)
    info = {}
    y1 = p3d.POCS(x, mask, None, niter=1, thresh_op="hard", thresh_model="linear", eps=0.0, results_dict=info, **kw)
    assert info["niterations"] == 1 and rel_l2(y1, ref1) <= RTOL
    # complex128 / float64 / Fortran-ordered inputs keep their dtype on return
    y128 = p3d.POCS(np.asfortranarray(x.astype(np.complex128)), mask, None, niter=5, **kw)
    assert y128.dtype == np.complex128
    xr = np.asfortranarray(x.real.astype(np.float64))
    yr = p3d.POCS(xr, mask, None, niter=5, **kw)
    assert yr.dtype == np.float64 and not np.iscomplexobj(yr)
    refr = orc.pocs_slice(xr, mask, niter=5)
    assert rel_l2(yr, refr) <= RTOL


def test_edge_many_small_cubes_with_own_masks_and_early_exit(p3d):
    """config-5 layout in miniature: a batch of cubes, each with its own mask, default eps."""
    per, ncubes, shape = 5, 6, (32, 32)
    xs, masks = [], []
    for cidx in range(ncubes):
        x, m = make_input(dict(seed=500 + cidx, shape=shape, keep=0.25 + 0.05 * cidx, nwaves=3))
        masks.append(m)
        xs += [x * (1 + 0.2 * s) for s in range(per)]
    x = np.stack(xs).astype(np.complex64)
    mask = np.stack(masks)
    params = dict(niter=40, thresh_op="garrote", thresh_model="exponential", eps=1e-9, alpha=1.0, p_max=0.99, p_min=1e-4)
    plan = p3d.PocsPlan(*shape, max_slices=7)             # forces several chunks that straddle cube boundaries
    y, info = plan.run(x, mask, slices_per_mask=per, **params)
    for i in range(x.shape[0]):
        oi = {}
        ref = orc.pocs_slice(x[i].astype(np.complex128), mask[i // per], info=oi, **params)
        assert int(info["niterations"][i]) == oi["niterations"], (i, info["niterations"][i], oi["niterations"])
        assert rel_l2(y[i], ref) <= RTOL, (i, rel_l2(y[i], ref))


@pytest.mark.gpu
def test_run_to_run_determinism_all_kernel_families():
    """every kernel family (register plans, mixed radix + Rader, generic + Bluestein, percentile, float64, kx-ky
    filter, envelope, time axis incl. the one-pass TMA kernels) repeated on the same input: bit-identical results
    (tools/sanitize_cases.py)."""
    import os
    import runpy
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    argv, sys.argv = sys.argv, ["sanitize_cases.py", "all"]
    try:
        glb = runpy.run_path(os.path.join(root, "tools", "sanitize_cases.py"), run_name="__main__")
    finally:
        sys.argv = argv
    assert glb["only"] == "all" and glb["CASES_RUN"] >= 17


# ------------------------------------------------------------------------------------------------
# escalating mode: the paths that real data rarely take
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,op,watch", [((200, 200), "hard", 1), ((200, 200), "hard", 0), ((256, 256), "garrote", 1),
                                            ((100, 40), "soft", 0), ((60, 847), "hard", 1)])
def test_escalating_mode_restart_from_any_verified_iteration(shape, op, watch, p3d):
    """The replay verifies the pilot's decisions; a failed verification at iteration i restarts complex128 from the last
    verified iterate.  Forced here at iterations 0, 3 and 9 for every slice (plan option `debug_fail_iter`), with the
    watch list on and off, register and generic kernels, an all-zero slice and an early-exit run in the batch: the
    result must still be the float64 mode's (a restart from ANY verified iteration is exact)."""
    x, mask = make_input(dict(seed=21, shape=shape, keep=0.3, nwaves=5, noise=0.01))
    x = np.stack([x, 0.5 * np.conj(x), np.zeros_like(x), 2.0 * x[::-1, ::-1] * mask]).astype(np.complex64)
    for eps in (0.0, 1e-7):
        params = dict(niter=24, thresh_op=op, thresh_model="exponential", eps=eps, alpha=1.0, p_max=0.99, p_min=1e-5)
        ref, iref = p3d.PocsPlan(*shape, precision=64).run(x, mask, **params)
        for fail_at in (-1, 0, 3, 9):
            plan = p3d.PocsPlan(*shape)
            plan.set_option("pilot_min_elems", 0)
            plan.set_option("watch_mode", watch)
            plan.set_option("debug_fail_iter", fail_at)
            y, info = plan.run(x, mask, **params)
            assert list(info["niterations"]) == list(iref["niterations"]), (fail_at, eps)
            assert rel_l2(y, ref) <= 2e-7, (fail_at, eps, rel_l2(y, ref))
            assert np.array_equal(y[2], x[2]) and info["niterations"][2] == 0
            plan.close()


def test_escalating_mode_many_slices_in_one_call(p3d):
    """more slices than one launch takes (band_max = 32768): the compacted lists are split across launches."""
    x, mask = make_input(dict(seed=4, shape=(8, 12), keep=0.5, nwaves=2))
    n = 33000
    scale = (1.0 + (np.arange(n) % 17) / 16.0).astype(np.float32)
    xs = (x.astype(np.complex64)[None] * scale[:, None, None]).astype(np.complex64)
    params = dict(niter=6, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)
    plan = p3d.PocsPlan(8, 12)
    plan.set_option("pilot_min_elems", 0)
    y, info = plan.run(xs, mask, **params)
    assert np.all(info["niterations"] == 6)
    for i in (0, 5, 16, 32767, 32768, 32999):
        ref = orc.pocs_slice(xs[i].astype(np.complex128), mask, **params)
        assert rel_l2(y[i], ref) <= RTOL, (i, rel_l2(y[i], ref))
    # slices that differ by a power-of-two factor are bit-identical up to that factor
    assert np.array_equal(y[16], 2 * y[0])
