"""Case list shared by ``make_golden.py`` (runs the reference) and the tests (replay).

Every case is regenerated from its seed, so the fixtures only need to store the
reference's outputs.  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np


def make_input(case):
    """Seeded sparse slice: a few plane waves (+ optional noise), Bernoulli mask."""
    rng = np.random.default_rng(case["seed"])
    n1, n2 = case["shape"]
    i = np.arange(n1)[:, None]
    j = np.arange(n2)[None, :]
    d = np.zeros((n1, n2), dtype=np.complex128)
    for _ in range(case.get("nwaves", 4)):
        k1 = rng.uniform(-0.2, 0.2)
        k2 = rng.uniform(-0.2, 0.2)
        amp = rng.uniform(0.3, 1.0) * np.exp(2j * np.pi * rng.random())
        d += amp * np.exp(2j * np.pi * (k1 * i + k2 * j))
    if case.get("noise", 0.0) > 0:
        d += case["noise"] * (rng.standard_normal(d.shape) + 1j * rng.standard_normal(d.shape))
    mask = (rng.random((n1, n2)) < case.get("keep", 0.4)).astype(np.uint8)
    if case.get("real_input"):
        d = d.real.copy()
    if case.get("all_zero"):
        d = np.zeros_like(d)
    x = d * mask
    # quantise like the real pipeline (complex64 / float32 on disk), then upcast: the
    # float64 reference of BASELINE.json is the reference run on this upcast input.
    x = x.astype(np.float32 if case.get("real_input") else np.complex64)
    return x, mask


CASES = [
    dict(name="hard_exp", seed=11, shape=(48, 40), params=dict(niter=20, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-5)),
    dict(name="hard_exp_pmin1e-3_noise", seed=12, shape=(64, 64), noise=0.02, params=dict(niter=30, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)),
    dict(name="soft_linear_prime", seed=13, shape=(53, 47), params=dict(niter=15, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-5)),
    dict(name="garrote_exp_earlyexit", seed=14, shape=(64, 64), params=dict(niter=80, thresh_op="garrote", thresh_model="exponential", eps=1e-9, alpha=1.0, p_max=0.99, p_min=1e-5)),
    dict(name="hard_datadriven_a07", seed=15, shape=(50, 60), noise=0.01, params=dict(niter=20, thresh_op="hard", thresh_model="data-driven", eps=0.0, alpha=0.7, p_max=0.99, p_min=1e-5)),
    dict(name="soft_exp_adaptive_pmin", seed=16, shape=(40, 56), noise=0.01, params=dict(niter=20, thresh_op="soft", thresh_model="exponential", eps=0.0, alpha=0.7, p_max=0.99, p_min="adaptive")),
    dict(name="real_input_hard", seed=17, shape=(40, 40), real_input=True, params=dict(niter=20, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-4)),
    dict(name="real_input_soft", seed=18, shape=(36, 44), real_input=True, params=dict(niter=12, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)),
    dict(name="apocs_a08", seed=19, shape=(48, 48), version="adaptive", params=dict(niter=20, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=0.8, p_max=0.99, p_min=1e-4)),
    dict(name="fpocs", seed=11, shape=(48, 40), version="fast", params=dict(niter=20, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-5)),
    dict(name="sqrt_decay_factors", seed=20, shape=(32, 48), params=dict(niter=10, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=1.0, p_max=4.0, p_min=0.25, sqrt_decay=True, decay_kind="factors")),
    dict(name="inverse_proportional", seed=21, shape=(40, 40), params=dict(niter=15, thresh_op="hard", thresh_model="inverse_proportional", eps=0.0, alpha=1.0)),
    dict(name="exponential_q2", seed=22, shape=(45, 35), params=dict(niter=18, thresh_op="garrote", thresh_model="exponential-2", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-4)),
    dict(name="default_eps_hard", seed=23, shape=(64, 50), params=dict(niter=100, thresh_op="hard", thresh_model="exponential", eps=1e-9, alpha=1.0, p_max=0.99, p_min=1e-5)),
    dict(name="all_zero", seed=24, shape=(16, 20), all_zero=True, params=dict(niter=10, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0)),
    dict(name="soft_percentile_linear", seed=26, shape=(48, 52), params=dict(niter=12, thresh_op="soft-percentile", thresh_model="linear", eps=0.0, alpha=1.0, p_max=99.5, p_min=20.0, decay_kind="factors")),
    dict(name="garrote_percentile_exp", seed=27, shape=(40, 64), noise=0.01, params=dict(niter=10, thresh_op="garrote-percentile", thresh_model="exponential", eps=0.0, alpha=0.8, p_max=99.0, p_min=5.0, decay_kind="factors")),
    dict(name="hard_percentile_exp", seed=28, shape=(56, 44), params=dict(niter=10, thresh_op="hard-percentile", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=99.9, p_min=50.0, decay_kind="factors")),
    dict(name="bluestein_sizes", seed=25, shape=(37, 58), params=dict(niter=12, thresh_op="soft", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)),
]

# parameter sets for schedule-table fixtures (run on the case-0 spectrum)
SCHEDULES = [
    dict(thresh_model="linear", niter=7, p_max=0.99, p_min=1e-5, kind="values"),
    dict(thresh_model="exponential", niter=9, p_max=0.99, p_min=1e-5, kind="values"),
    dict(thresh_model="exponential-3", niter=9, p_max=0.9, p_min=1e-3, kind="values"),
    dict(thresh_model="data-driven", niter=11, p_max=0.99, p_min=1e-5, kind="values"),
    dict(thresh_model="linear", niter=5, p_max=0.99, p_min="adaptive", kind="values"),
    dict(thresh_model="exponential", niter=5, p_max=0.99, p_min="adaptive", kind="values"),
    dict(thresh_model="exponential", niter=6, p_max=3.0, p_min=0.1, kind="factors"),
    dict(thresh_model="inverse_proportional", niter=8, p_max=0.99, p_min=1e-5, kind="values"),
    dict(thresh_model="inverse-proportional-2", niter=8, p_max=0.99, p_min=1e-5, kind="values"),
]
