"""Small representative invocations of every kernel family, each repeated and compared bit for bit (a shared-memory
race or an out-of-bounds read shows up as run-to-run differences); also usable under compute-sanitizer where that
is available:

    python tools/sanitize_cases.py [spec|mix|generic|f64|aux|time]
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pseudo_3d_interpolation_b200 as p3d                                      # noqa: E402
from pseudo_3d_interpolation_b200 import timeaxis, cube_postprocessing_3D as post    # noqa: E402
from oracle.golden_cases import make_input                                      # noqa: E402


CASES_RUN = 0


def run(shape, precision=32, **kw):
    global CASES_RUN
    CASES_RUN += 1
    x, mask = make_input(dict(seed=3, shape=shape, keep=0.3))
    x = np.stack([x, 0.5 * x]).astype(np.complex64)
    params = dict(niter=4, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)
    params.update(kw)
    plan = p3d.PocsPlan(*shape, precision=precision)
    y, info = plan.run(x, mask, **params)
    assert np.isfinite(y.view(np.float32)).all()
    for _ in range(3):
        y2, info2 = plan.run(x, mask, **params)
        assert np.array_equal(y, y2) and np.array_equal(info["niterations"], info2["niterations"]), ("not deterministic", shape, kw)
    print("ok", shape, precision, kw, flush=True)


only = sys.argv[1] if len(sys.argv) > 1 else "all"
if only in ("all", "spec"):
    run((200, 200)); run((256, 256), thresh_op="soft", thresh_model="data-driven", alpha=0.7)
    run((1000, 24)); run((24, 1000), version="adaptive", alpha=0.8); run((2000, 8)); run((8, 2000))
if only in ("all", "mix"):
    run((1201, 8), thresh_op="garrote"); run((8, 847), thresh_op="soft", thresh_model="linear"); run((1201, 847), niter=2)
if only in ("all", "generic"):
    run((37, 58)); run((121, 77), thresh_op="soft-percentile", decay_kind="factors", p_max=99.0, p_min=10.0, thresh_model="linear")
if only in ("all", "f64"):
    run((200, 200), precision=64); run((256, 40), precision=64, thresh_op="garrote"); run((40, 1000), precision=64); run((53, 47), precision=64)
if only in ("all", "aux"):
    rng = np.random.default_rng(0)
    for shape in ((256, 256), (1201, 16), (33, 47)):
        d = (rng.standard_normal((2,) + shape) + 1j * rng.standard_normal((2,) + shape)).astype(np.complex64)
        a = post.remove_acquisition_footprint(d, sigma=2)
        assert all(np.array_equal(a, post.remove_acquisition_footprint(d, sigma=2)) for _ in range(2))
    for nt, ntr in ((512, 37), (2048, 6), (300, 9), (75, 4)):
        x = rng.standard_normal((nt, ntr)).astype(np.float32)
        e = timeaxis.envelope(x, axis=0)
        assert all(np.array_equal(e, timeaxis.envelope(x, axis=0)) for _ in range(2))
        twt = 725.0 + 0.05 * np.arange(nt)
        F, _ = timeaxis.time_fft(x.reshape(nt, ntr, 1), twt, compute_real=True)
        timeaxis.time_ifft(F, 0.05, 725.0, compute_real=True)
    CASES_RUN += 1
    print("ok aux", flush=True)
if only in ("all", "time"):
    # one-pass time-axis kernels (TMA-staged tiles; trace counts divisible by 4, two CTAs so that every CTA refills its stage)
    grid_before = os.environ.get("P3D_TIME_GRID")
    os.environ["P3D_TIME_GRID"] = "2"
    rng = np.random.default_rng(1)
    for nt, ntr in ((512, 136), (1000, 72), (1024, 72), (2048, 72), (2000, 40), (4096, 24), (4000, 24), (2500, 12)):
        x = rng.standard_normal((nt, ntr)).astype(np.float32)
        twt = 725.0 + 0.05 * np.arange(nt)
        for real in (True, False):
            F, _ = timeaxis.time_fft(x.reshape(nt, ntr, 1), twt, compute_real=real)
            F2, _ = timeaxis.time_fft(x.reshape(nt, ntr, 1), twt, compute_real=real)
            assert np.array_equal(F, F2)
            Fin = F if real else np.fft.fftshift(F, axes=0)
            y = timeaxis.time_ifft(Fin, 0.05, 725.0, compute_real=real)
            assert np.array_equal(y, timeaxis.time_ifft(Fin, 0.05, 725.0, compute_real=real))
            assert np.abs(y[:, :, 0] - x).max() < 1e-4
        e = timeaxis.envelope(x, axis=0)
        assert np.array_equal(e, timeaxis.envelope(x, axis=0))
    if grid_before is None:
        del os.environ["P3D_TIME_GRID"]
    else:
        os.environ["P3D_TIME_GRID"] = grid_before
    CASES_RUN += 1
    print("ok time", flush=True)
