"""Print GPU-vs-reference errors for every golden case, next to the reference's own complex64 drift."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pseudo_3d_interpolation_b200 as p3d
from oracle.golden_cases import CASES, make_input
from oracle import pocs_oracle as orc

g = np.load(os.path.join(ROOT, "tests", "golden", "reference_pocs.npz"))
def rel(a, b):
    a = np.asarray(a, np.complex128); b = np.asarray(b, np.complex128)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
for case in CASES:
    x, mask = make_input(case)
    info = {}
    fn = {"regular": p3d.POCS, "fast": p3d.FPOCS, "adaptive": p3d.APOCS}[case.get("version", "regular")]
    y = fn(x, mask, None, transform=np.fft.fft2, itransform=np.fft.ifft2, transform_kind="FFT", results_dict=info, **case["params"])
    n = case["name"]
    fl = float(g[n + "__c64_drift"]) if n + "__c64_drift" in g else float("nan")
    print(f"{n:28s} gpu-vs-ref {rel(y, g[n + '__y']):.3e}  ref-c64-floor {fl:.3e}  its {info['niterations']} / {int(g[n + '__niterations'])}")
    if n in ("apocs_a08", "inverse_proportional"):
        plan = p3d.PocsPlan(*x.shape)
        kw = {k: v for k, v in case["params"].items() if k in ("niter", "thresh_model", "p_max", "p_min")}
        tau = plan.schedule(x, **kw)[0]
        X0 = np.fft.fft2(x.astype(np.complex128))
        ref = orc.threshold_table(X0, case["params"]["niter"], case["params"]["thresh_model"], case["params"].get("p_max", 0.99), case["params"].get("p_min", 1e-5))
        print("   tau gpu", tau[:3], tau[-2:]); print("   tau ref", ref[:3], ref[-2:])
        print("   min|X0|", np.abs(X0).min(), "max", np.abs(X0).max())
# spec-sized cases: error vs float64 oracle and the oracle's own complex64 drift
for shape, op, model, niter, pmin in [((256, 256), "hard", "exponential", 12, 1e-4), ((1000, 1000), "hard", "exponential", 8, 1e-4),
                                      ((256, 256), "hard", "exponential", 12, 1e-5), ((256, 256), "soft", "exponential", 12, 1e-4)]:
    x, mask = make_input(dict(seed=77, shape=shape, keep=0.3, nwaves=5))
    params = dict(niter=niter, thresh_op=op, thresh_model=model, eps=0.0, alpha=1.0, p_max=0.99, p_min=pmin)
    ref = orc.pocs_slice(x.astype(np.complex128), mask, **params)
    r32 = orc.pocs_slice(x, mask, **params)
    plan = p3d.PocsPlan(*shape)
    y, _ = plan.run(x, mask, **params)
    plan.set_option("force_generic", 1)
    yg, _ = plan.run(x, mask, **params)
    print(f"{shape} {op} {model} niter={niter} pmin={pmin}: spec {rel(y[0], ref):.3e} generic {rel(yg[0], ref):.3e} numpy-c64 {rel(r32, ref):.3e} dtype {r32.dtype}")
# raw FFT accuracy vs numpy complex64
rng = np.random.default_rng(0)
for shape in [(256, 256), (1000, 1000), (200, 200), (37, 58), (121, 77)]:
    x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    ref = np.fft.fft2(x.astype(np.complex128))
    plan = p3d.PocsPlan(*shape)
    print(f"fft2 {shape}: gpu {rel(plan.fft2(x), ref):.3e}  numpy-c64 {rel(np.fft.fft2(x), ref):.3e}")
