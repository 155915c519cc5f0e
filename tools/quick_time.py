"""Quick device-resident timing of the POCS iteration kernels (development helper).

    python tools/quick_time.py [n_il] [n_xl] [n_slices] [niter] [band_slices] [op] [force_generic] [spec_variant] [precision] [lanes] [spec_variant64]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pseudo_3d_interpolation_b200 as p3d           # noqa: E402
from pseudo_3d_interpolation_b200 import _lib        # noqa: E402


def main():
    a = sys.argv[1:]
    n1 = int(a[0]) if len(a) > 0 else 1000
    n2 = int(a[1]) if len(a) > 1 else 1000
    ns = int(a[2]) if len(a) > 2 else 64
    niter = int(a[3]) if len(a) > 3 else 20
    band = int(a[4]) if len(a) > 4 else 0
    op = a[5] if len(a) > 5 else "hard"
    force_generic = int(a[6]) if len(a) > 6 else 0
    rng = np.random.default_rng(0)
    i = np.arange(n1)[:, None]; j = np.arange(n2)[None, :]
    base = np.zeros((n1, n2), np.complex128)
    for _ in range(6):
        base += rng.uniform(0.3, 1) * np.exp(2j * np.pi * (rng.uniform(-.2, .2) * i + rng.uniform(-.2, .2) * j))
    mask = (rng.random((n1, n2)) < 0.2).astype(np.uint8)
    x1 = (base * mask).astype(np.complex64)
    x = np.broadcast_to(x1, (ns, n1, n2)).copy()
    x *= (1 + 0.01 * np.arange(ns, dtype=np.float32))[:, None, None]
    precision = int(a[8]) if len(a) > 8 else 32
    plan = p3d.PocsPlan(n1, n2, band_slices=band, precision=precision)
    if force_generic:
        plan.set_option("force_generic", 1)
    if len(a) > 7:
        plan.set_option("spec_variant", int(a[7]))
    if len(a) > 9:
        plan.set_option("lanes", int(a[9]))
    if len(a) > 10:
        plan.set_option("spec_variant64", int(a[10]))
    if os.environ.get("P3D_GUARD"):
        plan.set_option("guard_factor", int(float(os.environ["P3D_GUARD"])))
    for kv in filter(None, os.environ.get("P3D_OPTS", "").split(",")):
        plan.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    print(plan.describe())
    dx = _lib.DeviceBuffer(x.nbytes); dx.upload(x)
    dm = _lib.DeviceBuffer(mask.nbytes); dm.upload(mask)
    do = _lib.DeviceBuffer(x.nbytes)
    params = p3d.make_params(niter=niter, thresh_op=op, thresh_model="exponential", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-5)
    plan.run_device(dx.ptr, dm.ptr, do.ptr, ns, params)          # warm-up
    plan.set_profiling(True)
    plan.event_record(0)
    t0 = time.perf_counter()
    plan.run_device(dx.ptr, dm.ptr, do.ptr, ns, params)
    wall = time.perf_counter() - t0
    plan.event_record(1)
    dev_ms = plan.event_elapsed_ms(0, 1)
    prof = plan.get_profile()
    sit = ns * niter
    print(f"{n1}x{n2} slices={ns} niter={niter} band={band} op={op}: wall {wall*1e3:.2f} ms, device {dev_ms:.2f} ms, "
          f"{sit / (dev_ms * 1e-3):.0f} slice-it/s")
    ne = n1 * n2
    for k, v in prof.items():
        if v["launches"]:
            per = v["ms"] / v["launches"]
            bytes_alg = {"cols_iter": 16, "rows_iter": 24.125, "rows_init": 16, "cols_stats": 8, "cols_iter64": 32, "rows_iter64": 40.125}.get(k, 0) * ne * (band if band else ns)
            print(f"  {k:10s} launches={v['launches']:5d} total={v['ms']:9.3f} ms  avg={per*1e3:9.1f} us  alg GB/s={bytes_alg / (per * 1e-3) / 1e9 if per else 0:8.1f}")
    out = np.empty_like(x[:1]); do.download(out)
    print("checksum", float(np.abs(out).sum()))


if __name__ == "__main__":
    main()
