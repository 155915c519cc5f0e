"""Device-resident timing of the time-axis kernels (p3d_time_fft / p3d_time_ifft).

    python tools/bench_time_axis.py [nt] [n_il] [n_xl] [compute_real]
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_3d_interpolation_b200 import _lib      # noqa: E402


def main():
    a = sys.argv[1:]
    nt = int(a[0]) if len(a) > 0 else 2048
    n1 = int(a[1]) if len(a) > 1 else 1000
    n2 = int(a[2]) if len(a) > 2 else 1000
    real = int(a[3]) if len(a) > 3 else 1
    lib = _lib.load()
    _lib.require_gpu()
    ntr = n1 * n2
    nf = nt // 2 + 1 if real else nt
    rng = np.random.default_rng(0)
    blk = rng.standard_normal((nt, min(ntr, 4096))).astype(np.float32)
    x = np.tile(blk, (1, (ntr + blk.shape[1] - 1) // blk.shape[1]))[:, :ntr].copy()
    dx = _lib.DeviceBuffer(x.nbytes); dx.upload(x)
    dF = _lib.DeviceBuffer(nf * ntr * 8)
    dy = _lib.DeviceBuffer(x.nbytes)
    dt, t0 = 0.05, 725.0

    def fwd():
        _lib.check(lib.p3d_time_fft(0, C.c_void_p(dx.ptr), 1, C.c_void_p(dF.ptr), 1, nt, nt, ntr, dt, t0, real, None))

    def inv():
        _lib.check(lib.p3d_time_ifft(0, C.c_void_p(dF.ptr), 1, C.c_void_p(dy.ptr), 1, nt, nt, ntr, dt, t0, real, 0))

    def env():
        _lib.check(lib.p3d_time_envelope(0, C.c_void_p(dx.ptr), 1, C.c_void_p(dy.ptr), 1, nt, ntr))

    cases = [("time_fft", fwd, x.nbytes + nf * ntr * 8), ("time_ifft", inv, x.nbytes + nf * ntr * 8)]
    if os.environ.get("P3D_BENCH_ENVELOPE"):
        cases = [("envelope", env, 2 * x.nbytes)]
    for name, fn, nbytes in cases:
        fn(); fn()
        lib.p3d_device_synchronize(0)
        t = time.perf_counter()
        reps = 3
        for _ in range(reps):
            fn()
        lib.p3d_device_synchronize(0)
        el = (time.perf_counter() - t) / reps
        kms = C.c_double(); lib.p3d_time_last_kernel_ms(C.byref(kms))
        print(f"{name}: nt={nt} traces={ntr} real={real}: call {el*1e3:.2f} ms; kernels {kms.value:.2f} ms = {nbytes/max(kms.value,1e-9)/1e6:.0f} GB/s algorithmic ({nbytes/1e9:.1f} GB) [{lib.p3d_time_last_path().decode()}]")
    if os.environ.get("P3D_BENCH_ENVELOPE"):
        return
    y = np.empty((nt, 64), np.float32)
    full = np.empty_like(x); dy.download(full)
    err = np.linalg.norm(full[:, :4096] - x[:, :4096]) / np.linalg.norm(x[:, :4096])
    print(f"round trip rel-L2 {err:.2e}")


if __name__ == "__main__":
    main()
