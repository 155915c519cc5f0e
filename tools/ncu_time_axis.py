"""One forward + inverse time-axis transform per implementation (for an ncu capture of the kernels side by side).

    python tools/ncu_time_axis.py [nt] [n_il] [n_xl]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_3d_interpolation_b200 import _lib      # noqa: E402

a = sys.argv[1:]
nt = int(a[0]) if len(a) > 0 else 2048
n1 = int(a[1]) if len(a) > 1 else 1000
n2 = int(a[2]) if len(a) > 2 else 1000
lib = _lib.load()
_lib.require_gpu()
ntr = n1 * n2
nf = nt // 2 + 1
rng = np.random.default_rng(0)
blk = rng.standard_normal((nt, 4096)).astype(np.float32)
x = np.tile(blk, (1, (ntr + 4095) // 4096))[:, :ntr].copy()
dx = _lib.DeviceBuffer(x.nbytes); dx.upload(x)
dF = _lib.DeviceBuffer(nf * ntr * 8)
dy = _lib.DeviceBuffer(x.nbytes)
for path, asyn in ((None, None), ("direct", None), ("pipeline", None)):
    if path is None:
        os.environ.pop("P3D_TIME_PATH", None)
    else:
        os.environ["P3D_TIME_PATH"] = path
    if asyn is None:
        os.environ.pop("P3D_TIME_ASYNC", None)
    else:
        os.environ["P3D_TIME_ASYNC"] = asyn
    for rep in range(2):
        _lib.check(lib.p3d_time_fft(0, C.c_void_p(dx.ptr), 1, C.c_void_p(dF.ptr), 1, nt, nt, ntr, 0.05, 725.0, 1, None))
        kf = C.c_double(); lib.p3d_time_last_kernel_ms(C.byref(kf))
        _lib.check(lib.p3d_time_ifft(0, C.c_void_p(dF.ptr), 1, C.c_void_p(dy.ptr), 1, nt, nt, ntr, 0.05, 725.0, 1, 0))
        ki = C.c_double(); lib.p3d_time_last_kernel_ms(C.byref(ki))
    nbytes = x.nbytes + nf * ntr * 8
    print(f"{path} async={asyn}: fwd {kf.value:.2f} ms = {nbytes/kf.value/1e6:.0f} GB/s, inv {ki.value:.2f} ms = {nbytes/ki.value/1e6:.0f} GB/s ({lib.p3d_time_last_path().decode()})")
