import sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ntr = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
x = torch.randn((nt, ntr), device="cuda", dtype=torch.float32)
for name, fn in (("cufft rfft dim0", lambda: torch.fft.rfft(x, dim=0)),
                 ("transpose copy", lambda: x.t().contiguous()),
                 ("cufft rfft dim1 of transposed", None)):
    if fn is None:
        xt = x.t().contiguous()
        fn = lambda: torch.fft.rfft(xt, dim=1)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nbytes = x.numel() * 4 + y.numel() * y.element_size()
    print(f"{name}: {ms:.2f} ms, {nbytes / ms / 1e6:.0f} GB/s")
    del y
