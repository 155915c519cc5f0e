"""GPU: escalating mode vs the float64 state mode (which reproduces the float64 oracle) over a whole synthetic cube.

    python tools/sweep_guard.py CONFIG N_SLICES GUARDS [SEG] [NOISE]
"""
import os, sys, time, json
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import pseudo_3d_interpolation_b200 as p3d
from pseudo_3d_interpolation_b200 import synth

cfg = int(sys.argv[1]); ns = int(sys.argv[2]); guards = [int(a) for a in sys.argv[3].split(",")]
seg = int(sys.argv[4]) if len(sys.argv) > 4 else 4
noise = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
c = dict(synth.CONFIGS[cfg]); n1, n2 = c["n_il"], c["n_xl"]; niter = c["niter"]
dev = torch.device("cuda", 0)
x, mask = bench.synth_device(torch, dev, cfg, c, 0, ns, c["nt"] // 2 + 1)
if noise > 0:
    g = torch.Generator(device=dev); g.manual_seed(5)
    sig = noise * float(x.abs().max())
    x = x + (sig / np.sqrt(2)) * torch.complex(torch.randn(x.shape, generator=g, device=dev), torch.randn(x.shape, generator=g, device=dev)) * mask[None]
    x = x.contiguous()
params = p3d.make_params(niter=niter, thresh_op=c["thresh_op"], thresh_model=c["thresh_model"], eps=0.0, alpha=c["alpha"], p_max=0.99, p_min=1e-5)
spm = ns

def run(prec, guard=None):
    plan = p3d.PocsPlan(n1, n2, precision=prec)
    if guard is not None:
        plan.set_option("guard_factor", guard); plan.set_option("seg_iters", seg)
        for kv in filter(None, os.environ.get("P3D_OPTS", "").split(",")):
            plan.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    out = torch.empty_like(x)
    plan.run_device(x.data_ptr(), mask.data_ptr(), out.data_ptr(), ns, params)
    plan.event_record(0)
    plan.run_device(x.data_ptr(), mask.data_ptr(), out.data_ptr(), ns, params)
    plan.event_record(1)
    ms = plan.event_elapsed_ms(0, 1)
    esc = plan.escalation()
    plan.close()
    return out, ms, esc

ref, ms64, _ = run(64)
print(f"C{cfg} {ns} slices: precision 64: {ns * niter / ms64 * 1e3:.0f} slice-it/s", flush=True)
rn = torch.linalg.vector_norm(ref.reshape(ns, -1), dim=1).double()
def report(name, y, ms, esc):
    e = torch.linalg.vector_norm((y - ref).reshape(ns, -1), dim=1).double()
    rel = (e / rn.clamp_min(1e-300)).cpu().numpy()
    cube = float(torch.sqrt((e ** 2).sum()) / torch.sqrt((rn ** 2).sum()))
    print(f"{name}: {ns * niter / ms * 1e3:.0f} slice-it/s, cube rel-L2 {cube:.2e}, slices > 1e-5: {int((rel > 1e-5).sum())}, > 1e-4: {int((rel > 1e-4).sum())}, "
          f"max {rel.max():.2e}, median {np.median(rel):.1e}; escalated {esc[0]} slices, {esc[1] / (ns * niter) * 100:.1f} % of slice-its in complex128", flush=True)
y, ms, esc = run(32)
report("precision 32", y, ms, esc)
for g in guards:
    y, ms, esc = run(0, g)
    report(f"auto guard {g} seg {seg}", y, ms, esc)
