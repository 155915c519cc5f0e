"""GPU check of the escalating-precision mode against the float64 oracle on full-size slices.

    python tools/check_escalate.py CONFIG SLICE_IDS [GUARD_FACTORS] [PRECISIONS]
"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import pseudo_3d_interpolation_b200 as p3d
from pseudo_3d_interpolation_b200 import synth
from oracle import pocs_oracle as orc
from concurrent.futures import ProcessPoolExecutor

cfg = int(sys.argv[1]); sids = [int(a) for a in sys.argv[2].split(",")]
guards = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1024]
noise = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
d, fold, c = synth.sparse_freq_slices(cfg, sids, noise=noise)
mask = orc.mask_from_fold(fold)
params = dict(niter=c["niter"], thresh_op=c["thresh_op"], thresh_model=c["thresh_model"], eps=0.0, alpha=c["alpha"], p_max=0.99, p_min=1e-5)

def ref_slice(i):
    return orc.pocs_slice(d[i].astype(np.complex128), mask, **params).astype(np.complex64)

t0 = time.time()
with ProcessPoolExecutor(min(len(sids), 16)) as ex:
    ref = np.stack(list(ex.map(ref_slice, range(len(sids)))))
print(f"oracle: {time.time() - t0:.1f} s", flush=True)
obs = mask == 1
for prec, g in [(32, 0)] + [(0, g) for g in guards] + [(64, 0)]:
    plan = p3d.PocsPlan(d.shape[1], d.shape[2], precision=prec)
    if prec == 0:
        plan.set_option("guard_factor", g)
    t0 = time.time()
    y, info = plan.run(d, mask, params=p3d.make_params(**params))
    dt = time.time() - t0
    errs = [np.linalg.norm(y[i] - ref[i]) / np.linalg.norm(ref[i]) for i in range(len(sids))]
    ns, nit = plan.escalation()
    exact = bool(np.array_equal(y[:, obs], d[:, obs])) if c["alpha"] == 1.0 else None
    print(f"precision {prec} guard {g}: cube rel-L2 {np.linalg.norm(y - ref) / np.linalg.norm(ref):.2e}  max slice {max(errs):.2e}  "
          f"escalated {ns}/{len(sids)} slices, {nit} of {len(sids) * c['niter']} slice-its in complex128, observed exact {exact}, {dt:.2f} s", flush=True)
    print("   per slice:", " ".join(f"{e:.1e}" for e in errs), flush=True)
    plan.close()
