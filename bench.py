#!/usr/bin/env python
"""Benchmark of the FFT-POCS hot path (BASELINE.json metric: POCS slice-iterations / s).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 2] [--slices S]

One "step" = the POCS of S frequency slices (all iterations) of the synthetic config-2 cube
(1000 x 1000 slices, 80 % of the traces missing, hard threshold, exponential schedule,
100 iterations, alpha = 1, eps = 0) on every GPU; N > 1 runs one process per GPU under torchrun,
each rank owning its own contiguous band of S slices (weak scaling, no data-path collective).

Output: ONE JSON line on rank 0 (see the keys at the end of main()).
"""
from __future__ import annotations

import argparse
import json
import math
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_ALG_PER_ELEM = 73.0          # SURVEY 8d contract model: 4 streaming passes x 16 B + 8 B observed data + 1 B mask (complex64 state)
# bytes the FUSED kernels have to move per element and launch (the floor their DRAM traffic is compared with):
# complex64 state: column pass 8 + 8, row pass 8 + 8 + 8 (observed data) + 1/8 (packed mask);
# complex128 state (escalating / float64 modes): 16 + 16 and 16 + 16 + 8 + 1/8
KERNEL_BYTES_PER_ELEM = {"cols_iter": 16.0, "rows_iter": 24.125, "cols_iter64": 32.0, "rows_iter64": 40.125}
FUSED_FLOOR_32 = 40.125
FUSED_FLOOR_64 = 72.125


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--slices", type=int, default=0, help="slices per GPU per step (0 = config default)")
    ap.add_argument("--niter", type=int, default=0)
    ap.add_argument("--band", type=int, default=-1, help="band_slices override (-1 = plan default)")
    ap.add_argument("--lanes", type=int, default=0, help="copy/compute lanes of the host-buffer path (0 = library default)")
    ap.add_argument("--chunk", type=int, default=0, help="largest chunk of slices per lane in the host-buffer path (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-diag", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precision", default="auto", choices=["auto", "32", "64"],
                    help="auto = escalating fp32 -> complex128 (default, within 1e-4 of the float64 reference); 32 = fp32 only; 64 = complex128 throughout")
    ap.add_argument("--e2e-slices", type=int, default=0, help="slices per GPU per end-to-end step (0 = all at 1-2 GPUs, 512 beyond: pinned host memory)")
    return ap.parse_args()


def workload(args):
    from pseudo_3d_interpolation_b200 import synth
    c = dict(synth.CONFIGS[args.config])
    nf_full = c["nt"] // 2 + 1
    default_slices = {1: 257, 2: 1025, 3: 256, 4: 64, 5: 513 * 8}[args.config]
    ns = args.slices if args.slices > 0 else default_slices
    niter = args.niter if args.niter > 0 else c["niter"]
    return c, ns, niter, nf_full


# ---------------------------------------------------------------------------------------------------
# synthetic slices generated on the device (torch is plumbing here: memory + RNG + trig)
# ---------------------------------------------------------------------------------------------------
def synth_device(torch, dev, cfg_id, c, first, count, nf_full):
    from pseudo_3d_interpolation_b200 import synth
    ev, fold, rng, _ = synth.config_case(cfg_id)
    n1, n2 = c["n_il"], c["n_xl"]
    f_all = np.fft.rfftfreq(c["nt"], synth.DT_MS)
    ids = (first + np.arange(count)) % nf_full
    ids = np.where(ids == 0, 1, ids)                      # skip the (empty) DC slice
    f = torch.tensor(f_all[ids], dtype=torch.float64, device=dev)
    il = torch.arange(n1, dtype=torch.float64, device=dev) - (n1 - 1) / 2.0
    xl = torch.arange(n2, dtype=torch.float64, device=dev) - (n2 - 1) / 2.0
    mask = torch.tensor((fold > 0).astype(np.uint8), device=dev)
    out = torch.empty((count, n1, n2), dtype=torch.complex64, device=dev)
    step = max(1, int(2e8 // (n1 * n2)))
    for s0 in range(0, count, step):
        fs = f[s0:s0 + step]
        acc = torch.zeros((fs.numel(), n1, n2), dtype=torch.complex64, device=dev)
        for k in range(len(ev.amp)):
            rk = (2.0 / np.sqrt(np.pi)) * fs * fs / ev.fpk[k] ** 3 * torch.exp(-(fs / ev.fpk[k]) ** 2)
            # separable phase: exp(-2 pi i f (t0 + tau)) * exp(-2 pi i f p il) * exp(-2 pi i f q xl)
            ph0 = -2.0 * np.pi * fs * (synth.T0_MS + ev.tau[k])
            a0 = (ev.amp[k] * rk) * torch.exp(1j * torch.remainder(ph0, 2 * np.pi))
            a1 = torch.exp(-2j * np.pi * fs[:, None] * (ev.p[k] * il)[None, :])
            a2 = torch.exp(-2j * np.pi * fs[:, None] * (ev.q[k] * xl)[None, :])
            acc += ((a0[:, None] * a1)[:, :, None] * a2[:, None, :]).to(torch.complex64)
        out[s0:s0 + step] = acc * mask[None]
    return out, mask


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for nme, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle (port of the reference algorithm) over a bounded sample of slices
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    import numpy as _np
    from oracle import pocs_oracle as orc
    x, mask, params = job
    t0 = time.perf_counter()
    info = {}
    orc.pocs_slice(x, mask, info=info, **params)
    return info["niterations"], time.perf_counter() - t0


def cpu_sample(cfg_id, c, niter, n_slices, procs):
    """Time `n_slices` slices (complex64 input, as the real pipeline feeds numpy >= 2) on `procs` processes."""
    from pseudo_3d_interpolation_b200 import synth
    ids = list(range(8, 8 + n_slices))
    d, fold, _ = synth.sparse_freq_slices(cfg_id, ids)
    mask = np.minimum(fold, 1).astype(np.uint8)
    params = dict(niter=niter, thresh_op=c["thresh_op"], thresh_model=c["thresh_model"], eps=0.0, alpha=c["alpha"], p_max=0.99, p_min=1e-5)
    jobs = [(d[i], mask, params) for i in range(n_slices)]
    t0 = time.perf_counter()
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs, chunksize=1)
    else:
        res = [_cpu_worker(j) for j in jobs]
    wall = time.perf_counter() - t0
    its = sum(r[0] for r in res)
    return its / wall, wall, its


def cpu_niter_for_budget(c, procs, budget_s=20.0):
    # ~135 ms per 1000x1000 iteration per core, scaled by size
    per_it = 135e-3 * (c["n_il"] * c["n_xl"]) / 1e6
    return max(4, int(budget_s / per_it))


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    # Only the JSON line may reach stdout: C libraries (NCCL prints its version banner there)
    # are redirected to stderr by swapping file descriptor 1; the JSON goes to the saved one.
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(saved_fd, "w")

    def emit(obj):
        json_out.write(json.dumps(obj) + "\n")
        json_out.flush()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    c, ns, niter, nf_full = workload(args)
    n1, n2 = c["n_il"], c["n_xl"]
    cfg_workload = (f"C{args.config}: {n1}x{n2} iline x xline slices of the {c['nt']}-sample cube, "
                    f"{'line-pattern' if c['keep'] == 'lines' else str(int(round((1 - c['keep']) * 100))) + '%'} traces missing, "
                    f"FFT-POCS {c['thresh_op']} threshold, {c['thresh_model']} schedule, {niter} iterations, alpha={c['alpha']}, eps=0; "
                    f"{ns} rfft slices per GPU per step")

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        procs = min(os.cpu_count() or 1, 64)
        it_cpu = niter                                   # the configuration's own iteration count (same_config)
        vals = []
        for i in range(args.warmup + args.steps):
            v, wall, its = cpu_sample(args.config, c, it_cpu, procs, procs)
            if i >= args.warmup:
                vals.append((v, wall))
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([w for _, w in vals]) * 1e3)
        sample = f"{procs} slices x {it_cpu} iterations per step, one slice per process ({procs} processes), complex64 input"
        line = {"impl": "reference", "metric": "pocs_slice_iterations_per_s", "value": value, "unit": "slice-iterations/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
                "config": {"workload": cfg_workload, "inputs": "host numpy arrays"},
                "cpu_baseline": {"value": value, "unit": "slice-iterations/s", "cores": procs, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "slice-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import pseudo_3d_interpolation_b200 as p3d
    from pseudo_3d_interpolation_b200 import _lib

    _lib.require_gpu()
    numa = "numa: not bound (P3D_NUMA_BIND=0)"
    if os.environ.get("P3D_NUMA_BIND", "1") != "0" and world > 1:
        from pseudo_3d_interpolation_b200.distributed import bind_to_gpu_numa_node
        gi = local
        if os.environ.get("CUDA_VISIBLE_DEVICES"):
            try:
                gi = int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local])
            except Exception:          # noqa: BLE001
                gi = local
        numa = bind_to_gpu_numa_node(gi)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    params = p3d.make_params(niter=niter, thresh_op=c["thresh_op"], thresh_model=c["thresh_model"], eps=0.0, alpha=c["alpha"],
                             p_max=0.99, p_min=1e-5)
    x_dev, mask_dev = synth_device(torch, dev, args.config, c, rank * ns, ns, nf_full)
    out_dev = torch.empty_like(x_dev)
    torch.cuda.synchronize()
    precision = {"auto": 0, "32": 32, "64": 64}[args.precision]
    plan = p3d.PocsPlan(n1, n2, device=local, precision=precision)
    if args.band >= 0:
        plan.set_option("band_slices", args.band)
    desc = plan.describe()
    nit = np.zeros(ns, np.int32)

    def step_device():
        plan.run_device(x_dev.data_ptr(), mask_dev.data_ptr(), out_dev.data_ptr(), ns, params, nit=nit)

    for _ in range(args.warmup):
        step_device()
    gpu_index = local
    if os.environ.get("CUDA_VISIBLE_DEVICES"):
        try:
            gpu_index = int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local])
        except Exception:
            gpu_index = local
    sampler = ClockSampler(gpu_index)
    plan.set_profiling(True)
    plan.get_profile(reset=True)
    barrier()
    sampler.start()
    plan.event_record(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    plan.event_record(1)
    dev_ms = plan.event_elapsed_ms(0, 1)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    prof = plan.get_profile(reset=True)
    plan.set_profiling(False)
    slice_its = int(nit.sum())                          # per step, this rank
    assert slice_its == ns * niter, (slice_its, ns * niter)
    esc_slices, esc_its = plan.escalation() if precision == 0 else ((ns, ns * niter) if precision == 64 else (0, 0))
    f64_share = esc_its / max(1, ns * niter)            # share of the slice-iterations that ran on complex128 state

    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    tot_its = torch.tensor([float(slice_its)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_its, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    total_its_per_step = float(tot_its[0])
    value = total_its_per_step * args.steps / (dev_ms_max * 1e-3)

    # ---- strong scaling (N > 1): ONE cube of nf_full slices split into contiguous frequency bands, one band per GPU
    strong = None
    if world > 1:
        from pseudo_3d_interpolation_b200.pocs import band_bounds
        lo, hi = band_bounds(nf_full, world)[rank]
        ns_s = hi - lo
        if ns_s <= ns:
            nit_s = np.zeros(max(ns_s, 1), np.int32)

            def step_strong():
                if ns_s > 0:
                    plan.run_device(x_dev.data_ptr(), mask_dev.data_ptr(), out_dev.data_ptr(), ns_s, params, nit=nit_s)

            step_strong()
            barrier()
            plan.event_record(4)
            for _ in range(args.steps):
                step_strong()
            plan.event_record(5)
            s_ms = plan.event_elapsed_ms(4, 5)
            barrier()
            ts = torch.tensor([s_ms], dtype=torch.float64, device=dev)
            tn = torch.tensor([float(nit_s[:ns_s].sum())], dtype=torch.float64, device=dev)
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            dist.all_reduce(tn, op=dist.ReduceOp.SUM)
            strong = {"value": float(tn[0]) * args.steps / (float(ts[0]) * 1e-3), "unit": "slice-iterations/s", "scaling": "strong",
                      "cube_slices": nf_full, "slices_per_gpu": -(-nf_full // world), "ms_per_step": float(ts[0]) / args.steps,
                      "what": "one cube of nf_full rfft slices split into contiguous bands (pocs.band_bounds), device-resident, max over ranks"}
        else:
            strong = {"skipped": f"a band of the cube ({ns_s} slices) exceeds the {ns} slices resident per GPU; rerun with --slices {ns_s}"}

    # ---- end to end: pinned host buffers in, pinned host buffers out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        # bounded pinned footprint: at most 512 slices per rank per e2e step (the chunk pipeline
        # of p3d_pocs_run is in steady state long before that), so 8 ranks stay below 70 GB pinned
        ns_e = args.e2e_slices if args.e2e_slices > 0 else (ns if world <= 2 else min(ns, 512))
        ns_e = min(ns_e, ns)
        hx = _lib.PinnedArray((ns_e, n1, n2), np.complex64)
        ho = _lib.PinnedArray((ns_e, n1, n2), np.complex64)
        hm = _lib.PinnedArray((n1, n2), np.uint8)
        _lib.check(_lib.load().p3d_memcpy(local, _lib.ptr(hx.array), x_dev.data_ptr(), hx.nbytes, 1))
        _lib.check(_lib.load().p3d_memcpy(local, _lib.ptr(hm.array), mask_dev.data_ptr(), hm.nbytes, 1))
        del x_dev, out_dev
        torch.cuda.empty_cache()

        if args.lanes > 0:
            plan.set_option("lanes", args.lanes)
        if args.chunk > 0:
            plan.set_option("max_slices", args.chunk)

        def step_e2e():
            plan.run(hx.array, hm.array, out=ho.array, params=params)

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        plan.event_record(2)
        for _ in range(args.steps):
            step_e2e()
        plan.event_record(3)
        e_ms = plan.event_elapsed_ms(2, 3)
        barrier()
        te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_its_per_step * (ns_e / ns) * args.steps / (float(te[0]) * 1e-3), "unit": "slice-iterations/s",
               "h2d_bytes_per_step": int(hx.nbytes + hm.nbytes), "d2h_bytes_per_step": int(ho.nbytes + ns_e * (niter + 2) * 8),
               "ms_per_step": float(te[0]) / args.steps, "slices_per_gpu_per_step": ns_e,
               "api": "PocsPlan.run -> p3d_pocs_run(host pinned in/out)", "lanes": args.lanes if args.lanes > 0 else "auto"}
        checksum = float(np.abs(ho.array[0]).sum())
    else:
        checksum = float(out_dev[0].abs().sum())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (CUDA events around every launch, on the launch stream)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = max(KERNEL_BYTES_PER_ELEM, key=lambda k: prof[k]["ms"])
    launches = prof[dom]["launches"]
    avg_ms = prof[dom]["ms"] / max(launches, 1)
    # slices per launch of the dominant kernel: the complex128 kernels run over the slices already switched
    if dom.endswith("64"):
        slices_per_launch = esc_its * args.steps / max(launches, 1) if precision == 0 else ns * args.steps * niter / max(launches, 1)
    else:
        slices_per_launch = (ns * niter - (esc_its if precision == 0 else 0)) * args.steps / max(launches, 1)
    alg_bytes = KERNEL_BYTES_PER_ELEM[dom] * n1 * n2 * slices_per_launch
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp)).get(dom)
            traffic = tj["dram_bytes_per_slice"] * slices_per_launch if (tj and n1 == 1000 and n2 == 1000) else None
        except Exception:
            traffic = None
    moved = traffic if traffic else alg_bytes
    achieved = moved / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    floor_per_it = (FUSED_FLOOR_32 * (1.0 - f64_share) + FUSED_FLOOR_64 * f64_share) * n1 * n2
    rate1 = value / world
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "note": ("achieved = DRAM bytes the dominant kernel moves per launch (ncu dram__bytes_read + write, profiles/traffic.json; the "
                         "algorithmic bytes of the fused kernel when no capture exists for this shape) / its average launch time (CUDA events)"),
                "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms, "slices_per_launch": slices_per_launch,
                "complex128_share_of_slice_iterations": f64_share, "slices_switched_to_complex128": esc_slices,
                "whole_iteration": {
                    # what one slice-iteration has to move with the fused two-kernel structure (16 + 24 B per element in complex64,
                    # 32 + 40 B in complex128, weighted by the measured share of complex128 slice-iterations)
                    "fused_floor_bytes_per_slice_iteration": floor_per_it,
                    "frac_of_fused_floor": rate1 * floor_per_it / 1e9 / peak,
                    # the SURVEY 8(d) contract model (four streaming passes, complex64): 73 B per element and slice-iteration
                    "contract_73B_bytes_per_slice_iteration": B_ALG_PER_ELEM * n1 * n2,
                    "contract_73B_achieved_GBps": rate1 * B_ALG_PER_ELEM * n1 * n2 / 1e9,
                    "contract_73B_frac_of_measured_peak": rate1 * B_ALG_PER_ELEM * n1 * n2 / 1e9 / peak,
                    "contract_73B_frac_of_8TBps_nominal": rate1 * B_ALG_PER_ELEM * n1 * n2 / 1e9 / 8000.0},
                # FFT flop rate, 5 N log2 N convention: two 2-D transforms + ~30 flops of element-wise work per element
                "fft_tflops": rate1 * (2 * 5 * n1 * n2 * math.log2(n1 * n2) + 30.0 * n1 * n2) / 1e12,
                "kernel_share_of_step": {k: prof[k]["ms"] / max(1e-9, sum(v["ms"] for v in prof.values())) for k in prof if prof[k]["launches"]}}
    gpu_launches = int(sum(v["launches"] for v in prof.values()))

    # ---- diagnostic: cuFFT + elementwise chain (torch.fft) on the same device, same slice shape
    diag = None
    if not args.no_diag:
        try:
            nb = min(16, ns)
            xs = torch.randn((nb, n1, n2), dtype=torch.complex64, device=dev)
            msk = (torch.rand((n1, n2), device=dev) < 0.2).to(torch.float32)
            d0 = xs * msk
            xx = d0.clone()
            keep = (1.0 - msk)
            def chain(nit_):
                nonlocal xx
                for k in range(nit_):
                    X = torch.fft.fft2(xx)
                    X = torch.where(X.abs() < 10.0 + k, torch.zeros_like(X), X)
                    y = torch.fft.ifft2(X)
                    xx = y * keep + d0
                    _ = xx.abs().sum()
            chain(3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); chain(10); e1.record(); torch.cuda.synchronize()
            diag = {"what": "torch.fft.fft2/ifft2 (cuFFT) + torch elementwise chain, hard threshold, same shape",
                    "value": nb * 10 / (e0.elapsed_time(e1) * 1e-3), "unit": "slice-iterations/s", "slices": nb}
        except Exception as ex:          # noqa: BLE001
            diag = {"error": str(ex)[:200]}

    # ---- extras (outside the timed region, small samples): float64 state mode and the time-axis kernels
    extras = None
    if not args.no_diag:
        extras = {}
        try:
            # the two other precision modes on a sample of the same synthetic cube (device-resident)
            nsx = min(128, ns)
            xs, ms_ = synth_device(torch, dev, args.config, c, 300, nsx, nf_full)
            os_ = torch.empty_like(xs)
            for name, prec in (("fp32_only_mode", 32), ("float64_state_mode", 64), ("escalating_mode", 0)):
                if prec == precision:
                    continue
                px = p3d.PocsPlan(n1, n2, device=local, precision=prec)
                px.run_device(xs.data_ptr(), ms_.data_ptr(), os_.data_ptr(), nsx, params)
                px.event_record(0)
                px.run_device(xs.data_ptr(), ms_.data_ptr(), os_.data_ptr(), nsx, params)
                px.event_record(1)
                extras[name] = {"value": nsx * niter / (px.event_elapsed_ms(0, 1) * 1e-3), "unit": "slice-iterations/s",
                                "sample": f"{nsx} slices x {niter} iterations of the synthetic cube, precision={prec or 'auto'}",
                                "plan": px.describe().split("precision=")[-1]}
                px.close()
            del xs, os_
        except Exception as ex:          # noqa: BLE001
            extras["precision_modes"] = {"error": str(ex)[:200]}
        try:
            nsf = min(128, ns)
            xf = torch.randn((nsf, n1, n2), dtype=torch.complex64, device=dev)
            of = torch.empty_like(xf)
            hf = torch.rand((n1, n2), dtype=torch.float32, device=dev)
            plan.kxky_filter_device(xf.data_ptr(), hf.data_ptr(), of.data_ptr(), nsf)
            plan.event_record(2)
            plan.kxky_filter_device(xf.data_ptr(), hf.data_ptr(), of.data_ptr(), nsf)
            plan.event_record(3)
            msf = plan.event_elapsed_ms(2, 3)
            extras["kxky_filter"] = {"value": nsf / (msf * 1e-3), "unit": "slices/s", "algorithmic_GBps": 52.0 * n1 * n2 * nsf / msf / 1e6,
                                     "alg_bytes_per_element": 52, "sample": f"{nsf} slices, ifft2(filter * fft2(x)) in three fused passes"}
            del xf, of, hf
        except Exception as ex:          # noqa: BLE001
            extras["kxky_filter"] = {"error": str(ex)[:200]}
        try:
            import ctypes as C
            nt_, ntr_ = c["nt"], min(n1 * n2, 262144)
            xt = torch.randn((nt_, ntr_), dtype=torch.float32, device=dev)
            ft = torch.empty((nt_ // 2 + 1, ntr_), dtype=torch.complex64, device=dev)
            lib = _lib.load()
            res = {}
            for name, fn in (("time_fft", lambda: lib.p3d_time_fft(local, C.c_void_p(xt.data_ptr()), 1, C.c_void_p(ft.data_ptr()), 1, nt_, nt_, ntr_, 0.05, 725.0, 1, None)),
                             ("time_ifft", lambda: lib.p3d_time_ifft(local, C.c_void_p(ft.data_ptr()), 1, C.c_void_p(xt.data_ptr()), 1, nt_, nt_, ntr_, 0.05, 725.0, 1, 0))):
                _lib.check(fn()); _lib.check(fn())
                ms = C.c_double(); lib.p3d_time_last_kernel_ms(C.byref(ms))
                nbytes = xt.numel() * 4 + ft.numel() * 8
                res[name] = {"kernel_ms": ms.value, "algorithmic_GBps": nbytes / max(ms.value, 1e-9) / 1e6, "path": lib.p3d_time_last_path().decode()}
            et = torch.empty_like(xt)
            fn = lambda: lib.p3d_time_envelope(local, C.c_void_p(xt.data_ptr()), 1, C.c_void_p(et.data_ptr()), 1, nt_, ntr_)   # noqa: E731
            _lib.check(fn()); _lib.check(fn())
            ms = C.c_double(); lib.p3d_time_last_kernel_ms(C.byref(ms))
            res["envelope"] = {"kernel_ms": ms.value, "algorithmic_GBps": 2 * xt.numel() * 4 / max(ms.value, 1e-9) / 1e6, "path": lib.p3d_time_last_path().decode()}
            del et
            res["sample"] = f"{nt_} samples x {ntr_} traces, compute_real"
            extras["time_axis"] = res
            del xt, ft
        except Exception as ex:          # noqa: BLE001
            extras["time_axis"] = {"error": str(ex)[:200]}

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        procs = min(os.cpu_count() or 1, 64)
        it_cpu = min(niter, cpu_niter_for_budget(c, procs, 15.0))
        v, wall, its = cpu_sample(args.config, c, it_cpu, procs, procs)
        cpu = {"value": v, "unit": "slice-iterations/s", "cores": procs, "kind": "port",
               "sample": f"{procs} slices x {it_cpu} iterations, one slice per process ({procs} processes), complex64 input, {wall:.1f} s"}

    line = {"metric": "pocs_slice_iterations_per_s", "value": value, "unit": "slice-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "c64 -> c128 (fp32 pilot, complex128 state after the exact restart)", 32: "c64 (fp32 complex)", 64: "c128"}[precision], "data": "synthetic",
            "config": {"workload": cfg_workload, "l2": "inputs (%.1f GB per GPU) far larger than the 126 MB L2" % (ns * n1 * n2 * 8 / 1e9),
                       "precision": args.precision, "plan": desc, "host": numa + f"; {os.cpu_count()} cpus", "timing": "CUDA events on the library stream, max over ranks",
                       "wall_ms_per_step": wall_ms_max / args.steps},
            "clocks": clocks, "e2e": e2e, "strong": strong, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
            "diag_cufft_chain": diag, "extras": extras, "checksum": checksum}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
