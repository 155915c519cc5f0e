"""Host <-> device copy bandwidth and end-to-end POCS rate per rank, all ranks concurrently, without and with
NUMA binding (development diagnostic; run under torchrun on 1..8 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/diag_pcie.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pseudo_3d_interpolation_b200 as p3d                      # noqa: E402
from pseudo_3d_interpolation_b200 import _lib                   # noqa: E402
from pseudo_3d_interpolation_b200.distributed import bind_to_gpu_numa_node, gpu_numa_node   # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    log = []

    def say(*a):
        log.append(" ".join(str(x) for x in a))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) if os.path.isdir("/sys/devices/system/node") else []
    say(f"rank {rank}/{world}: cpus={os.cpu_count()} affinity={len(os.sched_getaffinity(0))} numa_nodes={nodes} gpu_node={gpu_numa_node(local)}")

    nb = 1 << 30
    d = torch.empty(nb, dtype=torch.uint8, device=dev)

    def bw(tag):
        h = torch.empty(nb, dtype=torch.uint8).pin_memory()
        h.fill_(1)
        res = {}
        for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
            fn(); barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                fn()
            torch.cuda.synchronize()
            res[name] = 4 * nb / (time.perf_counter() - t0) / 1e9
            barrier()
        # both directions at once on two streams
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        h2 = torch.empty(nb, dtype=torch.uint8).pin_memory(); d2 = torch.empty_like(d)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        res["bidir_each"] = 4 * nb / (time.perf_counter() - t0) / 1e9
        barrier()
        say(f"  [{tag}] concurrent on {world} ranks: H2D {res['h2d']:.1f} GB/s, D2H {res['d2h']:.1f} GB/s, simultaneous {res['bidir_each']:.1f} GB/s each way")

    def e2e(tag, ns=128, niter=25):
        n1 = n2 = 1000
        rng = np.random.default_rng(rank)
        hx = _lib.PinnedArray((ns, n1, n2), np.complex64); ho = _lib.PinnedArray((ns, n1, n2), np.complex64)
        mask = (rng.random((n1, n2)) < 0.2).astype(np.uint8)
        base = (rng.standard_normal((n1, n2)) + 1j * rng.standard_normal((n1, n2))).astype(np.complex64) * mask
        hx.array[:] = base[None]
        plan = p3d.PocsPlan(n1, n2, device=local)
        params = p3d.make_params(niter=niter, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0)
        plan.run(hx.array, mask, out=ho.array, params=params)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            plan.run(hx.array, mask, out=ho.array, params=params)
        dt = (time.perf_counter() - t0) / 2
        barrier()
        # device-resident for comparison
        dx = torch.from_numpy(np.ascontiguousarray(hx.array)).to(dev); do = torch.empty_like(dx); dm = torch.from_numpy(mask).to(dev)
        plan.run_device(dx.data_ptr(), dm.data_ptr(), do.data_ptr(), ns, params)
        barrier()
        t0 = time.perf_counter()
        plan.run_device(dx.data_ptr(), dm.data_ptr(), do.data_ptr(), ns, params)
        torch.cuda.synchronize()
        dd = time.perf_counter() - t0
        barrier()
        say(f"  [{tag}] {ns} slices x {niter} it: e2e {dt*1e3:.1f} ms ({ns*niter/dt:.0f} s-it/s), device-resident {dd*1e3:.1f} ms ({ns*niter/dd:.0f} s-it/s)")
        plan.close()

    bw("unbound")
    e2e("unbound")
    say("  " + bind_to_gpu_numa_node(local))
    bw("bound")
    e2e("bound")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"diag_pcie_w{world}_r{rank}.log"), "w") as f:
        f.write("\n".join(log) + "\n")
    if rank == 0:
        print("\n".join(log))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
