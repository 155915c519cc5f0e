"""Time the distributed chain 12 -> 13 -> 14 (development / profiles): per-phase CUDA-synchronised wall times.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29514 tools/bench_chain.py [nt n_il n_xl niter]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_3d_interpolation_b200 import distributed as pd          # noqa: E402
from pseudo_3d_interpolation_b200.pocs import band_bounds           # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    nt, n_il, n_xl, niter = (a + [2048, 1000, 1000, 20][len(a):])[:4]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nf = nt // 2 + 1
    il_blocks = band_bounds(n_il, world); f_bands = band_bounds(nf, world)
    ntr = [(b - a_) * n_xl for a_, b in il_blocks]
    i0, i1 = il_blocks[rank]
    twt = 725.0 + 0.05 * np.arange(nt)
    rng = np.random.default_rng(rank)
    fold = (np.random.default_rng(0).random((n_il, n_xl)) < 0.2).astype(np.uint8)
    x_local = (rng.standard_normal((nt, (i1 - i0) * n_xl)).astype(np.float32)) * fold[i0:i1].reshape(1, -1)
    params = dict(niter=niter, thresh_op="hard", thresh_model="exponential", eps=0.0, alpha=1.0)
    fft_fn, pocs_fn, ifft_fn = pd._gpu_steps(local, n_il, n_xl, twt, True, None, params)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    times = {}
    for it in range(2):
        sync(); t0 = time.perf_counter()
        F = fft_fn(x_local); sync(); t1 = time.perf_counter()
        band = pd.exchange_traces_to_band(F, f_bands, ntr, rank) if world > 1 else F
        sync(); t2 = time.perf_counter()
        band, nit = pocs_fn(band, fold); sync(); t3 = time.perf_counter()
        Fb = pd.exchange_band_to_traces(band, f_bands, ntr, rank) if world > 1 else band
        sync(); t4 = time.perf_counter()
        out = ifft_fn(Fb); sync(); t5 = time.perf_counter()
        times = dict(upload_fft=t1 - t0, a2a_fwd=t2 - t1, pocs=t3 - t2, a2a_back=t4 - t3, ifft_download=t5 - t4, total=t5 - t0)
        del F, band, Fb
    if rank == 0:
        gb = nf * n_il * n_xl * 8 / 1e9
        print(f"chain {nt}x{n_il}x{n_xl}, {niter} iterations, {world} GPU(s): " + ", ".join(f"{k} {v*1e3:.1f} ms" for k, v in times.items()) +
              f"; spectrum {gb:.2f} GB, all-to-all {gb * (world - 1) / max(world, 1) / max(times['a2a_fwd'], 1e-9):.0f} GB/s aggregate incl. repack")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
