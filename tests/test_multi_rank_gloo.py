"""world_size-2 (and 3) gloo runs of the N-GPU host logic on CPU: contiguous band sharding,
no data-path collective, gather of the finished bands.  The per-band compute is the numpy
oracle here (injected); on GPUs it is `pocs_cube` on device LOCAL_RANK."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ns, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pocs_oracle as orc
    from oracle.golden_cases import make_input
    from pseudo_3d_interpolation_b200 import distributed as pd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, mask = make_input(dict(seed=5, shape=(12, 10), keep=0.5))
    cube = np.stack([x * (1 + 0.1 * s) for s in range(ns)]).astype(np.complex64)
    params = dict(niter=6, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)
    calls = []

    def proc(band, fm, **kw):
        calls.append(band.shape[0])
        return orc.pocs_cube(band, fm, **kw)

    full, local = pd.pocs_cube_distributed(cube, mask, process_fn=proc, **params)
    lo, hi = pd.rank_band(ns, rank, world)
    ok = local.shape[0] == hi - lo and (calls == [hi - lo] or hi == lo)
    if rank == 0:
        ref = orc.pocs_cube(cube, mask, **params)
        ok = ok and full is not None and np.array_equal(full, ref)
    else:
        ok = ok and full is None
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ns", [(2, 7), (3, 5), (2, 1)])
def test_band_sharding_and_gather(world, ns):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, ns, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert [ret.get(r) for r in range(world)] == [True] * world


def _chain_worker(rank, world, port, shape, compute_real, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import pocs_oracle as orc, time_axis_oracle as tor
    from pseudo_3d_interpolation_b200 import distributed as pd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nt, n_il, n_xl = shape
    rng = np.random.default_rng(3)
    t = np.arange(nt)[:, None, None]
    x = (np.cos(0.3 * t + 0.2 * np.arange(n_il)[None, :, None]) * np.exp(-((t - nt / 2) / (0.2 * nt)) ** 2)
         + 0.1 * rng.standard_normal((nt, n_il, n_xl))).astype(np.float32)
    fold = (rng.random((n_il, n_xl)) < 0.5).astype(np.uint8) * 2          # fold = 2 exercises min(fold, 1)
    x *= (fold > 0)
    twt = 725.0 + 0.05 * np.arange(nt)
    params = dict(niter=5, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)
    dt, t0 = 0.05, 725.0
    nte = nt - nt % 2

    def fft_fn(xl):            # (nt, ntr_loc) -> (nf, ntr_loc)
        F, _ = tor.time_fft(xl.reshape(nte, -1, 1), twt[:nte], compute_real=compute_real)
        return torch.from_numpy(np.ascontiguousarray(F.reshape(F.shape[0], -1)))

    def pocs_fn(band, mask):
        b = band.numpy().reshape(-1, n_il, n_xl)
        out = orc.pocs_cube(b, mask, **params) if b.shape[0] else b
        return torch.from_numpy(np.ascontiguousarray(out.reshape(b.shape[0], -1))), np.full(b.shape[0], 5, np.int32)

    def ifft_fn(Fl):
        F = Fl.numpy().reshape(Fl.shape[0], -1, 1)
        return tor.time_ifft(F, dt, t0, compute_real=compute_real, ascending=False).reshape(nte, -1)

    res = {}
    full, local = pd.interpolate_time_cube_distributed(x, twt, fold, compute_real=compute_real, steps=(fft_fn, pocs_fn, ifft_fn),
                                                       results=res, **params)
    i0, i1 = pd.band_bounds(n_il, world)[rank]
    ok = local.shape == (nte, i1 - i0, n_xl) and res["band"] == pd.band_bounds(nte // 2 + 1 if compute_real else nte, world)[rank]
    if rank == 0:
        # single-process chain with the same oracle steps
        F, _ = tor.time_fft(x[:nte], twt[:nte], compute_real=compute_real)
        Y = orc.pocs_cube(F.astype(np.complex64), np.minimum(fold, 1), **params)
        ref = tor.time_ifft(Y, dt, t0, compute_real=compute_real, ascending=False)
        ok = ok and full is not None and full.shape == ref.shape and np.allclose(full, ref, rtol=0, atol=1e-5 * np.abs(ref).max())
    else:
        ok = ok and full is None
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,compute_real", [(2, (16, 5, 4), True), (3, (13, 7, 3), False), (2, (8, 1, 6), True)])
def test_distributed_chain_all_to_all(world, shape, compute_real):
    """steps 12 -> 13 -> 14 over `world` ranks: trace-sharded transforms, slice-sharded iterations, two all-to-alls
    (uneven iline blocks and frequency bands, a rank with no ilines, odd record length)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_chain_worker, args=(r, world, port, shape, compute_real, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert [ret.get(r) for r in range(world)] == [True] * world
