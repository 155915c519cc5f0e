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
