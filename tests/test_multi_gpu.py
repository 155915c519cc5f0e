"""N-GPU runs (one process per GPU, NCCL): the distributed chain 12 -> 13 -> 14 with its two all-to-alls against the
single-GPU chain.  Skipped when fewer than two GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _n_gpus():
    try:
        from pseudo_3d_interpolation_b200 import _lib
        return _lib.load().p3d_device_count()
    except Exception:          # noqa: BLE001
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(shape):
    nt, n_il, n_xl = shape
    rng = np.random.default_rng(5)
    t = np.arange(nt)[:, None, None]
    x = np.zeros(shape)
    for _ in range(4):
        f, p, q = rng.uniform(0.05, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2)
        x += np.cos(f * t + p * np.arange(n_il)[None, :, None] + q * np.arange(n_xl)[None, None, :]) * np.exp(-((t - nt * rng.uniform(0.3, 0.7)) / (0.2 * nt)) ** 2)
    fold = (rng.random((n_il, n_xl)) < 0.3).astype(np.uint8)
    return (x * fold).astype(np.float32), 725.0 + 0.05 * np.arange(nt), fold


PARAMS = dict(niter=12, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-3)


def _worker(rank, world, port, shape, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from pseudo_3d_interpolation_b200 import distributed as pd, pipeline
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    x, twt, fold = _make(shape)
    res = {}
    full, local = pd.interpolate_time_cube_distributed(x, twt, fold, compute_real=True, results=res, **PARAMS)
    ok = bool((res["niterations"] == PARAMS["niter"]).all())
    if rank == 0:
        ref = pipeline.interpolate_time_cube(x, twt, fold, compute_real=True, device=0, **PARAMS)
        err = float(np.linalg.norm(full - ref) / np.linalg.norm(ref))
        obs = fold > 0
        ok = ok and err <= 1e-5 and np.allclose(full[:, obs], x[: full.shape[0]][:, obs], atol=2e-5 * np.abs(x).max())
        ret["err"] = err
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(512, 37, 48), (300, 200, 21)])
def test_distributed_chain_nccl(shape):
    world = min(_n_gpus(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert [ret.get(r) for r in range(world)] == [True] * world, dict(ret)
