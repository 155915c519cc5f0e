"""Multi-process (one process per GPU) sharding of the slice axis.

Frequency slices are independent POCS problems (cube_POCS_interpolation_3D.py:303-340; the only
shared input is the 2-D mask), so the N-GPU path has no data-path collective: rank r owns the
contiguous band ``band_bounds(n_slices, world)[r]`` and runs every iteration locally.  The only
communication is the optional gather of the finished bands onto one rank (and whatever the caller
uses for timing barriers).  ``process_fn`` is injectable so the host-side logic can be tested on
CPU with the gloo backend.
"""
from __future__ import annotations

import numpy as np

from .pocs import band_bounds


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(device_index: int, sysfs="/sys"):
    """NUMA node of a CUDA device from its PCI address (``/sys/bus/pci/devices/<bdf>/numa_node``); None if unknown."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()
        if len(bdf.split(":")[0]) == 8:           # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bdf = bdf[4:]
        with open(f"{sysfs}/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:            # noqa: BLE001
        return None


def bind_to_gpu_numa_node(device_index: int, sysfs="/sys"):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned staging buffers are
    allocated, so that they are first-touched on the memory next to the GPU's PCIe root and H2D / D2H copies do
    not cross the socket interconnect (one process per GPU, as the reference's LocalCluster workers).
    Returns a short description of what was done; never raises."""
    import os
    node = gpu_numa_node(device_index, sysfs)
    if node is None:
        return "numa: unknown (not bound)"
    try:
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"numa: node {node} has no allowed cpus (not bound)"
        os.sched_setaffinity(0, allowed)
        return f"numa: gpu {device_index} -> node {node}, {len(allowed)} cpus"
    except Exception as ex:      # noqa: BLE001
        return f"numa: bind failed ({ex})"


def rank_band(n_slices: int, rank: int, world: int):
    return band_bounds(n_slices, world)[rank]


def pocs_cube_distributed(cube, fold_or_mask, process_fn=None, gather_to=0, group=None, **metadata):
    """Every rank passes the same host ``cube`` (or just its own band placed at the right
    offsets); returns the full result on rank ``gather_to`` (None elsewhere) and the local band.

    process_fn(band, fold_or_mask, **metadata) -> ndarray defaults to the GPU path
    (`pocs_cube` on device LOCAL_RANK).
    """
    import os
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    ns = cube.shape[0]
    lo, hi = rank_band(ns, rank, world)
    if process_fn is None:
        from .pocs import pocs_cube
        dev = int(os.environ.get("LOCAL_RANK", "0"))

        def process_fn(band, fm, **kw):
            return pocs_cube(band, fm, devices=[dev], **kw)
    local = process_fn(cube[lo:hi], fold_or_mask, **metadata) if hi > lo else np.empty((0,) + cube.shape[1:], cube.dtype)
    if world == 1:
        return local, local
    # gather variable-length bands: pad to the largest band, all_gather, trim (gloo and nccl both support it)
    per = -(-ns // world)
    is_c = np.iscomplexobj(local)
    flat = np.ascontiguousarray(local).view(np.float32 if local.dtype in (np.complex64, np.float32) else np.float64)
    pad = np.zeros((per,) + flat.shape[1:], dtype=flat.dtype)
    pad[: hi - lo] = flat
    t = torch.from_numpy(pad)
    backend = dist.get_backend(group)
    if backend == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    if rank != gather_to:
        return None, local
    out = np.empty(cube.shape, dtype=local.dtype)
    for r, (a, b) in enumerate(band_bounds(ns, world)):
        if b > a:
            arr = parts[r].cpu().numpy()[: b - a]
            out[a:b] = arr.view(local.dtype) if is_c else arr
    return out, local
