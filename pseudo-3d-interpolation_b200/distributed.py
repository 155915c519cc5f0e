"""Multi-process (one process per GPU) sharding of the slice axis.

Frequency slices are independent POCS problems (cube_POCS_interpolation_3D.py:303-340; the only
shared input is the 2-D mask), so the N-GPU path has no data-path collective: rank r owns the
contiguous band ``band_bounds(n_slices, world)[r]`` and runs every iteration locally.  The only
communication is the optional gather of the finished bands onto one rank (and whatever the caller
uses for timing barriers).  ``process_fn`` is injectable so the host-side logic can be tested on
CPU with the gloo backend.
"""
from __future__ import annotations

import os

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
        t = t.to(torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
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


# ------------------------------------------------------------------------------------------------------
# Steps 12 -> 13 -> 14 chained on N GPUs (second row, and 8f-1).
#
# The time-axis transforms are independent per TRACE, the POCS iterations independent per FREQUENCY SLICE.
# Rank r therefore owns a block of ilines for steps 12 / 14 and a contiguous band of frequency slices for
# step 13, and the spectrum changes hands twice: trace-sharded -> slice-sharded before the iterations and
# back after them.  That is the one real exchange step of the path: an all-to-all (NCCL over NVLink 5 /
# NVSwitch on GPUs, gloo in the CPU tests).  Config 2 moves 8.2 GB * (G-1)/G in total per direction.
# ------------------------------------------------------------------------------------------------------
def _a2a(dist, out_flat, in_flat, out_splits, in_splits, group):
    """all_to_all_single on the float32 view of complex64 buffers (NCCL has no complex type)."""
    import torch
    o = torch.view_as_real(out_flat).reshape(-1) if out_flat.is_complex() else out_flat
    i = torch.view_as_real(in_flat).reshape(-1) if in_flat.is_complex() else in_flat
    k = 2 if in_flat.is_complex() else 1
    dist.all_to_all_single(o, i, [s * k for s in out_splits], [s * k for s in in_splits], group=group)


def exchange_traces_to_band(F_local, f_bands, ntr_per_rank, rank, group=None):
    """``F_local`` (nf, ntr_local): all frequencies of this rank's traces  ->  (nf_r, ntr_total): this rank's
    frequency band of every trace (rank q's traces in columns ``sum(ntr_per_rank[:q]) ...``)."""
    import torch
    import torch.distributed as dist
    nf, ntr_loc = F_local.shape
    world = len(f_bands)
    nf_r = f_bands[rank][1] - f_bands[rank][0]
    in_splits = [(b - a) * ntr_loc for a, b in f_bands]                 # rows [a, b) are contiguous: no packing
    out_splits = [nf_r * n for n in ntr_per_rank]
    recv = torch.empty(sum(out_splits), dtype=F_local.dtype, device=F_local.device)
    _a2a(dist, recv, F_local.reshape(-1), out_splits, in_splits, group)
    band = torch.empty((nf_r, sum(ntr_per_rank)), dtype=F_local.dtype, device=F_local.device)
    off = col = 0
    for q in range(world):
        band[:, col:col + ntr_per_rank[q]] = recv[off:off + out_splits[q]].view(nf_r, ntr_per_rank[q])
        off += out_splits[q]; col += ntr_per_rank[q]
    return band


def exchange_band_to_traces(band, f_bands, ntr_per_rank, rank, group=None):
    """Inverse of :func:`exchange_traces_to_band`: (nf_r, ntr_total) -> (nf, ntr_local)."""
    import torch
    import torch.distributed as dist
    world = len(f_bands)
    nf_r = band.shape[0]
    nf = f_bands[-1][1]
    ntr_loc = ntr_per_rank[rank]
    send = torch.empty(band.numel(), dtype=band.dtype, device=band.device)
    in_splits, off, col = [], 0, 0
    for q in range(world):
        n = nf_r * ntr_per_rank[q]
        send[off:off + n].view(nf_r, ntr_per_rank[q]).copy_(band[:, col:col + ntr_per_rank[q]])
        in_splits.append(n); off += n; col += ntr_per_rank[q]
    out_splits = [(b - a) * ntr_loc for a, b in f_bands]                # received blocks stack along frequency
    F_local = torch.empty((nf, ntr_loc), dtype=band.dtype, device=band.device)
    _a2a(dist, F_local.reshape(-1), send, out_splits, in_splits, group)
    return F_local


def _gpu_steps(device, n_il, n_xl, twt, compute_real, precision, metadata):
    """Default step functions: the CUDA library on device tensors (torch only owns the memory)."""
    import ctypes as C
    import torch
    from . import _lib
    from .pocs import get_plan, make_params
    lib = _lib.load()
    _lib.require_gpu()
    params = make_params(**metadata)
    nt = len(twt)
    dt, t0 = float(twt[1] - twt[0]), float(twt[0])
    nf = nt // 2 + 1 if compute_real else nt
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)                 # collectives and `.cuda()` copies of this process target the rank's own GPU

    def drain():
        # the library runs on its own non-blocking streams: everything torch enqueued (all-to-all, repacking copies,
        # uploads) must have finished before a library call reads it, and vice versa the library calls return drained
        torch.cuda.current_stream(dev).synchronize()

    def fft_fn(x_local):                       # (nt, ntr_loc) float32 host -> (nf, ntr_loc) complex64 device
        xd = torch.from_numpy(x_local).to(dev)
        F = torch.empty((nf, x_local.shape[1]), dtype=torch.complex64, device=dev)
        drain()
        _lib.check(lib.p3d_time_fft(device, C.c_void_p(xd.data_ptr()), _lib.MEM_DEVICE, C.c_void_p(F.data_ptr()), _lib.MEM_DEVICE,
                                    nt, nt, x_local.shape[1], dt, t0, 1 if compute_real else 0, None))
        return F

    def pocs_fn(band, mask):                   # (nf_r, n_il * n_xl) complex64 device
        out = torch.empty_like(band)
        md = torch.from_numpy(mask).to(dev)
        nit = np.zeros(band.shape[0], np.int32)
        drain()
        if band.shape[0]:
            get_plan(n_il, n_xl, device, precision).run_device(band.data_ptr(), md.data_ptr(), out.data_ptr(), band.shape[0], params, nit=nit)
        return out, nit

    def ifft_fn(F_local):                      # (nf, ntr_loc) complex64 device -> (nt, ntr_loc) float32 host
        x = torch.empty((nt, F_local.shape[1]), dtype=torch.float32, device=dev)
        drain()
        _lib.check(lib.p3d_time_ifft(device, C.c_void_p(F_local.data_ptr()), _lib.MEM_DEVICE, C.c_void_p(x.data_ptr()), _lib.MEM_DEVICE,
                                     nt, nt, F_local.shape[1], dt, t0, 1 if compute_real else 0, 0))
        return x.cpu().numpy()

    return fft_fn, pocs_fn, ifft_fn


def interpolate_time_cube_distributed(x, twt, fold, compute_real=True, steps=None, gather_to=0, group=None,
                                      precision=None, results=None, **metadata):
    """Steps 12 -> 13 -> 14 on all ranks of ``group``: ``x`` (nt, n_il, n_xl) float32 sparse time cube (every rank
    passes the same host array, or at least its own iline block filled in), ``twt`` (nt,), ``fold`` (n_il, n_xl).

    Rank r transforms the traces of its iline block (``band_bounds(n_il, world)[r]``), receives its frequency band
    (``band_bounds(nf, world)[r]``) of ALL traces in one all-to-all, runs every POCS iteration locally, returns the
    band in a second all-to-all and inverse-transforms its own traces.  Returns ``(full, local)``: the interpolated
    time cube on rank ``gather_to`` (None elsewhere) and this rank's iline block ``(nt_even, n_il_r, n_xl)``.

    ``steps = (fft_fn, pocs_fn, ifft_fn)`` is injectable (CPU tests run the numpy oracle under gloo)."""
    import os
    import torch
    import torch.distributed as dist
    from .pocs import mask_from_fold

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    x = np.asarray(x)
    twt = np.asarray(twt, dtype=np.float64)
    if x.shape[0] % 2:                           # cube_apply_FFT.py:223-233
        x, twt = x[:-1], twt[:-1]
    nt, n_il, n_xl = x.shape
    nf = nt // 2 + 1 if compute_real else nt
    il_blocks = band_bounds(n_il, world)
    f_bands = band_bounds(nf, world)
    ntr_per_rank = [(b - a) * n_xl for a, b in il_blocks]
    i0, i1 = il_blocks[rank]
    mask = np.ascontiguousarray(mask_from_fold(fold), dtype=np.uint8)
    for k in ("transform", "itransform", "transform_kind", "auxiliary_data", "verbose", "results_dict", "path_results"):
        metadata.pop(k, None)
    if steps is None:
        steps = _gpu_steps(int(os.environ.get("LOCAL_RANK", "0")), n_il, n_xl, twt, compute_real, precision, metadata)
    fft_fn, pocs_fn, ifft_fn = steps

    x_local = np.ascontiguousarray(x[:, i0:i1, :], dtype=np.float32).reshape(nt, -1)
    F_local = fft_fn(x_local)                                                        # step 12, own traces
    if world > 1:
        band = exchange_traces_to_band(F_local, f_bands, ntr_per_rank, rank, group)  # trace-sharded -> slice-sharded
    else:
        band = F_local
    band, nit = pocs_fn(band, mask)                                                  # step 13, own frequency band
    if world > 1:
        F_local = exchange_band_to_traces(band, f_bands, ntr_per_rank, rank, group)  # and back
    else:
        F_local = band
    local = np.asarray(ifft_fn(F_local)).reshape(nt, i1 - i0, n_xl)                  # step 14, own traces
    if isinstance(results, dict):
        results["niterations"] = nit
        results["band"] = f_bands[rank]
    if world == 1:
        return local, local
    per = -(-n_il // world)
    pad = np.zeros((nt, per, n_xl), dtype=np.float32)
    pad[:, : i1 - i0] = local
    t = torch.from_numpy(pad)
    if dist.get_backend(group) == "nccl":
        t = t.to(torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    if rank != gather_to:
        return None, local
    full = np.empty((nt, n_il, n_xl), dtype=np.float32)
    for r, (a, b) in enumerate(il_blocks):
        if b > a:
            full[:, a:b] = parts[r].cpu().numpy()[:, : b - a]
    return full, local
