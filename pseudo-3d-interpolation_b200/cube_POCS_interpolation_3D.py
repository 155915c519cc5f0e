"""Step 13 mirror: interpolate a sparse 3-D cube with FFT-POCS on the GPU(s).

Same command line and YAML keys as the reference script
(cube_POCS_interpolation_3D.py:68-84, 210-236; docs/3D/3D_cube_interpolation.md:126-173):

    13_cube_interpolate_POCS path_cube --path_pocs_parameter cfg.yml [--path_output_dir D] [--verbose {0,1,2}]

YAML: ``dim``, ``var``, ``metadata{transform_kind, niter, eps, thresh_op, thresh_model, alpha,
p_max, p_min, sqrt_decay, decay_kind, version}``, ``output_runtime_results``; the dask keys
(``n_workers``, ``processes``, ``threads_per_worker``, ``memory_limit``, ``batch_chunk``) are
accepted and ignored: the slice loop of cube_POCS_interpolation_3D.py:303-340 is replaced by
contiguous frequency bands over the visible GPUs (``n_gpus`` key or all of them).

Extensions (not in the reference): the YAML key ``precision`` ("auto" = escalating fp32 -> complex128, the default;
32; 64) and the flag ``--fused``: ``path_cube`` is then the TIME-domain cube of step 11 and steps 12 -> 13 -> 14 run
chained on the device (``pipeline.interpolate_time_cube`` / ``distributed.interpolate_time_cube_distributed``
under torchrun): one upload, no frequency-domain files; ``--compute_real`` as in steps 12 / 14.  The output is the
interpolated time cube ``{file}_interp{suffix}`` (variable ``<var>_interp``), as step 14 would write it.

Outputs follow the reference's names: directory ``{file}_{TRANSFORM}_{thresh_op}_niter-{niter}``
with ``parameter_{prefix}.yml`` (and ``runtimes_{prefix}.txt``), and the merged cube
``{out_dir}{suffix}`` whose frequency axis is ascending like the reference's
``open_mfdataset`` merge (cube_POCS_interpolation_3D.py:394-405); complex results are stored as
``<var>_interp.real`` / ``.imag`` float32 (cube_POCS_interpolation_3D.py:160-164).
"""
from __future__ import annotations

import argparse
import datetime
import os
import sys
import time

import numpy as np
import yaml

from . import _lib
from .cube_io import Cube, open_cube, write_cube
from .pocs import mask_from_fold, pocs_cube

POCS_VERSIONS = {"POCS": "regular", "FPOCS": "fast", "APOCS": "adaptive"}


def define_input_args():  # noqa
    parser = argparse.ArgumentParser(description="Interpolate sparse 3D cube using POCS algorithm.")
    parser.add_argument("path_cube", type=str, help="Input path of 3D cube")
    parser.add_argument("--path_pocs_parameter", type=str, required=True,
                        help="Path of netCDF parameter file (YAML format).")
    parser.add_argument("--path_output_dir", type=str, help="Output directory for interpolated slices.")
    parser.add_argument("--verbose", "-V", type=int, nargs="?", default=0, choices=[0, 1, 2],
                        help="Level of output verbosity (default: 0)")
    # extension: steps 12 -> 13 -> 14 chained on the device, from and to the time-domain cube
    parser.add_argument("--fused", action="store_true",
                        help="path_cube is the time-domain cube: apply FFT, POCS and IFFT on the GPU without intermediate files.")
    parser.add_argument("--compute_real", action="store_true",
                        help="(with --fused) transform assuming real input, discarding the redundant negative frequencies.")
    return parser


def xprint(*args, kind="info", verbosity=0, **kwargs):
    """Verbosity-gated print with the reference's prefixes (functions/utils.py:57-76)."""
    levels = {"info": ("[INFO]  ", 1), "warning": ("[WARN]  ", 0), "error": ("[ERROR]  ", 0),
              "success": ("[SUCCESS]  ", 1), "debug": ("[DEBUG]  ", 2)}
    prefix, lvl = levels.get(kind, ("", 1))
    if lvl <= (1 if verbosity is True else verbosity):
        print(prefix, *args, **kwargs)


def interpolate(cube: Cube, cfg: dict, devices=None, verbose=0):
    """Numeric content of the reference's ``main`` on an in-memory cube -> (Cube, per-slice info)."""
    metadata = dict(cfg["metadata"])
    metadata["transform_kind"] = str(metadata["transform_kind"]).upper()
    if metadata["transform_kind"] != "FFT":
        raise ValueError(f'Transform < {metadata["transform_kind"]} > is not supported.')
    dim = cfg["dim"]
    var = cfg.get("var") or [v for v in cube.data_vars if v != "fold"][0]
    dims, data = cube.variables[var]
    if tuple(dims) != (dim, "iline", "xline"):
        order = [dims.index(d) for d in (dim, "iline", "xline")]      # apply_ufunc moves the core dims last
        data = np.transpose(data, order)
    fold = cube.data("fold")
    if cube.dims_of("fold") == ("xline", "iline"):
        fold = fold.T
    mask = mask_from_fold(fold)                                       # cube_POCS_interpolation_3D.py:242-244
    if devices is None:
        n = _lib.require_gpu()
        devices = list(range(min(int(cfg.get("n_gpus", n)), n)))
    results = {}
    t0 = time.perf_counter()
    out = pocs_cube(np.ascontiguousarray(data), mask, devices=devices, results=results, precision=cfg.get("precision"), **metadata)
    runtime = time.perf_counter() - t0
    xprint(f"POCS on {len(devices)} GPU(s): {data.shape[0]} slices in {runtime:.2f} s", kind="info", verbosity=verbose)

    # merged output with ascending coordinate (open_mfdataset(..., combine='by_coords'))
    coord = np.asarray(cube.coords[dim])
    order = np.argsort(coord, kind="stable")
    res = Cube(attrs=dict(cube.attrs), coord_attrs=dict(cube.coord_attrs), var_attrs=dict(cube.var_attrs))
    res.coords = {k: np.asarray(v) for k, v in cube.coords.items()}
    res.coords[dim] = coord[order]
    res.variables[f"{var}_interp"] = ((dim, "iline", "xline"), out[order])
    res.variables["fold"] = (("iline", "xline"), np.asarray(fold))
    res.var_attrs[f"{var}_interp"] = dict(cube.var_attrs.get(var, {}))
    script = os.path.basename(__file__)
    today = datetime.date.today().strftime("%Y-%m-%d")
    domain = "(frequency domain)" if "freq" in dim else "(time domain)"
    exclude = ("transform", "itransform", "results_dict", "path_results")
    res.attrs.update({
        "description": f'Interpolated pseudo-3D cube using {metadata["transform_kind"]} transform created from TOPAS profiles {domain}',
        "interp_params_keys": ";".join(k for k in metadata if k not in exclude),
        "interp_params_vals": ";".join(str(metadata[k]) for k in metadata if k not in exclude),
        "history": cube.attrs.get("history", "") + f'{script}:{metadata["transform_kind"]} {domain};',
        "text": cube.attrs.get("text", "") + f'\n{today}: {metadata["transform_kind"]} {domain.upper()}',
    })
    info = dict(niterations=results["niterations"][order], cost=results["cost"][order], runtime=runtime)
    return res, info


def interpolate_fused(cube: Cube, cfg: dict, compute_real=True, device=0, verbose=0):
    """Steps 12 -> 13 -> 14 on the device for a time-domain cube -> (Cube with ``<var>_interp``, info)."""
    from .pipeline import interpolate_time_cube
    metadata = dict(cfg["metadata"])
    metadata["transform_kind"] = str(metadata["transform_kind"]).upper()
    if metadata["transform_kind"] != "FFT":
        raise ValueError(f'Transform < {metadata["transform_kind"]} > is not supported.')
    var = cfg.get("var") or [v for v in cube.data_vars if v != "fold"][0]
    dims, data = cube.variables[var]
    if "twt" not in dims:
        raise ValueError("--fused needs the time-domain cube (dimension 'twt')")
    if tuple(dims) != ("twt", "iline", "xline"):
        data = np.transpose(data, [dims.index(d) for d in ("twt", "iline", "xline")])
    fold = cube.data("fold")
    if cube.dims_of("fold") == ("xline", "iline"):
        fold = fold.T
    results = {}
    t0 = time.perf_counter()
    out = interpolate_time_cube(np.ascontiguousarray(data, dtype=np.float32), np.asarray(cube.coords["twt"], dtype=np.float64), fold,
                                compute_real=compute_real, device=device, precision=cfg.get("precision"), results=results, **metadata)
    runtime = time.perf_counter() - t0
    xprint(f"FFT -> POCS -> IFFT on the device: {len(results['niterations'])} slices in {runtime:.2f} s", kind="info", verbosity=verbose)
    res = Cube(attrs=dict(cube.attrs), coord_attrs=dict(cube.coord_attrs), var_attrs=dict(cube.var_attrs))
    res.coords = {k: np.asarray(v) for k, v in cube.coords.items()}
    res.coords["twt"] = np.asarray(cube.coords["twt"])[: out.shape[0]]
    res.variables[f"{var}_interp"] = (("twt", "iline", "xline"), out)
    res.variables["fold"] = (("iline", "xline"), np.asarray(fold))
    res.var_attrs[f"{var}_interp"] = dict(cube.var_attrs.get(var, {}))
    script = os.path.basename(__file__)
    res.attrs["history"] = cube.attrs.get("history", "") + f"{script}:FFT({var});{script}:FFT (frequency domain);{script}:IFFT({var}_interp);"
    return res, dict(niterations=results["niterations"], cost=results["cost"], runtime=runtime)


def main(argv=sys.argv, return_dataset=False):
    """Interpolate sparse 3D cube."""
    args = define_input_args().parse_args(argv[1:])
    verbose = args.verbose
    if args.fused:
        with open(args.path_pocs_parameter, mode="r") as f:
            cfg = yaml.safe_load(f)
        cube = open_cube(args.path_cube)
        res, _ = interpolate_fused(cube, cfg, compute_real=args.compute_real, verbose=verbose)
        base, suffix = os.path.splitext(args.path_cube)
        out_dir = args.path_output_dir
        out_file = os.path.join(out_dir, os.path.basename(base) + "_interp" + suffix) if out_dir else base + "_interp" + suffix
        if out_dir and not os.path.isdir(out_dir):
            os.mkdir(out_dir)
        write_cube(out_file, res, split_complex=True)
        return res if return_dataset else None
    xprint("Load POCS parameter from config file", kind="info", verbosity=verbose)
    with open(args.path_pocs_parameter, mode="r") as f:
        cfg = yaml.safe_load(f)
    cfg["metadata"]["transform_kind"] = cfg["metadata"]["transform_kind"].upper()
    metadata = cfg["metadata"]
    transform = metadata["transform_kind"]

    path_cube = args.path_cube
    dir_work, file = os.path.split(path_cube)
    filename, suffix = os.path.splitext(file)
    prefix = f"{filename}_{transform}_{metadata['thresh_op']}_niter-{metadata['niter']}"
    out_path = args.path_output_dir if args.path_output_dir is not None else os.path.join(dir_work, prefix)
    if not os.path.isdir(out_path):
        os.mkdir(out_path)
    with open(os.path.join(out_path, f"parameter_{prefix}.yml"), mode="w", newline="\n") as f:
        yaml.safe_dump(dict(metadata), f)

    cube = open_cube(path_cube)
    res, info = interpolate(cube, cfg, verbose=verbose)

    if cfg.get("output_runtime_results"):
        # one line per slice: niterations;runtime;cost (per-slice runtimes do not exist on the GPU:
        # the band's wall time is divided evenly)
        per = info["runtime"] / max(len(info["niterations"]), 1)
        with open(os.path.join(out_path, f"runtimes_{prefix}.txt"), mode="w", newline="\n") as f:
            for n, c in zip(info["niterations"], info["cost"]):
                f.write(f"{int(n)};{per};{c}\n")

    xprint("Write combined file to disk", kind="info", verbosity=verbose)
    write_cube(f"{out_path}{suffix}", res, split_complex=True)
    if return_dataset:
        return res


if __name__ == "__main__":
    main()
