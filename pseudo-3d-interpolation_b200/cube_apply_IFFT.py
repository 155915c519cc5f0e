"""Step 14 mirror: inverse FFT along the frequency axis of a (pseudo-)3D cube on the GPU.

Same command line as the reference (cube_apply_IFFT.py:20-32):

    14_cube_apply_IFFT path_cube --params_netcdf Y [--compute_real] [--rescale-envelope] [-V]
"""
from __future__ import annotations

import argparse
import datetime
import os
import sys

import numpy as np
import yaml

from .cube_io import Cube, open_cube, write_cube
from .timeaxis import rescale_envelope, time_ifft


def define_input_args():  # noqa
    parser = argparse.ArgumentParser(description="Apply inverse FFT along frequency axis of (pseudo-)3D cube.")
    parser.add_argument("path_cube", type=str, help="Input path of 3D cube.")
    parser.add_argument("--params_netcdf", type=str, required=True, help="Path of netCDF parameter file (*.yaml).")
    parser.add_argument("--compute_real", action="store_true",
                        help="Compute IFFT assuming real input was used for previously applied FFT.")
    parser.add_argument("--rescale-envelope", action="store_true", help="Rescale envelope data to [0-1].")
    parser.add_argument("--verbose", "-V", type=int, nargs="?", default=0, const=1, choices=[0, 1, 2],
                        help="Level of output verbosity (default: 0)")
    return parser


def apply_ifft(cube: Cube, compute_real=False, rescale=False, kwargs_nc=None, script="cube_apply_IFFT", device=0):
    """Numeric content of the reference's ``main`` on an in-memory cube."""
    today = datetime.date.today().strftime("%Y-%m-%d")
    dim = cube.other_dim()
    prefix = dim.split("_")[0]
    names = cube.data_vars
    first = [v for v in names if prefix in v][0]
    var = cube.var_attrs.get(first.split(".")[0], {}).get("original_var", "_".join(first.split(".")[0].split("_")[1:]))
    var_real = [v for v in names if "real" in v]
    var_imag = [v for v in names if "imag" in v]
    if var_real and var_imag:                                   # cube_apply_IFFT.py:73-79
        dims = cube.dims_of(var_real[0])
        spec = cube.data(var_real[0]).astype(np.complex64)
        spec.imag = cube.data(var_imag[0])
    else:
        dims = cube.dims_of(first)
        spec = cube.data(first)
    if tuple(dims) != (dim, "iline", "xline"):
        spec = np.transpose(spec, [dims.index(d) for d in (dim, "iline", "xline")])
    f = np.asarray(cube.coords[dim], dtype=np.float64)
    ca = cube.coord_attrs.get(dim, {})
    nf = f.size
    nfft = 2 * (nf - 1) if compute_real else nf
    df = float(f[1] - f[0])
    dt = 1.0 / (nfft * df)
    # xrft stores the time of sample N/2 as `direct_lag` on the frequency coordinate; t0 = lag - (N/2) dt
    t0 = float(ca.get("direct_lag", 0.0)) - (nfft // 2) * dt
    ascending = bool(np.all(np.diff(f) > 0))
    x = time_ifft(np.ascontiguousarray(spec), dt, t0, compute_real=compute_real, ascending=ascending, device=device)
    if rescale:                                                 # cube_apply_IFFT.py:121-140
        x = rescale_envelope(x).astype(np.float32)

    out = Cube(attrs=dict(cube.attrs), coord_attrs={k: dict(v) for k, v in cube.coord_attrs.items() if k != dim},
               var_attrs={})
    out.coords = {k: np.asarray(v) for k, v in cube.coords.items() if k != dim}
    out.coords["twt"] = (t0 + dt * np.arange(nfft)).astype(np.float32)
    out.variables[var] = (("twt", "iline", "xline"), x)
    fdims, fold = cube.variables["fold"]
    out.variables["fold"] = (fdims, np.asarray(fold))
    out.attrs.update({
        "long_name": cube.attrs.get("long_name", "").split(" (")[0] + " (interpolated)",
        "history": cube.attrs.get("history", "") + f"{script}: IFFT({var});",
        "text": cube.attrs.get("text", "") + f"\n{today}: INVERSE FFT(FREQ -> TIME)",
    })
    out.coord_attrs["twt"] = {}
    if kwargs_nc is not None:
        out.var_attrs[var] = dict(kwargs_nc.get("attrs_time", {}).get(var.split("_")[0], {}))
        out.coord_attrs["twt"].update(kwargs_nc.get("attrs_time", {}).get("twt", {}))
    out.coord_attrs["twt"]["dt"] = float(f"{dt:g}")
    out.coord_attrs["twt"].pop("spacing", None)
    return out


def main(argv=sys.argv, return_dataset=False):  # noqa
    """Apply inverse FFT along _frequency_ axis wrapper function."""
    script = os.path.splitext(os.path.basename(__file__))[0]
    args = define_input_args().parse_args(argv[1:])
    path_in = args.path_cube
    dir_work, file = os.path.split(path_in)
    with open(args.params_netcdf, "r") as f_attrs:
        kwargs_nc = yaml.safe_load(f_attrs)
    cube = open_cube(path_in)
    prefix = cube.other_dim().split("_")[0]
    out = apply_ifft(cube, compute_real=args.compute_real, rescale=args.rescale_envelope, kwargs_nc=kwargs_nc, script=script)
    tsuffix = "_rescale-env" if args.rescale_envelope else ""
    basename, fsuffix = os.path.splitext(file)
    path_out = os.path.join(dir_work, basename.replace(prefix, "twt") + f"_interp-freq{tsuffix}{fsuffix}")
    write_cube(path_out, out)
    if return_dataset:
        return out


if __name__ == "__main__":
    main()
