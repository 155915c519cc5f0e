"""Step 12 mirror: forward FFT along the time axis of a (pseudo-)3D cube on the GPU.

Same command line as the reference (cube_apply_FFT.py:24-45):

    12_cube_apply_FFT path_cube --params_netcdf Y [--prefix freq] [--compute_real]
        [--upsampling-factor N] [--filter {lowpass,highpass,bandpass}] [--filter_freqs ...]
        [--drop-filtered-freq] [-V]
"""
from __future__ import annotations

import argparse
import datetime
import os
import sys
import warnings

import numpy as np
import yaml

from .cube_io import Cube, open_cube, write_cube
from .timeaxis import freq_filter_keep, freq_filter_window, time_fft


def define_input_args():  # noqa
    parser = argparse.ArgumentParser(description="Apply FFT along time axis of (pseudo-)3D cube.")
    parser.add_argument("path_cube", type=str, help="Input path of 3D cube")
    parser.add_argument("--params_netcdf", type=str, required=True, help="Path of netCDF parameter file (YAML format).")
    parser.add_argument("--prefix", type=str, default="freq", help="Prefix for new netCDF variable and coordinate.")
    parser.add_argument("--compute_real", action="store_true",
                        help="Compute FFT assuming real input and thus discarting redundant negative frequencies.")
    parser.add_argument("--upsampling-factor", type=int, default=1, help="Increase resolution of FFT by `upsampling-factor`.")
    parser.add_argument("--filter", type=str, default=None, choices=["lowpass", "highpass", "bandpass"],
                        help="Optional filter to apply prior to FFT computation.")
    parser.add_argument("--filter_freqs", type=int, nargs="+", help="Filter corner frequencies (in Hz).")
    parser.add_argument("--drop-filtered-freq", action="store_true", help="Drop filtered frequency samples.")
    parser.add_argument("--verbose", "-V", type=int, nargs="?", default=0, const=1, choices=[0, 1, 2],
                        help="Level of output verbosity (default: 0)")
    return parser


def apply_fft(cube: Cube, prefix="freq", compute_real=False, upsampling_factor=1, filter_type=None, filter_freqs=None,
              drop_filtered_freq=False, kwargs_nc=None, script="cube_apply_FFT", device=0):
    """Numeric content of the reference's ``main`` on an in-memory cube."""
    today = datetime.date.today().strftime("%Y-%m-%d")
    dim = cube.other_dim()
    var = [v for v in cube.data_vars if v not in ("fold", "amp_ref")][0]
    pre = f"{prefix}_"
    var_new, dim_new = f"{pre}{var}", f"{pre}{dim}"
    dims, data = cube.variables[var]
    if tuple(dims) != (dim, "iline", "xline"):
        data = np.transpose(data, [dims.index(d) for d in (dim, "iline", "xline")])
    twt = np.asarray(cube.coords[dim], dtype=np.float64)
    if data.shape[0] % 2 != 0:
        warnings.warn(f"Selected dim `{dim}` has odd length ({data.shape[0]}), which causes issues for inverse FFT. "
                      "Last slice will be removed!")
    nt_even = data.shape[0] - (data.shape[0] % 2)
    dt = float(twt[1] - twt[0])
    nfft = upsampling_factor * nt_even
    freqs = np.fft.rfftfreq(nfft, dt) if compute_real else np.fft.fftfreq(nfft, dt)

    window, history_filter, attrs_var = None, "", {}
    if filter_type is not None:
        if filter_freqs is None:
            raise ValueError("Filter frequencies must be specified!")
        units = cube.coord_attrs.get(dim, {}).get("units")
        divisor = 1000 if units == "ms" else 1
        ff = [f / divisor for f in filter_freqs]
        window = freq_filter_window(ff, freqs, filter_type)
        _s = "/".join(str(f) for f in filter_freqs)
        attrs_var = {"filter": filter_type, "filter_freq_Hz": _s}
        history_filter = f" {filter_type.upper()} ({_s} Hz)"

    spec, f_axis = time_fft(np.ascontiguousarray(data), twt, compute_real=compute_real,
                            upsampling_factor=upsampling_factor, window=window, device=device)

    out = Cube(attrs=dict(cube.attrs), coord_attrs=dict(cube.coord_attrs), var_attrs=dict(cube.var_attrs))
    out.coords = {k: np.asarray(v) for k, v in cube.coords.items() if k != dim}
    dim_attrs = {"direct_lag": float(twt[nt_even // 2]), "spacing": float(f_axis[1] - f_axis[0]) if len(f_axis) > 1 else 0.0}
    if filter_type is not None and drop_filtered_freq:
        if filter_type == "lowpass":
            dim_attrs["nfft"] = int(f_axis.size)
            keep = freq_filter_keep(f_axis, ff, filter_type)
            spec, f_axis = np.ascontiguousarray(spec[keep]), f_axis[keep]
        else:
            warnings.warn(f"Filter type `{filter_type}` does not support dropping of frequency slices")
    out.coords[dim_new] = f_axis
    out.coord_attrs[dim_new] = dim_attrs
    out.variables[var_new] = ((dim_new, "iline", "xline"), spec)
    fdims, fold = cube.variables["fold"]
    out.variables["fold"] = (fdims, np.asarray(fold))
    reso = f" FACTOR x{upsampling_factor}" if upsampling_factor > 1 else ""
    out.attrs.update({
        "long_name": cube.attrs.get("long_name", "") + " (frequency domain)",
        "description": cube.attrs.get("description", "") + " (frequency domain)",
        "history": cube.attrs.get("history", "") + f"{script}: FFT({var}){reso}{history_filter};",
        "text": cube.attrs.get("text", "") + f"\n{today}: FFT(TIME){reso}{history_filter}",
    })
    va = dict(cube.var_attrs.get(var, {}))
    va["original_var"] = var
    if kwargs_nc is not None:
        va.update(kwargs_nc.get("attrs_freq", {}).get("data", {}))
        va.update(attrs_var)
        out.coord_attrs[dim_new].update(kwargs_nc.get("attrs_freq", {}).get("new_dim", {}))
    out.var_attrs[var_new] = va
    return out


def main(argv=sys.argv, return_dataset=False):  # noqa
    """Apply FFT along _time_ axis wrapper function."""
    script = os.path.splitext(os.path.basename(__file__))[0]
    args = define_input_args().parse_args(argv[1:])
    path_cube = args.path_cube
    dir_work, filename = os.path.split(path_cube)
    basename, suffix = os.path.splitext(filename)
    fout = basename.replace("twt", f"{args.prefix}")
    fout += f"_up-{args.upsampling_factor}" if args.upsampling_factor > 1 else ""
    fout += "-trunc" if args.drop_filtered_freq else ""
    path_cube_freq = os.path.join(dir_work, fout + suffix)
    with open(args.params_netcdf, "r") as f_attrs:
        kwargs_nc = yaml.safe_load(f_attrs)
    cube = open_cube(path_cube)
    out = apply_fft(cube, prefix=args.prefix, compute_real=args.compute_real, upsampling_factor=args.upsampling_factor,
                    filter_type=args.filter, filter_freqs=args.filter_freqs, drop_filtered_freq=args.drop_filtered_freq,
                    kwargs_nc=kwargs_nc, script=script)
    write_cube(path_cube_freq, out, split_complex=False)
    if return_dataset:
        return out


if __name__ == "__main__":
    main()
