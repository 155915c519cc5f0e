"""Minimal cube container + file I/O for the three CLI mirrors.

The reference reads and writes netCDF-4 through xarray/h5netcdf
(cube_POCS_interpolation_3D.py:231-233, 370-405; cube_apply_FFT.py:207, 316-319;
cube_apply_IFFT.py:53-57, 149-152).  Neither package is present in this image, so:

* ``*.nc`` paths go through xarray + h5netcdf when they can be imported (same engine, same
  ``invalid_netcdf=True`` for complex variables, complex results split into ``.real`` /
  ``.imag`` float32 like cube_POCS_interpolation_3D.py:160-164);
* ``*.npz`` paths use a plain numpy archive with the same variable / coordinate / attribute
  names, which is what the tests and examples in this repo use.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Cube:
    """variables: name -> (dims tuple, ndarray); coords: dim -> 1-D ndarray."""
    variables: dict = field(default_factory=dict)
    coords: dict = field(default_factory=dict)
    attrs: dict = field(default_factory=dict)
    var_attrs: dict = field(default_factory=dict)
    coord_attrs: dict = field(default_factory=dict)

    def dims_of(self, name):
        return self.variables[name][0]

    def data(self, name):
        return self.variables[name][1]

    @property
    def data_vars(self):
        return list(self.variables)

    def other_dim(self):
        """The dimension that is neither iline nor xline (cube_apply_FFT.py:210)."""
        dims = []
        for d, _ in self.variables.values():
            for n in d:
                if n not in ("iline", "xline") and n not in dims:
                    dims.append(n)
        return dims[0]


def _have_xarray():
    try:
        import xarray  # noqa: F401
        import h5netcdf  # noqa: F401
        return True
    except Exception:
        return False


def open_cube(path: str) -> Cube:
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        meta = json.loads(str(z["__meta__"])) if "__meta__" in z else {}
        cube = Cube(attrs=meta.get("attrs", {}), var_attrs=meta.get("var_attrs", {}), coord_attrs=meta.get("coord_attrs", {}))
        dims = meta.get("dims", {})
        for k in z.files:
            if k == "__meta__":
                continue
            if k.startswith("coord__"):
                cube.coords[k[len("coord__"):]] = z[k]
            else:
                cube.variables[k] = (tuple(dims.get(k, ())), z[k])
        return cube
    if not _have_xarray():
        raise ImportError("reading netCDF needs xarray + h5netcdf (not installed); use a .npz cube instead")
    import xarray as xr
    ds = xr.open_dataset(path, engine="h5netcdf")
    cube = Cube(attrs=dict(ds.attrs))
    for name in ds.data_vars:
        cube.variables[name] = (tuple(ds[name].dims), np.asarray(ds[name].values))
        cube.var_attrs[name] = dict(ds[name].attrs)
    for name in ds.coords:
        cube.coords[name] = np.asarray(ds[name].values)
        cube.coord_attrs[name] = dict(ds[name].attrs)
    ds.close()
    return cube


def write_cube(path: str, cube: Cube, split_complex: bool = False):
    """``split_complex``: store complex variables as '<var>.real' / '<var>.imag' float32."""
    variables = {}
    for name, (dims, arr) in cube.variables.items():
        if split_complex and np.iscomplexobj(arr):
            variables[f"{name}.real"] = (dims, np.ascontiguousarray(arr.real, dtype=np.float32))
            variables[f"{name}.imag"] = (dims, np.ascontiguousarray(arr.imag, dtype=np.float32))
        else:
            variables[name] = (dims, arr)
    if path.endswith(".npz"):
        meta = dict(attrs=_jsonable(cube.attrs), var_attrs=_jsonable(cube.var_attrs), coord_attrs=_jsonable(cube.coord_attrs),
                    dims={k: list(d) for k, (d, _) in variables.items()})
        out = {k: a for k, (_, a) in variables.items()}
        out.update({f"coord__{k}": v for k, v in cube.coords.items()})
        out["__meta__"] = np.array(json.dumps(meta))
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        np.savez(path, **out)
        return
    if not _have_xarray():
        raise ImportError("writing netCDF needs xarray + h5netcdf (not installed); use a .npz path instead")
    import xarray as xr
    ds = xr.Dataset({k: (d, a, cube.var_attrs.get(k.split(".")[0], {})) for k, (d, a) in variables.items()},
                    coords={k: (k, v, cube.coord_attrs.get(k, {})) for k, v in cube.coords.items()}, attrs=cube.attrs)
    has_complex = any(np.iscomplexobj(a) for _, a in variables.values())
    ds.to_netcdf(path, engine="h5netcdf", invalid_netcdf=has_complex)


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, np.generic):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    return o
