"""Minimal labelled array used where the reference passes xarray.DataArray objects
(x (T,M,Y,X), y (T,Y,X): utils/dataloader.py:295-298).  xarray is optional: real DataArrays are
accepted anywhere (`as_labeled`), and results convert back with `.to_xarray()` when it is installed.
Only the operations the hot-path host code needs are provided."""
from __future__ import annotations

import numpy as np


class LabeledArray:
    def __init__(self, values, dims, coords=None, name=None):
        self.values = np.asarray(values)
        self.dims = tuple(dims)
        if self.values.ndim != len(self.dims):
            raise ValueError(f"{self.values.ndim}-d values but dims {self.dims}")
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        self.name = name

    # -- basic protocol
    @property
    def shape(self):
        return self.values.shape

    def __len__(self):
        return len(self.values)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[key]
        raise TypeError("use .isel / .values for positional indexing")

    def __setitem__(self, key, value):
        self.coords[key] = np.asarray(value)

    def axis(self, dim):
        return self.dims.index(dim)

    def copy(self):
        return LabeledArray(self.values.copy(), self.dims, {k: v.copy() for k, v in self.coords.items()}, self.name)

    def _like(self, values, dims=None, drop=()):
        dims = self.dims if dims is None else tuple(dims)
        coords = {k: v for k, v in self.coords.items() if k in dims and k not in drop}
        return LabeledArray(values, dims, coords, self.name)

    # -- the few xarray-style operations used by the host code
    def isel(self, **indexers):
        out = self
        for dim, idx in indexers.items():
            ax = out.axis(dim)
            idx = np.asarray(idx)
            vals = np.take(out.values, idx, axis=ax)
            coords = dict(out.coords)
            if dim in coords:
                coords[dim] = coords[dim][idx]
            for k, v in list(coords.items()):          # auxiliary coords along the same dim (e.g. week)
                if k != dim and getattr(v, "shape", ()) == (out.shape[ax],):
                    coords[k] = v[idx]
            out = LabeledArray(vals, out.dims, coords, out.name)
        return out

    def transpose(self, *dims):
        perm = [self.axis(d) for d in dims]
        return LabeledArray(np.transpose(self.values, perm), dims, self.coords, self.name)

    def mean(self, dim):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            v = np.nanmean(self.values, axis=self.axis(dim))
        return self._like(v, [d for d in self.dims if d != dim], drop=(dim,))

    def fillna(self, value):
        return self._like(np.where(np.isnan(self.values), value, self.values))

    def sortby(self, dim):
        order = np.argsort(self.coords[dim], kind="stable")
        return self.isel(**{dim: order})

    def notnull(self):
        return self._like(~np.isnan(self.values))

    def to_xarray(self):
        import xarray as xr
        return xr.DataArray(self.values, dims=self.dims, coords={k: ((k,), v) if v.ndim == 1 and k in self.dims else v
                                                                 for k, v in self.coords.items() if k in self.dims}, name=self.name)

    def to_netcdf(self, path, name=None):
        """Write the array as a NetCDF-3 (64-bit offset) file — the on-disk format of the reference's
        `xr.concat(rpss_list, dim='bootstrap').to_netcdf('outputs/.../unet_rpss_test_<week>.nc')`
        (tune_ECMWF_com.py:114-121), readable by `xr.open_dataarray` / Bar_plot.ipynb.  netCDF4 / h5py are not
        needed: scipy's pure-Python writer is used.  Datetime coordinates are stored CF-style as float64
        "days since 1970-01-01"; string coordinates (category) as a char matrix."""
        from scipy.io import netcdf_file
        name = name or self.name or "__xarray_dataarray_variable__"
        with netcdf_file(str(path), "w", version=2) as f:
            for d, n in zip(self.dims, self.shape):
                f.createDimension(d, n)
            for d in self.dims:
                if d not in self.coords:
                    continue
                c = np.asarray(self.coords[d])
                if np.issubdtype(c.dtype, np.datetime64):
                    v = f.createVariable(d, "d", (d,))
                    v[:] = (c.astype("datetime64[ns]").astype(np.int64) / 86400e9)
                    v.units = "days since 1970-01-01 00:00:00"
                    v.calendar = "proleptic_gregorian"
                elif c.dtype.kind in "US":
                    w = max(len(str(x)) for x in c)
                    f.createDimension(f"{d}_strlen", w)
                    v = f.createVariable(d, "c", (d, f"{d}_strlen"))
                    v[:] = np.array([list(str(x).ljust(w)) for x in c], dtype="S1")
                else:
                    c = c.astype(np.float64) if c.dtype.kind == "f" else c.astype(np.int32)
                    v = f.createVariable(d, c.dtype.char, (d,))
                    v[:] = c
            vals = self.values
            vals = vals.astype(np.float64) if vals.dtype == np.float64 else vals.astype(np.float32)
            v = f.createVariable(name, vals.dtype.char, self.dims)
            v[:] = vals
            v._FillValue = np.array(np.nan, vals.dtype)

    def __repr__(self):
        return f"LabeledArray(dims={self.dims}, shape={self.shape})"


def as_labeled(obj) -> LabeledArray:
    """LabeledArray from a LabeledArray or an xarray.DataArray (duck-typed: .dims/.values/.coords)."""
    if isinstance(obj, LabeledArray):
        return obj
    if hasattr(obj, "dims") and hasattr(obj, "values"):
        coords = {}
        for d in obj.dims:
            try:
                coords[d] = np.asarray(obj[d].values)
            except Exception:
                pass
        return LabeledArray(np.asarray(obj.values), tuple(obj.dims), coords, getattr(obj, "name", None))
    raise TypeError(f"expected a labelled array, got {type(obj).__name__}")


def concat(arrays, dim: str) -> LabeledArray:
    """xr.concat(list, dim=...): along an existing dimension (e.g. 'T') or a new leading one (e.g. 'bootstrap',
    tune_ECMWF_com.py:114-116)."""
    arrays = [as_labeled(a) for a in arrays]
    first = arrays[0]
    if dim in first.dims:
        ax = first.axis(dim)
        coords = dict(first.coords)
        if dim in coords:
            coords[dim] = np.concatenate([a.coords[dim] for a in arrays])
        return LabeledArray(np.concatenate([a.values for a in arrays], axis=ax), first.dims, coords, first.name)
    coords = {**first.coords, dim: np.arange(len(arrays))}
    return LabeledArray(np.stack([a.values for a in arrays]), (dim,) + first.dims, coords, first.name)


def open_netcdf(path, name=None) -> LabeledArray:
    """Read back a file written by LabeledArray.to_netcdf (or any NetCDF-3 file with one data variable)."""
    from scipy.io import netcdf_file
    with netcdf_file(str(path), "r", mmap=False) as f:
        dims = set(f.dimensions)
        names = [k for k in f.variables if k not in dims]
        name = name or names[0]
        var = f.variables[name]
        coords = {}
        for d in var.dimensions:
            if d not in f.variables:
                continue
            c = f.variables[d]
            data = np.array(c[:])
            units = getattr(c, "units", b"")
            units = units.decode() if isinstance(units, bytes) else units
            if units.startswith("days since 1970-01-01"):
                data = (np.round(data * 86400e9).astype(np.int64)).astype("datetime64[ns]")
            elif data.dtype.kind == "S" and data.ndim == 2:
                data = np.array([b"".join(r).decode().rstrip() for r in data])
            coords[d] = data
        return LabeledArray(np.array(var[:]), tuple(var.dimensions), coords, name)


def maybe_xarray(arr: LabeledArray):
    """Return an xarray.DataArray when xarray is importable (drop-in for the reference's callers)."""
    try:
        import xarray  # noqa: F401
    except Exception:
        return arr
    return arr.to_xarray()
