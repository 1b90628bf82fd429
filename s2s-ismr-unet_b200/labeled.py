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


def maybe_xarray(arr: LabeledArray):
    """Return an xarray.DataArray when xarray is importable (drop-in for the reference's callers)."""
    try:
        import xarray  # noqa: F401
    except Exception:
        return arr
    return arr.to_xarray()
