"""Mirror of utils/performance_metrics.py (climo_predict, rps, rpss) plus the ACC / CC and Grad-CAM
entry points BASELINE.json names, all as per-gridpoint CUDA reductions over the start-date axis.

Reference: performance_metrics.py:11-45 (xskillscore.rps with input_distributions='p');
ACC / CC: ACCs.ipynb:362-388 (xr.corr of ISO-week anomalies / raw fields)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pandas as pd

from s2s_ismr_unet_b200._lib import call
from s2s_ismr_unet_b200.labeled import LabeledArray, as_labeled
from s2s_ismr_unet_b200.runtime import DeviceBuffer, Stream

CATEGORIES = np.array(["below", "normal", "above"])
_stream = None


def _st() -> Stream:
    global _stream
    if _stream is None:
        _stream = Stream()
    return _stream


def _tdim(a: LabeledArray) -> str:
    return "MT" if "MT" in a.dims else "T"


def climo_predict(x, predictor="mean"):
    """Always predicts climatology: 1/3 per tercile wherever the ensemble-mean predictor is not null."""
    x = as_labeled(x)
    if "M" in x.dims:
        x = x.mean("M")
    t = _tdim(x)
    v = np.full(x.shape + (3,), 1.0 / 3.0, np.float32)
    v[np.isnan(x.values)] = np.nan
    return LabeledArray(v, (t, "Y", "X", "category"), {**{k: c for k, c in x.coords.items() if k in (t, "Y", "X")},
                                                       "category": CATEGORIES})


def _obs_onehot(obs: LabeledArray) -> np.ndarray:
    lab = obs.values
    oh = np.stack([(lab == k) for k in range(3)], -1).astype(np.float32)
    oh[np.isnan(lab)] = np.nan
    return oh


def _as_tyxc(a) -> np.ndarray:
    a = as_labeled(a)
    t = _tdim(a)
    return np.ascontiguousarray(a.transpose(t, "Y", "X", "category").values, np.float32)


def _out_map(like: LabeledArray, values) -> LabeledArray:
    return LabeledArray(values, ("Y", "X"), {k: v for k, v in like.coords.items() if k in ("Y", "X")})


def rps(obs, fcst, predictor="mean"):
    """obs: tercile labels (T,Y,X) in {0,1,2,NaN}; fcst: probabilities (T,Y,X,category) -> RPS (Y,X)."""
    obs = as_labeled(obs)
    o, p = _obs_onehot(obs), _as_tyxc(fcst)
    T, Y, X = obs.shape
    st = _st()
    d_o, d_p, d_out = DeviceBuffer.from_array(o, st), DeviceBuffer.from_array(p, st), DeviceBuffer(4 * Y * X)
    call("s2s_rps_map", C.c_void_p(d_p.ptr), C.c_void_p(d_o.ptr), T, Y, X, C.c_void_p(d_out.ptr), C.c_void_p(st.ptr))
    return _out_map(obs, d_out.download((Y, X), np.float32, st))


def rpss(reference, forecast, observations, predictor="mean"):
    """1 - RPS(forecast) / RPS(reference) per gridpoint, one fused pass (performance_metrics.py:44-45)."""
    obs = as_labeled(observations)
    o, f, r = _obs_onehot(obs), _as_tyxc(forecast), _as_tyxc(reference)
    T, Y, X = obs.shape
    st = _st()
    d_o, d_f, d_r = DeviceBuffer.from_array(o, st), DeviceBuffer.from_array(f, st), DeviceBuffer.from_array(r, st)
    d_out = DeviceBuffer(4 * Y * X)
    call("s2s_rpss_map", C.c_void_p(d_f.ptr), C.c_void_p(d_r.ptr), C.c_void_p(d_o.ptr), T, Y, X, C.c_void_p(d_out.ptr),
         C.c_void_p(st.ptr))
    return _out_map(obs, d_out.download((Y, X), np.float32, st))


def iso_week_groups(times):
    """(order, group_start, n_groups): start indices sorted by ISO week and the offsets of each week."""
    week = np.asarray(pd.DatetimeIndex(pd.to_datetime(np.asarray(times))).isocalendar().week, dtype=np.int64)
    weeks, gid = np.unique(week, return_inverse=True)
    order = np.argsort(gid, kind="stable").astype(np.int32)
    gstart = np.concatenate([[0], np.cumsum(np.bincount(gid, minlength=len(weeks)))]).astype(np.int32)
    return order, gstart, len(weeks)


def acc(x, y, week_index=None, return_cc=False):
    """Anomaly correlation per gridpoint (ACCs.ipynb:362-388): x (T,[M,]Y,X) forecast (ensemble-mean
    taken if M is present), y (T,Y,X) observations; anomalies are relative to the mean over all starts
    sharing the ISO week of `week_index` (default: the ISO week of the T coordinate); Pearson r over T
    with pairwise NaN skipping.  Returns the ACC map (Y,X) (and the raw CC map if return_cc)."""
    x, y = as_labeled(x), as_labeled(y)
    if "M" in x.dims:
        x = x.mean("M")
    T, Y, X = y.shape
    if week_index is None:
        order, gstart, ng = iso_week_groups(y["T"])
    else:
        weeks, gid = np.unique(np.asarray(week_index), return_inverse=True)
        order = np.argsort(gid, kind="stable").astype(np.int32)
        gstart = np.concatenate([[0], np.cumsum(np.bincount(gid, minlength=len(weeks)))]).astype(np.int32)
        ng = len(weeks)
    st = _st()
    d_x = DeviceBuffer.from_array(np.ascontiguousarray(x.transpose("T", "Y", "X").values, np.float32), st)
    d_y = DeviceBuffer.from_array(np.ascontiguousarray(y.transpose("T", "Y", "X").values, np.float32), st)
    d_o, d_g = DeviceBuffer.from_array(order, st), DeviceBuffer.from_array(gstart, st)
    d_acc, d_cc = DeviceBuffer(4 * Y * X), DeviceBuffer(4 * Y * X)
    call("s2s_acc_map", C.c_void_p(d_x.ptr), C.c_void_p(d_y.ptr), C.c_void_p(d_o.ptr), C.c_void_p(d_g.ptr), ng, T, Y, X,
         C.c_void_p(d_acc.ptr), C.c_void_p(d_cc.ptr), C.c_void_p(st.ptr))
    a = _out_map(y, d_acc.download((Y, X), np.float32, st))
    if return_cc:
        return a, _out_map(y, d_cc.download((Y, X), np.float32, st))
    return a


def gradcam(model, x, layer_name="bottleneck", category="above"):
    """Grad-CAM maps (N,Hl,Wl) of `model` for a tercile category at a named Keras layer
    (deep_nn_models.py:89,142,145,154,157,160).  The reference's notebook is missing; the definition
    is documented in DESIGN.md (score = spatial mean of the class probability)."""
    cls = int(np.nonzero(CATEGORIES == category)[0][0]) if isinstance(category, str) else int(category)
    return model.gradcam(x, layer_name, cls)
