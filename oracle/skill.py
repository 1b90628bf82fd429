"""ORACLE (test infrastructure only) — NumPy restatement of the reference's skill reductions and
tensor-layout helpers.  PARITY UNPINNED: xarray / xskillscore are not installed here and the
reference holds no golden values; pinned only by the analytic known-answer tests in tests/.

  rps / rpss / climo_predict   utils/performance_metrics.py:11-45  (xskillscore.rps, input_distributions='p')
  cc / acc                     ACCs.ipynb:362-388 (xr.corr over T; ISO-week anomalies)
  ensemble_mean                utils/preprocessing.py:21-23  (xt.mean('M'), NaN-skipping)
  mme_combine                  utils/training.py:344-350
  rolling tercile labels       utils/preprocessing.py:53-167
  to_categorical               tf.keras.utils.to_categorical (preprocessing.py:426-428)
"""
from __future__ import annotations

import numpy as np


def climo_forecast(shape_tyx, valid=None):
    f = np.full(tuple(shape_tyx) + (3,), 1.0 / 3.0, np.float64)
    if valid is not None:
        f[~valid] = np.nan
    return f


def onehot_obs(labels):
    """labels (T,Y,X) in {0,1,2,NaN} -> (T,Y,X,3) with NaN rows where the label is NaN."""
    labels = np.asarray(labels, np.float64)
    oh = np.stack([(labels == k).astype(np.float64) for k in range(3)], -1)
    oh[np.isnan(labels)] = np.nan
    return oh


def rps(obs_onehot, fcst):
    """mean_T sum_cat (cumsum p - cumsum o)^2 per gridpoint; starts with a NaN observation are
    skipped; all-NaN -> NaN.  (T,Y,X,3) x2 -> (Y,X)."""
    o = np.asarray(obs_onehot, np.float64)
    p = np.asarray(fcst, np.float64)
    d = np.cumsum(p, -1) - np.cumsum(o, -1)
    s = (d * d).sum(-1)                                  # (T,Y,X)
    valid = ~np.isnan(o).any(-1)
    cnt = valid.sum(0)
    tot = np.where(valid, s, 0.0).sum(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(cnt > 0, tot / cnt, np.nan)


def rpss(reference, forecast, obs_onehot):
    return 1.0 - rps(obs_onehot, forecast) / rps(obs_onehot, reference)


def pearson_t(x, y):
    """xr.corr(x, y, dim='T'): pairwise-NaN-skipping Pearson r with ddof=0.  (T,Y,X) x2 -> (Y,X)."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    valid = ~(np.isnan(x) | np.isnan(y))
    n = valid.sum(0).astype(np.float64)
    xv = np.where(valid, x, 0.0)
    yv = np.where(valid, y, 0.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        mx = xv.sum(0) / n
        my = yv.sum(0) / n
        dx = np.where(valid, x - mx, 0.0)
        dy = np.where(valid, y - my, 0.0)
        cov = (dx * dy).sum(0) / n
        sx = np.sqrt((dx * dx).sum(0) / n)
        sy = np.sqrt((dy * dy).sum(0) / n)
        return cov / (sx * sy)


def week_anomalies(x, week_id):
    """value - mean over all starts sharing the same ISO week (NaN-skipping mean), ACCs.ipynb:369-376."""
    x = np.asarray(x, np.float64)
    out = np.empty_like(x)
    week_id = np.asarray(week_id)
    for w in np.unique(week_id):
        sel = week_id == w
        with np.errstate(invalid="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                m = np.nanmean(x[sel], axis=0)
        out[sel] = x[sel] - m
    return out


def acc_cc(x, y, week_id):
    """(ACC, CC) maps: ACC = Pearson of ISO-week anomalies, CC = Pearson of the raw fields."""
    return pearson_t(week_anomalies(x, week_id), week_anomalies(y, week_id)), pearson_t(x, y)


def ensemble_mean(x_tmyx):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmean(np.asarray(x_tmyx, np.float64), axis=1)


def mme_combine(probs_list):
    p = np.mean(np.stack([np.asarray(q, np.float64) for q in probs_list], 0), 0)
    return p / p.sum(-1, keepdims=True)


def to_categorical(y, num_classes=3):
    y = np.asarray(y)
    out = np.zeros(y.shape + (num_classes,), np.float32)
    idx = np.nan_to_num(y, nan=0.0).astype(np.int64)
    np.put_along_axis(out, idx[..., None], 1.0, axis=-1)
    return out


def iso_week(times) -> np.ndarray:
    import pandas as pd
    return np.asarray(pd.DatetimeIndex(pd.to_datetime(np.asarray(times))).isocalendar().week, dtype=np.int64)


def rolling_tercile_edges(y_train, week_train, window=1):
    """{week: (2,Y,X) edges}: 1/3 and 2/3 quantiles (linear interpolation, NaN if any NaN — xarray's
    quantile default skipna for floats is True, the inputs are NaN-free after fillna(0)) over all
    training starts whose ISO week is within +-window of the week (weeks wrap at 53)."""
    y_train = np.asarray(y_train)          # dtype kept: numpy's quantile lerp subtracts in the input dtype
    edges = {}
    for w in np.unique(week_train):
        wins = [((int(w) + i) % 53) or 53 for i in range(-window, window + 1)]
        sel = np.isin(week_train, wins)
        edges[int(w)] = np.nanquantile(y_train[sel], [1 / 3, 2 / 3], axis=0)
    return edges


def apply_tercile_labels(y, week_y, edges):
    """labels (T,Y,X): 0 below the first edge, 2 above the second, else 1; NaN where an edge is NaN.
    The edges of the NEAREST training week are used (edges.sel(week=, method='nearest'))."""
    y = np.asarray(y)
    weeks = np.array(sorted(edges))
    out = np.empty(y.shape, np.float64)
    for t in range(y.shape[0]):
        d = np.abs(weeks - week_y[t])
        w = weeks[np.nonzero(d == d.min())[0][-1]]      # pandas method='nearest': ties prefer the larger value
        e = edges[int(w)]
        lab = np.where(y[t] < e[0], 0.0, np.where(y[t] > e[1], 2.0, 1.0))
        lab[np.isnan(e).any(0)] = np.nan
        out[t] = lab
    return out
