"""Mirror of the hot-path half of utils/preprocessing.py: year-wise bootstrap splits, the ISO-week
rolling tercile labeler, predictor-image layout and one-hot targets.

Same function names, arguments and return structure as the reference (preprocessing.py:21-49,
53-167, 335-449, 564-638).  Inputs may be xarray.DataArray or `LabeledArray` (x: (T,M,Y,X),
y: (T,Y,X), T = datetime64 start dates).  The tensor layout handed to the model is the reference's:
X (T,Y,X) float32 ensemble mean (or (T,Y,X,M) channels-last), Y one-hot (T,Y,X,3) float32.
The ELR-only helpers (preprocessing.py:172-333, 452-561) are out of scope."""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd

from s2s_ismr_unet_b200.keras_api.utils import to_categorical
from s2s_ismr_unet_b200.labeled import LabeledArray, as_labeled


# ------------------------------------------------------------------ predictor images (:21-49)
def create_mean_predictor_images(xt):
    return as_labeled(xt).mean("M").values


def create_multi_predictor_images(xt):
    return as_labeled(xt).transpose("T", "Y", "X", "M").values


def create_stacked_predictor_images(xt, yt):
    """Members become extra samples: (M,T) stacked to an 'MT' axis (M outer), y tiled M times."""
    xt, yt = as_labeled(xt), as_labeled(yt)
    v = xt.transpose("M", "T", "Y", "X").values
    M, T = v.shape[:2]
    stacked = LabeledArray(v.reshape((M * T,) + v.shape[2:]), ("MT", "Y", "X"),
                           {"MT": np.arange(M * T), "Y": xt.coords.get("Y", np.arange(v.shape[2])),
                            "X": xt.coords.get("X", np.arange(v.shape[3]))})
    yt_stacked = np.tile(yt.values, (M, 1, 1))
    yt_xr = LabeledArray(yt_stacked, ("MT", "Y", "X"), dict(stacked.coords))
    return stacked, yt_stacked, yt_xr


def convert_to_ndarray(xt, yt, type="mean"):
    if type == "mean":
        return create_mean_predictor_images(xt), as_labeled(yt).values
    if type == "multi_predictor":
        return create_multi_predictor_images(xt), as_labeled(yt).values
    if type == "stacked":
        return create_stacked_predictor_images(xt, yt)
    raise ValueError(f"unknown predictor image type {type!r}")


# ------------------------------------------------------------------ labelers (:11-19, 53-167)
def _iso_week(times) -> np.ndarray:
    return np.asarray(pd.DatetimeIndex(pd.to_datetime(np.asarray(times))).isocalendar().week, dtype=np.int64)


def _quantile_edges(v):
    """1/3 and 2/3 quantiles over the first axis (whole-period labeler, preprocessing.py:11-19): one window holding
    every start, computed by the same CUDA kernel as the rolling labeler."""
    v = np.asarray(v)
    return _gpu_edges(v, np.zeros(len(v), np.int64), np.array([0]), [np.arange(len(v))])


def make_tercile_labeler(observations):
    obs = as_labeled(observations)
    edges = _quantile_edges(obs.values)[0]

    def labeler(y):
        y = as_labeled(y)
        lab = np.where(y.values < edges[0], 0.0, np.where(y.values > edges[1], 2.0, 1.0))
        lab[np.isnan(y.values)] = np.nan
        return y._like(lab)
    return labeler


# ---- CUDA tercile kernels (csrc/prep.cu): no CPU fallback, the library must be loadable and a GPU present
_prep_stream = None


def _pst():
    global _prep_stream
    if _prep_stream is None:
        from s2s_ismr_unet_b200.runtime import Stream
        _prep_stream = Stream()
    return _prep_stream


def _as_dev_dtype(v: np.ndarray) -> np.ndarray:
    """The kernels take float32 or float64 fields as they are (numpy's quantile lerp subtracts in the input dtype)."""
    v = np.asarray(v)
    if v.dtype not in (np.float32, np.float64):
        v = v.astype(np.float64)
    return np.ascontiguousarray(v)


def _gpu_edges(values, week_values, weeks, window_index_lists):
    """edges [nW, 2, *grid] (float64) for the given per-week lists of start indices."""
    import ctypes as C
    from s2s_ismr_unet_b200._lib import call
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    del week_values
    st = _pst()
    v = _as_dev_dtype(values)
    grid = v.shape[1:]
    YX = int(np.prod(grid)) if grid else 1
    starts = np.zeros(len(weeks) + 1, np.int32)
    starts[1:] = np.cumsum([len(ix) for ix in window_index_lists])
    idx = np.concatenate([np.asarray(ix, np.int32) for ix in window_index_lists]).astype(np.int32)
    nmax = int(max(len(ix) for ix in window_index_lists))
    dv, ds, di = DeviceBuffer.from_array(v, st), DeviceBuffer.from_array(starts, st), DeviceBuffer.from_array(idx, st)
    de = DeviceBuffer(8 * len(weeks) * 2 * YX)
    call("s2s_tercile_edges", C.c_void_p(dv.ptr), int(v.dtype == np.float64), C.c_void_p(ds.ptr), C.c_void_p(di.ptr),
         len(weeks), C.c_int64(YX), nmax, C.c_void_p(de.ptr), C.c_void_p(st.ptr))
    edges = de.download((len(weeks), 2) + tuple(grid), np.float64, st)
    for b in (dv, ds, di, de):
        b.free()
    if v.dtype == np.float32 and YX > 0:
        # numpy quirk the reference inherits (xarray .quantile -> np.nanquantile -> np.apply_along_axis): the result
        # buffer takes the dtype of the FIRST gridpoint's result, which is float32 (np.full(.., dtype=arr.dtype)) when
        # that gridpoint is all-NaN in the window and float64 otherwise; the window's edges are then rounded through it.
        first = v.reshape(len(v), -1)[:, 0]
        for w, ix in enumerate(window_index_lists):
            if np.isnan(first[np.asarray(ix, np.int64)]).all():
                edges[w] = edges[w].astype(np.float32).astype(np.float64)
    return edges


def _gpu_label(values, week_slot, edges, want_onehot=False):
    """labels [T, *grid] float64 (0 / 1 / 2 / NaN) and optionally the float32 one-hot [T, *grid, 3], one kernel pass."""
    import ctypes as C
    from s2s_ismr_unet_b200._lib import call
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    st = _pst()
    v = _as_dev_dtype(values)
    T, grid = v.shape[0], v.shape[1:]
    YX = int(np.prod(grid)) if grid else 1
    dv = DeviceBuffer.from_array(v, st)
    dw = DeviceBuffer.from_array(np.asarray(week_slot, np.int32), st)
    de = DeviceBuffer.from_array(np.ascontiguousarray(edges, np.float64), st)
    dl = DeviceBuffer(4 * T * YX)
    do = DeviceBuffer(4 * T * YX * 3) if want_onehot else None
    call("s2s_tercile_label", C.c_void_p(dv.ptr), int(v.dtype == np.float64), C.c_void_p(dw.ptr), C.c_void_p(de.ptr), T,
         C.c_int64(YX), C.c_void_p(dl.ptr), C.c_void_p(do.ptr) if do else None, C.c_void_p(st.ptr))
    lab = dl.download((T,) + tuple(grid), np.float32, st).astype(np.float64)
    oh = do.download((T,) + tuple(grid) + (3,), np.float32, st) if do else None
    for b in (dv, dw, de, dl) + ((do,) if do else ()):
        b.free()
    return lab, oh


def rolling_labeler(observations, window=1):
    """Tercile edges per ISO week from all training starts within +-`window` weeks (weeks wrap at 53,
    preprocessing.py:112-126), computed per gridpoint by the CUDA quantile kernel (csrc/prep.cu); the returned
    labeler assigns 0 / 1 / 2 (NaN where an edge is NaN) using the edges of the nearest training week (:137, ties
    -> the later week as pandas does) and returns the array sorted by T (:165).  `labeler.edges` / `labeler.weeks`
    expose the fitted edges; `labeler(y, onehot=True)` also returns to_categorical(labels, 3) from the same pass."""
    obs = as_labeled(observations)
    week_values = _iso_week(obs["T"])
    weeks = np.unique(week_values)
    windows = []
    for week in weeks:
        window_weeks = [(int(week) + i) % 53 or 53 for i in range(-window, 1 + window)]
        windows.append(np.nonzero(np.isin(week_values, window_weeks))[0])
    edges = _gpu_edges(obs.values, week_values, weeks, windows)

    def labeler(y, onehot=False):
        y = as_labeled(y).sortby("T")
        wk = _iso_week(y["T"])
        dist = np.abs(weeks[None, :] - wk[:, None])
        slot = dist.shape[1] - 1 - np.argmin(dist[:, ::-1], axis=1)     # nearest training week, ties -> larger
        lab, oh = _gpu_label(y.values, slot, edges, want_onehot=onehot)
        out = y._like(lab)
        return (out, oh) if onehot else out
    labeler.edges, labeler.weeks = edges, weeks
    return labeler


def _nearest_slot(weeks, wk):
    dist = np.abs(weeks[None, :] - wk[:, None])
    return dist.shape[1] - 1 - np.argmin(dist[:, ::-1], axis=1)         # nearest training week, ties -> larger


def rolling_labeler_ELR(observations, window=1):
    """Mirror of preprocessing.py:226-333: labeler(y) -> (y_terciled (T,Y,X), edges_t (quantile,T,Y,X),
    y_elr (quantile,T,Y,X)) with y_elr = [y <= edge] and NaN where the week's edges are NaN, e0 == 0 or e0 == e1
    (:304-308).  Edges come from the CUDA quantile kernel; `labeler.edges / .weeks / .slots(y)` feed s2s_elr_fit_predict."""
    base = rolling_labeler(observations, window=window)
    edges, weeks = base.edges, base.weeks

    def labeler(y):
        y = as_labeled(y).sortby("T")
        slot = _nearest_slot(weeks, _iso_week(y["T"]))
        e_t = edges[slot]                                                 # (T, 2, Y, X)
        mask = np.isnan(e_t).any(1) | (e_t[:, 0] == 0) | (e_t[:, 0] == e_t[:, 1])
        v = y.values
        lab = np.where(v < e_t[:, 0], 0.0, np.where(v > e_t[:, 1], 2.0, 1.0))
        lab[mask] = np.nan
        elr = np.stack([(v <= e_t[:, 0]), (v <= e_t[:, 1])]).astype(np.float64)
        elr[:, mask] = np.nan
        qc = {"quantile": np.array([1 / 3, 2 / 3])}
        dims = ("quantile",) + y.dims
        coords = {**{k: c for k, c in y.coords.items() if k in y.dims}, **qc}
        return (y._like(lab), LabeledArray(np.moveaxis(e_t, 1, 0), dims, coords), LabeledArray(elr, dims, coords))
    labeler.edges, labeler.weeks = edges, weeks
    labeler.slots = lambda y: _nearest_slot(weeks, _iso_week(as_labeled(y).sortby("T")["T"]))
    return labeler


# ------------------------------------------------------------------ bootstrap splits (:335-391, 564-638)
def _years(arr) -> np.ndarray:
    return np.asarray(pd.DatetimeIndex(pd.to_datetime(np.asarray(arr["T"]))).year)


def _standardize(a: LabeledArray) -> LabeledArray:
    ax = a.axis("T")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        mu = np.nanmean(a.values, axis=ax, keepdims=True)
        sd = np.nanstd(a.values, axis=ax, keepdims=True)
    return a._like((a.values - mu) / (sd + 1e-6))


def _select_years(a: LabeledArray, years_of_a, wanted) -> LabeledArray:
    idx = np.nonzero(np.isin(years_of_a, wanted))[0]
    return a.isel(T=idx).sortby("T")


def _split_years(unique_years, i, frac_valid, frac_test):
    np.random.seed(i)          # per-bootstrap seed, exactly as preprocessing.py:360
    shuffled = np.random.permutation(unique_years)
    n = len(shuffled)
    n_valid, n_test = int(frac_valid * n), int(frac_test * n)
    return shuffled[n_valid + n_test:], shuffled[:n_valid], shuffled[n_valid:n_valid + n_test]


def bootstrap_splits(x, y, n_bootstraps=10, frac_valid=0.2, frac_test=0.1, standardize=False):
    x, y = as_labeled(x), as_labeled(y)
    if standardize:
        x, y = _standardize(x), _standardize(y)
    x, y = x.fillna(0), y.fillna(0)          # ocean / missing -> 0 (preprocessing.py:342-343)
    x["T"] = pd.to_datetime(x["T"]).values
    y["T"] = pd.to_datetime(y["T"]).values
    yx, yy = _years(x), _years(y)
    unique_years = np.unique(yx)
    out = ([], [], [], [], [], [])
    for i in range(n_bootstraps):
        train, valid, test = _split_years(unique_years, i, frac_valid, frac_test)
        for lst, arr, yrs, sel in ((out[0], x, yx, train), (out[1], y, yy, train), (out[2], x, yx, valid),
                                   (out[3], y, yy, valid), (out[4], x, yx, test), (out[5], y, yy, test)):
            lst.append(_select_years(arr, yrs, sel))
    return out


def bootstrap_splits_ELR(x, y, n_bootstraps=10, frac_test=0.3, standardize=False):
    """Year-wise train / test splits of the ELR baseline (preprocessing.py:452-492): no validation set, no fillna."""
    x, y = as_labeled(x), as_labeled(y)
    if standardize:
        x, y = _standardize(x), _standardize(y)
    x["T"] = pd.to_datetime(x["T"]).values
    y["T"] = pd.to_datetime(y["T"]).values
    yx, yy = _years(x), _years(y)
    unique_years = np.unique(yx)
    out = ([], [], [], [])
    for i in range(n_bootstraps):
        np.random.seed(i)
        shuffled = np.random.permutation(unique_years)
        n_test = int(len(shuffled) * frac_test)
        train_years, test_years = shuffled[:-n_test], shuffled[-n_test:]
        for lst, arr, yrs, sel in ((out[0], x, yx, train_years), (out[1], y, yy, train_years),
                                   (out[2], x, yx, test_years), (out[3], y, yy, test_years)):
            lst.append(_select_years(arr, yrs, sel))
    return out


def bootstrap_splits_ELR_mme(x_dict, y, n_bootstraps=10, frac_test=0.3, standardize=False):
    """Multi-model variant (preprocessing.py:494-561): the year split is drawn once per bootstrap from y's years and
    shared by every model -> (xtrain_dict, xtest_dict, ytrain_list, ytest_list)."""
    x_dict = {k: as_labeled(v) for k, v in x_dict.items()}
    y = as_labeled(y)
    if standardize:
        x_dict = {k: _standardize(v) for k, v in x_dict.items()}
        y = _standardize(y)
    y["T"] = pd.to_datetime(y["T"]).values
    for v in x_dict.values():
        v["T"] = pd.to_datetime(v["T"]).values
    yy = _years(y)
    unique_years = np.unique(yy)
    xtrain = {m: [] for m in x_dict}
    xtest = {m: [] for m in x_dict}
    ytrain, ytest = [], []
    for i in range(n_bootstraps):
        np.random.seed(i)
        shuffled = np.random.permutation(unique_years)
        n_test = int(len(shuffled) * frac_test)
        train_years, test_years = shuffled[:-n_test], shuffled[-n_test:]
        for m, xv in x_dict.items():
            ym = _years(xv)
            xtrain[m].append(_select_years(xv, ym, train_years))
            xtest[m].append(_select_years(xv, ym, test_years))
        ytrain.append(_select_years(y, yy, train_years))
        ytest.append(_select_years(y, yy, test_years))
    return xtrain, xtest, ytrain, ytest


def bootstrap_splits_mme(x_dict, y, n_bootstraps=10, frac_valid=0.2, frac_test=0.1, standardize=False):
    x_dict = {k: as_labeled(v) for k, v in x_dict.items()}
    y = as_labeled(y)
    if standardize:
        x_dict = {k: _standardize(v) for k, v in x_dict.items()}
        y = _standardize(y)
    x_dict = {k: v.fillna(0) for k, v in x_dict.items()}
    y = y.fillna(0)
    y["T"] = pd.to_datetime(y["T"]).values
    for v in x_dict.values():
        v["T"] = pd.to_datetime(v["T"]).values
    yy = _years(y)
    unique_years = np.unique(yy)
    xtrain = {m: [] for m in x_dict}
    xval = {m: [] for m in x_dict}
    xtest = {m: [] for m in x_dict}
    ytrain, yval, ytest = [], [], []
    for i in range(n_bootstraps):
        train, valid, test = _split_years(unique_years, i, frac_valid, frac_test)
        for m, xv in x_dict.items():
            ym = _years(xv)
            xtrain[m].append(_select_years(xv, ym, train))
            xval[m].append(_select_years(xv, ym, valid))
            xtest[m].append(_select_years(xv, ym, test))
        ytrain.append(_select_years(y, yy, train))
        yval.append(_select_years(y, yy, valid))
        ytest.append(_select_years(y, yy, test))
    return xtrain, xval, xtest, ytrain, yval, ytest


# ------------------------------------------------------------------ preprocess (:393-449)
def preprocess(xtrain, ytrain, xval, yval, xtest, ytest, predictor_type="mean"):
    """-> X_train, Y_train_oh, X_val, Y_val_oh, X_test, Y_test_oh, y_train_terciled, y_val_terciled,
    y_test_terciled.  predictor_type="multi_predictor" gives (T,Y,X,M) channel-stacked images
    (create_multi_predictor_images, unused by the reference's callers, used by the MME C>1 config)."""
    labeler_train = rolling_labeler(ytrain, window=1)
    num_classes = 3
    del num_classes            # labels and to_categorical(labels, 3) (:426-428) come from one CUDA pass
    y_train_terciled, Y_train_oh = labeler_train(ytrain, onehot=True)
    y_val_terciled, Y_val_oh = labeler_train(yval, onehot=True)
    y_test_terciled, Y_test_oh = labeler_train(ytest, onehot=True)
    X_train, _ = convert_to_ndarray(xtrain, y_train_terciled, predictor_type)
    X_val, _ = convert_to_ndarray(xval, y_val_terciled, predictor_type)
    X_test, _ = convert_to_ndarray(xtest, y_test_terciled, predictor_type)
    return (X_train.astype(np.float32), Y_train_oh, X_val.astype(np.float32), Y_val_oh, X_test.astype(np.float32), Y_test_oh,
            y_train_terciled, y_val_terciled, y_test_terciled)


def preprocess_stacked(xtrain, ytrain, xval, yval, xtest, ytest):
    labeler_train = rolling_labeler(ytrain, window=1)
    num_classes = 3
    y_train_terciled = labeler_train(ytrain)
    y_val_terciled = labeler_train(yval)
    y_test_terciled = labeler_train(ytest)
    X_train, Y_train_terciled, y_train_terciled = convert_to_ndarray(xtrain, y_train_terciled, "stacked")
    X_val, Y_val_terciled, y_val_terciled = convert_to_ndarray(xval, y_val_terciled, "stacked")
    X_test, Y_test_terciled, y_test_terciled = convert_to_ndarray(xtest, y_test_terciled, "stacked")
    return (X_train, to_categorical(Y_train_terciled, num_classes), X_val, to_categorical(Y_val_terciled, num_classes),
            X_test, to_categorical(Y_test_terciled, num_classes), y_train_terciled, y_val_terciled, y_test_terciled)
