"""GPU tercile labeler (csrc/prep.cu) against the NumPy oracle (oracle/skill.py), bit-exact:
edges are float64 order statistics with numpy's own lerp arithmetic, labels are small integers."""
import numpy as np
import pandas as pd
import pytest

from oracle import skill as so

pytestmark = pytest.mark.gpu


def make_xy(years=range(2003, 2019), M=4, Y=8, X=8, seed=0, dtype=np.float32):
    from s2s_ismr_unet_b200.labeled import LabeledArray
    rng = np.random.default_rng(seed)
    T = np.concatenate([pd.date_range(f"{y}-05-01", f"{y}-09-30", freq="7D").values for y in years])
    x = rng.gamma(2.0, 3.0, size=(len(T), M, Y, X)).astype(np.float32)
    y = rng.gamma(2.0, 3.0, size=(len(T), Y, X)).astype(dtype)
    y[:, 0, 0] = np.nan                                            # an ocean point
    co = {"T": T, "Y": np.arange(Y), "X": np.arange(X)}
    return LabeledArray(x, ("T", "M", "Y", "X"), {**co, "M": np.arange(M)}), LabeledArray(y, ("T", "Y", "X"), co)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("window", [0, 1, 3])
def test_rolling_edges_and_labels_bit_exact(dtype, window):
    from s2s_ismr_unet_b200.utils import preprocessing as pp
    _, y = make_xy(Y=16, X=24, dtype=dtype, seed=3)
    v = y.values.copy()
    v[5:40, 3, 3] = np.nan                                         # partially missing point: nanquantile skips
    v[:, 2, 2] = 0.0                                               # constant point: both edges 0, everything is class 1
    v[::3, 4, 4] = v[1, 4, 4]                                      # ties at an edge
    y = y._like(v)
    lab = pp.rolling_labeler(y, window=window)
    wk = so.iso_week(y["T"])
    edges = so.rolling_tercile_edges(y.values, wk, window=window)
    for i, w in enumerate(lab.weeks):
        np.testing.assert_array_equal(lab.edges[i], edges[int(w)])
    assert np.isnan(lab.edges[:, :, 0, 0]).all()
    # label a shifted period (other ISO weeks, incl. ones absent from training -> nearest week)
    rng = np.random.default_rng(1)
    T2 = pd.date_range("2021-03-01", "2021-11-30", freq="5D").values
    y2 = y._like(rng.gamma(2.0, 3.0, size=(len(T2),) + y.shape[1:]).astype(dtype))
    y2.coords["T"] = T2
    got, oh = lab(y2, onehot=True)
    want = so.apply_tercile_labels(y2.values, so.iso_week(T2), edges)
    np.testing.assert_array_equal(got.values, want)
    ok = ~np.isnan(want)
    np.testing.assert_array_equal(oh[ok], np.eye(3, dtype=np.float32)[want[ok].astype(int)])
    assert np.isnan(oh[~ok]).all()
    # the training period labelled with its own edges
    np.testing.assert_array_equal(lab(y).values, so.apply_tercile_labels(y.values, wk, edges))


def test_whole_period_labeler_matches_numpy():
    from s2s_ismr_unet_b200.utils import preprocessing as pp
    _, y = make_xy(seed=5)
    lab = pp.make_tercile_labeler(y)(y).values
    e = np.nanquantile(y.values, [1 / 3, 2 / 3], axis=0)
    want = np.where(y.values < e[0], 0.0, np.where(y.values > e[1], 2.0, 1.0))
    want[np.isnan(y.values)] = np.nan
    np.testing.assert_array_equal(lab, want)


def test_large_grid_properties():
    """0.25-degree grid (256x256), 30 years: terciles of the training period are balanced and edges are ordered."""
    from s2s_ismr_unet_b200.utils import preprocessing as pp
    from s2s_ismr_unet_b200.labeled import LabeledArray
    rng = np.random.default_rng(7)
    T = np.concatenate([pd.date_range(f"{y}-05-01", f"{y}-09-30", freq="7D").values for y in range(1991, 2021)])
    v = rng.gamma(2.0, 3.0, size=(len(T), 256, 256)).astype(np.float32)
    y = LabeledArray(v, ("T", "Y", "X"), {"T": T, "Y": np.arange(256), "X": np.arange(256)})
    lab = pp.rolling_labeler(y, window=1)
    assert (lab.edges[:, 0] <= lab.edges[:, 1]).all()
    out = lab(y).values
    frac = [(out == k).mean() for k in range(3)]
    assert all(abs(f - 1 / 3) < 0.02 for f in frac), frac
    # spot-check 64 random gridpoints against numpy
    wk = so.iso_week(T)
    for _ in range(64):
        i, j, w = rng.integers(0, 256), rng.integers(0, 256), rng.integers(0, len(lab.weeks))
        wins = [((int(lab.weeks[w]) + d) % 53) or 53 for d in (-1, 0, 1)]
        np.testing.assert_array_equal(lab.edges[w, :, i, j], np.quantile(v[np.isin(wk, wins), i, j], [1 / 3, 2 / 3]))


def test_preprocess_layout_and_labels_match_the_oracle_labeler():
    from s2s_ismr_unet_b200.utils import preprocessing as pp
    x, y = make_xy()
    xtr, ytr, xva, yva, xte, yte = [l[0] for l in pp.bootstrap_splits(x, y, n_bootstraps=1)]
    X_train, Y_train_oh, X_val, Y_val_oh, X_test, Y_test_oh, ytr_t, yva_t, yte_t = pp.preprocess(xtr, ytr, xva, yva, xte, yte)
    assert X_train.shape == (len(xtr), 8, 8) and X_train.dtype == np.float32
    assert Y_train_oh.shape == (len(xtr), 8, 8, 3) and Y_train_oh.dtype == np.float32
    np.testing.assert_allclose(X_train, xtr.values.mean(1), rtol=1e-6)
    np.testing.assert_allclose(Y_val_oh.sum(-1), 1.0)
    wk_tr, wk_va = so.iso_week(ytr["T"]), so.iso_week(yva["T"])
    edges = so.rolling_tercile_edges(ytr.values, wk_tr, window=1)
    np.testing.assert_array_equal(yva_t.values, so.apply_tercile_labels(yva.values, wk_va, edges))
    np.testing.assert_array_equal(ytr_t.values, so.apply_tercile_labels(ytr.values, wk_tr, edges))
    # terciles of the training labels are roughly balanced away from the zero-filled ocean point
    frac = [(ytr_t.values[:, 1:, 1:] == k).mean() for k in range(3)]
    assert all(0.25 < f < 0.42 for f in frac)
    multi, _ = pp.convert_to_ndarray(xtr, ytr_t, "multi_predictor")
    assert multi.shape == (len(xtr), 8, 8, 4)                                         # channels-last (T,Y,X,M)
    st, yst, _ = pp.convert_to_ndarray(xtr, ytr_t, "stacked")
    assert st.shape == (4 * len(xtr), 8, 8) and yst.shape == (4 * len(xtr), 8, 8)


