"""Batched ELR baseline (csrc/elr.cu) against the NumPy restatement of training.py:402-530 +
statsmodels' GLM IRLS (oracle/elr.py).  Tolerance: probabilities 1e-7 absolute (both sides iterate IRLS to a deviance
change <= 1e-8; the WLS is lstsq in the oracle and centred normal equations in the kernel)."""
import numpy as np
import pandas as pd
import pytest

from oracle import elr as eo
from oracle import skill as so

pytestmark = pytest.mark.gpu


def make_xy(years=range(1998, 2021), M=3, Y=8, X=8, seed=0):
    from s2s_ismr_unet_b200.labeled import LabeledArray
    rng = np.random.default_rng(seed)
    T = np.concatenate([pd.date_range(f"{y}-05-01", f"{y}-09-30", freq="7D").values for y in years])
    sig = rng.gamma(2.0, 3.0, size=(len(T), Y, X))
    x = (sig[:, None] * 0.6 + 0.4 * rng.gamma(2.0, 3.0, size=(len(T), M, Y, X))).astype(np.float32)
    y = (0.5 * sig + 0.5 * rng.gamma(2.0, 3.0, size=(len(T), Y, X))).astype(np.float32)
    y[:, 0, 0] = np.nan                                            # ocean: NaN predictand -> gridpoint skipped
    y[:, 1, 1] = 0.0                                               # dry point: e0 == 0 -> every row masked -> skipped
    y[::2, 2, 2] = 0.0                                             # half dry: e0 == 0 in every window
    co = {"T": T, "Y": np.arange(Y), "X": np.arange(X)}
    return LabeledArray(x, ("T", "M", "Y", "X"), {**co, "M": np.arange(M)}), LabeledArray(y, ("T", "Y", "X"), co)


def test_elr_matches_the_statsmodels_restatement():
    from s2s_ismr_unet_b200.utils import preprocessing as pp, training
    x, y = make_xy()
    xtr, ytr, xte, yte = [l[0] for l in pp.bootstrap_splits_ELR(x, y, n_bootstraps=1)]
    p_tr, p_te, ytr_t, yte_t = training.train_single_bootstrap_ELR(xtr, ytr, xte, yte)
    wtr, wte = so.iso_week(ytr["T"]), so.iso_week(yte["T"])
    o_tr, o_te, iters = eo.train_single_bootstrap_elr(xtr.values.mean(1), ytr.values, wtr, xte.values.mean(1), wte)
    assert iters.max() < 30 and (iters > 0).sum() >= 60
    for got, want in ((p_tr.values, o_tr), (p_te.values, o_te)):
        assert got.shape == want.shape
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert np.abs(got[ok] - want[ok]).max() <= 1e-7
        np.testing.assert_allclose(got[ok].reshape(-1, 3).sum(1), 1.0, atol=1e-12)
    assert np.isnan(p_tr.values[:, 0, 0]).all() and np.isnan(p_tr.values[:, 1, 1]).all()
    # terciled predictands come from the same edges as the U-Net path's labeler
    edges = so.rolling_tercile_edges(ytr.values, wtr, window=1)
    lab = so.apply_tercile_labels(yte.values, wte, edges)
    e_t, mask = eo.elr_masks(edges, wte)
    lab[mask] = np.nan
    np.testing.assert_array_equal(yte_t.values, lab)


def test_elr_recovers_a_known_logistic_model():
    """Synthetic truth: P(y <= edge_q) = sigmoid(b0 + b1 x + b2 q): the fitted probabilities converge to it."""
    from s2s_ismr_unet_b200.labeled import LabeledArray
    from s2s_ismr_unet_b200.utils import training
    rng = np.random.default_rng(3)
    T = np.concatenate([pd.date_range(f"{y}-06-01", f"{y}-09-30", freq="7D").values for y in range(1950, 2020)])
    Y = X = 4
    xs = rng.normal(5.0, 2.0, size=(len(T), 1, Y, X)).astype(np.float32)
    y = (xs[:, 0] + rng.logistic(0.0, 1.0, size=(len(T), Y, X))).astype(np.float32) + 20.0
    co = {"T": T, "Y": np.arange(Y), "X": np.arange(X)}
    xa, ya = LabeledArray(xs, ("T", "M", "Y", "X"), {**co, "M": np.arange(1)}), LabeledArray(y, ("T", "Y", "X"), co)
    p_tr, p_te, _, _ = training.train_single_bootstrap_ELR(xa, ya, xa, ya)
    pb = p_tr.values[..., 0]
    # below-normal probability decreases with the forecast, and the three classes are balanced on average
    hi, lo = xs[:, 0] > 6.5, xs[:, 0] < 3.5
    assert pb[hi].mean() < 0.2 < 0.5 < pb[lo].mean()
    np.testing.assert_allclose(p_tr.values.mean((0, 1, 2)), 1 / 3, atol=0.02)


def test_train_elr_rpss_shapes_and_skill():
    from s2s_ismr_unet_b200.utils import preprocessing as pp, training
    x, y = make_xy(seed=2)
    splits = pp.bootstrap_splits_ELR(x, y, n_bootstraps=2)
    rpss_tr, rpss_te, preds, y_oh = training.train_elr(*splits)
    assert len(rpss_tr) == len(rpss_te) == len(preds) == len(y_oh) == 2
    assert rpss_te[0].shape == (8, 8) and y_oh[0].shape == preds[0].shape
    assert np.nanmean(rpss_tr[0].values) > 0.02          # the forecast carries signal by construction


def test_train_elr_mme_averages_the_models():
    from s2s_ismr_unet_b200.utils import preprocessing as pp, training
    x1, y = make_xy(seed=4)
    x2, _ = make_xy(seed=5)
    xtr, xte, ytr, yte = pp.bootstrap_splits_ELR_mme({"GEFS": x1, "IITM": x2}, y, n_bootstraps=1)
    rpss_tr, rpss_te, preds, y_oh = training.train_elr_mme(xtr, ytr, xte, yte)
    pa = training.train_single_bootstrap_ELR(xtr["GEFS"][0], ytr[0], xte["GEFS"][0], yte[0])[1].values
    pb = training.train_single_bootstrap_ELR(xtr["IITM"][0], ytr[0], xte["IITM"][0], yte[0])[1].values
    want = (pa + pb) / 2
    want = want / want.sum(-1, keepdims=True)
    ok = ~np.isnan(want)
    np.testing.assert_allclose(preds[0].values[ok], want[ok], atol=2e-6)
    assert rpss_te[0].shape == (8, 8) and y_oh[0].shape == preds[0].shape
