"""The reference-facing API end to end on the GPU: utils.training.train_deepnet / train_deepnet_mme with the
reference's own keyword arguments (tune_ECMWF_com.py:94-109, tune_MME.py:118-131) on a small synthetic
hindcast set; checks shapes, file naming, probabilities and that the RPSS maps equal the oracle's."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import skill as so

pytestmark = pytest.mark.gpu


def synth(years=range(2003, 2013), M=3, Y=16, X=16, seed=0):
    from s2s_ismr_unet_b200.labeled import LabeledArray
    rng = np.random.default_rng(seed)
    T = np.concatenate([pd.date_range(f"{y}-06-01", f"{y}-08-31", freq="7D").values for y in years])
    x = rng.gamma(2.0, 3.0, size=(len(T), M, Y, X)).astype(np.float32)
    y = (0.5 * x.mean(1) + 0.5 * rng.gamma(2.0, 3.0, size=(len(T), Y, X))).astype(np.float32)
    co = {"T": T, "Y": np.arange(Y), "X": np.arange(X)}
    return LabeledArray(x, ("T", "M", "Y", "X"), {**co, "M": np.arange(M)}), LabeledArray(y, ("T", "Y", "X"), co)


def test_train_deepnet_tune_path(tmp_path, monkeypatch):
    from s2s_ismr_unet_b200.utils import preprocessing, training
    monkeypatch.chdir(tmp_path)
    x, y = synth()
    splits = preprocessing.bootstrap_splits(x, y, n_bootstraps=2)
    grid = {"n_blocks": [3, 4], "n_filters": [2], "ct_kernels": [(3, 3)], "batch_sizes": [16], "learning_rates": [1e-3], "patience": 2}
    out = training.train_deepnet(*splits, training_type="tune", architecture="unet", architecture_params=None, tuning_grid=grid,
                                 predictor="mean", obs="IMD", modname="GEFS", week="wk3-4", epochs=3, batch_size=16, dir="T/")
    rpss_train, rpss_val, rpss_test, preds, y_oh = out
    assert len(rpss_test) == 2 and rpss_test[0].shape == (16, 16) and np.isfinite(rpss_test[0].values).all()
    p = preds[0]
    assert p.dims == ("T", "Y", "X", "category") and p.shape[1:] == (16, 16, 3)
    np.testing.assert_allclose(p.values.sum(-1), 1.0, atol=1e-5)
    assert os.path.exists("models/T/GEFS_IMD/wk3-4/best_model_unet_0_tuned.keras")
    assert os.path.exists("models/T/GEFS_IMD/wk3-4/best_model_unet_bootstrap_2_trial_1.keras")
    # RPSS equals the oracle's on the same predictions / labels
    xte, yte = splits[4][0], splits[5][0]
    ytr = splits[1][0]
    edges = so.rolling_tercile_edges(ytr.values, so.iso_week(ytr["T"]))
    lab = so.apply_tercile_labels(yte.values, so.iso_week(yte["T"]), edges)
    ref = so.rpss(so.climo_forecast(lab.shape), p.values, so.onehot_obs(lab))
    np.testing.assert_allclose(rpss_test[0].values, ref, atol=1e-4)
    # "load" re-uses the tuned model and reproduces the same predictions
    out2 = training.train_deepnet(*splits, training_type="load", architecture="unet", predictor="mean", obs="IMD", modname="GEFS",
                                  week="wk3-4", dir="T/")
    np.testing.assert_array_equal(out2[3][0].values, p.values)


def test_train_deepnet_mme_averages_and_renormalises(tmp_path, monkeypatch):
    from s2s_ismr_unet_b200.utils import preprocessing, training
    monkeypatch.chdir(tmp_path)
    x1, y = synth(seed=1)
    x2, _ = synth(seed=2)
    xtr, xva, xte, ytr, yva, yte = preprocessing.bootstrap_splits_mme({"GEFS": x1, "IITM": x2}, y, n_bootstraps=1)
    out = training.train_deepnet_mme(xtr, ytr, xva, yva, xte, yte, training_type="train", architecture="unet",
                                     architecture_params={"n_blocks": 3, "filters": 2, "ct_kernel": (2, 2)}, predictor="mean",
                                     obs="IMD", week="wk2", epochs=2, batch_size=16, learning_rate=1e-3, dir="M/")
    rpss_train, rpss_val, rpss_test, preds, y_oh = out
    np.testing.assert_allclose(preds[0].values.sum(-1), 1.0, atol=1e-5)
    assert rpss_val[0].shape == (16, 16) and os.path.exists("models/M/IITM_IMD/wk2/best_model_unet_0.keras")


def test_tune_with_eight_concurrent_trial_threads(tmp_path, monkeypatch):
    """The tuning loop runs S2S_TRIAL_WORKERS trials on their own host threads and streams (graph captures, data-set uploads,
    cached device buffers handed from one thread's trial to another's): the full 18-trial grid of tune_2MME.py on a 32x32
    grid, 8 workers — a device-wide synchronisation in any of those paths breaks another thread's capture."""
    from s2s_ismr_unet_b200.utils import preprocessing, training
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("S2S_TRIAL_WORKERS", "8")
    x, y = synth(Y=32, X=32, seed=3)
    splits = preprocessing.bootstrap_splits(x, y, n_bootstraps=1)
    grid = {"n_blocks": [3, 4, 5], "n_filters": [2, 3], "ct_kernels": [(2, 2), (3, 3), (5, 5)], "batch_sizes": [16],
            "learning_rates": [1e-3], "patience": 2}
    out = training.train_deepnet(*splits, training_type="tune", architecture="unet", tuning_grid=grid, predictor="mean", obs="IMD",
                                 modname="GEFS", week="wk3-4", epochs=3, batch_size=16, dir="W/")
    assert np.isfinite(out[2][0].values).all()
    np.testing.assert_allclose(out[3][0].values.sum(-1), 1.0, atol=1e-5)
