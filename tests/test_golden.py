"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them (CPU), and the CUDA path must match them on the GPU box."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import elr as eo
from oracle import keras_unet as ko
from oracle import skill as so

G = Path(__file__).parent / "golden"
spec = importlib.util.spec_from_file_location("make_golden", G / "make_golden.py")
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def test_oracle_reproduces_unet_golden_vectors():
    z = np.load(G / "unet_small.npz")
    cfg = ko.UnetConfig(**mg.CFG)
    w = ko.random_init(cfg, mg.SEED_W)
    x, y = mg.inputs()
    net = ko.UnetOracle(cfg, w, dtype=torch.float64)
    np.testing.assert_allclose(net.predict(x), z["predict"], atol=1e-7)
    net.compile(lr=1e-3)
    loss0, acc0, g = net.backward(x, y)
    assert abs(loss0 - float(z["loss0"])) < 1e-12 and abs(acc0 - float(z["acc0"])) < 1e-12
    for k in z.files:
        if k.startswith("grad:"):
            np.testing.assert_allclose(g[k[5:]].numpy(), z[k], atol=1e-12)
    net = ko.UnetOracle(cfg, w, dtype=torch.float64)
    net.compile(lr=1e-3)
    np.testing.assert_allclose([net.train_step(x, y)[0] for _ in range(mg.STEPS)], z["losses"], rtol=1e-12)


def test_oracle_reproduces_skill_golden_vectors():
    z = np.load(G / "skill_small.npz")
    np.testing.assert_allclose(so.rps(z["o"], z["p"]), z["rps"], atol=1e-12, equal_nan=True)
    acc, cc = so.acc_cc(z["fx"], z["fy"], z["week"])
    np.testing.assert_allclose(acc, z["acc"], atol=1e-12)
    np.testing.assert_allclose(cc, z["cc"], atol=1e-12)


@pytest.mark.gpu
def test_cuda_path_matches_unet_golden_vectors():
    from s2s_ismr_unet_b200.keras_api.optimizers import Adam
    from s2s_ismr_unet_b200.model import Model
    z = np.load(G / "unet_small.npz")
    cfg = ko.UnetConfig(**mg.CFG)
    w = ko.random_init(cfg, mg.SEED_W)
    x, y = mg.inputs()
    m = Model((cfg.H, cfg.W, cfg.Cin), filters=cfg.filters, n_blocks=cfg.n_blocks, ct_kernel=cfg.ct_kernel, max_batch=mg.N, weights=w)
    assert rel_l2(m.predict(x), z["predict"]) <= 1e-5                     # BASELINE: fp32 forward rel-L2 <= 1e-5
    m.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
    loss0, acc0 = m.backward_on_batch(x, y)
    assert abs(loss0 - float(z["loss0"])) <= 1e-5 * float(z["loss0"]) and abs(acc0 - float(z["acc0"])) < 1e-6
    g = m.get_gradients()
    for k in z.files:
        if k.startswith("grad:"):
            assert rel_l2(g[k[5:]], z[k]) <= 2e-4, k
    m2 = Model((cfg.H, cfg.W, cfg.Cin), filters=cfg.filters, n_blocks=cfg.n_blocks, ct_kernel=cfg.ct_kernel, max_batch=mg.N, weights=w)
    m2.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
    losses = [m2.train_on_batch(x, y)[0] for _ in range(mg.STEPS)]
    np.testing.assert_allclose(losses, z["losses"], rtol=1e-4)            # BASELINE: per-step loss <= 1e-4 relative
    wa = m2.get_weights()
    for k in z.files:
        if k.startswith("after:"):
            assert rel_l2(wa[k[6:]], z[k]) <= 1e-4, k
    cam = m2.gradcam(x, "bottleneck", 2)
    assert np.abs(cam - z["gradcam_bottleneck_above"]).max() <= 1e-4 * max(np.abs(z["gradcam_bottleneck_above"]).max(), 1e-12) + 1e-9


@pytest.mark.gpu
def test_cuda_skill_matches_golden_vectors():
    from s2s_ismr_unet_b200.labeled import LabeledArray
    from s2s_ismr_unet_b200.utils import performance_metrics as pm
    z = np.load(G / "skill_small.npz")
    T, Y, X = z["fx"].shape
    lab = np.where(np.isnan(z["o"][..., 0]), np.nan, z["o"].argmax(-1).astype(np.float64))
    co = {"T": np.arange(T), "Y": np.arange(Y), "X": np.arange(X)}
    obs = LabeledArray(lab, ("T", "Y", "X"), co)
    fc = LabeledArray(z["p"], ("T", "Y", "X", "category"), co)
    np.testing.assert_allclose(pm.rps(obs, fc).values, z["rps"], atol=1e-5, equal_nan=True)
    clim = pm.climo_predict(LabeledArray(z["fx"], ("T", "Y", "X"), co))
    np.testing.assert_allclose(pm.rpss(clim, fc, obs).values, z["rpss"], atol=1e-4, equal_nan=True)
    acc, cc = pm.acc(LabeledArray(z["fx"], ("T", "Y", "X"), co), LabeledArray(z["fy"], ("T", "Y", "X"), co), week_index=z["week"], return_cc=True)
    np.testing.assert_allclose(acc.values, z["acc"], atol=1e-4)           # BASELINE: ACC maps <= 1e-4 absolute
    np.testing.assert_allclose(cc.values, z["cc"], atol=1e-4)


def test_oracle_reproduces_prep_and_elr_golden_vectors():
    z = np.load(G / "prep_elr_small.npz")
    got = mg.prep_elr_case()
    for k in z.files:
        if k.startswith("elr_t"):
            np.testing.assert_allclose(got[k], z[k], atol=1e-10, equal_nan=True)
        else:
            np.testing.assert_array_equal(got[k], z[k])


@pytest.mark.gpu
def test_cuda_labeler_and_elr_match_golden_vectors():
    """Tercile edges / labels bit-exact, ELR probabilities within 1e-7 of the committed vectors."""
    from s2s_ismr_unet_b200.labeled import LabeledArray
    from s2s_ismr_unet_b200.utils import preprocessing as pp, training
    z = np.load(G / "prep_elr_small.npz")
    T, x, y, week, test = mg.prep_elr_inputs()
    co = {"Y": np.arange(4), "X": np.arange(5)}
    mk = lambda v, sel: LabeledArray(v[sel], ("T", "Y", "X"), {**co, "T": T[sel]})
    ytr, yte = mk(y, ~test), mk(y, test)
    lab = pp.rolling_labeler(ytr, window=1)
    np.testing.assert_array_equal(lab.weeks, z["weeks"])
    np.testing.assert_array_equal(lab.edges, z["edges"])
    np.testing.assert_array_equal(lab(ytr).values, z["labels_train"])
    np.testing.assert_array_equal(lab(yte).values, z["labels_test"])
    xm = lambda sel: LabeledArray(x[sel][:, None], ("T", "M", "Y", "X"), {**co, "T": T[sel], "M": np.arange(1)})
    p_tr, p_te, _, _ = training.train_single_bootstrap_ELR(xm(~test), ytr, xm(test), yte)
    for got, want in ((p_tr.values, z["elr_train"]), (p_te.values, z["elr_test"])):
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert np.abs(got[ok] - want[ok]).max() <= 1e-7
