"""Generates tests/golden/unet_small.npz, skill_small.npz and prep_elr_small.npz from the oracle (run from the repo root):

    python tests/golden/make_golden.py

The reference (Keras/TensorFlow) cannot be imported in this image, so these vectors pin the ORACLE's
restatement (PARITY UNPINNED against real Keras, see oracle/keras_unet.py); they guard both the oracle and the
CUDA path against regressions and travel to the GPU box, where /root/reference does not exist."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import elr as eo  # noqa: E402
from oracle import keras_unet as ko  # noqa: E402
from oracle import skill as so  # noqa: E402

CFG = dict(H=16, W=16, Cin=3, filters=2, n_blocks=3, ct_kernel=3)
SEED_W, SEED_X, N, STEPS = 11, 12, 4, 3


def inputs():
    rng = np.random.default_rng(SEED_X)
    x = (rng.gamma(2.0, 3.0, size=(N, 16, 16, 3)) / 6.0).astype(np.float32)
    y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=(N, 16, 16))]
    return x, y


def prep_elr_inputs():
    """12 years of weekly May-Sep starts on a 4x5 grid: predictand y (float32), ensemble-mean forecast x, ISO weeks;
    one all-NaN (ocean) point, one dry point (e0 == 0), one point with ties.  Years 2015-16 are the test period."""
    import pandas as pd
    rng = np.random.default_rng(21)
    T = np.concatenate([pd.date_range(f"{yr}-05-01", f"{yr}-09-30", freq="7D").values for yr in range(2005, 2017)])
    sig = rng.gamma(2.0, 3.0, size=(len(T), 4, 5))
    x = (0.6 * sig + 0.4 * rng.gamma(2.0, 3.0, size=sig.shape)).astype(np.float32)
    y = (0.5 * sig + 0.5 * rng.gamma(2.0, 3.0, size=sig.shape)).astype(np.float32)
    y[:, 0, 0] = np.nan
    y[:, 1, 1] = 0.0
    y[::3, 2, 2] = y[1, 2, 2]
    week = so.iso_week(T)
    test = np.asarray(pd.DatetimeIndex(T).year) >= 2015
    return T, x, y, week, test


def prep_elr_case():
    T, x, y, week, test = prep_elr_inputs()
    tr = ~test
    edges = so.rolling_tercile_edges(y[tr], week[tr], window=1)
    weeks = np.array(sorted(edges))
    out = {"weeks": weeks, "edges": np.stack([edges[int(w)] for w in weeks]),
           "labels_train": so.apply_tercile_labels(y[tr], week[tr], edges),
           "labels_test": so.apply_tercile_labels(y[test], week[test], edges)}
    p_tr, p_te, iters = eo.train_single_bootstrap_elr(x[tr], y[tr], week[tr], x[test], week[test])
    out.update(elr_train=p_tr, elr_test=p_te, elr_iters=iters)
    return out


def main():
    cfg = ko.UnetConfig(**CFG)
    w = ko.random_init(cfg, SEED_W)
    x, y = inputs()
    net = ko.UnetOracle(cfg, w, dtype=torch.float64)
    out = {"predict": net.predict(x)}
    net.compile(lr=1e-3)
    loss0, acc0, g = net.backward(x, y)
    out["loss0"], out["acc0"] = np.float64(loss0), np.float64(acc0)
    for k in ("down_conv1_1/kernel", "up_conv2_1/kernel", "batch_normalization/gamma", "conv2d_1/kernel", "bottleneck/bias"):
        out["grad:" + k] = g[k].numpy()
    net = ko.UnetOracle(cfg, w, dtype=torch.float64)
    net.compile(lr=1e-3)
    out["losses"] = np.array([net.train_step(x, y)[0] for _ in range(STEPS)])
    wa = net.get_weights()
    for k in ("down_conv1_1/kernel", "conv2d_1/bias", "batch_normalization/moving_variance", "up_conv1_3/kernel"):
        out["after:" + k] = wa[k]
    out["gradcam_bottleneck_above"] = net.gradcam(x, "bottleneck", 2)
    np.savez_compressed(ROOT / "tests" / "golden" / "unet_small.npz", **out)

    rng = np.random.default_rng(3)
    T, Y, X = 40, 6, 5
    p = rng.dirichlet([1, 1, 1], size=(T, Y, X)).astype(np.float32)
    lab = rng.integers(0, 3, size=(T, Y, X)).astype(np.float64)
    lab[rng.random((T, Y, X)) < 0.1] = np.nan
    o = so.onehot_obs(lab).astype(np.float32)
    week = rng.integers(20, 26, size=T)
    fx = rng.gamma(2.0, 3.0, size=(T, Y, X)).astype(np.float32)
    fy = (0.4 * fx + rng.gamma(2.0, 3.0, size=(T, Y, X))).astype(np.float32)
    acc, cc = so.acc_cc(fx, fy, week)
    np.savez_compressed(ROOT / "tests" / "golden" / "skill_small.npz", p=p, o=o, week=week, fx=fx, fy=fy,
                        rps=so.rps(o, p), rpss=so.rpss(so.climo_forecast((T, Y, X)), p, o), acc=acc, cc=cc)
    np.savez_compressed(ROOT / "tests" / "golden" / "prep_elr_small.npz", **prep_elr_case())
    print("golden vectors written")


if __name__ == "__main__":
    main()
