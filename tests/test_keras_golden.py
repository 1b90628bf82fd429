"""Consumes golden vectors generated from the REAL reference under real Keras by tools/make_keras_golden.py
(tests/golden/keras/keras_unet_*.npz).  Keras / TensorFlow are not installable in this image, so no such file is
committed yet and these tests SKIP; the day someone runs the generator they pin the oracle (and, on a GPU box, the CUDA
path) against the reference itself.  Tolerances: BASELINE.json (forward rel-L2 <= 1e-5, per-step loss <= 1e-4)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import keras_unet as ko

GOLDEN = sorted((Path(__file__).parent / "golden" / "keras").glob("keras_unet_*.npz"))
pytestmark = pytest.mark.skipif(not GOLDEN, reason="no Keras-generated goldens (run tools/make_keras_golden.py where Keras exists)")


def _load(path):
    z = np.load(path)
    H, W, C = (int(v) for v in z["config.shape"])
    ct = z["config.ct_kernel"]
    cfg = ko.UnetConfig(H=H, W=W, Cin=C, filters=int(z["config.filters"]), n_blocks=int(z["config.n_blocks"]),
                        ct_kernel=int(np.ravel(ct)[0]), apool=bool(z["config.apool"]) if "config.apool" in z.files else True,
                        bn=bool(z["config.bn"]) if "config.bn" in z.files else True)
    grab = lambda pre: {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    return z, cfg, grab("w0/"), grab("w1/")


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_oracle_reproduces_keras(path):
    z, cfg, w0, w1 = _load(path)
    assert sorted(w0) == sorted(n for n, _, _ in ko.param_specs(cfg)), "layer / variable names differ from Keras'"
    oracle = ko.UnetOracle(cfg, w0, dtype=torch.float64)
    assert rel_l2(oracle.predict(z["x"]), z["predict0"]) <= 1e-5
    oracle.compile(lr=float(z["config.lr"]))
    loss, acc = oracle.train_step(z["x"], z["y"])
    assert abs(loss - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    after = oracle.get_weights()
    for k, v in w1.items():
        assert rel_l2(after[k], v) <= 1e-4, k
    assert rel_l2(oracle.predict(z["x"]), z["predict1"]) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_cuda_reproduces_keras(path):
    from s2s_ismr_unet_b200.keras_api.optimizers import Adam
    from s2s_ismr_unet_b200.model import Model
    z, cfg, w0, w1 = _load(path)
    m = Model((cfg.H, cfg.W, cfg.Cin), filters=cfg.filters, n_blocks=cfg.n_blocks, ct_kernel=cfg.ct_kernel, apool=cfg.apool,
              bn=cfg.bn, max_batch=len(z["x"]), weights=w0)
    assert rel_l2(m.predict(z["x"]), z["predict0"]) <= 1e-5
    m.compile(optimizer=Adam(learning_rate=float(z["config.lr"])), loss="categorical_crossentropy")
    loss, _ = m.train_on_batch(z["x"], z["y"])
    assert abs(loss - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    after = m.get_weights()
    for k, v in w1.items():
        assert rel_l2(after[k], v) <= 1e-4, k
    assert rel_l2(m.predict(z["x"]), z["predict1"]) <= 1e-5
