"""Golden vectors from the REAL reference under real Keras (closes the "parity unpinned" gap wherever Keras exists).

This image has no TensorFlow / Keras, so the script cannot run here; it is committed so that anyone with the reference's
environment (dependencies-windows.txt: keras 3.5.0 / tensorflow 2.18.0) can pin the oracle in one command:

    python tools/make_keras_golden.py --reference /path/to/s2s-ismr-unet [--out tests/golden/keras]

It imports the UNMODIFIED reference module utils/deep_nn_models.py (Unet.build_model, :17-163), compiles the model
exactly as utils/training.py:66-67 does (Adam(learning_rate), "categorical_crossentropy", ['accuracy']) and dumps, per
configuration, one .npz holding

    config.*                 the Unet kwargs / input shape / learning rate
    w0/<layer>/<var>         every variable before the step (Keras names: kernel, bias, gamma, beta, moving_mean, ...)
    x, y                     one synthetic batch (seeded NumPy, NHWC / one-hot)
    predict0                 model.predict(x) before the step (inference-mode BatchNorm)
    loss, accuracy           model.train_on_batch(x, y)  == one optimiser step of model.fit (training.py:102)
    w1/<layer>/<var>         every variable after the step (incl. the updated BN moving statistics)
    predict1                 model.predict(x) after the step
    fit_val_loss             history.history['val_loss'] of a 2-epoch model.fit(shuffle=False) from the w1 state

tests/test_keras_golden.py consumes these files when present (and is skipped otherwise): the oracle restatement
(oracle/keras_unet.py) must reproduce them within BASELINE.json's tolerances, which pins the Keras-specific choices
(BatchNorm eps / momentum / biased moving variance, CCE renormalise + clip, Adam epsilon placement) that nothing in
this container can cross-check.
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import numpy as np

CONFIGS = {
    # name: (input shape (H, W, C), Unet kwargs, batch, learning rate)
    "default_c1": ((64, 64, 1), dict(filters=2, n_blocks=3, ct_kernel=(3, 3)), 16, 1e-3),
    "mme_c3": ((64, 64, 3), dict(filters=2, n_blocks=3, ct_kernel=(3, 3)), 16, 1e-3),
    "ecmwf24_f3_ct5": ((24, 24, 1), dict(filters=3, n_blocks=3, ct_kernel=(5, 5)), 5, 1e-4),
    "nb4_ct2_maxpool": ((32, 32, 1), dict(filters=2, n_blocks=4, ct_kernel=(2, 2), apool=False), 8, 1e-3),
    "nobn": ((32, 32, 2), dict(filters=2, n_blocks=3, ct_kernel=(3, 3), bn=False), 8, 1e-3),
}


def synthetic_batch(shape, n, seed):
    rng = np.random.default_rng(seed)
    H, W, C = shape
    x = (rng.gamma(2.0, 3.0, size=(n, H, W, C)) / 6.0).astype(np.float32)
    y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=(n, H, W))]
    return x, y


def variables(model) -> dict:
    """{'<layer name>/<variable name>': array} in layer-creation order (the names the oracle / CUDA layouts use)."""
    out = {}
    for layer in model.layers:
        for v in layer.weights:
            name = getattr(v, "name", None) or v.path.split("/")[-1]
            name = name.split("/")[-1].split(":")[0]
            out[f"{layer.name}/{name}"] = np.asarray(v.numpy() if hasattr(v, "numpy") else v)
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("S2S_REFERENCE", "/root/reference"))
    ap.add_argument("--out", default=str(Path(__file__).resolve().parent.parent / "tests" / "golden" / "keras"))
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import keras                                   # noqa: F401  (fails here: that is expected, see the docstring)
    from keras import optimizers
    from utils import deep_nn_models               # the unmodified reference module

    keras.utils.set_random_seed(42)                # training.py:23-27 seeds everything with 42
    out_dir = Path(args.out)
    out_dir.mkdir(parents=True, exist_ok=True)
    for name, (shape, kw, n, lr) in CONFIGS.items():
        model = deep_nn_models.Unet("", train_patches=False, weighted_loss=False, **kw).build_model(shape, dg_train_weight_target=None)
        model.compile(optimizer=optimizers.Adam(learning_rate=lr), loss="categorical_crossentropy", metrics=["accuracy"])
        # non-degenerate BatchNorm state / biases so that every term of the arithmetic is exercised
        rng = np.random.default_rng(7)
        for layer in model.layers:
            ws = layer.get_weights()
            if not ws:
                continue
            new = []
            for v, w in zip(layer.weights, ws):
                vn = (getattr(v, "name", "") or "").split("/")[-1].split(":")[0]
                if vn in ("bias", "beta", "moving_mean"):
                    w = rng.normal(0, 0.1, size=w.shape).astype(np.float32)
                elif vn == "gamma":
                    w = rng.uniform(0.5, 1.5, size=w.shape).astype(np.float32)
                elif vn == "moving_variance":
                    w = rng.uniform(0.5, 2.0, size=w.shape).astype(np.float32)
                new.append(w)
            layer.set_weights(new)
        x, y = synthetic_batch(shape, n, seed=1)
        rec = {"x": x, "y": y, "config.shape": np.asarray(shape), "config.lr": np.asarray(lr), "config.batch": np.asarray(n)}
        for k, v in kw.items():
            rec[f"config.{k}"] = np.asarray(v)
        for k, v in variables(model).items():
            rec["w0/" + k] = v
        rec["predict0"] = model.predict(x, verbose=0)
        res = model.train_on_batch(x, y, return_dict=True)
        rec["loss"], rec["accuracy"] = np.asarray(res["loss"]), np.asarray(res.get("accuracy", res.get("compile_metrics", np.nan)))
        for k, v in variables(model).items():
            rec["w1/" + k] = v
        rec["predict1"] = model.predict(x, verbose=0)
        xv, yv = synthetic_batch(shape, n, seed=2)
        hist = model.fit(x=x, y=y, validation_data=(xv, yv), epochs=2, batch_size=n, shuffle=False, verbose=0)
        rec["xv"], rec["yv"] = xv, yv
        rec["fit_loss"], rec["fit_val_loss"] = np.asarray(hist.history["loss"]), np.asarray(hist.history["val_loss"])
        rec["versions"] = np.asarray([f"keras {keras.__version__}"])
        np.savez_compressed(out_dir / f"keras_unet_{name}.npz", **rec)
        print(f"wrote {out_dir / f'keras_unet_{name}.npz'}: loss {float(rec['loss']):.6f}")


if __name__ == "__main__":
    main()
