"""Keras-3 `.keras` archives for the U-Net of utils/deep_nn_models.py (training.py:114-115, 127-131).

A `.keras` file is a zip of
    config.json          the serialised Functional model (module / class_name / config / compile_config)
    metadata.json        {"keras_version", "date_saved"}
    model.weights.h5     HDF5: layers/<snake_case(class)[_k]>/vars/<i>  and  optimizer/vars/<i>
where the layer groups are named after the layer CLASS with a per-class counter in model.layers order
(keras/src/saving/saving_lib.py, _save_container_state) and a layer's variables are numbered in
trainable-then-non-trainable order (Conv2D: kernel, bias; BatchNormalization: gamma, beta, moving_mean,
moving_variance).  The optimizer group holds [iteration, learning_rate, m_0, v_0, m_1, v_1, ...] over
model.trainable_variables (keras/src/optimizers/adam.py, build).

`read_keras_archive` loads a file written by the reference (real Keras) into this library's weight dict + `Model`
kwargs; `write_keras_archive` writes the same layout, so that `keras.models.load_model` can read a model trained
here.  h5py does not exist in this image: the HDF5 container is read / written by keras_api/hdf5_min.py, and the
round trip is tested against this reader only (tests/test_host_logic.py) — a load by real Keras is untested here.
"""
from __future__ import annotations

import datetime
import io
import json
import re
import zipfile

import numpy as np

from .hdf5_min import read_hdf5, write_hdf5

KERAS_VERSION = "3.5.0"         # dependencies-windows.txt:190


def to_snake_case(name: str) -> str:
    """keras/src/utils/naming.py:to_snake_case."""
    name = re.sub(r"\W+", "", name)
    name = re.sub("(.)([A-Z][a-z]+)", r"\1_\2", name)
    name = re.sub("([a-z])([A-Z])", r"\1_\2", name).lower()
    return name


# ------------------------------------------------------------------------------------------------ the layer list
def unet_layers(cfg: dict) -> list:
    """[(class_name, layer name, layer config extras, inbound layer names)] in the creation order of
    Unet.build_model (deep_nn_models.py:75-105, 139-163).  Unnamed layers get Keras' automatic names."""
    H, W, C = cfg["input_shape"]
    f, nb, k = cfg["filters"], cfg["n_blocks"], cfg["ct_kernel"]
    act = cfg.get("activation", "elu")
    counters: dict = {}

    def auto(cls):
        base = to_snake_case(cls)
        n = counters.get(base, 0)
        counters[base] = n + 1
        return base if n == 0 else f"{base}_{n}"

    layers = []
    add = lambda cls, name, extra, inbound: (layers.append((cls, name, extra, inbound)), name)[1]
    cur = add("InputLayer", auto("InputLayer"), {"batch_shape": [None, H, W, C]}, [])
    skips = []
    for b in range(nb):
        ch = f * 4 * 2 ** b
        c = add("Conv2D", f"down_conv{b + 1}_1", {"filters": ch, "kernel_size": [3, 3], "activation": act}, [cur])
        c = add("Dropout", auto("Dropout"), {"rate": 0}, [c])
        c = add("Conv2D", f"down_conv{b + 1}_2", {"filters": ch, "kernel_size": [3, 3], "activation": act}, [c])
        if cfg["bn"]:
            c = add("BatchNormalization", auto("BatchNormalization"), {}, [c])
        skips.append(c)
        pool = "AveragePooling2D" if cfg["apool"] else "MaxPooling2D"
        cur = add(pool, auto(pool), {"pool_size": [2, 2]}, [c])
    ch = f * 4 * 2 ** nb
    cur = add("Conv2D", "bottleneck", {"filters": ch, "kernel_size": [3, 3], "activation": act}, [cur])
    cur = add("Conv2D", auto("Conv2D"), {"filters": ch, "kernel_size": [3, 3], "activation": act}, [cur])
    if cfg["bn"]:
        cur = add("BatchNormalization", auto("BatchNormalization"), {}, [cur])
    for b in range(nb - 1, -1, -1):
        ch = f * 4 * 2 ** b
        u = add("Conv2DTranspose", f"up_conv{b + 1}_1", {"filters": ch, "kernel_size": [k, k], "strides": [2, 2], "activation": "linear"}, [cur])
        u = add("Concatenate", auto("Concatenate"), {"axis": -1}, [skips[b], u])
        u = add("Conv2D", f"up_conv{b + 1}_2", {"filters": ch, "kernel_size": [3, 3], "activation": act}, [u])
        u = add("Dropout", auto("Dropout"), {"rate": 0}, [u])
        u = add("Conv2D", f"up_conv{b + 1}_3", {"filters": ch, "kernel_size": [3, 3], "activation": act}, [u])
        if cfg["bn"] and b > 0:
            u = add("BatchNormalization", auto("BatchNormalization"), {}, [u])
        cur = u
    head = {"filters": 3, "kernel_size": [1, 1], "activation": "softmax"} if cfg["output"] == "proba" else \
           {"filters": 1, "kernel_size": [1, 1], "activation": "relu"}
    add("Conv2D", auto("Conv2D"), head, [cur])
    return layers


VARS_OF = {"Conv2D": ("kernel", "bias"), "Conv2DTranspose": ("kernel", "bias"),
           "BatchNormalization": ("gamma", "beta", "moving_mean", "moving_variance")}
TRAINABLE = {"kernel", "bias", "gamma", "beta"}


def h5_groups(layers) -> list:
    """[(h5 group name, class, layer name)]: class-name-indexed group names in model.layers order."""
    used, out = {}, []
    for cls, name, _, _ in layers:
        g = to_snake_case(cls)
        if g in used:
            used[g] += 1
            g = f"{g}_{used[g]}"
        else:
            used[g] = 0
        out.append((g, cls, name))
    return out


# ------------------------------------------------------------------------------------------------ write
def _dtype_policy():
    return {"module": "keras", "class_name": "DTypePolicy", "config": {"name": "float32"}, "registered_name": None}


def _layer_config(cls, name, extra, inbound, shapes):
    c = {"name": name, "trainable": True, "dtype": _dtype_policy()}
    l2 = {"module": "keras.regularizers", "class_name": "L2", "config": {"l2": 0.0}, "registered_name": None}
    init = lambda k, **kw: {"module": "keras.initializers", "class_name": k, "config": kw, "registered_name": None}
    if cls == "InputLayer":
        c = {"batch_shape": extra["batch_shape"], "dtype": "float32", "sparse": False, "name": name}
    elif cls in ("Conv2D", "Conv2DTranspose"):
        c.update(filters=extra["filters"], kernel_size=extra["kernel_size"], strides=extra.get("strides", [1, 1]), padding="same",
                 data_format="channels_last", dilation_rate=[1, 1], activation=extra["activation"], use_bias=True,
                 kernel_initializer=init("GlorotUniform", seed=None), bias_initializer=init("Zeros"),
                 kernel_regularizer=l2, bias_regularizer=l2, activity_regularizer=None, kernel_constraint=None, bias_constraint=None)
        if cls == "Conv2D":
            c["groups"] = 1
        else:
            c["output_padding"] = None
    elif cls == "BatchNormalization":
        c.update(axis=-1, momentum=0.99, epsilon=0.001, center=True, scale=True, beta_initializer=init("Zeros"),
                 gamma_initializer=init("Ones"), moving_mean_initializer=init("Zeros"), moving_variance_initializer=init("Ones"),
                 beta_regularizer=None, gamma_regularizer=None, beta_constraint=None, gamma_constraint=None, synchronized=False)
    elif cls in ("AveragePooling2D", "MaxPooling2D"):
        c.update(pool_size=extra["pool_size"], padding="valid", strides=extra["pool_size"], data_format="channels_last")
    elif cls == "Dropout":
        c.update(rate=extra["rate"], seed=None, noise_shape=None)
    elif cls == "Concatenate":
        c.update(axis=-1)
    args = [{"class_name": "__keras_tensor__", "config": {"shape": shapes[i], "dtype": "float32", "keras_history": [i, 0, 0]}} for i in inbound]
    if cls == "Concatenate":
        args = [args]
    entry = {"module": "keras.layers", "class_name": cls, "config": c, "registered_name": None, "name": name,
             "inbound_nodes": [{"args": args, "kwargs": {}}] if inbound else []}
    if cls != "InputLayer" and inbound:
        entry["build_config"] = {"input_shape": [shapes[i] for i in inbound] if cls == "Concatenate" else shapes[inbound[0]]}
    return entry


def _shapes(layers) -> dict:
    shp = {}
    for cls, name, extra, inbound in layers:
        if cls == "InputLayer":
            shp[name] = extra["batch_shape"]
            continue
        s = list(shp[inbound[0]])
        if cls == "Conv2D":
            s[3] = extra["filters"]
        elif cls == "Conv2DTranspose":
            s = [s[0], s[1] * 2, s[2] * 2, extra["filters"]]
        elif cls in ("AveragePooling2D", "MaxPooling2D"):
            s = [s[0], s[1] // 2, s[2] // 2, s[3]]
        elif cls == "Concatenate":
            s[3] = sum(shp[i][3] for i in inbound)
        shp[name] = s
    return shp


def keras_config(cfg: dict, optimizer=None) -> dict:
    layers = unet_layers(cfg)
    shapes = _shapes(layers)
    out = {"module": "keras", "class_name": "Functional",
           "config": {"name": "functional", "trainable": True,
                      "layers": [_layer_config(c, n, e, i, shapes) for c, n, e, i in layers],
                      "input_layers": [[layers[0][1], 0, 0]], "output_layers": [[layers[-1][1], 0, 0]]},
           "registered_name": "Functional", "build_config": {"input_shape": None}}
    if optimizer is not None:
        out["compile_config"] = {
            "optimizer": {"module": "keras.optimizers", "class_name": "Adam",
                          "config": {"name": "adam", "learning_rate": float(optimizer.learning_rate), "weight_decay": None, "clipnorm": None,
                                     "global_clipnorm": None, "clipvalue": None, "use_ema": False, "ema_momentum": 0.99,
                                     "ema_overwrite_frequency": None, "loss_scale_factor": None, "gradient_accumulation_steps": None,
                                     "beta_1": float(optimizer.beta_1), "beta_2": float(optimizer.beta_2), "epsilon": float(optimizer.epsilon),
                                     "amsgrad": False}, "registered_name": None},
            "loss": "categorical_crossentropy" if cfg["output"] == "proba" else "mse", "loss_weights": None, "metrics": ["accuracy"],
            "weighted_metrics": None, "run_eagerly": False, "steps_per_execution": 1, "jit_compile": False}
    return out


def weights_tree(cfg: dict, weights: dict, opt_state: dict | None, layout: list | None, optimizer=None) -> dict:
    """The HDF5 tree of model.weights.h5 for a weight dict {'<layer>/<var>': array}."""
    layers = unet_layers(cfg)
    tree = {"layers": {}, "vars": {}}
    trainable = []
    for (g, cls, name) in h5_groups(layers):
        vs = {}
        for i, var in enumerate(VARS_OF.get(cls, ())):
            arr = np.asarray(weights[f"{name}/{var}"], np.float32)
            vs[str(i)] = arr
            if var in TRAINABLE:
                trainable.append(f"{name}/{var}")
        tree["layers"][g] = {"vars": vs}
    if opt_state is not None and layout is not None:
        ov = {"0": np.asarray(int(opt_state["step"]), np.int64),
              "1": np.asarray(float(optimizer.learning_rate) if optimizer is not None else 1e-3, np.float32)}
        by_name = {d["name"]: d for d in layout if d["arena"] == 0}
        m, v = np.asarray(opt_state["m"]), np.asarray(opt_state["v"])
        for j, tn in enumerate(trainable):
            d = by_name[tn]
            sl = slice(d["offset"], d["offset"] + d["count"])
            ov[str(2 + 2 * j)] = m[sl].reshape(d["shape"]).astype(np.float32)
            ov[str(3 + 2 * j)] = v[sl].reshape(d["shape"]).astype(np.float32)
        tree["optimizer"] = {"vars": ov}
    return tree


def write_keras_archive(path, cfg: dict, weights: dict, opt_state=None, optimizer=None, layout=None, extra_members: dict | None = None) -> None:
    with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
        z.writestr("metadata.json", json.dumps({"keras_version": KERAS_VERSION,
                                                "date_saved": datetime.datetime.now().strftime("%Y-%m-%d@%H:%M:%S")}))
        z.writestr("config.json", json.dumps(keras_config(cfg, optimizer)))
        z.writestr("model.weights.h5", write_hdf5(weights_tree(cfg, weights, opt_state, layout, optimizer)))
        for name, blob in (extra_members or {}).items():
            z.writestr(name, blob)


# ------------------------------------------------------------------------------------------------ read
def parse_keras_config(kc: dict) -> tuple[dict, list]:
    """Keras Functional config of the reference's U-Net -> (Model kwargs, [(class, layer name)] in layer order)."""
    if kc.get("class_name") not in ("Functional", "Model"):
        raise ValueError(f"not a Keras functional model: class_name={kc.get('class_name')!r}")
    layers = kc["config"]["layers"]
    seq = [(L["class_name"], L["config"].get("name", L.get("name"))) for L in layers]
    cfgs = [L["config"] for L in layers]
    convs = [c for (cls, _), c in zip(seq, cfgs) if cls == "Conv2D"]
    convts = [c for (cls, _), c in zip(seq, cfgs) if cls == "Conv2DTranspose"]
    inp = next(c for (cls, _), c in zip(seq, cfgs) if cls == "InputLayer")
    shape = inp.get("batch_shape") or inp.get("batch_input_shape")
    if not convs or not convts or shape is None:
        raise ValueError("this archive is not the U-Net of utils/deep_nn_models.py (no Conv2D / Conv2DTranspose / InputLayer)")
    names = {n for _, n in seq}
    for need in ("down_conv1_1", "bottleneck", "up_conv1_3"):
        if need not in names:
            raise ValueError(f"this archive is not the U-Net of utils/deep_nn_models.py (layer {need!r} is missing)")
    head = convs[-1]
    if head["filters"] == 3 and head.get("activation") == "softmax":
        output = "proba"
    elif head["filters"] == 1 and head.get("activation") == "relu":
        output = "deterministic"
    else:
        raise ValueError(f"unsupported output layer: {head['filters']} filters, activation {head.get('activation')!r}")
    kw = dict(input_shape=[int(shape[1]), int(shape[2]), int(shape[3])], filters=int(convs[0]["filters"]) // 4, n_blocks=len(convts),
              ct_kernel=int(convts[0]["kernel_size"][0]), apool=any(cls == "AveragePooling2D" for cls, _ in seq),
              bn=any(cls == "BatchNormalization" for cls, _ in seq), output=output, activation=convs[0].get("activation", "elu"))
    return kw, seq


def read_keras_archive(path) -> dict:
    """{'config': Model kwargs, 'weights': {'<layer>/<var>': array}, 'optimizer': {...} | None, 'opt_vars': [...]}."""
    with zipfile.ZipFile(path) as z:
        kc = json.loads(z.read("config.json"))
        tree = read_hdf5(z.read("model.weights.h5"))
    kw, seq = parse_keras_config(kc)
    if "layers" not in tree:
        raise ValueError(f"model.weights.h5 has no 'layers' group (found {sorted(tree)})")
    weights, trainable = {}, []
    for (g, cls, name) in h5_groups([(cls, n, None, None) for cls, n in seq]):
        names = VARS_OF.get(cls, ())
        if not names:
            continue
        if g not in tree["layers"] or "vars" not in tree["layers"][g]:
            raise ValueError(f"model.weights.h5 lacks layers/{g}/vars for layer {name!r}")
        vs = tree["layers"][g]["vars"]
        if len(vs) != len(names):
            raise ValueError(f"layer {name!r} ({cls}) expected {len(names)} variables, the file holds {len(vs)}")
        for i, var in enumerate(names):
            weights[f"{name}/{var}"] = np.asarray(vs[str(i)], np.float32)
            if var in TRAINABLE:
                trainable.append(f"{name}/{var}")
    out = {"config": kw, "weights": weights, "optimizer": None, "opt_vars": None, "trainable_order": trainable}
    cc = kc.get("compile_config") or {}
    oc = (cc.get("optimizer") or {}).get("config") if isinstance(cc.get("optimizer"), dict) else None
    if oc:
        lr = oc.get("learning_rate", 1e-3)
        out["optimizer"] = dict(learning_rate=float(lr) if not isinstance(lr, dict) else 1e-3, beta_1=float(oc.get("beta_1", 0.9)),
                                beta_2=float(oc.get("beta_2", 0.999)), epsilon=float(oc.get("epsilon", 1e-7)))
    ov = (tree.get("optimizer") or {}).get("vars") or {}
    if len(ov) == 2 + 2 * len(trainable):
        out["opt_vars"] = dict(step=int(np.asarray(ov["0"]).ravel()[0]),
                               m={tn: np.asarray(ov[str(2 + 2 * j)], np.float32) for j, tn in enumerate(trainable)},
                               v={tn: np.asarray(ov[str(3 + 2 * j)], np.float32) for j, tn in enumerate(trainable)})
    return out
