"""Mirror of the U-Net half of utils/training.py: reset_random_seeds, train_single_bootstrap_deepnet,
train_deepnet, train_deepnet_mme — same signatures, keyword names, file naming and return tuples as
the reference (training.py:23-27, 30-242, 245-287, 305-375).  The Keras calls are served by the CUDA
`Model`; RPSS by the CUDA reductions in performance_metrics.  The ELR baseline (terciled_to_ohe_xr,
train_single_bootstrap_ELR, train_elr, train_elr_mme; training.py:377-628) runs as one batched IRLS kernel (csrc/elr.cu).

Deviations (documented, not silent):
  * predictor="stacked": the reference overwrites the tuned architecture with a default Unet
    (training.py:168) and its "train" path loads a file it never saved (:228); here the requested
    architecture is used and "train" loads the checkpoint it wrote.
  * architecture "cnn"/"mlp" (never selected by any tune_*.py) raise NotImplementedError.
  * every trial's Model handle is closed as soon as it is superseded (device memory is per-handle).
  * the trials of a tuning grid are independent, so S2S_TRIAL_WORKERS (default 8) of them run concurrently on one
    GPU (one host thread + CUDA stream each; a single small U-Net fills well under half of a B200).  Each trial
    draws its weight init / shuffling from its own generator seeded by (42, bootstrap, trial), so results do not
    depend on the schedule (the reference's init stream depends on the sequential trial order and is not
    reproducible outside TensorFlow anyway)."""
from __future__ import annotations

import itertools
import os
import random

import numpy as np

os.environ["PYTHONHASHSEED"] = str(42)

import s2s_ismr_unet_b200.utils.deep_nn_models as deep_nn_models  # noqa: E402
import s2s_ismr_unet_b200.utils.performance_metrics as performance_metrics  # noqa: E402
import s2s_ismr_unet_b200.utils.preprocessing as preprocessing  # noqa: E402
from s2s_ismr_unet_b200 import model as _model  # noqa: E402
from s2s_ismr_unet_b200.keras_api import models, optimizers  # noqa: E402
from s2s_ismr_unet_b200.keras_api.callbacks import EarlyStopping, ModelCheckpoint  # noqa: E402
from s2s_ismr_unet_b200.labeled import LabeledArray  # noqa: E402

CATEGORIES = ["below", "normal", "above"]


def reset_random_seeds():
    os.environ["PYTHONHASHSEED"] = str(42)
    _model.set_seed(42)          # tf.random.set_seed(42)
    np.random.seed(42)
    random.seed(42)


def _build(architecture, architecture_params, input_shape, ct_kernel=(3, 3), n_blocks=3, filters=2, max_batch=32, rng=None):
    if architecture != "unet":
        raise NotImplementedError(f"architecture={architecture!r}: only the U-Net is on the B200 path "
                                  "(every tune_*.py passes architecture='unet')")
    if architecture_params is not None:
        ct_kernel = architecture_params["ct_kernel"]
        n_blocks = architecture_params["n_blocks"]
        filters = architecture_params["filters"]
    return deep_nn_models.Unet("", ct_kernel=ct_kernel, n_blocks=n_blocks, filters=filters, train_patches=False,
                               weighted_loss=False).build_model(input_shape, dg_train_weight_target=None, max_batch=max_batch,
                                                                rng=rng)


def _wrap(pred, dims, like: LabeledArray):
    t = dims[0]
    coords = {"category": np.array(CATEGORIES)}
    for k in (t, "Y", "X"):
        if k in like.coords:
            coords[k] = like.coords[k]
    return LabeledArray(pred, dims, coords)


def train_single_bootstrap_deepnet(i, xtrain_list, ytrain_list, xval_list, yval_list, xtest_list, ytest_list,
                                   architecture_params, tuning_grid, architecture, training_type,
                                   predictor, modname, obs, week, epochs, batch_size, learning_rate, dir):
    xtrain, ytrain = xtrain_list[i], ytrain_list[i]
    xval, yval = xval_list[i], yval_list[i]
    xtest, ytest = xtest_list[i], ytest_list[i]
    best_params_dict = {}
    reset_random_seeds()

    stacked = predictor == "stacked"
    if predictor not in ("mean", "stacked"):
        raise ValueError(f"predictor must be 'mean' or 'stacked', got {predictor!r}")
    prep = preprocessing.preprocess_stacked if stacked else preprocessing.preprocess
    (X_train, Y_train_oh, X_val, Y_val_oh, X_test, Y_test_oh,
     ytrain_terciled, yval_terciled, ytest_terciled) = prep(xtrain, ytrain, xval, yval, xtest, ytest)
    val = (lambda a: a.values) if stacked else (lambda a: a)
    tag = "stacked_" if stacked else ""
    input_shape = (X_train.shape[1], X_train.shape[2], 1)
    base = "models/" + (dir or "") + f"{modname}_{obs}/{week}/"

    if training_type == "tune":
        best_val_loss, best_model_path, best_params = float("inf"), None, None
        grid = list(itertools.product(tuning_grid["batch_sizes"], tuning_grid["learning_rates"], tuning_grid["ct_kernels"],
                                      tuning_grid["n_filters"], tuning_grid["n_blocks"]))
        patience = tuning_grid["patience"]
        from s2s_ismr_unet_b200.runtime import current_device, set_device
        trial_device = current_device()         # worker threads start on device 0 whatever the caller selected

        def run_trial(item):
            set_device(trial_device)
            trial_num, (bs, lr, ct_kernel, n_filter, n_block) = item
            print(f"Trial {trial_num + 1}/ {len(grid)}")
            print(f"Tuning Combination: Batch size={bs}, LR={lr}, Kernel={ct_kernel}, Filters={n_filter}, Blocks={n_block}")
            model = _build(architecture, None, input_shape, ct_kernel, n_block, n_filter, max_batch=max(bs, 32),
                           rng=np.random.default_rng([42, i, trial_num]))
            model.compile(optimizer=optimizers.Adam(learning_rate=lr), loss="categorical_crossentropy", metrics=["accuracy"])
            checkpoint_path = base + f"best_model_{tag}{architecture}_bootstrap_{i + 1}_trial_{trial_num + 1}.keras"
            checkpoint = ModelCheckpoint(checkpoint_path, save_best_only=True, save_weights_only=False, monitor="val_loss",
                                         mode="min", verbose=0)
            early_stopping = EarlyStopping(monitor="val_loss", patience=patience, restore_best_weights=True)
            history = model.fit(x=val(X_train), y=Y_train_oh, validation_data=(val(X_val), Y_val_oh), epochs=epochs,
                                batch_size=bs, callbacks=[checkpoint, early_stopping], shuffle=True, verbose=0)
            model.close()
            val_loss = min(history.history["val_loss"])
            print(f"Validation loss for bootstrap {i + 1}, trial {trial_num + 1}: {val_loss}")
            return val_loss, checkpoint_path, (bs, lr, ct_kernel, n_filter, n_block)

        workers = max(1, int(os.environ.get("S2S_TRIAL_WORKERS", "8")))
        if workers > 1 and len(grid) > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(min(workers, len(grid))) as ex:
                results = list(ex.map(run_trial, enumerate(grid)))
        else:
            results = [run_trial(it) for it in enumerate(grid)]
        for val_loss, checkpoint_path, params in results:      # first minimum wins, as in the sequential loop
            if val_loss < best_val_loss:
                best_val_loss, best_model_path, best_params = val_loss, checkpoint_path, params
        best_model = models.load_model(best_model_path)
        if not stacked:
            best_model.save(base + f"best_model_{architecture}_{i}_tuned.keras")
        best_params_dict[i] = {"batch_size": best_params[0], "lr": best_params[1], "ct_kernel": best_params[2],
                               "filters": best_params[3], "blocks": best_params[4], "val_loss": best_val_loss}
        print(f"Best hyperparameters for bootstrap {i + 1}: {best_params_dict[i]}")
    elif training_type == "train":
        model = _build(architecture, architecture_params, input_shape, max_batch=max(batch_size, 32))
        model.compile(optimizer=optimizers.Adam(learning_rate=learning_rate), loss="categorical_crossentropy", metrics=["accuracy"])
        path = base + f"best_model_{tag}{architecture}_{i}.keras"
        checkpoint = ModelCheckpoint(path, save_best_only=True, save_weights_only=False, monitor="val_loss", mode="min", verbose=0)
        model.fit(x=val(X_train), y=Y_train_oh, validation_data=(val(X_val), Y_val_oh), epochs=epochs, batch_size=batch_size,
                  callbacks=[checkpoint], shuffle=True, verbose=0)
        model.close()
        best_model = models.load_model(path)
    elif training_type == "load":
        try:
            best_model = models.load_model(base + f"best_model_{architecture}_{i}_tuned.keras")
        except Exception:
            best_model = models.load_model(base + f"best_model_{architecture}_{i}.keras")
    else:
        raise ValueError(f"training_type must be 'tune', 'train' or 'load', got {training_type!r}")

    predictions = best_model.predict(val(X_test), verbose=0)
    train_predictions = best_model.predict(val(X_train), verbose=0)
    val_predictions = best_model.predict(val(X_val), verbose=0)
    best_model.close()

    t = "MT" if stacked else "T"
    dims = (t, "Y", "X", "category")
    return (_wrap(train_predictions, dims, ytrain_terciled), _wrap(val_predictions, dims, yval_terciled),
            _wrap(predictions, dims, ytest_terciled), _wrap(Y_test_oh, dims, ytest_terciled),
            X_train, X_val, X_test, ytrain_terciled, yval_terciled, ytest_terciled)


def _skill(i, predictor, xtrain_list, xval_list, xtest_list, X_train, X_val, X_test, preds, terciled):
    if predictor == "mean":
        fc = [performance_metrics.climo_predict(a[i], predictor) for a in (xtrain_list, xval_list, xtest_list)]
    else:
        fc = [performance_metrics.climo_predict(a, predictor) for a in (X_train, X_val, X_test)]
    return [performance_metrics.rpss(f, p, t, predictor) for f, p, t in zip(fc, preds, terciled)]


def train_deepnet(xtrain_list, ytrain_list, xval_list, yval_list, xtest_list, ytest_list,
                  architecture_params=None, tuning_grid=None, architecture="unet",
                  training_type="train", predictor="mean", modname="GEFS", obs="IMD", week="wk3-4",
                  epochs=100, batch_size=16, learning_rate=1e-3, dir=None):
    rpss_test_list, rpss_train_list, rpss_val_list, predictions_list, y_test_oh_list = [], [], [], [], []
    for i in range(len(xtrain_list)):
        print(f"Bootstrap {i + 1}")
        (tr, va, te, Y_test_oh_xr, X_train, X_val, X_test, ytr, yva, yte) = train_single_bootstrap_deepnet(
            i, xtrain_list, ytrain_list, xval_list, yval_list, xtest_list, ytest_list, architecture_params, tuning_grid,
            architecture, training_type, predictor, modname, obs, week, epochs, batch_size, learning_rate, dir)
        predictions_list.append(te)
        y_test_oh_list.append(Y_test_oh_xr)
        r_tr, r_va, r_te = _skill(i, predictor, xtrain_list, xval_list, xtest_list, X_train, X_val, X_test, (tr, va, te), (ytr, yva, yte))
        rpss_train_list.append(r_tr), rpss_val_list.append(r_va), rpss_test_list.append(r_te)
    return rpss_train_list, rpss_val_list, rpss_test_list, predictions_list, y_test_oh_list


def _mme_mean(preds):
    """xr.concat(..., 'model').mean('model') then / sum('category') (training.py:344-350), on the GPU."""
    import ctypes as C
    from s2s_ismr_unet_b200._lib import call
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    st = performance_metrics._st()
    stack = np.ascontiguousarray(np.stack([p.values for p in preds], 0), np.float32)
    n_points = int(np.prod(stack.shape[1:-1]))
    d_in, d_out = DeviceBuffer.from_array(stack, st), DeviceBuffer(4 * n_points * 3)
    call("s2s_mme_combine", C.c_void_p(d_in.ptr), len(preds), C.c_int64(n_points), C.c_void_p(d_out.ptr), C.c_void_p(st.ptr))
    return preds[0]._like(d_out.download(stack.shape[1:], np.float32, st))


def train_deepnet_mme(xtrain_dict, ytrain_list, xval_dict, yval_list, xtest_dict, ytest_list,
                      architecture_params=None, tuning_grid=None, architecture="unet",
                      training_type="train", predictor="mean", obs="IMD", week="wk3-4",
                      epochs=100, batch_size=16, learning_rate=1e-3, dir=None):
    rpss_test_list, rpss_train_list, rpss_val_list, predictions_list, y_test_oh_list = [], [], [], [], []
    for i in range(len(ytrain_list)):
        print(f"### Bootstrap {i + 1} ####")
        tr_l, va_l, te_l = [], [], []
        for name in xtrain_dict:
            print(f"----- Model {name}")
            xtrain_list, xval_list, xtest_list = xtrain_dict[name], xval_dict[name], xtest_dict[name]
            (tr, va, te, Y_test_oh_xr, X_train, X_val, X_test, ytr, yva, yte) = train_single_bootstrap_deepnet(
                i=i, xtrain_list=xtrain_list, ytrain_list=ytrain_list, xval_list=xval_list, yval_list=yval_list,
                xtest_list=xtest_list, ytest_list=ytest_list, architecture_params=architecture_params, tuning_grid=tuning_grid,
                architecture=architecture, training_type=training_type, modname=name, predictor=predictor, obs=obs, week=week,
                epochs=epochs, batch_size=batch_size, learning_rate=learning_rate, dir=dir)
            tr_l.append(tr), va_l.append(va), te_l.append(te)
        train_preds, val_preds, test_preds = _mme_mean(tr_l), _mme_mean(va_l), _mme_mean(te_l)
        predictions_list.append(test_preds)
        y_test_oh_list.append(Y_test_oh_xr)
        r_tr, r_va, r_te = _skill(i, predictor, xtrain_list, xval_list, xtest_list, X_train, X_val, X_test,
                                  (train_preds, val_preds, test_preds), (ytr, yva, yte))
        rpss_train_list.append(r_tr), rpss_val_list.append(r_va), rpss_test_list.append(r_te)
    return rpss_train_list, rpss_val_list, rpss_test_list, predictions_list, y_test_oh_list


# ------------------------------------------------------------------ ELR baseline (training.py:377-571)
def terciled_to_ohe_xr(y):
    """One-hot (T,Y,X,category) of a terciled predictand, NaN where the label is NaN (training.py:377-398)."""
    y = preprocessing.as_labeled(y)
    lab = y.values
    oh = np.full(lab.shape + (3,), np.nan)
    ok = ~np.isnan(lab)
    oh[ok] = np.eye(3)[lab[ok].astype(int)]
    return LabeledArray(oh, y.dims + ("category",), {**{k: c for k, c in y.coords.items() if k in y.dims},
                                                     "category": np.array(CATEGORIES)})


def train_single_bootstrap_ELR(xtrain, ytrain, xtest, ytest):
    """Extended logistic regression per gridpoint (training.py:402-530) -> (train_predictions, test_predictions,
    y_train_terciled, y_test_terciled); predictions are (T,Y,X,category) float64.  The reference's Python loop over
    gridpoints around statsmodels' GLM is one batched IRLS kernel here (s2s_elr_fit_predict)."""
    import ctypes as C
    from s2s_ismr_unet_b200._lib import call
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    xtrain, ytrain = preprocessing.as_labeled(xtrain).sortby("T"), preprocessing.as_labeled(ytrain).sortby("T")
    xtest, ytest = preprocessing.as_labeled(xtest).sortby("T"), preprocessing.as_labeled(ytest).sortby("T")
    labeler = preprocessing.rolling_labeler_ELR(ytrain, window=1)
    y_train_terciled, _, _ = labeler(ytrain)
    y_test_terciled, _, _ = labeler(ytest)
    xm_tr = np.ascontiguousarray(xtrain.mean("M").values, np.float64)
    xm_te = np.ascontiguousarray(xtest.mean("M").values, np.float64)
    yv = preprocessing._as_dev_dtype(ytrain.values)
    T, Tt = len(xm_tr), len(xm_te)
    grid = xm_tr.shape[1:]
    YX = int(np.prod(grid))
    st = preprocessing._pst()
    bufs = [DeviceBuffer.from_array(a, st) for a in (xm_tr, yv, labeler.slots(ytrain).astype(np.int32), xm_te,
                                                     labeler.slots(ytest).astype(np.int32), np.ascontiguousarray(labeler.edges))]
    dptr, dpte = DeviceBuffer(8 * T * YX * 3), DeviceBuffer(8 * Tt * YX * 3)
    call("s2s_elr_fit_predict", C.c_void_p(bufs[0].ptr), C.c_void_p(bufs[1].ptr), int(yv.dtype == np.float64),
         C.c_void_p(bufs[2].ptr), C.c_void_p(bufs[3].ptr), C.c_void_p(bufs[4].ptr), C.c_void_p(bufs[5].ptr), T, Tt,
         C.c_int64(YX), C.c_void_p(dptr.ptr), C.c_void_p(dpte.ptr), None, C.c_void_p(st.ptr))
    p_tr = dptr.download((T,) + tuple(grid) + (3,), np.float64, st)
    p_te = dpte.download((Tt,) + tuple(grid) + (3,), np.float64, st)
    for b in bufs + [dptr, dpte]:
        b.free()
    cat = {"category": np.array(CATEGORIES)}
    dims = ("T", "Y", "X", "category")
    return (LabeledArray(p_tr, dims, cat), LabeledArray(p_te, dims, cat), y_train_terciled, y_test_terciled)


def train_elr(xtrain_list_elr, ytrain_list_elr, xtest_list_elr, ytest_list_elr):
    """ELR over the bootstraps (training.py:533-571) -> rpss_train_list, rpss_test_list, predictions_list, y_test_oh_list."""
    rpss_test_list, rpss_train_list, predictions_list, y_test_oh_list = [], [], [], []
    for i in range(len(xtrain_list_elr)):
        xtrain, ytrain, xtest, ytest = xtrain_list_elr[i], ytrain_list_elr[i], xtest_list_elr[i], ytest_list_elr[i]
        p_train, p_test, y_train_terciled, y_test_terciled = train_single_bootstrap_ELR(xtrain, ytrain, xtest, ytest)
        predictions_list.append(p_test)
        y_test_oh_list.append(terciled_to_ohe_xr(y_test_terciled))
        fcast_test = performance_metrics.climo_predict(xtest)
        fcast_train = performance_metrics.climo_predict(xtrain)
        for p, yt in ((p_train, y_train_terciled), (p_test, y_test_terciled)):     # predictions carry the starts' coords
            p.coords.update({k: c for k, c in yt.coords.items() if k in ("T", "Y", "X")})
        rpss_train_list.append(performance_metrics.rpss(fcast_train, p_train, y_train_terciled))
        rpss_test_list.append(performance_metrics.rpss(fcast_test, p_test, y_test_terciled))
    return rpss_train_list, rpss_test_list, predictions_list, y_test_oh_list


def train_elr_mme(xtrain_dict_elr, ytrain_list_elr, xtest_dict_elr, ytest_list_elr):
    """Multi-model ELR (training.py:575-628): one ELR per model and bootstrap, probabilities averaged over the models
    and renormalised over the categories (the same s2s_mme_combine kernel as the U-Net MME), then RPSS."""
    rpss_test_list, rpss_train_list, predictions_list, y_test_oh_list = [], [], [], []
    for i in range(len(ytrain_list_elr)):
        train_preds_list, test_preds_list = [], []
        for name, xtrain_list in xtrain_dict_elr.items():
            xtrain, xtest = xtrain_list[i], xtest_dict_elr[name][i]
            p_train, p_test, y_train_terciled, y_test_terciled = train_single_bootstrap_ELR(xtrain, ytrain_list_elr[i], xtest,
                                                                                            ytest_list_elr[i])
            train_preds_list.append(p_train)
            test_preds_list.append(p_test)
        train_preds, test_preds = _mme_mean(train_preds_list), _mme_mean(test_preds_list)
        for p, yt in ((train_preds, y_train_terciled), (test_preds, y_test_terciled)):
            p.coords.update({k: c for k, c in yt.coords.items() if k in ("T", "Y", "X")})
        predictions_list.append(test_preds)
        y_test_oh_list.append(terciled_to_ohe_xr(y_test_terciled))
        rpss_train_list.append(performance_metrics.rpss(performance_metrics.climo_predict(xtrain), train_preds, y_train_terciled))
        rpss_test_list.append(performance_metrics.rpss(performance_metrics.climo_predict(xtest), test_preds, y_test_terciled))
    return rpss_train_list, rpss_test_list, predictions_list, y_test_oh_list
