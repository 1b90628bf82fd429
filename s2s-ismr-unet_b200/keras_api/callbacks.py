"""keras.callbacks.ModelCheckpoint / EarlyStopping with the Keras-3 semantics the reference relies
on (utils/training.py:69,98-100): best weights are tracked in a HOST-side copy (one download per improving
epoch, shared by both callbacks) and the checkpoint file is written once, when fit() ends — also when fit() ends
with an exception (Model.fit runs on_train_end in a `finally`), so a best-so-far checkpoint is never lost.  The
observable state after fit is identical; the per-epoch zip write that dominates the reference's wall time is gone.
Files keep the reference's `.keras` names and are Keras-3 archives (zip of config.json + metadata.json +
model.weights.h5, written by keras_api/keras_archive.py) that also carry this library's own fast-load members."""
from __future__ import annotations

import numpy as np


class Callback:
    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass


def _weights_of(model):
    """One shared download per epoch when the model offers it (Model._snapshot_weights), else a plain get_weights()."""
    snap = getattr(model, "_snapshot_weights", None)
    return snap() if snap is not None else model.get_weights()


def _monitor_op(mode, monitor):
    if mode == "max" or (mode == "auto" and ("acc" in monitor or monitor.startswith("fmeasure"))):
        return np.greater, -np.inf
    return np.less, np.inf


class ModelCheckpoint(Callback):
    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", save_freq="epoch", **_ignored):
        self.filepath = str(filepath)
        self.monitor, self.verbose = monitor, verbose
        self.save_best_only, self.save_weights_only = save_best_only, save_weights_only
        self.op, self.best = _monitor_op(mode, monitor)
        self._snapshot = None

    def on_train_begin(self, logs=None):
        self._snapshot = None

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if self.save_best_only:
            if cur is None or not self.op(cur, self.best):
                return
            self.best = cur
        m = self.model
        self._snapshot = (_weights_of(m), m._get_opt_state() if m.optimizer is not None else None)

    def on_train_end(self, logs=None):
        if self._snapshot is None:
            return
        from ..model import save_weights_file
        import os
        os.makedirs(os.path.dirname(os.path.abspath(self.filepath)), exist_ok=True)
        w, opt = self._snapshot
        save_weights_file(self.filepath, self.model.config, w, opt, self.model.optimizer)


class EarlyStopping(Callback):
    """Keras-3 EarlyStopping: `wait` counts epochs since the best value; stop when wait >= patience;
    with restore_best_weights the best weights are restored at train end (also without an early stop)."""

    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", baseline=None,
                 restore_best_weights=False, start_from_epoch=0):
        self.monitor, self.patience, self.verbose = monitor, int(patience), verbose
        self.min_delta = abs(float(min_delta))
        self.baseline, self.restore_best_weights, self.start_from_epoch = baseline, restore_best_weights, start_from_epoch
        self.op, self._init_best = _monitor_op(mode, monitor)
        if self.op is np.less:
            self.min_delta *= -1
        self.wait, self.stopped_epoch, self.best, self.best_weights, self.best_epoch = 0, 0, self._init_best, None, 0

    def on_train_begin(self, logs=None):
        self.wait, self.stopped_epoch, self.best, self.best_weights, self.best_epoch = 0, 0, self._init_best, None, 0

    def _improved(self, cur, ref):
        return self.op(cur - self.min_delta, ref)

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None or epoch < self.start_from_epoch:
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = _weights_of(self.model)
            self.best_epoch = epoch
        self.wait += 1
        if self._improved(cur, self.best):
            self.best, self.best_epoch = cur, epoch
            if self.restore_best_weights:
                self.best_weights = _weights_of(self.model)
            if self.baseline is None or self._improved(cur, self.baseline):
                self.wait = 0
            return
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True

    def on_train_end(self, logs=None):
        if self.restore_best_weights and self.best_weights is not None:
            self.model.set_weights(self.best_weights)
