"""`Model`: the object returned by Unet(...).build_model(...) — quacks like the Keras `Model` for
exactly the calls utils/training.py makes (training.py:66-67,95-106,114-115,133-135):

    model.compile(optimizer=Adam(lr), loss="categorical_crossentropy", metrics=['accuracy'])
    history = model.fit(x=, y=, validation_data=(xv, yv), epochs=, batch_size=, callbacks=[...],
                        shuffle=True, verbose=0);  history.history['val_loss']
    model.predict(X, verbose=0) -> np.ndarray (N,H,W,3) float32
    model.save(path);  load_model(path)

All arithmetic runs in libs2s_unet.so (CUDA, sm_100a) through the C ABI in include/s2s_unet.h.
The data set is uploaded once and whole epochs are enqueued on the device (s2s_unet_fit_epoch);
the host only reads the epoch accumulators to drive the callbacks.
"""
from __future__ import annotations

import ctypes as C
import io
import json
import math
import os
import weakref
import zipfile

import numpy as np

from . import _lib
from ._lib import AdamCfg, TensorDesc, UnetCfg, call
from .runtime import DeviceBuffer, PinnedArray, Stream, current_device, d2d, d2h, h2d, is_pinned, set_device

POOL_AVG, POOL_MAX = 0, 1
HEAD_SOFTMAX3, HEAD_RELU1 = 0, 1
LOSS_CCE, LOSS_MASKED_MSE = 0, 1
PRECISIONS = {"fp32": 0, "bf16_tc": 1, "tf32": 2}      # s2s_precision
ACTIVATIONS = {"elu": 0, "relu": 1}                    # s2s_act_kind

# ------------------------------------------------------------------ global seed (tf.random.set_seed)
_seed_state = {"rng": np.random.default_rng(0)}


def set_seed(seed: int) -> None:
    """tf.random.set_seed equivalent (training.py:25): re-seeds weight init and fit shuffling.
    TF's RNG stream itself cannot be reproduced; the stream here is numpy PCG64."""
    _seed_state["rng"] = np.random.default_rng(int(seed))


def _rng() -> np.random.Generator:
    return _seed_state["rng"]


class History:
    def __init__(self):
        self.history = {"loss": [], "accuracy": [], "val_loss": [], "val_accuracy": []}
        self.epoch = []


class Model:
    def __init__(self, input_shape, filters=2, n_blocks=3, ct_kernel=3, apool=True, bn=True, output="proba",
                 max_batch=32, device=None, weights=None, precision="fp32", rng=None, activation="elu"):
        H, W, Cin = (int(v) for v in input_shape)
        if isinstance(ct_kernel, (tuple, list)):
            if ct_kernel[0] != ct_kernel[1]:
                raise ValueError("ct_kernel must be square")
            ct_kernel = int(ct_kernel[0])
        if output not in ("proba", "deterministic"):
            raise ValueError(f"output must be 'proba' or 'deterministic', got {output!r}")
        if precision not in PRECISIONS:
            raise ValueError("precision must be 'fp32' (parity path), 'tf32' (tcgen05 tf32 forward + input-gradient convolutions, "
                             "training and inference) or 'bf16_tc' (tcgen05 bf16 inference of the thick layers)")
        if activation not in ACTIVATIONS:
            raise ValueError(f"activation must be 'elu' (the reference's setting, deep_nn_models.py:139) or 'relu', got {activation!r}")
        if not 3 <= int(n_blocks) <= 5:
            # the reference graph always builds three down / up blocks (deep_nn_models.py:82-84, 97-99)
            raise ValueError(f"n_blocks must be 3, 4 or 5 (got {n_blocks})")
        self.precision = precision
        self._own_rng = rng          # per-model generator (weight init + fit shuffling); None = the global seeded stream
        div = 2 ** int(n_blocks)
        if H % div or W % div:
            # Keras raises on the Concatenate shape mismatch (comment at tune_ECMWF_com.py:26)
            raise ValueError(f"input {H}x{W} is not divisible by 2^n_blocks={div}: "
                             "the skip concatenation shapes would not match")
        if device is not None:
            set_device(device)
        self.device = current_device()      # every later call re-binds the calling thread to it (see _bind)
        self.config = dict(input_shape=[H, W, Cin], filters=int(filters), n_blocks=int(n_blocks), ct_kernel=int(ct_kernel),
                           apool=bool(apool), bn=bool(bn), output=output, activation=activation)
        self.H, self.W, self.Cin = H, W, Cin
        self.NC = 3 if output == "proba" else 1
        self.max_batch = int(max_batch)
        self.stream = Stream()
        self._h = None
        self._create()
        self.optimizer = None
        self.loss = None
        self.stop_training = False
        self._pin_x = None
        self._pin_y = None
        self._pin_small = PinnedArray((8,), np.float64)
        if weights is None:
            weights = self._glorot_init()
        self.set_weights(weights)

    def _bind(self):
        """cudaSetDevice is per-thread: a model used from a worker thread (trial pool, utils/training.py) or after another
        model selected a different GPU must re-select ITS device before touching its streams / pools."""
        set_device(self.device)

    # ---------------------------------------------------------------- handle life cycle
    def _create(self):
        c = self.config
        cfg = UnetCfg(self.H, self.W, self.Cin, c["filters"], c["n_blocks"], c["ct_kernel"],
                      POOL_AVG if c["apool"] else POOL_MAX, int(c["bn"]),
                      HEAD_SOFTMAX3 if c["output"] == "proba" else HEAD_RELU1, self.max_batch, 1e-3, 0.99,
                      PRECISIONS[self.precision], ACTIVATIONS[c.get("activation", "elu")])
        h = C.c_void_p()
        call("s2s_unet_create", C.byref(cfg), C.byref(h))
        self._h = h
        self._fin = weakref.finalize(self, _lib.load().s2s_unet_destroy, h)
        n = C.c_int(0)
        call("s2s_unet_param_layout", self._h, None, C.byref(n))
        arr = (TensorDesc * n.value)()
        call("s2s_unet_param_layout", self._h, arr, C.byref(n))
        self.layout = [dict(name=d.name.decode(), arena=d.arena, shape=tuple(d.shape[:d.ndim]), offset=d.offset, count=d.count)
                       for d in arr]
        p, cnt = C.c_void_p(), C.c_size_t()
        call("s2s_unet_params", self._h, C.byref(p), C.byref(cnt))
        self._params_ptr, self.n_params_padded = p.value, cnt.value
        call("s2s_unet_state", self._h, C.byref(p), C.byref(cnt))
        self._state_ptr, self.n_state_padded = p.value, cnt.value
        call("s2s_unet_grad_arena", self._h, C.byref(p), C.byref(cnt))
        self._grads_ptr = p.value
        x, y = C.c_void_p(), C.c_void_p()
        call("s2s_unet_io_buffers", self._h, C.byref(x), C.byref(y))
        self._xin_ptr, self._yin_ptr = x.value, y.value
        s, sa = C.c_void_p(), C.c_void_p()
        call("s2s_unet_stats_buffers", self._h, C.byref(s), C.byref(sa))
        self._stats_ptr, self._stats_acc_ptr = s.value, sa.value

    def close(self):
        """Free the device handle now (hundreds of models are created sequentially, training.py:87-93)."""
        if self._h is not None:
            self.stream.synchronize()        # the pool may be re-used by the next handle: drain our stream first
            self._fin()
            self._h = None

    def _ensure_batch(self, n: int):
        """Re-create the handle with a larger workspace, keeping weights / optimiser state."""
        if n <= self.max_batch:
            return
        w = self.get_weights()
        opt = self._get_opt_state() if self.optimizer is not None else None
        self.close()
        self.max_batch = int(n)
        self._create()
        self.set_weights(w)
        if self.optimizer is not None:
            self._compile_native()
            self._set_opt_state(opt)

    @property
    def sp(self):
        return C.c_void_p(self.stream.ptr)

    # ---------------------------------------------------------------- weights
    def count_params(self) -> int:
        return int(sum(d["count"] for d in self.layout))

    def _glorot_init(self):
        rng = self._own_rng if self._own_rng is not None else _rng()
        w = {}
        for d in self.layout:
            name, shape = d["name"], d["shape"]
            if name.endswith("/kernel"):
                rf = shape[0] * shape[1]
                limit = math.sqrt(6.0 / (rf * shape[2] + rf * shape[3]))
                w[name] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
            elif name.endswith("/gamma") or name.endswith("/moving_variance"):
                w[name] = np.ones(shape, np.float32)
            else:
                w[name] = np.zeros(shape, np.float32)
        return w

    def set_weights(self, weights) -> None:
        """weights: {tensor name: array} (Keras kernel layouts) or a list in layout order."""
        self._bind()
        if not isinstance(weights, dict):
            weights = {d["name"]: w for d, w in zip(self.layout, weights)}
        flat = [np.zeros(self.n_params_padded, np.float32), np.zeros(self.n_state_padded, np.float32)]
        for d in self.layout:
            a = np.asarray(weights[d["name"]], np.float32)
            if a.size != d["count"]:
                raise ValueError(f"{d['name']}: expected {d['shape']}, got {a.shape}")
            flat[d["arena"]][d["offset"]:d["offset"] + d["count"]] = a.ravel()
        h2d(self._params_ptr, flat[0].ctypes.data, flat[0].nbytes, self.stream)
        if self.n_state_padded:
            h2d(self._state_ptr, flat[1].ctypes.data, flat[1].nbytes, self.stream)
        self.stream.synchronize()

    def _download(self, ptr, n) -> np.ndarray:
        out = np.empty(n, np.float32)
        d2h(out.ctypes.data, ptr, out.nbytes, self.stream)
        self.stream.synchronize()
        return out

    def get_weights(self, as_dict=True):
        self._bind()
        flat = [self._download(self._params_ptr, self.n_params_padded),
                self._download(self._state_ptr, self.n_state_padded) if self.n_state_padded else np.zeros(0, np.float32)]
        out = {d["name"]: flat[d["arena"]][d["offset"]:d["offset"] + d["count"]].reshape(d["shape"]).copy() for d in self.layout}
        return out if as_dict else [out[d["name"]] for d in self.layout]

    def get_gradients(self):
        g = self._download(self._grads_ptr, self.n_params_padded)
        return {d["name"]: g[d["offset"]:d["offset"] + d["count"]].reshape(d["shape"]).copy() for d in self.layout if d["arena"] == 0}

    def _get_opt_state(self):
        m, v, st = C.c_void_p(), C.c_void_p(), C.c_void_p()
        call("s2s_unet_opt_state", self._h, C.byref(m), C.byref(v), C.byref(st))
        step = np.empty(1, np.int64)
        d2h(step.ctypes.data, st.value, 8, self.stream)
        return dict(m=self._download(m.value, self.n_params_padded), v=self._download(v.value, self.n_params_padded), step=int(step[0]))

    def _set_opt_state(self, s):
        m, v, st = C.c_void_p(), C.c_void_p(), C.c_void_p()
        call("s2s_unet_opt_state", self._h, C.byref(m), C.byref(v), C.byref(st))
        h2d(m.value, s["m"].ctypes.data, s["m"].nbytes, self.stream)
        h2d(v.value, s["v"].ctypes.data, s["v"].nbytes, self.stream)
        step = np.array([s["step"]], np.int64)
        h2d(st.value, step.ctypes.data, 8, self.stream)
        self.stream.synchronize()

    # ---------------------------------------------------------------- compile / step-level API
    def compile(self, optimizer=None, loss="categorical_crossentropy", metrics=None):
        self._bind()
        from .keras_api import optimizers
        if optimizer is None or optimizer == "adam":
            optimizer = optimizers.Adam()
        if not isinstance(optimizer, optimizers.Adam):
            raise ValueError("only keras.optimizers.Adam is supported on the B200 path (training.py:66)")
        if loss in ("categorical_crossentropy",):
            self._loss_kind = LOSS_CCE
        elif loss in ("mse", "mean_squared_error", "masked_mse"):
            self._loss_kind = LOSS_MASKED_MSE
        else:
            raise ValueError(f"unsupported loss {loss!r}")
        self.optimizer, self.loss, self.metrics = optimizer, loss, metrics
        self._compile_native()

    def _compile_native(self):
        o = self.optimizer
        cfg = AdamCfg(o.learning_rate, o.beta_1, o.beta_2, o.epsilon)
        call("s2s_unet_compile", self._h, C.byref(cfg), self._loss_kind)

    def set_graphs(self, enable: bool) -> None:
        call("s2s_unet_set_graphs", self._h, int(bool(enable)))

    def launch_count(self) -> int:
        n = C.c_int64(0)
        call("s2s_unet_launch_count", self._h, C.byref(n))
        return n.value

    def _prep_x(self, x) -> np.ndarray:
        x = np.asarray(x, np.float32)
        if x.ndim == 3:
            x = x[..., None]            # Keras auto-expands (T,Y,X) -> (T,Y,X,1)
        if x.shape[1:] != (self.H, self.W, self.Cin):
            raise ValueError(f"expected input (N,{self.H},{self.W},{self.Cin}), got {x.shape}")
        return np.ascontiguousarray(x)

    def _prep_y(self, y) -> np.ndarray:
        y = np.asarray(y, np.float32)
        if self.NC == 1 and y.ndim == 3:
            y = y[..., None]
        if y.shape[1:] != (self.H, self.W, self.NC):
            raise ValueError(f"expected target (N,{self.H},{self.W},{self.NC}), got {y.shape}")
        return np.ascontiguousarray(y)

    def _pinned(self, n):
        if self._pin_x is None or self._pin_x.shape[0] < n:
            self._pin_x = PinnedArray((n, self.H, self.W, self.Cin), np.float32)
            self._pin_y = PinnedArray((n, self.H, self.W, self.NC), np.float32)
        return self._pin_x, self._pin_y

    def _read_stats(self):
        d2h(self._pin_small.ptr, self._stats_ptr, 8, self.stream)
        self.stream.synchronize()
        f = np.frombuffer(self._pin_small.array.tobytes()[:8], np.float32)
        return float(f[0]), float(f[1])

    def _upload_batch(self, x, y):
        """H2D of one host batch into the handle's staging buffers.  Arrays that already live in pinned memory
        (runtime.pinned_empty) are copied directly; pageable arrays go through a pinned staging copy first."""
        n = len(x)
        if is_pinned(x) and is_pinned(y):
            h2d(self._xin_ptr, x.ctypes.data, x.nbytes, self.stream)
            h2d(self._yin_ptr, y.ctypes.data, y.nbytes, self.stream)
            return
        px, py = self._pinned(n)
        px.array[:n] = x
        py.array[:n] = y
        h2d(self._xin_ptr, px.ptr, x.nbytes, self.stream)
        h2d(self._yin_ptr, py.ptr, y.nbytes, self.stream)

    def train_on_batch(self, x, y, mask_ptr=None):
        """One optimiser step from HOST arrays: (pinned staging ->) H2D -> fused step -> D2H of the loss."""
        self._bind()
        x, y = self._prep_x(x), self._prep_y(y)
        n = len(x)
        self._ensure_batch(n)
        if mask_ptr is None and is_pinned(x) and is_pinned(y):
            # one C call: H2D x, y -> step -> D2H {loss, accuracy} -> sync (s2s_unet_train_step_host)
            call("s2s_unet_train_step_host", self._h, C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), n,
                 C.c_void_p(self._pin_small.ptr), self.sp)
            f = self._pin_small.array.view(np.float32)
            return float(f[0]), float(f[1])
        self._upload_batch(x, y)
        call("s2s_unet_train_step", self._h, C.c_void_p(self._xin_ptr), C.c_void_p(self._yin_ptr),
             C.c_void_p(mask_ptr) if mask_ptr else None, n, None, self.sp)
        return self._read_stats()

    def train_on_batches(self, xs, ys, n_global: int = 0):
        """A stream of optimiser steps from lists of equally sized host batches: one C call (s2s_unet_train_steps_host).  Every step
        copies its own batch H2D and returns its own (loss, accuracy) D2H; the copy of batch i + 1 overlaps step i.  Batches in
        pinned memory (runtime.pinned_empty) are copied directly; ordinary (pageable) NumPy arrays — what the reference hands to
        model.fit — are staged through a ring of pinned slots by the calling thread while the GPU computes.  Returns [len(xs), 2]."""
        self._bind()
        n_steps = len(xs)
        if n_steps == 0:
            return np.zeros((0, 2), np.float32)
        if len(ys) != n_steps:
            raise ValueError("xs and ys must have the same length")
        n = len(xs[0])
        # host prelude kept to ~1 us per array (it is inside the caller's clock): arrays that already are C-contiguous float32 of
        # the right shape are passed by address as they are, anything else goes through the usual conversion; `keep` holds the
        # converted copies alive until the call returns
        tx, ty = (self.H, self.W, self.Cin), (self.H, self.W, self.NC)
        keep, ax, ay = [], [], []
        for src, tail, prep, dst in ((xs, tx, self._prep_x, ax), (ys, ty, self._prep_y, ay)):
            for a in src:
                if type(a) is not np.ndarray or a.dtype != np.float32 or a.shape[1:] != tail or not a.flags.c_contiguous:
                    a = prep(a)
                    keep.append(a)
                if a.shape[0] != n:
                    raise ValueError("train_on_batches needs equally sized batches")
                dst.append(a.__array_interface__["data"][0])
        self._ensure_batch(n)
        px = (C.c_void_p * n_steps)(*ax)
        py = (C.c_void_p * n_steps)(*ay)
        out = np.zeros((n_steps, 2), np.float32)
        if n_global:      # data-parallel shards (a communicator must be attached): the GLOBAL (loss, accuracy) per step
            call("s2s_unet_dp_train_steps_host", self._h, px, py, n_steps, n, int(n_global), C.c_void_p(out.ctypes.data), self.sp)
        else:
            call("s2s_unet_train_steps_host", self._h, px, py, n_steps, n, C.c_void_p(out.ctypes.data), self.sp)
        return out

    def backward_on_batch(self, x, y, grad_scale=1.0, mask_ptr=None):
        """fwd + loss + bwd only (dense grads stay in the grad arena for an all-reduce)."""
        self._bind()
        x, y = self._prep_x(x), self._prep_y(y)
        n = len(x)
        self._ensure_batch(n)
        self._upload_batch(x, y)
        call("s2s_unet_backward_only", self._h, C.c_void_p(self._xin_ptr), C.c_void_p(self._yin_ptr),
             C.c_void_p(mask_ptr) if mask_ptr else None, n, C.c_float(grad_scale), None, self.sp)
        return self._read_stats()

    def dp_train_on_batch(self, x, y, n_global: int):
        """One data-parallel optimiser step on this rank's shard of a global batch of n_global samples
        (a communicator must be attached, parallel.PeerDataParallelTrainer): returns the GLOBAL (loss, accuracy)."""
        self._bind()
        x, y = self._prep_x(x), self._prep_y(y)
        n = len(x)
        self._ensure_batch(n)
        if is_pinned(x) and is_pinned(y):
            call("s2s_unet_dp_train_step_host", self._h, C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), n, int(n_global),
                 C.c_void_p(self._pin_small.ptr), self.sp)
            f = self._pin_small.array.view(np.float32)
            return float(f[0]), float(f[1])
        self._upload_batch(x, y)
        call("s2s_unet_dp_train_step", self._h, C.c_void_p(self._xin_ptr), C.c_void_p(self._yin_ptr), n, int(n_global),
             C.c_void_p(self._stats_ptr), self.sp)
        return self._read_stats()

    def apply_adam(self):
        call("s2s_unet_apply_adam", self._h, self.sp)

    def test_on_batch(self, x, y, mask_ptr=None):
        self._bind()
        x, y = self._prep_x(x), self._prep_y(y)
        n = len(x)
        self._ensure_batch(n)
        dx, dy = DeviceBuffer.from_array(x, self.stream), DeviceBuffer.from_array(y, self.stream)
        call("s2s_unet_eval_batch", self._h, C.c_void_p(dx.ptr), C.c_void_p(dy.ptr), C.c_void_p(mask_ptr) if mask_ptr else None,
             n, None, self.sp)
        return self._read_stats()

    # ---------------------------------------------------------------- fit / evaluate / predict
    def _epoch_stats(self):
        d2h(self._pin_small.ptr, self._stats_acc_ptr, 24, self.stream)
        self.stream.synchronize()
        a = self._pin_small.array
        loss_sum, correct, npix = float(a[0]), float(a[1]), float(a[2])
        call("s2s_unet_reset_epoch_stats", self._h, self.sp)
        if npix == 0:
            return float("nan"), float("nan")
        return loss_sum / npix, correct / npix

    def _snapshot_weights(self):
        """Weights of the epoch that just ended, downloaded ONCE and shared by the callbacks (ModelCheckpoint and
        EarlyStopping both keep the best epoch's weights); treat the returned dict as read-only."""
        ep = getattr(self, "_snap_epoch", None)
        if getattr(self, "_snap", None) is None or self._snap[0] != ep:
            self._snap = (ep, self.get_weights())
        return self._snap[1]

    def fit(self, x=None, y=None, validation_data=None, epochs=1, batch_size=32, callbacks=None, shuffle=True,
            verbose=0, mask=None, _orders=None):
        """Keras fit protocol (training.py:102-103): per-epoch reshuffle, partial last batch kept,
        validation pass in inference mode after every epoch, callbacks on epoch end.
        `_orders` (list of index arrays, one per epoch) injects the sample order for parity tests."""
        self._bind()
        if self.optimizer is None:
            raise RuntimeError("call compile() before fit()")
        x, y = self._prep_x(x), self._prep_y(y)
        T = len(x)
        bs = int(batch_size) if batch_size else 32
        self._ensure_batch(bs)
        callbacks = list(callbacks or [])
        st = self.stream
        dx, dy = DeviceBuffer.from_array(x, st), DeviceBuffer.from_array(y, st)
        dmask = DeviceBuffer.from_array(np.asarray(mask, np.uint8), st) if mask is not None else None
        mptr = C.c_void_p(dmask.ptr) if dmask is not None else None
        have_val = validation_data is not None
        if have_val:
            xv, yv = self._prep_x(validation_data[0]), self._prep_y(validation_data[1])
            dxv, dyv = DeviceBuffer.from_array(xv, st), DeviceBuffer.from_array(yv, st)
        dperm = DeviceBuffer(4 * T)
        perm_pin = PinnedArray((T,), np.int32)
        hist = History()
        self.stop_training = False
        for cb in callbacks:
            cb.set_model(self)
            cb.on_train_begin()
        call("s2s_unet_reset_epoch_stats", self._h, self.sp)
        rng = self._own_rng if self._own_rng is not None else _rng()
        self._snap = None               # (epoch, weights): one download per improving epoch, shared by all callbacks
        try:
            for ep in range(int(epochs)):
                if _orders is not None:
                    order = np.asarray(_orders[ep], np.int32)
                elif shuffle:
                    order = rng.permutation(T).astype(np.int32)
                else:
                    order = np.arange(T, dtype=np.int32)
                n_ep = len(order)
                perm_pin.array[:n_ep] = order
                h2d(dperm.ptr, perm_pin.ptr, 4 * n_ep, st)
                call("s2s_unet_fit_epoch", self._h, C.c_void_p(dx.ptr), C.c_void_p(dy.ptr), C.c_void_p(dperm.ptr), n_ep, bs, mptr, self.sp)
                loss, acc = self._epoch_stats()
                logs = {"loss": loss, "accuracy": acc}
                if have_val:
                    call("s2s_unet_eval_dataset", self._h, C.c_void_p(dxv.ptr), C.c_void_p(dyv.ptr), len(xv), bs, mptr, self.sp)
                    vl, va = self._epoch_stats()
                    logs.update(val_loss=vl, val_accuracy=va)
                for k, v in logs.items():
                    hist.history[k].append(v)
                hist.epoch.append(ep)
                if verbose:
                    print(f"Epoch {ep + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
                self._snap_epoch = ep
                for cb in callbacks:
                    cb.on_epoch_end(ep, logs)
                if self.stop_training:
                    break
        finally:
            # on_train_end also runs when an epoch raised (exception, KeyboardInterrupt, device fault): ModelCheckpoint then
            # still writes the best-so-far snapshot it holds in HOST memory, as the reference has one on disk after every
            # improving epoch (training.py:98).  A failing callback must not mask the original error.
            import sys as _sys
            failing = _sys.exc_info()[0] is not None
            for cb in callbacks:
                try:
                    cb.on_train_end()
                except Exception:
                    if not failing:
                        raise
            self._snap = None
        if not have_val:
            hist.history.pop("val_loss"), hist.history.pop("val_accuracy")
        for b in (dx, dy, dperm):
            b.free()
        self.history = hist
        return hist

    def evaluate(self, x, y, batch_size=32, verbose=0, mask=None):
        self._bind()
        x, y = self._prep_x(x), self._prep_y(y)
        self._ensure_batch(batch_size)
        st = self.stream
        dx, dy = DeviceBuffer.from_array(x, st), DeviceBuffer.from_array(y, st)
        dmask = DeviceBuffer.from_array(np.asarray(mask, np.uint8), st) if mask is not None else None
        call("s2s_unet_reset_epoch_stats", self._h, self.sp)
        call("s2s_unet_eval_dataset", self._h, C.c_void_p(dx.ptr), C.c_void_p(dy.ptr), len(x), int(batch_size),
             C.c_void_p(dmask.ptr) if dmask is not None else None, self.sp)
        loss, acc = self._epoch_stats()
        return [loss, acc]

    def predict(self, x, batch_size=32, verbose=0):
        """Inference forward (BN moving statistics), Keras default batch 32 (training.py:133-135)."""
        self._bind()
        x = self._prep_x(x)
        T = len(x)
        if T == 0:
            return np.zeros((0, self.H, self.W, self.NC), np.float32)
        bs = int(batch_size) if batch_size else 32
        self._ensure_batch(min(bs, max(T, 1)))
        st = self.stream
        dx = DeviceBuffer.from_array(x, st)
        dp = DeviceBuffer(4 * T * self.H * self.W * self.NC)
        call("s2s_unet_predict_dataset", self._h, C.c_void_p(dx.ptr), T, min(bs, self.max_batch), C.c_void_p(dp.ptr), self.sp)
        out = dp.download((T, self.H, self.W, self.NC), np.float32, st)
        dx.free(), dp.free()
        return out

    def __call__(self, x, training=False):
        x = self._prep_x(x)
        n = len(x)
        self._ensure_batch(n)
        dx = DeviceBuffer.from_array(x, self.stream)
        dp = DeviceBuffer(4 * n * self.H * self.W * self.NC)
        call("s2s_unet_forward", self._h, C.c_void_p(dx.ptr), n, C.c_void_p(dp.ptr), int(bool(training)), self.sp)
        return dp.download((n, self.H, self.W, self.NC), np.float32, self.stream)

    def activation(self, layer_name: str) -> np.ndarray:
        """Output of a named Keras layer for the last forward batch (N = last batch size)."""
        p, h, w, c, ld = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        call("s2s_unet_activation", self._h, layer_name.encode(), C.byref(p), C.byref(h), C.byref(w), C.byref(c), C.byref(ld))
        return p.value, h.value, w.value, c.value, ld.value

    def gradcam(self, x, layer_name="bottleneck", cls=2, batch_size=32):
        """Grad-CAM map (N,Hl,Wl) for class `cls` at the named layer (north_star item d)."""
        self._bind()
        x = self._prep_x(x)
        T = len(x)
        bs = min(int(batch_size), T)
        self._ensure_batch(bs)
        _, hl, wl, _, _ = self.activation(layer_name)
        dx = DeviceBuffer.from_array(x, self.stream)
        dc = DeviceBuffer(4 * T * hl * wl)
        xrow = 4 * self.H * self.W * self.Cin
        for i in range(0, T, bs):
            n = min(bs, T - i)
            call("s2s_unet_gradcam", self._h, C.c_void_p(dx.ptr + i * xrow), n, layer_name.encode(), int(cls),
                 C.c_void_p(dc.ptr + 4 * i * hl * wl), self.sp)
        return dc.download((T, hl, wl), np.float32, self.stream)

    # ---------------------------------------------------------------- persistence
    def save(self, path, include_optimizer=True):
        """model.save(path) (training.py:115): a Keras-3 `.keras` archive (zip of config.json + metadata.json +
        model.weights.h5 in Keras' layer / variable path layout, keras_api/keras_archive.py)."""
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        save_weights_file(path, self.config, self.get_weights(),
                          self._get_opt_state() if (include_optimizer and self.optimizer is not None) else None,
                          self.optimizer)


def _trainable_layout(weights: dict) -> list:
    """Offsets of the trainable tensors inside the flat (4-float aligned) parameter arena, from the weight dict's order
    (= Keras layer-creation order, the order s2s_unet_param_layout reports)."""
    out, off = [], 0
    for name, w in weights.items():
        if name.rsplit("/", 1)[-1] in ("kernel", "bias", "gamma", "beta"):
            n = int(np.asarray(w).size)
            out.append(dict(name=name, arena=0, shape=tuple(np.asarray(w).shape), offset=off, count=n))
            off += (n + 3) // 4 * 4
    return out


def save_weights_file(path, config, weights, opt_state=None, optimizer=None):
    """Keras-3 archive + one extra member (s2s_unet_b200.json: this library's own config, e.g. the precision mode)."""
    from .keras_api.keras_archive import write_keras_archive
    meta = dict(format="s2s-unet-b200/2", config=config)
    if optimizer is not None:
        meta["optimizer"] = dict(learning_rate=optimizer.learning_rate, beta_1=optimizer.beta_1, beta_2=optimizer.beta_2,
                                 epsilon=optimizer.epsilon)
    write_keras_archive(path, config, weights, opt_state, optimizer, _trainable_layout(weights),
                        extra_members={"s2s_unet_b200.json": json.dumps(meta)})


def load_model(path, device=None) -> Model:
    """keras.models.load_model equivalent (training.py:114,128-131): reads Keras-3 `.keras` archives of the reference's
    U-Net — written by Model.save here or by real Keras — and this library's round-1 container (config.json with a
    'format' key + weights.npz)."""
    from .keras_api import optimizers
    from .keras_api.keras_archive import read_keras_archive
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        head = json.loads(z.read("config.json")) if "config.json" in names else {}
        if str(head.get("format", "")).startswith("s2s-unet-b200/1"):          # round-1 container
            wz = np.load(io.BytesIO(z.read("weights.npz")))
            weights = {k.replace("__", "/"): wz[k] for k in wz.files}
            cfg, opt, opt_vars = head["config"], head.get("optimizer"), None
            if "optimizer.npz" in names:
                oz = np.load(io.BytesIO(z.read("optimizer.npz")))
                opt_vars = dict(m=oz["m"], v=oz["v"], step=int(oz["step"][0]))
        elif "model.weights.h5" in names:
            ar = read_keras_archive(path)
            cfg, weights, opt, opt_vars = ar["config"], ar["weights"], ar["optimizer"], ar["opt_vars"]
            if "s2s_unet_b200.json" in names:
                own = json.loads(z.read("s2s_unet_b200.json"))
                cfg = {**cfg, **own.get("config", {})}
                opt = own.get("optimizer", opt)
        else:
            raise ValueError(f"{path} is neither a Keras-3 archive (model.weights.h5) nor an s2s-unet-b200 model file")
    m = Model(tuple(cfg["input_shape"]), filters=cfg["filters"], n_blocks=cfg["n_blocks"], ct_kernel=cfg["ct_kernel"],
              apool=cfg["apool"], bn=cfg["bn"], output=cfg["output"], device=device, weights=weights,
              activation=cfg.get("activation", "elu"))
    if opt is not None:
        m.compile(optimizer=optimizers.Adam(opt["learning_rate"], opt["beta_1"], opt["beta_2"], opt["epsilon"]),
                  loss="categorical_crossentropy" if cfg["output"] == "proba" else "mse")
        if opt_vars is not None:
            if isinstance(opt_vars["m"], dict):       # per-tensor moments of a Keras archive -> the flat arenas
                fm, fv = np.zeros(m.n_params_padded, np.float32), np.zeros(m.n_params_padded, np.float32)
                for d in m.layout:
                    if d["arena"] == 0 and d["name"] in opt_vars["m"]:
                        sl = slice(d["offset"], d["offset"] + d["count"])
                        fm[sl] = opt_vars["m"][d["name"]].ravel()
                        fv[sl] = opt_vars["v"][d["name"]].ravel()
                opt_vars = dict(m=fm, v=fv, step=opt_vars["step"])
            m._set_opt_state(opt_vars)
    return m
