"""Mirror of utils/deep_nn_models.py: the `Unet` hyper-parameter holder and `build_model`.

Same constructor signature and defaults as the reference (deep_nn_models.py:19-21, 73); the graph of
deep_nn_models.py:73-163 is realised by the CUDA handle behind `Model`.  The reference's CNN / MLP
alternatives (deep_nn_models.py:166-203) are never selected by any tune script and are out of scope."""
from __future__ import annotations

from s2s_ismr_unet_b200.model import Model


class Unet:
    def __init__(self, v, train_patches=False, weighted_loss=False, ct_kernel=(3, 3), ct_stride=(2, 2), n_blocks=3,
                 filters=2, apool=True, bn=True):
        if train_patches or weighted_loss:
            raise NotImplementedError("train_patches / weighted_loss are never enabled by the reference's callers "
                                      "(training.py:58-60,91-93) and are not built")
        if tuple(ct_stride) != (2, 2):
            raise ValueError("ct_stride is fixed to (2, 2) (every caller uses the default)")
        self.train_patches = train_patches
        self.model_architecture = "unet"
        self.weighted_loss = weighted_loss
        self.input_dims = 0
        self.output_dims = 0
        self.n_bins = 3
        self.region = "global"
        self.filters = filters
        self.apool = apool
        self.n_blocks = n_blocks
        self.bn = bn
        self.ct_kernel = ct_kernel
        self.ct_stride = ct_stride
        self.optimizer_str = "adam"
        self.call_back = True
        # unused training hints kept for attribute compatibility (deep_nn_models.py:47-71)
        if v == "tp":
            self.learn_rate, self.decay_rate, self.delayed_early_stop = 0.001, 0.005, True
        else:
            self.learn_rate, self.decay_rate, self.delayed_early_stop = 1e-4, 0, False
        self.bs, self.ep, self.patience, self.start_epoch = 16, 50, 10, 5

    def build_model(self, dg_train_shape, dg_train_weight_target=None, output="proba", max_batch=32, device=None, rng=None,
                    precision="fp32", activation="elu"):
        """dg_train_shape = (H, W, C) -> Model mapping (N,H,W,C) -> (N,H,W,3) softmax | (N,H,W,1) relu.
        `precision` ("fp32" | "tf32" | "bf16_tc") and `activation` ("elu" = the reference's down()/up() default,
        deep_nn_models.py:139,152 | "relu") are extensions with the reference's behaviour as default."""
        return Model((dg_train_shape[0], dg_train_shape[1], dg_train_shape[2]), filters=self.filters,
                     n_blocks=self.n_blocks, ct_kernel=self.ct_kernel, apool=self.apool, bn=self.bn, output=output,
                     max_batch=max_batch, device=device, rng=rng, precision=precision, activation=activation)
