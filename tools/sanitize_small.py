"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck): smoke train/predict, a larger-batch
step (16x32 tiles, pixel-split wgrad), Grad-CAM, skill maps, tercile labeler, ELR, world-1 peer-DP step."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.smoke()
from s2s_ismr_unet_b200.labeled import LabeledArray  # noqa: E402
from s2s_ismr_unet_b200.model import Model  # noqa: E402
from s2s_ismr_unet_b200.parallel import PeerDataParallelTrainer  # noqa: E402
from s2s_ismr_unet_b200.utils import performance_metrics as pm, preprocessing as pp, training  # noqa: E402

rng = np.random.default_rng(0)
m = Model((64, 64, 3), filters=2, n_blocks=3, ct_kernel=5, max_batch=32)
m.compile(loss="categorical_crossentropy")
x = rng.gamma(2.0, 3.0, size=(32, 64, 64, 3)).astype(np.float32)
y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (32, 64, 64))]
print("batch 32 loss", m.train_on_batch(x, y))
print("gradcam", m.gradcam(x[:4], "bottleneck", 2).shape)
tr = PeerDataParallelTrainer(m, sync_bn=True)
print("dp world 1", tr.train_on_batch(x[:16], y[:16]))
tr.close()
m.close()
T = np.concatenate([pd.date_range(f"{yr}-05-01", f"{yr}-09-30", freq="7D").values for yr in range(2005, 2017)])
co = {"T": T, "Y": np.arange(12), "X": np.arange(20)}
xs = LabeledArray(rng.gamma(2.0, 3.0, size=(len(T), 3, 12, 20)).astype(np.float32), ("T", "M", "Y", "X"), {**co, "M": np.arange(3)})
ys = LabeledArray(rng.gamma(2.0, 3.0, size=(len(T), 12, 20)).astype(np.float32), ("T", "Y", "X"), co)
xtr, ytr, xte, yte = [l[0] for l in pp.bootstrap_splits_ELR(xs, ys, n_bootstraps=1)]
p_tr, p_te, ytr_t, yte_t = training.train_single_bootstrap_ELR(xtr, ytr, xte, yte)
print("elr", p_te.shape, float(np.nanmean(p_te.values)))
print("rpss", float(np.nanmean(pm.rpss(pm.climo_predict(xte), p_te._like(p_te.values.astype(np.float32)), yte_t).values)))
print("sanitize pass done")
