"""compute-sanitizer (memcheck) pass over the tensor-core training mode: the largest tuning-grid point (flat geometry, chunked
channels, transposed convs on tcgen05, tcwgrad) and the default net at two batch sizes, a ragged last batch included."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2s_ismr_unet_b200.model import Model  # noqa: E402

rng = np.random.default_rng(0)
for kw, nb in ((dict(filters=3, n_blocks=5, ct_kernel=5), 8), (dict(filters=2, n_blocks=3, ct_kernel=3), 16), (dict(filters=3, n_blocks=4, ct_kernel=2), 8)):
    m = Model((64, 64, 3), max_batch=nb, precision="tf32", **kw)
    m.compile(loss="categorical_crossentropy")
    m.set_graphs(False)
    for n in (nb, 5):
        x = rng.gamma(2.0, 3.0, size=(n, 64, 64, 3)).astype(np.float32)
        y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (n, 64, 64))]
        print(kw, n, "loss", m.train_on_batch(x, y))
    print("predict", m.predict(x, batch_size=nb).shape)
    m.close()
print("sanitize tf32 pass done")
