"""tf.keras.utils.to_categorical (utils/preprocessing.py:426-428; training.py:389)."""
import numpy as np


def to_categorical(y, num_classes=None):
    """Integer labels -> one-hot float32 on a new last axis.  NaN labels (ocean / missing tercile
    edges) cast to class 0 exactly like `np.array(y, dtype='int64')` does for NaN after the reference's
    fillna(0) has removed them in practice (preprocessing.py:342-343)."""
    y = np.asarray(y)
    idx = np.nan_to_num(y, nan=0.0).astype(np.int64)
    if num_classes is None:
        num_classes = int(idx.max()) + 1
    out = np.zeros(idx.shape + (int(num_classes),), np.float32)
    np.put_along_axis(out, idx[..., None], 1.0, axis=-1)
    return out
