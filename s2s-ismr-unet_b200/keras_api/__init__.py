"""The slice of the `keras` namespace that utils/training.py imports (training.py:13-14,21):
optimizers.Adam, callbacks.ModelCheckpoint / EarlyStopping, models.load_model, utils.to_categorical."""
from . import callbacks, models, optimizers, utils  # noqa: F401
