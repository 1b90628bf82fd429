"""`keras` import name for running the reference's utils/training.py call sites against the B200 library
(training.py:13-14,21: `from keras import layers, models, optimizers`, `from keras.callbacks import ...`)."""
from s2s_ismr_unet_b200.keras_api import callbacks, models, optimizers, utils  # noqa: F401

layers = None   # imported by the reference (training.py:13) but never used on the U-Net path
