"""`tensorflow` import name: only what the reference's U-Net path touches —
tf.random.set_seed (training.py:25), tf.keras.utils.to_categorical (preprocessing.py:426-428),
tf.config.experimental.list_physical_devices('GPU') (tune_ECMWF_com.py:20)."""
from types import SimpleNamespace

from s2s_ismr_unet_b200.keras_api import utils as _utils
from s2s_ismr_unet_b200.model import set_seed as _set_seed
from s2s_ismr_unet_b200.runtime import device_count as _device_count

random = SimpleNamespace(set_seed=_set_seed)
keras = SimpleNamespace(utils=_utils)
config = SimpleNamespace(experimental=SimpleNamespace(
    list_physical_devices=lambda kind="GPU": [f"/physical_device:GPU:{i}" for i in range(_device_count())] if kind == "GPU" else []),
    list_physical_devices=lambda kind="GPU": [f"/physical_device:GPU:{i}" for i in range(_device_count())] if kind == "GPU" else [])
