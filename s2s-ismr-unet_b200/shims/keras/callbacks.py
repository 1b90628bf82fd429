from s2s_ismr_unet_b200.keras_api.callbacks import Callback, EarlyStopping, ModelCheckpoint  # noqa: F401
