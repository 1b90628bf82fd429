from s2s_ismr_unet_b200.keras_api.optimizers import Adam  # noqa: F401
