from s2s_ismr_unet_b200.keras_api.models import Model, load_model  # noqa: F401
