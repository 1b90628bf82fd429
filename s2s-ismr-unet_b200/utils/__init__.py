"""Host-side mirror of the reference's `utils` package for the U-Net hot path: same module and
function names as utils/deep_nn_models.py, utils/training.py, utils/preprocessing.py and
utils/performance_metrics.py, backed by the CUDA library.  Put `s2s-ismr-unet_b200/` (and
`s2s-ismr-unet_b200/shims/` for the keras / tensorflow names) first on sys.path to run the
reference's tune_*.py against it."""
