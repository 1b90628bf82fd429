"""keras.models.load_model (utils/training.py:114,128-131)."""
from ..model import Model, load_model  # noqa: F401
