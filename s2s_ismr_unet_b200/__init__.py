"""Importable alias of the `s2s-ismr-unet_b200/` package directory (a hyphen cannot appear in a
Python module name).  All code lives in ../s2s-ismr-unet_b200/; this file only redirects the
package search path there and re-exports the public host API."""
from pathlib import Path as _Path

__path__ = [str(_Path(__file__).resolve().parent.parent / "s2s-ismr-unet_b200")]

from ._lib import S2SError, load as load_library  # noqa: E402,F401
from .model import Model, History, load_model, set_seed  # noqa: E402,F401
from . import keras_api  # noqa: E402,F401
