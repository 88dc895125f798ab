"""Import alias: ``madeleine.*`` resolves to the B200-native package ``madeleine_b200.*`` so code written against the
reference (``from madeleine.models.Model import MADELEINE``) runs unchanged."""
import importlib
import sys

_TARGET = "madeleine_b200"
_SUBMODULES = ["models", "models.Model", "models.abmil", "models.factory", "utils", "utils.loss", "utils.trainer", "utils.utils",
               "datasets", "datasets.wsi_dataset"]

_pkg = importlib.import_module(_TARGET)
for _name in _SUBMODULES:
    _mod = importlib.import_module(f"{_TARGET}.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
models = sys.modules[f"{__name__}.models"]
utils = sys.modules[f"{__name__}.utils"]
datasets = sys.modules[f"{__name__}.datasets"]
