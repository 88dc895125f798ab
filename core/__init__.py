"""Import alias: ``core.*`` resolves to the B200-native package ``madeleine_b200.*`` so code written against the
reference (``from core.models.Model import MADELEINE, as bin/pretrain.py does``) runs unchanged."""
import importlib
import sys

_TARGET = "madeleine_b200"
_SUBMODULES = ["models", "models.Model", "models.abmil", "models.factory", "utils", "utils.loss", "utils.trainer", "utils.utils", "utils.file_utils",
               "datasets", "datasets.wsi_dataset"]

_pkg = importlib.import_module(_TARGET)
for _name in _SUBMODULES:
    _mod = importlib.import_module(f"{_TARGET}.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
models = sys.modules[f"{__name__}.models"]
utils = sys.modules[f"{__name__}.utils"]
datasets = sys.modules[f"{__name__}.datasets"]

# everything that is not rebuilt here (setup_components, process_args, datasets.modalities, preprocessing, ...) falls
# through to a checkout of the reference when one is reachable ($MADELEINE_REFERENCE_ROOT or sys.path)
from madeleine_b200 import _overlay  # noqa: E402
reference_root = _overlay.attach()
if reference_root is not None:
    import os as _os
    _ref_pkg = _os.path.join(reference_root, "madeleine")
    if _ref_pkg not in __path__:
        __path__.append(_ref_pkg)          # root-level fall-through, e.g. madeleine.preprocessing


def __getattr__(name):
    # late attach: the scripts extend sys.path (sys.path.append('../')) before their first `core.*` import
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    _overlay.attach()
    try:
        return importlib.import_module(f"{_TARGET}.{name}")
    except ImportError as e:
        raise AttributeError(name) from e
