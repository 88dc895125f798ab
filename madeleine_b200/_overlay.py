"""Overlay of the B200-native hot path onto a checkout of the reference.

``madeleine`` / ``core`` (the name bin/*.py import) resolve to this package for everything on the hot path: models, losses,
trainer, the utils and datasets modules.  Whatever is NOT rebuilt here — argument parsing, ``setup_components`` (model /
optimiser / dataloader wiring), ``datasets.modalities``, the preprocessing pipeline — falls through to the reference's own
files when a checkout is reachable: its ``madeleine/<sub>`` directories are appended to the ``__path__`` of the matching
subpackages here, so e.g. ``from core.utils.setup_components import setup_model`` loads the reference's file, whose
``from madeleine.models.Model import MADELEINE`` in turn lands on the B200 model.  The checkout is taken from
``$MADELEINE_REFERENCE_ROOT`` or found on ``sys.path`` (the scripts append ``'../'``).
"""
from __future__ import annotations

import os
import sys
from typing import Optional

_SUBPACKAGES = ("", "models", "utils", "datasets")
_MARKER = os.path.join("madeleine", "utils", "setup_components.py")


def find_reference_root() -> Optional[str]:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = []
    env = os.environ.get("MADELEINE_REFERENCE_ROOT")
    if env:
        cands.append(env)
    cands.extend(p or os.getcwd() for p in sys.path)
    for c in cands:
        root = os.path.abspath(c)
        if root == here:
            continue
        if os.path.isfile(os.path.join(root, _MARKER)):
            return root
    return None


def attach(root: Optional[str] = None) -> Optional[str]:
    """Append the reference's package directories to this package's search paths (idempotent).  Returns the root used."""
    import importlib
    root = root or find_reference_root()
    if root is None:
        return None
    for sub in _SUBPACKAGES:
        mod = importlib.import_module("madeleine_b200" + ("." + sub if sub else ""))
        ref_dir = os.path.join(root, "madeleine", sub) if sub else os.path.join(root, "madeleine")
        if os.path.isdir(ref_dir) and ref_dir not in list(mod.__path__):
            mod.__path__.append(ref_dir)
    return root
