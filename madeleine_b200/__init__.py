"""madeleine_b200 — B200-native (sm_100a) implementation of the MADELEINE hot path.

Public surface mirrors the reference package layout (``models.Model``, ``models.abmil``, ``models.factory``,
``utils.loss``, ``utils.trainer``, ``utils.utils``); ``madeleine`` and ``core`` at the repo root alias it so the
reference's scripts import it unchanged.  Importing the package does not require a GPU; running it does.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def library_path() -> str:
    return _lib.LIB_PATH
