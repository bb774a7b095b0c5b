"""Import the *real* reference (``/root/reference/src/psa``) in the build container.

TEST INFRASTRUCTURE ONLY - used by ``oracle/make_golden.py`` and by the optional
differential tests that run when the reference tree is present.  The GPU box has
no ``/root/reference``; nothing that runs there may depend on this module
succeeding (``load_reference()`` returns ``None`` when the tree is absent).

The reference imports matplotlib at package import time
(reference: src/psa/__init__.py:17 -> visualization/sed_plotter.py:5), which this
image does not have, so a permissive dummy module is registered first
(SURVEY.md appendix B).
"""
from __future__ import annotations

import logging
import os
import sys
import types
from pathlib import Path
from typing import Optional


class _Dummy(types.ModuleType):
    """Module whose every attribute is another dummy and which can be called."""

    def __getattr__(self, name: str):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Dummy(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *args, **kwargs):
        return self


def _stub_matplotlib() -> None:
    try:
        import matplotlib  # noqa: F401  (a real install wins)
        return
    except Exception:
        pass
    root = _Dummy("matplotlib")
    root.__path__ = []  # behave like a package
    sys.modules.setdefault("matplotlib", root)
    for sub in ("pyplot", "colors", "cm", "gridspec", "ticker", "animation", "figure",
                "backends", "backends.backend_tkagg", "patches", "lines"):
        mod = _Dummy(f"matplotlib.{sub}")
        mod.__path__ = []
        sys.modules.setdefault(f"matplotlib.{sub}", mod)
        setattr(root, sub.split(".")[0], sys.modules[f"matplotlib.{sub.split('.')[0]}"])


def reference_root() -> Optional[Path]:
    for cand in (os.environ.get("PSA_REFERENCE_SRC"), "/root/reference/src"):
        if cand and (Path(cand) / "psa" / "core" / "sed_calculator.py").exists():
            return Path(cand)
    return None


def load_reference():
    """Return the imported reference package ``psa`` or ``None`` if it is not on this machine."""
    root = reference_root()
    if root is None:
        return None
    _stub_matplotlib()
    sys.dont_write_bytecode = True          # the reference tree is read-only
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    prev = logging.root.manager.disable
    logging.disable(logging.ERROR)          # silences the OVITO import error line
    try:
        import psa  # type: ignore
    finally:
        logging.disable(prev)
    return psa
