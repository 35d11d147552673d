"""Import shim for the UNMODIFIED reference (cca_zoo at /root/reference).

TEST INFRASTRUCTURE ONLY.  Used in the authoring container to (a) validate the
numpy restatement in ``oracle/restatement.py`` and (b) generate the golden
fixtures under ``tests/golden`` (``oracle/make_golden.py``).  ``/root/reference``
does not exist on the GPU box, so nothing here may be imported by the product,
by ``bench.py`` or by ``-m gpu`` tests.

Two obstacles are shimmed without touching the reference tree:
  * ``cca_zoo/__init__.py:10`` asks importlib.metadata for the installed version
    (the tree is not pip-installed)            -> answer a dummy version;
  * ``cca_zoo/linear/__init__.py:26`` imports ``_tcca`` which imports
    ``tensorly`` (absent)                      -> register stub modules.
"""
from __future__ import annotations

import importlib.metadata as _ilm
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CCA_ZOO_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cca_zoo"))


def install() -> None:
    """Make ``import cca_zoo`` resolve to the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    real_version = _ilm.version

    def _version(name: str) -> str:
        if name == "cca_zoo":
            return "0.0.0+reference"
        return real_version(name)

    _ilm.version = _version  # type: ignore[assignment]
    if "tensorly" not in sys.modules:
        try:
            import tensorly  # noqa: F401
        except Exception:
            tl = types.ModuleType("tensorly")
            tl.set_backend = lambda *a, **k: None  # type: ignore[attr-defined]
            dec = types.ModuleType("tensorly.decomposition")
            dec.parafac = None  # type: ignore[attr-defined]
            tl.decomposition = dec  # type: ignore[attr-defined]
            sys.modules["tensorly"] = tl
            sys.modules["tensorly.decomposition"] = dec
