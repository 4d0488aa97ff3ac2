"""Import the unmodified reference package (only where it exists: the dev container).

TEST INFRASTRUCTURE ONLY.  The reference tree is read-only and lacks the
setuptools-scm generated ``rkstiff/__version__.py`` (rkstiff/__init__.py:4), so a stub
module is registered before import.  Search order: $RKSTIFF_REF, /root/reference.
Nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.
"""
from __future__ import annotations

import os
import sys
import types


def reference_root():
    for cand in (os.environ.get("RKSTIFF_REF"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "rkstiff")):
            return cand
    return None


def load_reference():
    """Return the imported reference ``rkstiff`` package or None if it is not present."""
    root = reference_root()
    if root is None:
        return None
    if "rkstiff" in sys.modules:
        return sys.modules["rkstiff"]
    stub = types.ModuleType("rkstiff.__version__")
    stub.version = "0+ref"
    sys.modules["rkstiff.__version__"] = stub
    sys.path.insert(0, root)
    try:
        import rkstiff  # noqa: F401
        for mod in ("etd", "etd4", "etd5", "etd34", "etd35", "if4", "if34", "if45dp", "models", "grids",
                    "solveras", "solvercs"):
            __import__(f"rkstiff.{mod}")
    finally:
        sys.path.remove(root)
    return sys.modules["rkstiff"]
