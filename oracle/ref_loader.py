"""Load the UNMODIFIED reference package from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  The reference targets numpy < 1.24: ``azulnet/azul.py:19-26,71,
93-98`` use ``np.int`` / ``np.bool``, which numpy 2.x removed.  Aliasing the two names before
the import is the only shim; no reference source is copied or modified.

The GPU box has no /root/reference -- nothing under tests ``-m gpu``, ``smoke()`` or
``bench.py`` may call this module.  It exists to (a) generate ``tests/golden`` and (b) let the
CPU test-suite cross-check the C restatement against the live reference when it is present.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("AZUL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "azulnet", "azul.py"))


def load_reference():
    """Return the imported reference ``azulnet`` package (raises if absent)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int          # removed in numpy 1.24; azul.py:19
    if not hasattr(np, "bool"):
        np.bool = bool        # removed in numpy 1.24; azul.py:23
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mod = importlib.import_module("azulnet")
    origin = os.path.realpath(mod.__file__)
    if not origin.startswith(os.path.realpath(REFERENCE_ROOT)):
        raise RuntimeError("'azulnet' resolved to %s, not the reference" % origin)
    return mod
