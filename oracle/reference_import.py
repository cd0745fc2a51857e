"""Import the UNMODIFIED reference decoder from /root/reference -- TEST INFRASTRUCTURE ONLY.

The reference tree only exists in the build container (never on the GPU box), so everything here
is optional: ``load_reference_decoder()`` returns None when the tree is absent.  The single stub
needed is ``matplotlib`` (tools/alignment_decoder.py:5 imports tools/plot.py:1, which imports it
at module scope; plotting is never called by the tests).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HFA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tools", "alignment_decoder.py"))


def load_reference_decoder():
    """Returns the reference ``AlignmentDecoder`` class, or None if the tree is not present."""
    if not reference_available():
        return None
    try:
        import numba  # noqa: F401  (the reference needs it; absent -> no reference)
    except Exception:
        return None
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own package is not called "tools", so there is no name clash
    from tools.alignment_decoder import AlignmentDecoder  # type: ignore

    return AlignmentDecoder
