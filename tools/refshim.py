"""Import shim for the upstream reference (tooling only; never imported by the product path).

The reference needs matplotlib (absent here) and numpy<2's ``np.math``; both are stubbed before
``/root/reference`` is put on sys.path.  Used only by the golden-vector generator and ad-hoc probes
run in the authoring container.
"""
import math
import os
import sys
import types


def install(ref_root=None):
    ref_root = ref_root or os.environ.get("DIB_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(ref_root):
        raise RuntimeError("reference tree not found at %s" % ref_root)
    import numpy as np
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.pyplot"].imread = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    if not hasattr(np, "math"):
        np.math = math
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    return ref_root
