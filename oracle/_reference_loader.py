"""Import the reference's own ``functions/POCS.py`` from /root/reference (this container only).

matplotlib is absent here and is only used by the reference's plotting helpers
(functions/POCS.py:7-8), so empty stub modules are registered first.  Nothing is copied:
the module object is loaded from where it lies.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("P3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pseudo_3D_interpolation", "functions", "POCS.py"))


def load():
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = lambda ax: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from pseudo_3D_interpolation.functions import POCS as ref
    return ref
