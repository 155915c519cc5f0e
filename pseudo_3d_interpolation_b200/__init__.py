"""Import shim: exposes the package that lives in ``pseudo-3d-interpolation_b200/`` (a
directory name Python cannot import) under the importable name ``pseudo_3d_interpolation_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pseudo-3d-interpolation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
