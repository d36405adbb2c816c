"""Importable alias of the `pointnav-vo_b200/` package directory (a hyphen cannot appear in a Python
module name).  All sources live in `pointnav-vo_b200/`; this shim only extends the package path."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pointnav-vo_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
