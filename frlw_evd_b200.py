"""Import shim: the package directory is ``frlw-evd_b200/`` (not a valid Python
identifier), so this module makes it importable as ``frlw_evd_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "frlw-evd_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__, "r", encoding="utf-8") as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
del _fh
