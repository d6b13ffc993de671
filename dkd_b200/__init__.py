"""Import name of the package whose sources live in `dl-dkd_b200/` (the directory name the project layout
prescribes carries a hyphen, which Python cannot import).  `import dkd_b200` with the repository root on
sys.path resolves every submodule (`dkd_b200.ops`, `dkd_b200.engine`, ...) to `dl-dkd_b200/<name>.py`; the
shared library is loaded from that directory as well (`dl-dkd_b200/libdkd_b200.so`)."""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dl-dkd_b200")
__path__.insert(0, _SRC)
with open(_os.path.join(_SRC, "__init__.py")) as _f:
    _code = _f.read()
__doc__ = None
exec(compile(_code, _os.path.join(_SRC, "__init__.py"), "exec"))
del _f, _code
