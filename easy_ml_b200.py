"""Import shim: the package directory is `easy-ml_b200/` (not an identifier), so `import easy_ml_b200`
resolves to it through importlib.  Sub-modules are aliased too, so `easy_ml_b200._ffi` and
`easy-ml_b200._ffi` are one module object (one set of exception classes, one loaded library)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = "easy-ml_b200"
_pkg = importlib.import_module(_real)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_real + "."):
        sys.modules[__name__ + _name[len(_real):]] = _mod
sys.modules[__name__] = _pkg
