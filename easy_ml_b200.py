"""Import shim: the package directory is `easy-ml_b200/` (not an identifier), so `import easy_ml_b200`
resolves to it through importlib."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("easy-ml_b200")
