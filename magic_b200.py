"""Import shim: `import magic_b200` -> the package directory `vln-magic_b200/` (its name is not a valid
Python identifier, so it is loaded through importlib and aliased here)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("vln-magic_b200")
sys.modules[__name__] = _pkg
