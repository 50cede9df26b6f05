"""Import shim: `import magic_b200` -> the package directory `vln-magic_b200/` (its name is not a valid
Python identifier, so it is loaded through importlib and aliased here).

Every submodule is imported once under its real name and ALIASED as `magic_b200.<sub>`: without the aliases
`from magic_b200.optim import ...` would load a second copy of the submodule (and, through its relative imports, a
second copy of `_lib` with its own launch counters and profiling state)."""
import importlib
import os
import pkgutil
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("vln-magic_b200")
sys.modules[__name__] = _pkg
for _m in pkgutil.iter_modules(_pkg.__path__):
    if not _m.ispkg:
        sys.modules[__name__ + "." + _m.name] = importlib.import_module("vln-magic_b200." + _m.name)
