"""Import alias: ``import pic_b200`` == the package directory
``efficient-pic-with-variance-aware-masking_b200/`` (whose name is not a Python identifier).
Sub-modules are aliased too (``pic_b200.ops`` is the same module object as the package's)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_REAL = "efficient-pic-with-variance-aware-masking_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules["pic_b200" + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
