"""Import shim: the package directory is `ac-tsr_b200/` (hyphen); `import ac_tsr_b200` returns it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('ac-tsr_b200')
sys.modules[__name__] = _pkg
