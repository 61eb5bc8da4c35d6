"""Import shim: the package directory is called ``snch-lbvh_b200`` (hyphen); this makes ``import snch_lbvh_b200`` work."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("snch-lbvh_b200")
sys.modules[__name__] = _pkg
