"""Import shim: the package directory carries the reference's repository name
(``joint-regressor-refinement_b200``), which is not a Python identifier."""
import importlib
import sys

_pkg = importlib.import_module("joint-regressor-refinement_b200")
sys.modules[__name__] = _pkg
