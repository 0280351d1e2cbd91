"""Drop-in import path: `from alpha_eigen_modules.group import OneGroup` (the reference's package
name) resolves to the B200 implementation in alpha_eigen_b200."""
import importlib
import sys

for _name in ("phi", "eigenvalue", "material", "slab", "zone", "one_direction", "group", "time_unit", "inputs"):
    _mod = importlib.import_module("alpha_eigen_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
