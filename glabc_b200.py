"""Importable alias of the package directory `gl-abc-mcmc_b200/` (a hyphen cannot be written in an
`import` statement).  `import glabc_b200` gives the package object itself."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("gl-abc-mcmc_b200")
sys.modules[__name__] = _pkg
