"""Importable alias of the package directory `gl-abc-mcmc_b200/` (a hyphen cannot be written in an
`import` statement).  `import glabc_b200` gives the package object itself, and `glabc_b200.<sub>` is the SAME module
object as the package's own `.<sub>` (one engine cache, one loaded libglabc.so), whichever way it is imported."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_REAL, _ALIAS = "gl-abc-mcmc_b200", __name__
_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """`glabc_b200.x.y` -> the module `gl-abc-mcmc_b200.x.y`"""

    def find_spec(self, name, path=None, target=None):
        if name.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(name, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
sys.modules[__name__] = _pkg
