"""glabc-b200: B200-native GL-ABC-MCMC sampler inner loop (drop-in for `glabcmcmc`'s hot path).

The directory name carries a hyphen (repo convention), so import it through the alias module at
the repo root: `import glabc_b200` (or `importlib.import_module("gl-abc-mcmc_b200")`).
Export list mirrors the reference's `glabcmcmc/__init__.py:1-14`.
"""
from . import _abi  # noqa: F401
