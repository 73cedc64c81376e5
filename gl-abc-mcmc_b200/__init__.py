"""glabc-b200: B200-native GL-ABC-MCMC sampler inner loop (drop-in for `glabcmcmc`'s hot path).

The directory name carries a hyphen (repo convention), so import it through the alias module at
the repo root: `import glabc_b200` (or `importlib.import_module("gl-abc-mcmc_b200")`).
Export list mirrors the reference's `glabcmcmc/__init__.py:1-14`.
"""
from . import _abi  # noqa: F401
from .AGLMCMC import AGLMCMC  # noqa: F401
from .distribution import DiagGaussian, Gamma, GaussianMixture, Uniform  # noqa: F401
from .ESJD import esjd  # noqa: F401
from .GLMALA import GLMALA  # noqa: F401
from .GLMCMC import GLMCMC  # noqa: F401
from .GLMCMC_NFs import GLMCMC_NF  # noqa: F401
from .GlobalMCMC import GlobalMCMC  # noqa: F401
from .kernel_density import KernelDensity  # noqa: F401
from .MCMCRunner import MCMCRunner  # noqa: F401
from .models import AbsNormalModel, Mixture_set, UserModel  # noqa: F401
