"""KernelDensity — reference glabcmcmc/kernel_density.py:4-177: weighted Gaussian KDE with a diagonal bandwidth
(Silverman / Scott factor x weighted unbiased std, or a given scalar / vector).  fit / log_prob / sample run in the
hand-written kernels of csrc/kde.cuh (training points tiled through shared memory, one thread per query, streaming
log-sum-exp) — the reference materialises an [n_points, n_samples, d] tensor (80 GB at 1e5 x 1e5)."""
import torch

from . import _abi
from .engine import get_engine

_RULES = {"silverman": _abi.BW_SILVERMAN, "scott": _abi.BW_SCOTT}


class KernelDensity:
    def __init__(self, bandwidth="silverman", device=None, arith="fast"):
        if isinstance(bandwidth, str) and bandwidth not in _RULES:
            raise ValueError("bandwidth should be 'silverman', 'scott' or a float")   # kernel_density.py:33
        self.bandwidth = bandwidth
        self._eng = get_engine(device)
        self.device = self._eng.device
        self._arith = {"fast": _abi.ARITH_FAST, "strict": _abi.ARITH_STRICT}[arith]
        self.X = None
        self.weights = None
        self.n_samples = 0
        self.dim = None
        self._fitted = False

    def fit(self, X, weights=None):
        """kernel_density.py:70-94"""
        self.X = torch.as_tensor(X, dtype=torch.float32).to(self.device).contiguous()
        self.n_samples, self.dim = self.X.shape
        rule = _RULES[self.bandwidth] if isinstance(self.bandwidth, str) else _abi.BW_SILVERMAN
        w, bw = self._eng.kde_fit(self.X, None if weights is None else torch.as_tensor(weights), rule=rule)
        self.weights = w
        if isinstance(self.bandwidth, str):
            self.bandwidth = bw
        elif isinstance(self.bandwidth, torch.Tensor):
            self.bandwidth = self.bandwidth.to(self.device, torch.float32)
        self._fitted = True
        return self

    def _bw(self):
        if isinstance(self.bandwidth, (int, float)):
            return torch.ones(self.dim, device=self.device) * float(self.bandwidth)   # kernel_density.py:108-109
        return self.bandwidth.reshape(-1).expand(self.dim).contiguous()

    def log_prob(self, x):
        """kernel_density.py:96-128"""
        if not self._fitted:
            raise RuntimeError("Must call fit() before computing probabilities")
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device).reshape(-1, self.dim)
        return self._eng.kde_log_prob(self.X, self.weights, self._bw(), x, arith=self._arith)

    def sample(self, n_samples=1, return_log_prob=False, seed=None):
        """kernel_density.py:130-156"""
        if not self._fitted:
            raise RuntimeError("Must call fit() before sampling")
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)   # advances torch's global RNG
        samples = self._eng.kde_sample(self.X, self.weights, self._bw(), n_samples, seed=seed)
        if return_log_prob:
            return samples, self.log_prob(samples)
        return samples

    def forward(self, n_samples=1, seed=None):
        """kernel_density.py:158-177"""
        return self.sample(n_samples, return_log_prob=True, seed=seed)
