"""Proposal / prior distribution classes — the plugin surface of `glabcmcmc/distribution.py`.

Same constructors, same `forward(n) -> (z, log_p)`, `log_prob(z)`, `sample(n)` (reference
distribution.py:16-48) and the same formulas, evaluated with torch ops on whatever device the
parameters live on.  Each class additionally lowers itself to the POD the fused kernels take
(`lower()` -> `_abi.DistPOD`, include/glabc.h `glabc_dist_t`), with every constant evaluated in
float32 the way the reference evaluates it (e.g. scale = exp(log_scale), distribution.py:170).
"""
import math

import numpy as np
import torch

from . import _abi


class BaseDistribution:
    """reference distribution.py:7-48"""

    def forward(self, num_samples=1):
        raise NotImplementedError

    def log_prob(self, z):
        raise NotImplementedError

    def sample(self, num_samples=1, **kwargs):
        z, _ = self.forward(num_samples, **kwargs)
        return z

    def lower(self):
        raise NotImplementedError(f"{type(self).__name__} has no fused lowering")


def _as_tensor(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(x, dtype=dtype or torch.float32)


class DiagGaussian(BaseDistribution):
    """Multivariate Gaussian with diagonal covariance — reference distribution.py:143-203."""

    def __init__(self, shape, loc, log_scale):
        if isinstance(shape, int):
            shape = (shape,)
        self.shape = tuple(shape)
        self.n_dim = len(self.shape)
        self.d = int(np.prod(self.shape))
        self.loc = _as_tensor(loc)
        self.log_scale = _as_tensor(log_scale)

    def forward(self, num_samples=1, context=None):
        eps = torch.randn((num_samples,) + self.shape, dtype=self.loc.dtype, device=self.loc.device)
        z = self.loc + torch.exp(self.log_scale) * eps
        log_p = -0.5 * self.d * math.log(2 * math.pi) - torch.sum(
            self.log_scale + 0.5 * torch.pow(eps, 2), list(range(1, self.n_dim + 1)))
        return z, log_p

    def log_prob(self, z, context=None):
        return -0.5 * self.d * math.log(2 * math.pi) - torch.sum(
            self.log_scale + 0.5 * torch.pow((z - self.loc) / torch.exp(self.log_scale), 2),
            list(range(1, self.n_dim + 1)))

    def cdf(self, z):
        """joint CDF under independence — reference distribution.py:183-200"""
        normal = torch.distributions.Normal(self.loc, torch.exp(self.log_scale))
        return torch.prod(normal.cdf(z), dim=-1)

    def register_buffer(self, param, param1):  # reference distribution.py:202-203 (no-op)
        pass

    def to(self, device):
        return DiagGaussian(self.shape, self.loc.to(device), self.log_scale.to(device))

    def lower(self):
        if self.n_dim != 1 or self.d > _abi.MAX_DIM:
            raise NotImplementedError("fused DiagGaussian needs a 1-D event shape of at most 8")
        d = self.d
        loc = self.loc.detach().float().cpu().reshape(-1)
        ls = self.log_scale.detach().float().cpu().reshape(-1)
        loc = loc.expand(d) if loc.numel() == 1 else loc
        ls = ls.expand(d) if ls.numel() == 1 else ls
        if loc.numel() != d or ls.numel() != d:
            raise ValueError("loc / log_scale do not broadcast to the event shape")
        pod = _abi.DistPOD(kind=_abi.DIST_DIAG_GAUSSIAN, dim=d)
        _abi.fill(pod.a, loc.tolist())
        _abi.fill(pod.b, ls.tolist())
        _abi.fill(pod.c, torch.exp(ls).tolist())  # float32 exp, as the reference evaluates it per call
        return pod


class Uniform(BaseDistribution):
    """Multivariate box uniform — reference distribution.py:50-86."""

    def __init__(self, shape, low=None, high=None):
        if isinstance(shape, int):
            shape = (shape,)
        self.shape = tuple(shape)
        self.low = _as_tensor([-2.0] if low is None else low)
        self.high = _as_tensor([2.0] if high is None else high)
        self.log_prob_val = -torch.log(torch.prod(self.high - self.low))

    def forward(self, num_samples=1, context=None):
        eps = torch.rand((num_samples,) + self.shape, dtype=self.low.dtype, device=self.low.device)
        z = self.low + (self.high - self.low) * eps
        log_p = self.log_prob_val * torch.ones(num_samples, device=self.low.device)
        return z, log_p

    def log_prob(self, z, context=None):
        log_p = self.log_prob_val * torch.ones(z.shape[0], device=z.device)
        out_range = torch.logical_or(z < self.low, z > self.high)
        ind_inf = torch.any(torch.reshape(out_range, (z.shape[0], -1)), dim=-1)
        log_p[ind_inf] = -np.inf
        return log_p

    def lower(self):
        d = int(np.prod(self.shape))
        pod = _abi.DistPOD(kind=_abi.DIST_UNIFORM, dim=d)
        _abi.fill(pod.a, self.low.float().cpu().reshape(-1).expand(d).tolist())
        _abi.fill(pod.b, self.high.float().cpu().reshape(-1).expand(d).tolist())
        pod.c[0] = float(self.log_prob_val)
        return pod


class Gamma(BaseDistribution):
    """Independent Gamma — reference distribution.py:90-137 (scipy.stats.gamma there, float64).

    Here sampling uses torch.distributions.Gamma (Marsaglia–Tsang) and log_prob the closed form,
    both in float64 like the reference's outputs."""

    def __init__(self, Shape, Rate):
        self.Shape = _as_tensor(Shape, torch.float64)
        self.Rate = _as_tensor(Rate, torch.float64)

    def forward(self, num_samples=1, context=None):
        z = torch.distributions.Gamma(self.Shape, self.Rate).sample((num_samples,))
        return z, self.log_prob(z)

    def log_prob(self, z, context=None):
        z = z.to(torch.float64)
        a, b = self.Shape.to(z.device), self.Rate.to(z.device)
        zs = torch.clamp(z, min=torch.finfo(torch.float64).tiny)
        lp = a * torch.log(b) - torch.lgamma(a) + (a - 1.0) * torch.log(zs) - b * zs
        # pdf == 0 -> -inf (reference distribution.py:133-136): outside the support, and at 0 for shape > 1
        zero = (z < 0) | ((z == 0) & (a > 1.0))
        lp = torch.where(zero, torch.full_like(lp, -np.inf), lp)
        at0 = (z == 0) & (a == 1.0)
        lp = torch.where(at0, torch.log(b).expand_as(lp), lp)
        return torch.sum(lp, dim=1)

    def lower(self):
        d = self.Shape.numel()
        pod = _abi.DistPOD(kind=_abi.DIST_GAMMA, dim=d)
        _abi.fill(pod.a, self.Shape.reshape(-1).tolist())
        _abi.fill(pod.b, self.Rate.reshape(-1).tolist())
        return pod


class GaussianMixture(BaseDistribution):
    """Mixture of diagonal Gaussians — reference distribution.py:206-293 (float64 parameters)."""

    def __init__(self, n_modes, dim, loc=None, scale=None, weights=None):
        self.n_modes, self.dim = n_modes, dim
        if loc is None:
            loc = np.random.randn(n_modes, dim)
        loc = np.array(loc)[None, ...]
        scale = np.ones((n_modes, dim)) if scale is None else scale
        scale = np.array(scale)[None, ...]
        weights = np.ones(n_modes) if weights is None else weights
        weights = np.array(weights, dtype=np.float64)[None, ...]
        weights = weights / weights.sum(1)
        self.loc = torch.tensor(1.0 * loc)
        self.log_scale = torch.tensor(np.log(1.0 * scale))
        self.weight_scores = torch.tensor(np.log(1.0 * weights))

    def _log_p(self, z, weights):
        eps = (z[:, None, :] - self.loc) / torch.exp(self.log_scale)
        log_p = (-0.5 * self.dim * np.log(2 * np.pi) + torch.log(weights)
                 - 0.5 * torch.sum(torch.pow(eps, 2), 2) - torch.sum(self.log_scale, 2))
        return torch.logsumexp(log_p, 1)

    def forward(self, num_samples=1):
        weights = torch.softmax(self.weight_scores, 1)
        mode = torch.multinomial(weights[0, :], num_samples, replacement=True)
        eps_ = torch.randn(num_samples, self.dim, dtype=self.loc.dtype, device=self.loc.device)
        z = eps_ * torch.exp(self.log_scale)[0, mode] + self.loc[0, mode]
        return z, self._log_p(z, weights)

    def log_prob(self, z):
        weights = torch.softmax(self.weight_scores, 1)
        if self.dim == 1 and z.dim() == 1:
            z = z[:, None]
        return self._log_p(z.to(self.loc.dtype), weights)

    def lower(self):
        if self.n_modes > _abi.MAX_MODES or self.dim > _abi.MAX_DIM:
            raise NotImplementedError("fused GaussianMixture: at most 8 modes x 8 dims")
        pod = _abi.DistPOD(kind=_abi.DIST_GAUSSIAN_MIXTURE, dim=self.dim, n_modes=self.n_modes)
        w = torch.softmax(self.weight_scores, 1)[0]
        for m in range(self.n_modes):
            _abi.fill(pod.mix_loc[m], self.loc[0, m].tolist())
            _abi.fill(pod.mix_log_scale[m], self.log_scale[0, m].tolist())
            _abi.fill(pod.mix_scale[m], torch.exp(self.log_scale[0, m].float()).tolist())
        _abi.fill(pod.mix_log_w, torch.log(w).tolist())
        _abi.fill(pod.mix_w, w.tolist())
        return pod
