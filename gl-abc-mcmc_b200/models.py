"""ABC model plugin — the duck-typed object every sampler takes (reference examples/Mixture.py:5-53,
README.md:66-104): attributes `epsilon, theta_dim, y_obs, y_dim`, methods `generate_samples`,
`prior_log_prob`, `discrepancy`, `calculate_log_kernel`, `calculate_log_kernel_dis`.

`AbsNormalModel` is the fused family the reference's `Mixture_set` belongs to:
    y | theta ~ N(link(theta) + noise_loc, diag(noise_scale^2)),  link = |.| or identity,
    prior DiagGaussian, discrepancy = L2 distance to y_obs, Gaussian ABC kernel of width epsilon.
Its torch methods follow the reference formulas (usable on CPU or CUDA tensors, batched); its
`lower()` gives the POD the kernels consume.  `lower_model()` also recognises a *reference*
`Mixture_set` instance (duck-typed probe of its deterministic methods) so existing user scripts can
pass their own model object unchanged.
"""
import torch

from . import _abi
from .distribution import DiagGaussian


class AbsNormalModel:
    def __init__(self, epsilon, y_obs=(1.5, 1.5), noise_var=0.05, prior_loc=None, prior_log_scale=None,
                 noise_loc=None, link="abs", device=None):
        self.epsilon = epsilon
        self.y_obs = torch.as_tensor(y_obs, dtype=torch.float32).view(1, -1)
        self.theta_dim = self.y_obs.shape[1]
        self.y_dim = self.theta_dim
        d = self.theta_dim
        nv = torch.as_tensor(noise_var, dtype=torch.float32).reshape(-1).expand(d).clone()
        self.noise_log_scale = torch.log(nv.sqrt())             # Mixture.py:19
        self.noise_loc = torch.zeros(d) if noise_loc is None else torch.as_tensor(noise_loc, dtype=torch.float32)
        self.prior_loc = torch.zeros(d) if prior_loc is None else torch.as_tensor(prior_loc, dtype=torch.float32)
        self.prior_log_scale = torch.zeros(d) if prior_log_scale is None else torch.as_tensor(prior_log_scale, dtype=torch.float32)
        if link not in ("abs", "identity"):
            raise ValueError("link must be 'abs' or 'identity'")
        self.link = link
        if device is not None:
            self.to(device)

    def to(self, device):
        for name in ("y_obs", "noise_log_scale", "noise_loc", "prior_loc", "prior_log_scale"):
            setattr(self, name, getattr(self, name).to(device))
        return self

    # -- plugin methods (reference Mixture.py:13-53) ------------------------------------------
    def generate_samples(self, theta, num_samples=1):
        theta = theta.to(self.y_obs.device)
        mean = torch.abs(theta) if self.link == "abs" else theta
        lik = DiagGaussian(self.theta_dim, self.noise_loc, self.noise_log_scale)
        if theta.dim() == 1:
            return mean + lik.sample(num_samples)
        n = theta.shape[0]
        if n == 1:
            return mean + lik.sample(num_samples)
        if num_samples == 1:
            return mean + lik.sample(n)
        return mean.unsqueeze(1).repeat(1, num_samples, 1) + lik.sample(num_samples * n).view(n, num_samples, -1)

    def prior_log_prob(self, samples):
        samples = samples.to(self.y_obs.device).view(-1, self.theta_dim)
        return DiagGaussian(self.theta_dim, self.prior_loc, self.prior_log_scale).log_prob(samples)

    def discrepancy(self, y):
        y = y.to(self.y_obs.device).view(-1, self.y_dim)
        return torch.sqrt(torch.sum((y - self.y_obs) ** 2, dim=1))

    def calculate_log_kernel_dis(self, dis, epsilon=None):
        eps = self.epsilon if epsilon is None else epsilon
        eps_t = torch.as_tensor([float(eps)], dtype=torch.float32, device=dis.device)
        return DiagGaussian(1, torch.zeros(1, device=dis.device), torch.log(eps_t)).log_prob(dis.view(-1, 1))

    def calculate_log_kernel(self, y, epsilon=None):
        return self.calculate_log_kernel_dis(self.discrepancy(y), epsilon)

    # -- lowering -----------------------------------------------------------------------------
    def lower(self):
        d = self.theta_dim
        if d > 4:
            raise NotImplementedError("fused kernels are instantiated for theta_dim 1..4")
        pod = _abi.ModelPOD(family=_abi.MODEL_ABS_NORMAL if self.link == "abs" else _abi.MODEL_ID_NORMAL,
                            theta_dim=d, y_dim=d)
        cpu = lambda t: t.detach().float().cpu().reshape(-1)  # noqa: E731
        _abi.fill(pod.y_obs, cpu(self.y_obs).tolist())
        _abi.fill(pod.noise_loc, cpu(self.noise_loc).tolist())
        _abi.fill(pod.noise_scale, torch.exp(cpu(self.noise_log_scale)).tolist())
        _abi.fill(pod.prior_loc, cpu(self.prior_loc).tolist())
        _abi.fill(pod.prior_log_scale, cpu(self.prior_log_scale).tolist())
        _abi.fill(pod.prior_scale, torch.exp(cpu(self.prior_log_scale)).tolist())
        eps_t = torch.tensor([float(self.epsilon)], dtype=torch.float32)
        pod.eps_log_scale = float(torch.log(eps_t))            # Mixture.py:43
        pod.eps_scale = float(torch.exp(torch.log(eps_t)))     # distribution.py:178 (0.05 -> 0.049999997)
        pod.epsilon = float(self.epsilon)                      # GLMALA.py:90 squares the Python float
        return pod


class Mixture_set(AbsNormalModel):
    """The README / examples model (reference Mixture.py:5-11): 2-D theta, y_obs = (1.5, 1.5),
    simulator noise variance 0.05, prior N(0, I)."""

    def __init__(self, epsilon, device=None):
        super().__init__(epsilon, y_obs=(1.5, 1.5), noise_var=0.05, device=device)


class UserModel:
    """An ABC model given as CUDA C++ source: the three plugin methods every sampler calls —
    `generate_samples`, `prior_log_prob`, `discrepancy` (examples/Mixture.py:13-36, README.md:66-104) — written as device
    functions, compiled at run time INTO the fused GlobalMCMC step kernel (csrc/user_model.cu):

        __device__ void  glabc_user_simulate(const float* theta, const float* noise, const float* params, float* y);
        __device__ float glabc_user_prior_log_prob(const float* theta, const float* params);
        __device__ float glabc_user_discrepancy(const float* y, const float* params);

    `noise` holds `n_noise` independent N(0,1) draws per simulator call, `params` the model's constants.  The Gaussian
    ABC kernel of width `epsilon` (Mixture.py:38-53) stays the library's.  `check()` compiles the source without a GPU."""

    def __init__(self, source, theta_dim, y_dim, n_noise, epsilon, params=()):
        self.source, self.theta_dim, self.y_dim, self.n_noise = str(source), int(theta_dim), int(y_dim), int(n_noise)
        self.epsilon = float(epsilon)
        self.params = [float(p) for p in params]
        if len(self.params) > _abi.USER_MAX_PARAMS:
            raise ValueError(f"at most {_abi.USER_MAX_PARAMS} params")

    def user_pod(self):
        pod = _abi.UserModelPOD(theta_dim=self.theta_dim, y_dim=self.y_dim, n_noise=self.n_noise, n_params=len(self.params),
                                source=self.source.encode(), epsilon=self.epsilon)
        _abi.fill(pod.params, self.params)
        return pod

    def check(self, cc=100):
        """compile-only validation (NVRTC, no GPU needed); raises ValueError with the compiler log"""
        import ctypes as C
        log = C.create_string_buffer(8192)
        pod = self.user_pod()
        st = _abi.load().glabc_user_model_check(C.byref(pod), int(cc), log, len(log))
        if st != _abi.OK:
            raise ValueError(f"user model rejected (status {st}): {log.value.decode(errors='replace')}")
        return True


# examples/Mixture.py:13-36 written as a user model (params = {y_obs0, y_obs1, noise_sd}): the worked example of the
# run-time compiled path, and what bench.py times beside the built-in family
MIXTURE_USER_SOURCE = r"""
__device__ void glabc_user_simulate(const float* theta, const float* noise, const float* params, float* y)
{
    y[0] = fabsf(theta[0]) + params[2] * noise[0];
    y[1] = fabsf(theta[1]) + params[2] * noise[1];
}
__device__ float glabc_user_prior_log_prob(const float* theta, const float* params)
{
    return -1.8378770664093453f - 0.5f * (theta[0] * theta[0] + theta[1] * theta[1]);
}
__device__ float glabc_user_discrepancy(const float* y, const float* params)
{
    const float a = y[0] - params[0], b = y[1] - params[1];
    return sqrtf(a * a + b * b);
}
"""


def mixture_user_model(epsilon=0.05):
    return UserModel(MIXTURE_USER_SOURCE, theta_dim=2, y_dim=2, n_noise=2, epsilon=epsilon, params=[1.5, 1.5, 0.05 ** 0.5])


def lower_model(abc_set):
    """POD for `abc_set`: our own model classes lower themselves; a foreign object is accepted only
    if it behaves exactly like the AbsNormal family on a deterministic probe (so a reference
    `Mixture_set` instance drops in).  Anything else raises — there is no interpreted fallback."""
    if hasattr(abc_set, "lower"):
        return abc_set.lower()
    for attr in ("epsilon", "theta_dim", "y_obs", "prior_log_prob", "discrepancy", "calculate_log_kernel", "generate_samples"):
        if not hasattr(abc_set, attr):
            raise NotImplementedError(f"ABC model lacks `{attr}`; cannot lower it to a fused family")
    d = int(abc_set.theta_dim)
    y_obs = torch.as_tensor(abc_set.y_obs, dtype=torch.float32).reshape(-1)
    if y_obs.numel() != d:
        raise NotImplementedError("fused family needs y_dim == theta_dim")
    # probe: recover prior loc/scale, noise scale and link from the object's own methods
    g = torch.Generator().manual_seed(1234)
    th = torch.randn(64, d, generator=g) * 2.0
    # prior: fit a diagonal Gaussian through 2d+1 evaluations, then verify on the probe set
    zero = torch.zeros(1, d)
    p0 = abc_set.prior_log_prob(zero).float().reshape(-1)[0]
    loc, ls = torch.zeros(d), torch.zeros(d)
    for i in range(d):
        e = torch.zeros(1, d)
        e[0, i] = 1.0
        pp = abc_set.prior_log_prob(e).float().reshape(-1)[0]
        pm = abc_set.prior_log_prob(-e).float().reshape(-1)[0]
        inv_var = -(pp + pm - 2 * p0)          # second difference of -0.5 (z-loc)^2 / s^2
        if not torch.isfinite(inv_var) or inv_var <= 0:
            raise NotImplementedError("prior is not a diagonal Gaussian; no fused lowering")
        loc[i] = (pp - pm) / (2 * inv_var)
        ls[i] = -0.5 * torch.log(inv_var)
    loc = torch.where(loc.abs() < 1e-6, torch.zeros_like(loc), loc)
    ls = torch.where(ls.abs() < 1e-6, torch.zeros_like(ls), ls)
    # simulator: same seed, compare against |theta| + s*eps and theta + s*eps
    state = torch.get_rng_state()
    try:
        torch.manual_seed(4321)
        y = abc_set.generate_samples(th, 1).float().reshape(64, d)
        torch.manual_seed(4321)
        eps = torch.randn(64, d)
    finally:
        torch.set_rng_state(state)
    link = None
    for name, mean in (("abs", th.abs()), ("identity", th)):
        resid = y - mean
        s = (resid / eps).median(dim=0).values
        if torch.allclose(resid, s * eps, rtol=1e-4, atol=1e-5):
            link, noise_scale = name, s
            break
    if link is None:
        raise NotImplementedError("simulator is not y = link(theta) + scale*eps; no fused lowering")
    model = AbsNormalModel(abc_set.epsilon, y_obs=y_obs.tolist(), noise_var=(noise_scale ** 2).tolist(),
                           prior_loc=loc.tolist(), prior_log_scale=ls.tolist(), link=link)
    # verify the deterministic plugin methods agree before trusting the lowering
    ok = (torch.allclose(model.prior_log_prob(th), abc_set.prior_log_prob(th).float(), rtol=1e-5, atol=1e-5)
          and torch.allclose(model.discrepancy(y), abc_set.discrepancy(y).float(), rtol=1e-5, atol=1e-6)
          and torch.allclose(model.calculate_log_kernel(y), abc_set.calculate_log_kernel(y).float(), rtol=1e-5, atol=1e-4))
    if not ok:
        raise NotImplementedError("model does not match the AbsNormal family on the probe set; no fused lowering")
    return model.lower()


def lower_proposal(dist):
    if hasattr(dist, "lower"):
        return dist.lower()
    # a reference glabcmcmc.distribution.DiagGaussian: same attributes
    if all(hasattr(dist, a) for a in ("loc", "log_scale", "shape", "d")) and type(dist).__name__ == "DiagGaussian":
        return DiagGaussian(tuple(dist.shape), dist.loc, dist.log_scale).lower()
    raise NotImplementedError(f"proposal {type(dist).__name__} has no fused lowering")
