"""RealNVP importance proposal of GLMCMC-NFs — the `normflows` pieces the reference composes at
GLMCMC_NFs.py:51-61 (restated from the package's public behaviour; normflows itself is not
installable here, so this restatement is unpinned against it and pinned instead by its own invariants: identity at
init, invertibility, log_prob(sample) consistency):

    32 x [AffineCouplingBlock(MLP([1, 128, 128, 2], init_zeros=True)), Permute(2, mode='swap')]
    over nf.distributions.base.DiagGaussian(2) (trainable loc / log_scale).

`RealNVP` is a plain torch module: it owns the parameters and provides the fp32 autograd path used for the (at most
`Train_step`) forward-KL Adam steps.  Inference — sample() / log_prob() over millions of candidates — runs in the
tcgen05 tensor-core kernel csrc/flow.cuh through `fused_sample` / `fused_log_prob`."""
import math

import torch
from torch import nn

from . import _abi
from .engine import get_engine


class RealNVP(nn.Module):
    def __init__(self, n_blocks=32, hidden=128, base_loc=None, base_log_scale=None, device=None):
        super().__init__()
        L, H = n_blocks, hidden
        self.n_blocks, self.hidden = L, H

        def lin(out_f, in_f):  # torch.nn.Linear's default initialisation, stacked over the blocks
            w = torch.empty(L, out_f, in_f)
            b = torch.empty(L, out_f)
            for i in range(L):
                nn.init.kaiming_uniform_(w[i], a=math.sqrt(5))
                bound = 1 / math.sqrt(in_f)
                nn.init.uniform_(b[i], -bound, bound)
            return nn.Parameter(w), nn.Parameter(b)

        self.w1, self.b1 = lin(H, 1)
        self.w2, self.b2 = lin(H, H)
        self.w3 = nn.Parameter(torch.zeros(L, 2, H))   # init_zeros=True: the untrained flow is the identity
        self.b3 = nn.Parameter(torch.zeros(L, 2))
        self.loc = nn.Parameter(torch.zeros(1, 2) if base_loc is None else torch.as_tensor(base_loc, dtype=torch.float32).reshape(1, 2))
        self.log_scale = nn.Parameter(torch.zeros(1, 2) if base_log_scale is None
                                      else torch.as_tensor(base_log_scale, dtype=torch.float32).reshape(1, 2))
        if device is not None:
            self.to(device)
        self._bound_version = None

    # ---- fp32 torch path (training / reference) ------------------------------------------------------------
    def _params(self, l, z1):
        h = torch.relu(z1 @ self.w1[l].t() + self.b1[l])
        h = torch.relu(h @ self.w2[l].t() + self.b2[l])
        p = h @ self.w3[l].t() + self.b3[l]
        return p[:, 0:1], p[:, 1:2]          # shift = param[:, 0::2], log-scale = param[:, 1::2]

    def base_forward(self, eps):
        z = self.loc + torch.exp(self.log_scale) * eps
        log_p = -0.5 * 2 * math.log(2 * math.pi) - torch.sum(self.log_scale + 0.5 * eps ** 2, 1)
        return z, log_p

    def base_log_prob(self, z):
        return -0.5 * 2 * math.log(2 * math.pi) - torch.sum(self.log_scale + 0.5 * ((z - self.loc) / torch.exp(self.log_scale)) ** 2, 1)

    def sample_from(self, eps):
        """NormalizingFlow.sample with the base normals given: z, log_q = q0(n); for flow: z, ld = flow(z); log_q -= ld"""
        z, log_q = self.base_forward(eps)
        z1, z2 = z[:, 0:1], z[:, 1:2]
        for l in range(self.n_blocks):
            shift, s = self._params(l, z1)
            z2 = z2 * torch.exp(s) + shift
            log_q = log_q - s[:, 0]
            z1, z2 = z2, z1                  # Permute(2, 'swap')
        return torch.cat([z1, z2], 1), log_q

    def sample(self, num_samples=1):
        eps = torch.randn(num_samples, 2, device=self.loc.device)
        return self.sample_from(eps)

    def log_prob(self, x):
        """NormalizingFlow.log_prob: inverse through the flows in reverse, log_q += ld, + q0.log_prob(z)"""
        x = x.reshape(-1, 2)
        z1, z2 = x[:, 0:1], x[:, 1:2]
        log_q = torch.zeros(x.shape[0], device=x.device)
        for l in reversed(range(self.n_blocks)):
            z1, z2 = z2, z1
            shift, s = self._params(l, z1)
            z2 = (z2 - shift) * torch.exp(-s)
            log_q = log_q - s[:, 0]
        return log_q + self.base_log_prob(torch.cat([z1, z2], 1))

    def forward_kld(self, x):
        return -torch.mean(self.log_prob(x))

    # ---- fused tensor-core path ----------------------------------------------------------------------------
    def bind(self, eng=None):
        """copy the current weights into the engine's context (packs W2 into the UMMA shared-memory layout)"""
        import ctypes as C
        eng = eng or get_engine(self.loc.device)
        t = {k: getattr(self, k).detach().float().contiguous() for k in ("w1", "b1", "w2", "b2", "w3", "b3")}
        pod = _abi.FlowPOD(n_blocks=self.n_blocks, hidden=self.hidden, dim=2,
                           w1=t["w1"].data_ptr(), b1=t["b1"].data_ptr(), w2=t["w2"].data_ptr(), b2=t["b2"].data_ptr(),
                           w3=t["w3"].data_ptr(), b3=t["b3"].data_ptr())
        for i in range(2):
            pod.base_loc[i] = float(self.loc.detach()[0, i])
            pod.base_log_scale[i] = float(self.log_scale.detach()[0, i])
        eng.ctx.check(eng.lib.glabc_flow_set(eng.ctx.handle, C.byref(pod), C.sizeof(pod), eng._stream()))
        torch.cuda.current_stream(eng.device).synchronize()   # the staging tensors in `t` may be freed now
        return eng

    # ---- native training step (csrc/flow_train.cuh) ---------------------------------------------------------
    _FLAT = ("w1", "b1", "w2", "b2", "w3", "b3", "loc", "log_scale")      # FlowParamLayout (include/glabc.h)

    def flat_params(self):
        return torch.cat([getattr(self, k).detach().reshape(-1).float() for k in self._FLAT])

    def load_flat(self, flat):
        """copy a flat parameter vector (glabc_flow_get) back into the module's tensors"""
        off = 0
        with torch.no_grad():
            for k in self._FLAT:
                p = getattr(self, k)
                p.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()

    def train_init(self, eng=None, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        """bind the current weights and (re)start Adam in the engine's context (GLMCMC_NFs.py:63)"""
        eng = self.bind(eng)
        eng.ctx.check(eng.lib.glabc_flow_train_init(eng.ctx.handle, float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay)))
        return eng

    def grad(self, x, eng=None):
        """d forward_kld(x) / d parameters as one flat device tensor + the loss (device scalar): glabc_flow_grad"""
        eng = eng or get_engine(self.loc.device)
        x = x.to(eng.device, torch.float32).reshape(-1, 2).contiguous()
        n_par = int(eng.lib.glabc_flow_param_count(self.n_blocks))
        g = torch.empty(n_par, device=eng.device)
        loss = torch.empty(1, device=eng.device)
        eng.ctx.check(eng.lib.glabc_flow_grad(eng.ctx.handle, eng._ptr(x), x.shape[0], eng._ptr(g), eng._ptr(loss), eng._stream()))
        return g, loss

    def adam_step(self, g, loss, eng=None, sync_module=True):
        """one Adam update of the context's master parameters from the flat gradient `g`; the module's tensors follow"""
        eng = eng or get_engine(self.loc.device)
        eng.ctx.check(eng.lib.glabc_flow_adam_step(eng.ctx.handle, eng._ptr(g), eng._ptr(loss), eng._stream()))
        if sync_module:
            flat = torch.empty_like(g)
            eng.ctx.check(eng.lib.glabc_flow_get(eng.ctx.handle, eng._ptr(flat), eng._stream()))
            self.load_flat(flat)

    def fused_sample_from(self, eps, eng=None, theta=None, log_q=None, precision=None):
        """`precision`: "precise" / "fast" sets the context's flow precision for this and the following calls (None: keep it)"""
        eng = eng or get_engine(self.loc.device)
        if precision is not None:
            eng.flow_precision(precision)
        eps = eps.to(eng.device, torch.float32).contiguous()
        n = eps.shape[0]
        theta = torch.empty(n, 2, device=eng.device) if theta is None else theta
        log_q = torch.empty(n, device=eng.device) if log_q is None else log_q
        assert theta.is_contiguous() and log_q.is_contiguous() and theta.numel() == 2 * n and log_q.numel() == n
        eng.ctx.check(eng.lib.glabc_flow_sample(eng.ctx.handle, eng._ptr(eps), n, eng._ptr(theta), eng._ptr(log_q), eng._stream()))
        return theta, log_q

    def fused_sample(self, n, seed, eng=None, theta=None, log_q=None, precision=None):
        """NormalizingFlow.sample(n) with the base normals drawn inside the kernel (Philox keyed by `seed`)"""
        eng = eng or get_engine(self.loc.device)
        if precision is not None:
            eng.flow_precision(precision)
        theta = torch.empty(n, 2, device=eng.device) if theta is None else theta
        log_q = torch.empty(n, device=eng.device) if log_q is None else log_q
        assert theta.is_contiguous() and log_q.is_contiguous() and theta.numel() == 2 * n and log_q.numel() == n
        eng.ctx.check(eng.lib.glabc_flow_sample_native(eng.ctx.handle, int(seed) & 0xFFFFFFFFFFFFFFFF, n, eng._ptr(theta), eng._ptr(log_q),
                                                       eng._stream()))
        return theta, log_q

    def fused_log_prob(self, x, eng=None, precision=None):
        eng = eng or get_engine(self.loc.device)
        if precision is not None:
            eng.flow_precision(precision)
        x = x.to(eng.device, torch.float32).reshape(-1, 2).contiguous()
        log_q = torch.empty(x.shape[0], device=eng.device)
        eng.ctx.check(eng.lib.glabc_flow_log_prob(eng.ctx.handle, eng._ptr(x), x.shape[0], eng._ptr(log_q), eng._stream()))
        return log_q
