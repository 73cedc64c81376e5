import torch, sys
sys.path.insert(0, '/root/repo')
import glabc_b200
from glabc_b200.flows import RealNVP
torch.manual_seed(0)
f = RealNVP(device="cuda")
with torch.no_grad():
    f.w3.copy_(0.05 * torch.randn_like(f.w3))
f.bind()
eps = torch.randn(50000, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(50000))
th, lq = f.fused_sample_from(eps)
with torch.no_grad():
    th_r, lq_r = f.sample_from(eps)
e = (th - th_r).abs().max(1).values
print("theta err median %.2e p99 %.2e max %.2e ; lq err median %.2e p99 %.2e max %.2e" % (e.median(), e.quantile(0.99), e.max(), (lq-lq_r).abs().median(), (lq-lq_r).abs().quantile(0.99), (lq-lq_r).abs().max()))
n = 50000
x = th_r + 0.1 * torch.randn(n, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(n + 1))
with torch.no_grad():
    lp, lp_r = f.fused_log_prob(x), f.log_prob(x)
ok = torch.isfinite(lp_r)
err = (lp[ok] - lp_r[ok]).abs() / (1 + 0.01 * lp_r[ok].abs())
print("log_prob err median %.2e p99 %.2e max %.2e finite-match %s" % (err.median(), err.quantile(0.99), err.max(), torch.equal(torch.isfinite(lp), ok)))
th2, lq2 = f.fused_sample_from(eps)
print("consistency |log_prob(sample) - lq| median %.2e max %.2e" % ((f.fused_log_prob(th2) - lq2).abs().median(), (f.fused_log_prob(th2) - lq2).abs().max()))
