"""GPU tests of K4 (RealNVP flow, tcgen05 tensor cores).  normflows is not installable here: parity is
UNPINNED against the package; the kernel is pinned against the fp32 torch restatement of Appendix C and by invariants."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def flow():
    from glabc_b200.flows import RealNVP
    torch.manual_seed(0)
    f = RealNVP(device="cuda")
    with torch.no_grad():       # a "trained" flow: small random last layers so the map is far from the identity
        f.w3.copy_(0.05 * torch.randn_like(f.w3))
        f.b3.copy_(0.02 * torch.randn_like(f.b3))
        f.loc.copy_(torch.tensor([[0.1, -0.2]]))
        f.log_scale.copy_(torch.tensor([[0.05, -0.1]]))
    f.bind()
    return f


def test_identity_at_init():
    """init_zeros=True: the untrained flow is the identity, so sample == base draw and log_prob == base density exactly"""
    from glabc_b200.flows import RealNVP
    f = RealNVP(device="cuda")
    f.bind()
    eps = torch.randn(5000, 2, device="cuda")
    th, lq = f.fused_sample_from(eps)
    assert torch.equal(th, eps)
    want = -np.log(2 * np.pi) - 0.5 * (eps ** 2).sum(1)
    assert torch.allclose(lq, want, rtol=0, atol=2e-6)
    assert torch.allclose(f.fused_log_prob(eps), want, rtol=0, atol=2e-6)


@pytest.mark.parametrize("n", [1, 127, 128, 1024, 1025, 50000])
def test_matches_fp32_torch(flow, n):
    """sample / log_prob against the fp32 autograd path; tolerance = TF32 operand rounding through 32 blocks"""
    eps = torch.randn(n, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(n))
    th, lq = flow.fused_sample_from(eps)
    with torch.no_grad():
        th_r, lq_r = flow.sample_from(eps)
        assert torch.allclose(th, th_r, rtol=5e-3, atol=5e-3), float((th - th_r).abs().max())
        assert torch.allclose(lq, lq_r, rtol=0, atol=2e-2), float((lq - lq_r).abs().max())
        # queries around the flow's own samples (far outside its support exp(-s) overflows in fp32 for both paths)
        x = th_r + 0.1 * torch.randn(n, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(n + 1))
        lp, lp_r = flow.fused_log_prob(x), flow.log_prob(x)
        ok = torch.isfinite(lp_r)
        assert ok.float().mean() > 0.99 and torch.equal(torch.isfinite(lp), ok)
        err = (lp[ok] - lp_r[ok]).abs() / (1 + 0.01 * lp_r[ok].abs())
        # TF32 operand rounding (2^-11 relative) through 32 blocks: tight in the bulk, amplified by exp(-s) in the tails
        assert float(err.median()) < 3e-3 and float(err.quantile(0.99)) < 5e-2 and float(err.max()) < 0.5, float(err.max())


def test_sample_log_prob_consistency(flow):
    """the kernel's own pair: log_prob(sample(eps)) reproduces the log q returned with the sample.  sample()'s log q is the
    exact density of the map that produced theta (same s values in the transform and the log-det); log_prob() re-derives
    each block's input to ~1e-7, which can flip the TF32 rounding of a hidden unit, hence the 1e-3-level tolerance."""
    eps = torch.randn(20000, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    th, lq = flow.fused_sample_from(eps)
    lp = flow.fused_log_prob(th)
    err = (lp - lq).abs()
    assert float(err.median()) < 2e-3 and float(err.quantile(0.99)) < 3e-2, (float(err.median()), float(err.quantile(0.99)))
    assert torch.isfinite(th).all() and torch.isfinite(lq).all()


def test_density_integrates_to_one(flow):
    """exp(log_prob) on a grid sums to ~1 and to the same mass as the fp32 torch path (far cells overflow exp(-s) in
    fp32 for both paths and are dropped the same way)"""
    g = torch.linspace(-6, 6, 601, device="cuda")
    xx, yy = torch.meshgrid(g, g, indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1)], 1)
    with torch.no_grad():
        lp, lp_r = flow.fused_log_prob(pts), flow.log_prob(pts)
    cell = float((g[1] - g[0]) ** 2)
    mass = float(torch.exp(torch.nan_to_num(lp.double(), nan=-1e30, posinf=-1e30)).sum() * cell)
    mass_r = float(torch.exp(torch.nan_to_num(lp_r.double(), nan=-1e30, posinf=-1e30)).sum() * cell)
    assert abs(mass - mass_r) < 5e-3 and abs(mass - 1.0) < 3e-2, (mass, mass_r)
