"""The flow's training step as kernels of the extension (csrc/flow_train.cuh: forward KL, backward through the 32 coupling
blocks on tcgen05, Adam) against torch autograd + torch.optim.Adam on the fp32 restatement of the same network (flows.RealNVP)
— the path GLMCMC_NFs.py:63,112-124 takes in the reference."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def trained_flow(seed=0, n_blocks=32):
    from glabc_b200.flows import RealNVP
    torch.manual_seed(seed)
    f = RealNVP(n_blocks=n_blocks, device="cuda")
    with torch.no_grad():
        f.w3.copy_(0.05 * torch.randn_like(f.w3))
        f.b3.copy_(0.02 * torch.randn_like(f.b3))
        f.loc.copy_(torch.tensor([[0.1, -0.2]]))
        f.log_scale.copy_(torch.tensor([[0.05, -0.1]]))
    return f


def posterior_like(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    signs = torch.randint(0, 2, (n, 2), device="cuda", generator=g) * 2 - 1
    return (signs * (1.0 + 0.4 * torch.randn(n, 2, device="cuda", generator=g))).float()


def groups(flow, flat):
    out, off = {}, 0
    for k in flow._FLAT:
        p = getattr(flow, k)
        out[k] = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return out


@pytest.mark.parametrize("n,n_blocks", [(512, 32), (1000, 4), (65536, 32), (777, 1)])
def test_gradient_matches_autograd(n, n_blocks):
    """d forward_kld / d every parameter: the kernels against FLOAT64 autograd of the same network.  Per parameter group the
    relative error (norm) is 6e-7 .. 3e-5 (fp32 autograd: 2e-7 .. 2e-6): the GEMMs run split-precision (~22 significant bits per
    operand) and the backward sweep rebuilds each block's input from its output instead of storing it, so a ReLU whose
    pre-activation sits within rounding of zero can take the other side for one sample — visible in small batches."""
    flow = trained_flow(1, n_blocks)
    x = posterior_like(n, 3)
    eng = flow.train_init()
    g, loss = flow.grad(x, eng)
    f64 = copy.deepcopy(flow).double()
    l64 = f64.forward_kld(x.double())
    l64.backward()
    f32 = copy.deepcopy(flow)
    l32 = f32.forward_kld(x)
    l32.backward()
    assert abs(float(loss) - float(l64.detach())) <= 1e-5 * abs(float(l64.detach())), (float(loss), float(l64.detach()))
    got = groups(flow, g)
    for k in flow._FLAT:
        ref = getattr(f64, k).grad
        err = float((got[k].double() - ref).norm() / ref.norm().clamp_min(1e-30))
        err32 = float((getattr(f32, k).grad.double() - ref).norm() / ref.norm().clamp_min(1e-30))
        assert err < 1e-4, (k, err, err32)
    # element-wise: every entry within 1e-4 of the group's largest gradient
    for k in flow._FLAT:
        ref = getattr(f64, k).grad
        assert float((got[k].double() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-12, k


def test_fifty_adam_steps_follow_torch():
    """50 training steps on fresh batches (GLMCMC_NFs.py:112-124 with train_steps = 50, lr 5e-4, weight_decay 1e-5): the loss
    trajectory and the final weights of the native step against torch autograd + torch.optim.Adam started from the same flow.
    Adam's update is lr * m / sqrt(v) — lr * sign(g) at the first step — so an entry whose gradient sits at rounding level moves
    by +-lr per step whichever way the rounding falls: two correct implementations agree on the losses and on almost every
    weight, not on every weight.  The yardstick is torch itself: the same 50 steps in FLOAT64 autograd, against which torch's
    own float32 run and the native run are both measured — the native run must be about as close as torch float32 is."""
    lr, wd, steps, n = 5e-4, 1e-5, 50, 8192
    flow = trained_flow(2)
    ref = copy.deepcopy(flow)
    ref64 = copy.deepcopy(flow).double()
    start = flow.flat_params().clone()
    target = trained_flow(7)                     # the data: draws of ANOTHER flow, shrunk towards its centre
    opt = torch.optim.Adam(ref.parameters(), lr=lr, weight_decay=wd)
    opt64 = torch.optim.Adam(ref64.parameters(), lr=lr, weight_decay=wd)
    eng = flow.train_init(lr=lr, weight_decay=wd)
    losses, losses_ref, losses_64 = [], [], []
    for s in range(steps):
        with torch.no_grad():
            x = 0.7 * target.sample_from(torch.randn(n, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(100 + s)))[0]
        g, loss = flow.grad(x, eng)
        flow.adam_step(g, loss, eng)
        losses.append(float(loss))
        for model, o, out, xx in ((ref, opt, losses_ref, x), (ref64, opt64, losses_64, x.double())):
            o.zero_grad()
            l = model.forward_kld(xx)
            l.backward()
            o.step()
            out.append(float(l.detach()))
    losses, losses_ref, losses_64 = np.array(losses), np.array(losses_ref), np.array(losses_64)
    assert np.isfinite(losses_64).all() and losses_64[-1] < losses_64[0] - 0.02, losses_64      # it trains
    scale = np.abs(losses_64).max()
    assert np.abs(losses - losses_64).max() <= 1e-4 * scale, np.abs(losses - losses_64).max()
    assert np.abs(losses - losses_ref).max() <= 1e-4 * scale
    a, b, c = flow.flat_params().double(), ref.flat_params().double(), ref64.flat_params()
    d_native, d_torch32 = (a - c).abs(), (b - c).abs()
    within = lambda d: float((d <= 1e-4).double().mean())   # noqa: E731
    print(f"weights within 1e-4 of the float64 run: native {within(d_native):.5f}, torch float32 {within(d_torch32):.5f}; "
          f"p99.9 |diff| native {float(d_native.quantile(0.999)):.2e} torch32 {float(d_torch32.quantile(0.999)):.2e}; "
          f"max native {float(d_native.max()):.2e} torch32 {float(d_torch32.max()):.2e}")
    assert within(d_native) >= 0.99 and within(d_native) >= within(d_torch32) - 0.01
    assert float(d_native.max()) <= 2 * lr * steps
    moved = (c - start.double()).abs()
    assert float(moved.median()) > 10 * float(d_native.median())                 # the agreement is not that of two unmoved flows
    # the kernels evaluate the updated weights (packed W2 and base parameters refreshed after every step)
    eps = torch.randn(4096, 2, device="cuda")
    th, lq = flow.fused_sample_from(eps, eng, precision="precise")
    with torch.no_grad():
        th_r, lq_r = flow.sample_from(eps)
    assert torch.allclose(th, th_r, rtol=1e-4, atol=1e-4) and torch.allclose(lq, lq_r, rtol=1e-4, atol=1e-4)


def test_non_finite_loss_leaves_the_flow_untouched():
    """GLMCMC_NFs.py:120-122: backward() is skipped for a NaN / inf loss and Adam.step() then has nothing to apply"""
    flow = trained_flow(3, 4)
    eng = flow.train_init()
    before = flow.flat_params().clone()
    x = posterior_like(256, 5)
    x[7, 0] = float("nan")
    g, loss = flow.grad(x, eng)
    assert not np.isfinite(float(loss))
    flow.adam_step(g, loss, eng)
    assert torch.equal(flow.flat_params(), before)
    x[7, 0] = 0.3
    g, loss = flow.grad(x, eng)
    flow.adam_step(g, loss, eng)
    assert np.isfinite(float(loss)) and not torch.equal(flow.flat_params(), before)


def test_training_is_deterministic():
    """per-CTA partial gradients folded in CTA order, fixed-order column sums: two runs give bit-identical gradients"""
    flow = trained_flow(4, 8)
    x = posterior_like(5000, 9)
    eng = flow.train_init()
    g1, l1 = flow.grad(x, eng)
    g2, l2 = flow.grad(x, eng)
    assert torch.equal(g1, g2) and torch.equal(l1, l2)
