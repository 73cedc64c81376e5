"""GPU tests of the device versions of the reference's distribution classes (distribution.py: Uniform, Gamma,
DiagGaussian, GaussianMixture) and of GlobalMCMC with them as proposals (the general kernel, csrc/step_generic.cuh)."""
import os

import numpy as np
import pytest
import torch
from scipy import stats as sst

from helpers import GOLDEN, abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from glabc_b200.engine import Engine
    return Engine()


@pytest.fixture(scope="module")
def dists():
    import glabc_b200 as g
    z = np.load(os.path.join(GOLDEN, "misc.npz"))
    t = torch.from_numpy
    return z, dict(
        diag=g.DiagGaussian(2, t(z["diag/loc"]), t(z["diag/log_scale"])),
        uniform=g.Uniform(2, t(z["uniform/low"]), t(z["uniform/high"])),
        gamma=g.Gamma(t(np.asarray(z["gamma/shape"])), t(np.asarray(z["gamma/rate"]))),
        mix=g.GaussianMixture(3, 2, loc=z["mix/loc"], scale=np.exp(z["mix/log_scale"]), weights=np.exp(z["mix/weight_scores"])))


@pytest.mark.parametrize("name", ["diag", "uniform", "gamma", "mix"])
def test_log_prob_matches_the_reference(eng, dists, name):
    """device log_prob against the values the reference's own classes produced (tests/golden/misc.npz), incl. -inf
    outside the Uniform box / Gamma support; float32 vs the reference's float64 for Gamma and GaussianMixture"""
    z, d = dists
    eng.bind_proposal(abi.SLOT_GLOBAL, d[name])
    got = eng.dist_log_prob(abi.SLOT_GLOBAL, torch.from_numpy(z[f"{name}/z"]).float()).cpu().numpy().astype(np.float64)
    want = z[f"{name}/log_prob"].astype(np.float64)
    assert np.array_equal(np.isneginf(got), np.isneginf(want))
    fin = np.isfinite(want)
    assert fin.any() and np.allclose(got[fin], want[fin], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("name", ["diag", "uniform", "gamma", "mix"])
def test_forward_draws_follow_the_law(eng, dists, name):
    """device forward(): the returned log-density is log_prob of the returned draw, and the draws follow the distribution"""
    _, d = dists
    dist = d[name]
    eng.bind_proposal(abi.SLOT_LOCAL, dist)
    n = 200000
    zs, lp = eng.dist_sample(abi.SLOT_LOCAL, n, 2, seed=7)
    assert torch.allclose(lp, eng.dist_log_prob(abi.SLOT_LOCAL, zs), rtol=2e-5, atol=2e-5)
    x = zs.cpu().numpy().astype(np.float64)
    if name == "uniform":
        lo, hi = dist.low.numpy(), dist.high.numpy()
        for i in range(2):
            assert sst.kstest(x[:, i], sst.uniform(lo[i], hi[i] - lo[i]).cdf).statistic < 0.005
    elif name == "gamma":
        a, b = dist.Shape.numpy(), dist.Rate.numpy()
        for i in range(2):
            assert sst.kstest(x[:, i], sst.gamma(a[i], scale=1 / b[i]).cdf).statistic < 0.005
    elif name == "diag":
        loc, sc = dist.loc.reshape(-1).numpy(), np.exp(dist.log_scale.reshape(-1).numpy())
        for i in range(2):
            assert sst.kstest(x[:, i], sst.norm(loc[i], sc[i]).cdf).statistic < 0.005
    else:
        w = torch.softmax(dist.weight_scores, 1)[0].numpy()
        loc, sc = dist.loc[0].numpy(), np.exp(dist.log_scale[0].numpy())
        cdf0 = lambda v: sum(w[m] * sst.norm(loc[m, 0], sc[m, 0]).cdf(v) for m in range(3))  # noqa: E731
        assert sst.kstest(x[:, 0], cdf0).statistic < 0.005
        # importance-sampling identity E_q[p/q] = 1 with p = a narrow Gaussian inside mode 0 (bounded ratio): the
        # returned density is the normalised density of the draws
        logp = (sst.norm(loc[0, 0], 0.5 * sc[0, 0]).logpdf(x[:, 0]) + sst.norm(loc[0, 1], 0.5 * sc[0, 1]).logpdf(x[:, 1]))
        assert abs(np.exp(logp - lp.cpu().numpy().astype(np.float64)).mean() - 1.0) < 0.03
    small = eng.dist_sample(abi.SLOT_LOCAL, 5, 2, seed=8)[0]
    assert not torch.equal(small, zs[:5])
    if name == "gamma":   # shape < 1 takes the U^(1/a) boost
        import glabc_b200 as g
        eng.bind_proposal(abi.SLOT_LOCAL, g.Gamma(torch.tensor([0.4, 0.9]), torch.tensor([2.0, 1.0])))
        xs = eng.dist_sample(abi.SLOT_LOCAL, n, 2, seed=9)[0].cpu().numpy().astype(np.float64)
        assert sst.kstest(xs[:, 0], sst.gamma(0.4, scale=0.5).cdf).statistic < 0.006


def readme():
    import glabc_b200 as g
    model = g.Mixture_set(epsilon=0.05)
    lp = g.DiagGaussian(2, loc=torch.zeros(1, 2), log_scale=torch.log(torch.tensor([0.35, 0.35])))
    return g, model, lp


def check_posterior(theta, tol=0.04):
    a = theta.abs().cpu().numpy().astype(np.float64)
    for i in range(2):
        assert sst.kstest(a[:, i], sst.norm(1.42518, np.sqrt(0.049881)).cdf).statistic < tol
    quad = ((theta[:, 0] > 0).long() * 2 + (theta[:, 1] > 0).long()).bincount(minlength=4).cpu().numpy() / theta.shape[0]
    assert np.abs(quad - 0.25).max() < tol


def test_global_mcmc_with_mixture_and_uniform_proposals(eng):
    """GlobalMCMC.py:37-68 with non-Gaussian proposals: a 4-mode GaussianMixture independence proposal (accepts several times
    more often than N(0, I); the ABC kernel of the fresh simulation caps the rate), a Uniform box independence proposal, and a Uniform random-walk local proposal —
    each must leave the closed-form ABC posterior invariant"""
    g, model, lp = readme()
    modes = [[1.425, 1.425], [1.425, -1.425], [-1.425, 1.425], [-1.425, -1.425]]
    gm = g.GaussianMixture(4, 2, loc=modes, scale=[[0.3, 0.3]] * 4, weights=[1, 1, 1, 1])
    _, st = g.GlobalMCMC(model, 1501, torch.zeros(2), None, gm, None, 0.5, lp, num_chains=8192, seed=2, trace="none", return_stats=True)
    acc_g = float((st.accepted_global / st.global_steps.clamp(min=1)).mean())
    assert acc_g > 0.02, acc_g                      # N(0, I) proposal: ~0.5 %
    out = g.GlobalMCMC(model, 1501, torch.zeros(2), None, gm, None, 0.5, lp, num_chains=8192, seed=3, trace="time")
    check_posterior(out[-1])
    box = g.Uniform(2, torch.tensor([-3.0, -3.0]), torch.tensor([3.0, 3.0]))
    out = g.GlobalMCMC(model, 6001, torch.zeros(2), None, box, None, 0.5, lp, num_chains=8192, seed=4, trace="time")
    check_posterior(out[-1], tol=0.05)
    rw = g.Uniform(2, torch.tensor([-0.6, -0.6]), torch.tensor([0.6, 0.6]))      # symmetric box random walk
    out = g.GlobalMCMC(model, 1501, torch.zeros(2), None, gm, None, 0.3, rw, num_chains=8192, seed=5, trace="time")
    check_posterior(out[-1])


def test_non_gaussian_proposals_are_refused_where_not_fused(eng):
    """the STRICT / replay GLMALA and AGLMCMC kernels are fused for a DiagGaussian importance proposal, and AGLMCMC for a
    DiagGaussian Local_Proposal: any other kind is refused with a message, not run through some other path
    (run_global_mcmc, run_glmcmc and the FAST native run_mala / run_aglmcmc take every kind of importance proposal)"""
    g, model, lp = readme()
    box = g.Uniform(2, torch.tensor([-3.0, -3.0]), torch.tensor([3.0, 3.0]))
    with pytest.raises(abi.GlabcError, match="DiagGaussian"):
        g.GLMALA(model, 100, torch.zeros(2), None, 0.3, 10, None, 0.8, box, 5, num_chains=64, arith="strict")
    with pytest.raises(abi.GlabcError, match="DiagGaussian"):
        g.AGLMCMC(model, 100, torch.zeros(2), None, lp, box, None, 1.0, 10, 5, 0.8, 0.2, num_chains=64, arith="strict")
    with pytest.raises(abi.GlabcError, match="DiagGaussian"):
        g.AGLMCMC(model, 100, torch.zeros(2), None, box, lp, None, 0.5, 10, 5, 0.8, 0.2, num_chains=64)


def test_aglmcmc_with_non_gaussian_initial_proposal(eng):
    """run_aglmcmc (AGLMCMC.py:84-112,137-149) with a Uniform / GaussianMixture Initial_ISIR_prop: the initial candidate
    block and the first rounds' weights use it, the per-chain KDEs take over afterwards; the closed-form ABC posterior
     is left invariant"""
    g, model, lp = readme()
    box = g.Uniform(2, torch.tensor([-3.0, -3.0]), torch.tensor([3.0, 3.0]))
    out = g.AGLMCMC(model, 3001, torch.zeros(2), None, lp, box, None, 1.0, 200, 5, 0.8, 0.2, num_chains=4096, seed=2, trace="time")
    check_posterior(out[-1], tol=0.04)
    modes = [[1.425, 1.425], [1.425, -1.425], [-1.425, 1.425], [-1.425, -1.425]]
    gm = g.GaussianMixture(4, 2, loc=modes, scale=[[0.5, 0.5]] * 4, weights=[1, 1, 1, 1])
    out, st = g.AGLMCMC(model, 2001, torch.zeros(2), None, lp, gm, None, 0.9, 100, 5, 0.8, 0.2, num_chains=4096, seed=3, trace="time",
                        return_stats=True)
    check_posterior(out[-1], tol=0.04)
    assert float(st.move_rate.mean()) > 0.005


def test_pooled_aglmcmc_with_uniform_initial_proposal(eng):
    """AGLMCMC(pooled=True) (BASELINE config 5) with a Uniform Initial_ISIR_prop: the pooled path draws the initial blocks
    through glabc_dist_sample / glabc_dist_log_prob, which take every class; posterior check as the DiagGaussian case"""
    g, model, lp = readme()
    box = g.Uniform(2, torch.tensor([-3.0, -3.0]), torch.tensor([3.0, 3.0]))
    out = g.AGLMCMC(model, 1001, torch.zeros(2), None, lp, box, None, 0.9, 50, 5, 0.8, 0.2, num_chains=2048, seed=5, trace="time",
                    pooled=True, kde_train=20000)
    check_posterior(out[-1], tol=0.05)


def test_glmala_with_non_gaussian_importance_proposals(eng):
    """run_mala (GLMALA.py:151-180) with a GaussianMixture / Uniform Importance_Proposal through k_mala_fast<GIP>: the
    closed-form ABC posterior is left invariant, the mode-covering mixture moves far more often than
    N(0, I) does, and two launches continue one launch bit-identically"""
    g, model, lp = readme()
    modes = [[1.425, 1.425], [1.425, -1.425], [-1.425, 1.425], [-1.425, -1.425]]
    gm = g.GaussianMixture(4, 2, loc=modes, scale=[[0.3, 0.3]] * 4, weights=[1, 1, 1, 1])
    out, st = g.GLMALA(model, 801, torch.zeros(2), None, 0.15, 10, None, 0.5, gm, 5, num_chains=8192, seed=2, trace="time",
                       return_stats=True)
    check_posterior(out[-1])
    assert float(st.move_rate.mean()) > 0.05
    box = g.Uniform(2, torch.tensor([-3.0, -3.0]), torch.tensor([3.0, 3.0]))
    out = g.GLMALA(model, 2501, torch.zeros(2), None, 0.15, 10, None, 0.5, box, 8, num_chains=8192, seed=3, trace="time")
    check_posterior(out[-1], tol=0.05)
    eng.bind_model(model)
    eng.bind_proposal(abi.SLOT_IMPORTANCE, gm)

    def fresh():
        aux = torch.zeros(70, abi.AUX_SLOTS, device="cuda")
        aux[:, abi.AUX_LOCAL] = 1.0
        y0 = torch.randn(70, 2, generator=torch.Generator().manual_seed(1)).cuda() * 0.2236
        return torch.zeros(70, 2, device="cuda"), y0, aux, torch.zeros(70, abi.STATE64_SLOTS, dtype=torch.float64, device="cuda")
    kw = dict(gf=0.6, seed=9, K=4, num_grad=6, tau=0.2, trace_layout=abi.TRACE_TIME_MAJOR)
    th, yy, ax, s64 = fresh()
    full = eng.run("mala", theta=th, y=yy, aux=ax, state64=s64, n_steps=120, **kw)
    th2, yy2, ax2, s642 = fresh()
    buf = torch.zeros(121, 70, 2, device="cuda")
    eng.run("mala", theta=th2, y=yy2, aux=ax2, state64=s642, n_steps=50, trace=buf, trace_rows=121, **kw)
    eng.run("mala", theta=th2, y=yy2, aux=ax2, state64=s642, n_steps=70, step_base=50, trace=buf, trace_rows=121, write_row0=False, **kw)
    assert torch.equal(buf, full) and torch.equal(th2, th) and torch.equal(ax2, ax)


def test_glmcmc_with_non_gaussian_proposals(eng):
    """run_glmcmc (GLMCMC.py:58-104) with a Uniform / GaussianMixture importance proposal and a Uniform local proposal through
    the general iSIR kernel (k_isir_generic): the closed-form ABC posterior is left invariant, the
    mode-covering mixture proposal moves far more often than N(0, I), and chunked launches continue bit-identically"""
    g, model, lp = readme()
    modes = [[1.425, 1.425], [1.425, -1.425], [-1.425, 1.425], [-1.425, -1.425]]
    gm = g.GaussianMixture(4, 2, loc=modes, scale=[[0.3, 0.3]] * 4, weights=[1, 1, 1, 1])
    out, st = g.GLMCMC(model, 1201, torch.zeros(2), None, lp, None, 0.9, gm, 5, num_chains=8192, seed=2, trace="time", return_stats=True)
    check_posterior(out[-1])
    assert float(st.move_rate.mean()) > 0.05                      # N(0, I) importance proposal: 0.9 %
    box = g.Uniform(2, torch.tensor([-3.0, -3.0]), torch.tensor([3.0, 3.0]))
    rw = g.Uniform(2, torch.tensor([-0.6, -0.6]), torch.tensor([0.6, 0.6]))
    out = g.GLMCMC(model, 4001, torch.zeros(2), None, rw, None, 0.8, box, 8, num_chains=8192, seed=3, trace="time")
    check_posterior(out[-1], tol=0.05)
    # one launch == two launches (aux carries the cached log-weight and the flags)
    eng.bind_model(model)
    eng.bind_proposal(abi.SLOT_LOCAL, rw)
    eng.bind_proposal(abi.SLOT_IMPORTANCE, gm)

    def fresh():
        aux = torch.zeros(70, abi.AUX_SLOTS, device="cuda")
        aux[:, abi.AUX_LOCAL] = 1.0
        y0 = torch.randn(70, 2, generator=torch.Generator().manual_seed(1)).cuda() * 0.2236
        return torch.zeros(70, 2, device="cuda"), y0, aux
    th, yy, ax = fresh()
    full = eng.run("isir", theta=th, y=yy, aux=ax, n_steps=200, gf=0.7, seed=9, K=4, trace_layout=abi.TRACE_TIME_MAJOR)
    th2, yy2, ax2 = fresh()
    buf = torch.zeros(201, 70, 2, device="cuda")
    eng.run("isir", theta=th2, y=yy2, aux=ax2, n_steps=90, gf=0.7, seed=9, K=4, trace=buf, trace_rows=201, trace_layout=abi.TRACE_TIME_MAJOR)
    eng.run("isir", theta=th2, y=yy2, aux=ax2, n_steps=110, step_base=90, gf=0.7, seed=9, K=4, trace=buf, trace_rows=201,
            trace_layout=abi.TRACE_TIME_MAJOR, write_row0=False)
    assert torch.equal(buf, full) and torch.equal(th2, th) and torch.equal(ax2, ax)


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_glmcmc_replay_with_reference_recordings(eng, ci):
    """tests/golden/glmcmc_generic.npz: the REFERENCE's GLMCMC run with Uniform / GaussianMixture / Gamma proposals in the
    Local / Importance slots, every draw recorded (incl. float64-weight steps after the state was promoted, and `None`
    resamples of underflowed float32 weights).  Fed the same draws, the general iSIR kernel must take the same branch, the same
    resample index (or None) and the same accept decision at every step, evaluate the weights in the same dtype, and write the
    same float32 trace bit for bit; log-weights / log-densities within 1e-5 of their terms."""
    import glabc_b200 as g
    z = np.load(os.path.join(GOLDEN, "glmcmc_generic.npz"))
    p = lambda k: z[f"case{ci}/{k}"]  # noqa: E731
    model = g.Mixture_set(float(p("epsilon")))
    eng.bind_model(model)
    eng.bind_proposal(abi.SLOT_LOCAL, _dist_from_golden(z, ci, "lp"))
    eng.bind_proposal(abi.SLOT_IMPORTANCE, _dist_from_golden(z, ci, "ip"))
    T, Cn, K = int(p("T")), p("theta0").shape[0], int(p("K"))
    theta, y = torch.from_numpy(p("theta0")).cuda(), torch.from_numpy(p("y0")).cuda()
    aux = torch.zeros(Cn, abi.AUX_SLOTS, device="cuda")
    aux[:, abi.AUX_LOCAL] = 1.0
    tape32, tape64 = torch.from_numpy(p("tape32")).cuda().contiguous(), torch.from_numpy(p("tape64")).cuda().contiguous()
    debug = torch.zeros(T - 1, abi.DEBUG_SLOTS, Cn, device="cuda")
    got = eng.run("isir", theta=theta, y=y, aux=aux, n_steps=T - 1, gf=float(p("gf")), K=K, rng_mode=abi.RNG_REPLAY, tape32=tape32,
                  tape64=tape64, debug=debug, trace_layout=abi.TRACE_TIME_MAJOR).cpu().numpy()
    dbg, rec = debug.cpu().numpy().astype(np.float64), p("rec")
    assert np.array_equal(dbg[:, 0], rec[:, 0])                 # branch, move, resample index (+1; 0 = None), weight dtype
    assert np.array_equal(got, p("trace"))                      # the chains, bit for bit
    flags = rec[:, 0].astype(int)
    glob = (flags & 1) == 1
    for j in range(K):                                          # candidate log-weights: prior + log K - log q, O(1..100) terms
        a, b = dbg[:, 4 + j][glob], rec[:, 4 + j][glob]
        fin = np.isfinite(b)
        assert np.array_equal(np.isneginf(a), np.isneginf(b))
        assert np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 10.0)) < 1e-5
    for slot, name in ((1, "prior'"), (2, "kernel'"), (3, "log_acc")):
        a, b = dbg[:, slot][~glob], rec[:, slot][~glob]
        scale = np.maximum(np.abs(b), 1.0) if slot != 3 else np.maximum(np.abs(rec[:, 2][~glob]) + np.abs(b), 1.0)
        assert np.max(np.abs(a - b) / scale) < 1e-5, name
    a, b = dbg[:, 1][glob], rec[:, 1][glob]                     # log-weight of the current state
    assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 10.0)) < 1e-5
    assert glob.any() and (~glob).any() and ((flags >> 1) & 1).sum() > 20


def _dist_from_golden(z, ci, prefix):
    import glabc_b200 as g
    kind = ["gauss", "uniform", "gamma", "mixture"][int(z[f"case{ci}/{prefix}_kind"])]
    get = lambda k: z[f"case{ci}/{prefix}_{k}"]  # noqa: E731
    t = lambda k: torch.from_numpy(get(k)).float()  # noqa: E731
    if kind == "gauss":
        return g.DiagGaussian(2, t("loc").view(1, 2), torch.log(t("sigma")))
    if kind == "uniform":
        return g.Uniform(2, low=t("low"), high=t("high"))
    if kind == "gamma":
        return g.Gamma(t("shape"), t("rate"))
    return g.GaussianMixture(get("loc").shape[0], 2, loc=get("loc"), scale=get("scale"), weights=get("weights"))


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_global_mcmc_replay_with_reference_recordings(eng, ci):
    """tests/golden/global_generic.npz: the REFERENCE's GlobalMCMC run with GaussianMixture / Uniform / Gamma proposals
    (distribution.py:50-137,206-293), every draw recorded.  Fed the same draws, the general kernel must take the same
    branch and the same accept / reject decision at every step, write the same float32 trace bit for bit (float64 state
    promotion reproduced), and agree on log prior, log kernel and log_acc within 1e-5 relative."""
    import glabc_b200 as g
    z = np.load(os.path.join(GOLDEN, "global_generic.npz"))
    p = lambda k: z[f"case{ci}/{k}"]  # noqa: E731
    model = g.Mixture_set(float(p("epsilon")))
    eng.bind_model(model)
    eng.bind_proposal(abi.SLOT_LOCAL, _dist_from_golden(z, ci, "lp"))
    eng.bind_proposal(abi.SLOT_GLOBAL, _dist_from_golden(z, ci, "gp"))
    T, Cn = int(p("T")), p("theta0").shape[0]
    theta, y = torch.from_numpy(p("theta0")).cuda(), torch.from_numpy(p("y0")).cuda()
    tape32, tape64 = torch.from_numpy(p("tape32")).cuda().contiguous(), torch.from_numpy(p("tape64")).cuda().contiguous()
    debug = torch.zeros(T - 1, abi.DEBUG_SLOTS, Cn, device="cuda")
    got = eng.run("global", theta=theta, y=y, n_steps=T - 1, gf=float(p("gf")), rng_mode=abi.RNG_REPLAY, tape32=tape32,
                  tape64=tape64, debug=debug, trace_layout=abi.TRACE_TIME_MAJOR).cpu().numpy()
    dbg, rec = debug.cpu().numpy().astype(np.float64), p("rec")
    assert np.array_equal(dbg[:, 0], rec[:, 0])                 # branch + accept decision of every step
    assert np.array_equal(got, p("trace"))                      # the chains, bit for bit
    for slot, name in ((1, "log prior"), (2, "log kernel"), (3, "log_acc")):
        a, b = dbg[:, slot], rec[:, slot]
        fin = np.isfinite(b)
        assert np.array_equal(np.isneginf(a), np.isneginf(b)), name
        # log_acc is a difference of O(100) terms: relative to the size of the terms, not of the result
        scale = np.maximum(np.abs(b[fin]), 1.0) if slot != 3 else np.maximum(np.abs(rec[:, 2][fin]) + np.abs(b[fin]), 1.0)
        assert np.max(np.abs(a[fin] - b[fin]) / scale) < 1e-5, name
    flags = rec[:, 0].astype(int)
    assert ((flags & 1) == 1).any() and ((flags & 1) == 0).any() and ((flags >> 1) & 1).sum() > 50
