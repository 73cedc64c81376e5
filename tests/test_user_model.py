"""User-supplied models compiled at run time into the fused GlobalMCMC kernel (csrc/user_model.cu).
CPU: the source assembles and compiles with NVRTC for sm_100a (no GPU needed), compile errors surface with the log.
GPU: the README Mixture model written as a user model reproduces the closed-form ABC posterior and the statistics of the
built-in fused family; a model OUTSIDE the family (a 3-parameter g-and-k-like quantile simulator) runs."""
import numpy as np
import pytest
import torch

MIXTURE_SRC = r"""
// examples/Mixture.py:13-36 as device functions; params = {y_obs0, y_obs1, noise_sd}
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

QUANTILE_SRC = r"""
// a model outside the built-in family: theta = (A, B, g), y = 5 order statistics-like summaries of a skewed
// location-scale quantile function evaluated at noise draws; uniform prior box; L1 discrepancy
__device__ void glabc_user_simulate(const float* theta, const float* noise, const float* params, float* y)
{
    for (int k = 0; k < 5; ++k) {
        const float z = noise[k];
        const float skew = (1.0f - __expf(-theta[2] * z)) / (1.0f + __expf(-theta[2] * z));
        y[k] = theta[0] + theta[1] * (1.0f + 0.8f * skew) * z;
    }
}
__device__ float glabc_user_prior_log_prob(const float* theta, const float* params)
{
    const bool in = theta[0] > 0.f && theta[0] < 10.f && theta[1] > 0.f && theta[1] < 10.f && theta[2] > 0.f && theta[2] < 10.f;
    return in ? -6.907755f : -INFINITY;
}
__device__ float glabc_user_discrepancy(const float* y, const float* params)
{
    float s = 0.f;
    for (int k = 0; k < 5; ++k) s += fabsf(y[k] - params[k]);
    return s * 0.2f;
}
"""


def test_user_model_compiles_without_a_gpu():
    import glabc_b200 as g
    m = g.UserModel(MIXTURE_SRC, theta_dim=2, y_dim=2, n_noise=2, epsilon=0.05, params=[1.5, 1.5, 0.05 ** 0.5])
    assert m.check()
    assert g.UserModel(QUANTILE_SRC, theta_dim=3, y_dim=5, n_noise=5, epsilon=0.5, params=[3.0] * 5).check()
    bad = g.UserModel(MIXTURE_SRC.replace("fabsf(theta[1])", "fabsf(theta[1]) +* 2"), 2, 2, 2, 0.05, [1.5, 1.5, 0.2])
    with pytest.raises(ValueError, match="does not compile"):
        bad.check()
    with pytest.raises(ValueError, match="glabc_user_discrepancy"):
        g.UserModel(MIXTURE_SRC.replace("glabc_user_discrepancy", "my_discrepancy"), 2, 2, 2, 0.05, [1.5, 1.5, 0.2]).check()


@pytest.mark.gpu
def test_user_mixture_matches_closed_form_and_builtin():
    from scipy import stats as sst
    import glabc_b200 as g
    lp = g.DiagGaussian(2, torch.zeros(1, 2), torch.log(torch.tensor([0.35, 0.35])))
    gp = g.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    um = g.UserModel(MIXTURE_SRC, theta_dim=2, y_dim=2, n_noise=2, epsilon=0.05, params=[1.5, 1.5, 0.05 ** 0.5])
    Cn, T = 16384, 6001
    y0 = torch.randn(Cn, 2, generator=torch.Generator().manual_seed(1)) * 0.2236
    out, st = g.GlobalMCMC(um, T, torch.zeros(2), y0, gp, None, 0.5, lp, num_chains=Cn, seed=3, trace="time", return_stats=True)
    assert out.shape == (T, Cn, 2) and torch.equal(out[0], torch.zeros(Cn, 2, device="cuda"))
    a = out[-1].abs().cpu().numpy().astype(np.float64)
    for i in range(2):   # |theta_i| ~ N(1.42518, 0.049881)
        assert sst.kstest(a[:, i], sst.norm(1.42518, np.sqrt(0.049881)).cdf).statistic < 0.02
    quad = ((out[-1][:, 0] > 0).long() * 2 + (out[-1][:, 1] > 0).long()).bincount(minlength=4).cpu().numpy() / Cn
    assert np.abs(quad - 0.25).max() < 0.02
    # same law as the built-in fused family: move rate and ESJD agree within sampling noise
    _, st2 = g.GlobalMCMC(g.Mixture_set(0.05), T, torch.zeros(2), y0, gp, None, 0.5, lp, num_chains=Cn, seed=4, trace="none",
                          return_stats=True)
    assert abs(float(st.move_rate.mean()) / float(st2.move_rate.mean()) - 1) < 0.03
    assert abs(float(st.global_steps.mean()) / (T - 1) - 0.5) < 0.005
    assert abs(float(st.esjd().mean()) / float(st2.esjd().mean()) - 1) < 0.06
    # layouts, chunked continuation and shards give the same chains (Philox keyed by global chain id and step)
    from glabc_b200 import _abi as abi
    from glabc_b200.engine import get_engine
    eng = get_engine()
    th, yy = torch.zeros(64, 2, device="cuda"), y0[:64].cuda()
    full = eng.run_user(um, theta=th, y=yy, n_steps=200, gf=0.5, seed=9, trace_layout=abi.TRACE_TIME_MAJOR)
    th2, yy2 = torch.zeros(64, 2, device="cuda"), y0[:64].cuda()
    buf = torch.zeros(64, 201, 2, device="cuda")
    eng.run_user(um, theta=th2, y=yy2, n_steps=77, gf=0.5, seed=9, trace=buf, trace_rows=201, trace_layout=abi.TRACE_CHAIN_MAJOR)
    r = abi.RunPOD  # continuation: rows 78..200
    th3 = th2.clone()
    rest = eng.run_user(um, theta=th2, y=yy2, n_steps=123, step_base=77, gf=0.5, seed=9, trace_layout=abi.TRACE_TIME_MAJOR,
                        trace_rows=201, trace=torch.zeros(201, 64, 2, device="cuda"), write_row0=False)
    assert torch.equal(buf[:, :78].permute(1, 0, 2), full[:78]) and torch.equal(rest[78:], full[78:]) and torch.equal(th2, th)
    lo = eng.run_user(um, theta=torch.zeros(32, 2, device="cuda"), y=y0[32:64].cuda(), n_steps=200, gf=0.5, seed=9, chain_id_base=32,
                      trace_layout=abi.TRACE_TIME_MAJOR)
    assert torch.equal(lo, full[:, 32:])
    del th3, r


@pytest.mark.gpu
def test_user_model_outside_the_family_runs(tmp_path):
    import glabc_b200 as g
    lp = g.DiagGaussian(3, torch.zeros(1, 3), torch.log(torch.tensor([0.2, 0.2, 0.2])))
    gp = g.DiagGaussian(3, torch.tensor([3.0, 1.0, 2.0]), torch.log(torch.tensor([1.0, 0.5, 1.0])))
    obs = [3.0 + 1.0 * z for z in (-1.2, -0.5, 0.0, 0.6, 1.4)]
    um = g.UserModel(QUANTILE_SRC, theta_dim=3, y_dim=5, n_noise=5, epsilon=0.5, params=obs)
    theta0 = torch.tensor([3.0, 1.0, 2.0])
    y0 = torch.tensor(obs)
    out, st = g.GlobalMCMC(um, 3001, theta0, y0, gp, None, 0.3, lp, num_chains=4096, seed=1, trace="time", return_stats=True)
    last = out[-1]
    assert bool(((last > 0) & (last < 10)).all())                         # the prior box is respected
    assert 0.01 < float(st.move_rate.mean()) < 0.9
    assert abs(float(last[:, 0].mean()) - 3.0) < 0.5                       # location is identified by the summaries
    runner = g.MCMCRunner(um, output_dir=str(tmp_path))                   # the public façade accepts the model object
    chain = runner.run_global_mcmc(500, theta0, y0, 0.3, lp, gp, output_file="user.csv", verbose=False)
    assert chain.shape == (500, 3) and (tmp_path / "user.csv").exists()
    with pytest.raises(Exception, match="does not compile"):
        g.GlobalMCMC(g.UserModel("__device__ int nothing;", 2, 2, 2, 0.05), 10, torch.zeros(2), torch.zeros(1, 2),
                     g.DiagGaussian(2, torch.zeros(2), torch.zeros(2)), None, 0.5,
                     g.DiagGaussian(2, torch.zeros(1, 2), torch.zeros(2)))


@pytest.mark.gpu
def test_user_model_glmcmc_isir():
    """run_glmcmc with a user model (glabc_run_isir_user): the README model as CUDA source reproduces the closed-form
    posterior and the statistics of the built-in iSIR kernel; chunked continuation carries the cached log-weight."""
    from scipy import stats as sst
    import glabc_b200 as g
    from glabc_b200 import _abi as abi
    from glabc_b200.engine import get_engine
    lp = g.DiagGaussian(2, torch.zeros(1, 2), torch.log(torch.tensor([0.35, 0.35])))
    ip = g.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    um = g.UserModel(MIXTURE_SRC, theta_dim=2, y_dim=2, n_noise=2, epsilon=0.05, params=[1.5, 1.5, 0.05 ** 0.5])
    Cn, T = 16384, 3001
    y0 = torch.randn(Cn, 2, generator=torch.Generator().manual_seed(1)) * 0.2236
    out, st = g.GLMCMC(um, T, torch.zeros(2), y0, lp, None, 0.9, ip, 5, num_chains=Cn, seed=3, trace="time", return_stats=True)
    a = out[-1].abs().cpu().numpy().astype(np.float64)
    for i in range(2):
        assert sst.kstest(a[:, i], sst.norm(1.42518, np.sqrt(0.049881)).cdf).statistic < 0.02
    _, st2 = g.GLMCMC(g.Mixture_set(0.05), T, torch.zeros(2), y0, lp, None, 0.9, ip, 5, num_chains=Cn, seed=4, trace="none",
                      return_stats=True)
    assert abs(float(st.move_rate.mean()) / float(st2.move_rate.mean()) - 1) < 0.05
    assert abs(float(st.global_steps.mean()) / (T - 1) - 0.9) < 0.005
    assert abs(float(st.esjd().mean()) / float(st2.esjd().mean()) - 1) < 0.08
    runner = g.MCMCRunner(um)
    chain = runner.run_glmcmc(300, torch.zeros(2), y0[:1], 0.9, lp, ip, 5, output_file=None, verbose=False)
    assert chain.shape == (300, 2)
    # one launch == two chunks (the aux state carries the cached log-weight and the `local` flag)
    eng = get_engine()
    eng.bind_proposal(abi.SLOT_LOCAL, lp)
    eng.bind_proposal(abi.SLOT_IMPORTANCE, ip)

    def fresh():
        aux = torch.zeros(64, abi.AUX_SLOTS, device="cuda")
        aux[:, abi.AUX_LOCAL] = 1.0
        return torch.zeros(64, 2, device="cuda"), y0[:64].cuda(), aux
    th, yy, ax = fresh()
    full = eng.run_user(um, theta=th, y=yy, aux=ax, n_steps=200, gf=0.8, seed=9, K=4, sampler="isir", trace_layout=abi.TRACE_TIME_MAJOR)
    th2, yy2, ax2 = fresh()
    buf = torch.zeros(201, 64, 2, device="cuda")
    eng.run_user(um, theta=th2, y=yy2, aux=ax2, n_steps=90, gf=0.8, seed=9, K=4, sampler="isir", trace=buf, trace_rows=201,
                 trace_layout=abi.TRACE_TIME_MAJOR)
    eng.run_user(um, theta=th2, y=yy2, aux=ax2, n_steps=110, step_base=90, gf=0.8, seed=9, K=4, sampler="isir", trace=buf, trace_rows=201,
                 trace_layout=abi.TRACE_TIME_MAJOR, write_row0=False)
    assert torch.equal(buf, full) and torch.equal(th2, th) and torch.equal(ax2[:, :2], ax[:, :2])


@pytest.mark.gpu
def test_user_model_glmala():
    """run_glmala with a user model (glabc_run_mala_user: a thread per chain, the finite-difference gradient of
    GLMALA.py:46-95 dealt over the warp): the README model as CUDA source reproduces the closed-form posterior and the
    statistics of the built-in GLMALA kernel; chunks and shards give the same chains."""
    from scipy import stats as sst
    import glabc_b200 as g
    from glabc_b200 import _abi as abi
    from glabc_b200.engine import get_engine
    ip = g.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    um = g.UserModel(MIXTURE_SRC, theta_dim=2, y_dim=2, n_noise=2, epsilon=0.05, params=[1.5, 1.5, 0.05 ** 0.5])
    Cn, T = 8192, 2501
    y0 = torch.randn(Cn, 2, generator=torch.Generator().manual_seed(1)) * 0.2236
    out, st = g.GLMALA(um, T, torch.zeros(2), y0, 0.3, 100, None, 0.8, ip, 5, num_chains=Cn, seed=3, trace="time", return_stats=True)
    assert out.shape == (T, Cn, 2)
    a = out[-1].abs().cpu().numpy().astype(np.float64)
    for i in range(2):
        assert sst.kstest(a[:, i], sst.norm(1.42518, np.sqrt(0.049881)).cdf).statistic < 0.03
    _, st2 = g.GLMALA(g.Mixture_set(0.05), T, torch.zeros(2), y0, 0.3, 100, None, 0.8, ip, 5, num_chains=Cn, seed=4, trace="none",
                      return_stats=True)
    assert abs(float(st.move_rate.mean()) / float(st2.move_rate.mean()) - 1) < 0.08
    assert abs(float(st.global_steps.mean()) / (T - 1) - 0.8) < 0.005
    assert abs(float(st.accepted_local.mean()) / float(st2.accepted_local.mean()) - 1) < 0.15      # the MALA move itself
    assert abs(float(st.esjd().mean()) / float(st2.esjd().mean()) - 1) < 0.12
    runner = g.MCMCRunner(um)
    chain = runner.run_glmala(200, torch.zeros(2), y0[:1], 0.8, ip, 5, 0.3, 100, output_file=None, verbose=False)
    assert chain.shape == (200, 2)
    eng = get_engine()
    eng.bind_proposal(abi.SLOT_IMPORTANCE, ip)

    def fresh(lo, hi):
        aux = torch.zeros(hi - lo, abi.AUX_SLOTS, device="cuda")
        aux[:, abi.AUX_LOCAL] = 1.0
        return torch.zeros(hi - lo, 2, device="cuda"), y0[lo:hi].cuda(), aux
    kw = dict(gf=0.7, seed=9, K=4, sampler="mala", num_grad=33, tau=0.3, trace_layout=abi.TRACE_TIME_MAJOR)
    th, yy, ax = fresh(0, 70)
    full = eng.run_user(um, theta=th, y=yy, aux=ax, n_steps=120, **kw)
    th2, yy2, ax2 = fresh(0, 70)
    buf = torch.zeros(121, 70, 2, device="cuda")
    eng.run_user(um, theta=th2, y=yy2, aux=ax2, n_steps=50, trace=buf, trace_rows=121, **kw)
    eng.run_user(um, theta=th2, y=yy2, aux=ax2, n_steps=70, step_base=50, trace=buf, trace_rows=121, write_row0=False, **kw)
    assert torch.equal(buf, full) and torch.equal(th2, th) and torch.equal(ax2, ax)
    th3, yy3, ax3 = fresh(37, 70)          # a shard: other lanes serve the gradients, the chains are the same
    part = eng.run_user(um, theta=th3, y=yy3, aux=ax3, n_steps=120, chain_id_base=37, **kw)
    assert torch.equal(part, full[:, 37:])


BOX_SRC = """
// theta in (0, 2)^2 or "outside the support": the reference's convention for a prior that wants the local proposal redrawn
__device__ __forceinline__ void glabc_user_simulate(const float* theta, const float* z, const float* p, float* y)
{
    y[0] = theta[0] + p[2] * z[0];
    y[1] = theta[1] + p[2] * z[1];
}
__device__ __forceinline__ float glabc_user_prior_log_prob(const float* theta, const float* p)
{
    const bool in = theta[0] > 0.f && theta[0] < 2.f && theta[1] > 0.f && theta[1] < 2.f;
    return in ? -1.3862944f : OUTSIDE;
}
__device__ __forceinline__ float glabc_user_discrepancy(const float* y, const float* p)
{
    const float a = y[0] - p[0], b = y[1] - p[1];
    return sqrtf(a * a + b * b);
}
"""


@pytest.mark.gpu
def test_prior_sentinel_redraws_the_local_proposal():
    """GLMCMC.py:92-93: `while prior_log_prob(theta') == 7 * log(1e-10): redraw`.  A box prior that answers with the sentinel
    never sees a local proposal outside the box simulated or rejected: with observations at the box's corner and a wide
    random walk the chains move more often than with the same prior answering -inf (which just rejects), and both stay in
    the box."""
    import glabc_b200 as g
    lp = g.DiagGaussian(2, torch.zeros(1, 2), torch.log(torch.tensor([0.5, 0.5])))
    ip = g.DiagGaussian(2, torch.tensor([1.0, 1.0]), torch.log(torch.tensor([1.0, 1.0])))
    Cn, T = 8192, 1501
    th0, y0 = torch.tensor([0.2, 0.2]), torch.tensor([[0.1, 0.1]])
    res = {}
    for name, outside in (("sentinel", "GLABC_PRIOR_SENTINEL"), ("minus_inf", "(-INFINITY)")):
        um = g.UserModel(BOX_SRC.replace("OUTSIDE", outside), theta_dim=2, y_dim=2, n_noise=2, epsilon=0.3, params=[0.1, 0.1, 0.1])
        out, st = g.GLMCMC(um, T, th0, y0, lp, None, 0.0, ip, 5, num_chains=Cn, seed=3, trace="time", return_stats=True)
        assert bool(((out > 0) & (out < 2)).all())
        res[name] = float(st.accepted_local.mean()) / (T - 1)
    assert res["sentinel"] > 1.25 * res["minus_inf"], res
