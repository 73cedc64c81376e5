"""Edge cases through the public entry points and the C-ABI: degenerate sizes (one iteration, no transitions, zero / one /
ragged chain counts), maximum sizes (K = 16 candidates, a 4096-candidate AGLMCMC block), and the error behaviour of the
boundary (status + message, nothing launched)."""
import numpy as np
import pytest
import torch

from helpers import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def objs():
    import glabc_b200 as g
    model = g.Mixture_set(0.05)
    lp = g.DiagGaussian(2, torch.zeros(1, 2), torch.log(torch.tensor([0.35, 0.35])))
    gp = g.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    return g, model, lp, gp


def samplers(g, model, lp, gp):
    z = torch.zeros(2)
    return {
        "global": lambda n, **kw: g.GlobalMCMC(model, n, z, None, gp, None, 0.5, lp, **kw),
        "glmcmc": lambda n, **kw: g.GLMCMC(model, n, z, None, lp, None, 0.9, gp, 5, **kw),
        "glmala": lambda n, **kw: g.GLMALA(model, n, z, None, 0.3, 10, None, 0.8, gp, 5, **kw),
        "aglmcmc": lambda n, **kw: g.AGLMCMC(model, n, z, None, lp, gp, None, 1.0, 10, 5, 0.8, 0.2, **kw),
        "glmcmc_nf": lambda n, **kw: g.GLMCMC_NF(model, n, z, None, lp, None, 0.5, 10, 5, None, 1, **kw),
    }


@pytest.mark.parametrize("name", ["global", "glmcmc", "glmala", "aglmcmc", "glmcmc_nf"])
def test_one_iteration_and_ragged_chain_counts(objs, name):
    run = samplers(*objs)[name]
    one = run(1, num_chains=5, seed=1)                       # num_ite = 1: only row 0, the initial theta
    assert one.shape == (5, 1, 2) and torch.equal(one, torch.zeros(5, 1, 2, device="cuda"))
    for c in (1, 31, 33):                                    # below / across a warp
        for layout, shape in (("chain", (c, 40, 2)), ("time", (40, c, 2))):
            out, st = run(40, num_chains=c, seed=2, trace=layout, return_stats=True)
            assert out.shape == shape and bool(torch.isfinite(out).all())
            first = out[:, 0] if layout == "chain" else out[0]
            assert torch.equal(first, torch.zeros(c, 2, device="cuda"))
            assert torch.equal(st.steps, torch.full((c,), 39.0, dtype=torch.float64, device="cuda"))
    # the chains of a ragged launch are the chains of a bigger one (Philox keyed by global chain id)
    if name in ("global", "glmcmc", "glmala"):
        big = run(40, num_chains=33, seed=2, trace="time")
        small = run(40, num_chains=31, seed=2, trace="time")
        assert torch.equal(big[:, :31], small)


def test_zero_chains_and_zero_steps(objs):
    g, model, lp, gp = objs
    from glabc_b200.engine import get_engine
    eng = get_engine()
    eng.bind_model(model)
    eng.bind_proposal(abi.SLOT_LOCAL, lp)
    eng.bind_proposal(abi.SLOT_GLOBAL, gp)
    eng.bind_proposal(abi.SLOT_IMPORTANCE, gp)
    empty = torch.zeros(0, 2, device="cuda")
    out = eng.run("global", theta=empty, y=empty.clone(), n_steps=10, gf=0.5, trace_layout=abi.TRACE_TIME_MAJOR)
    assert out.shape == (11, 0, 2)
    th, yy = torch.ones(7, 2, device="cuda"), torch.ones(7, 2, device="cuda")
    st = torch.zeros(7, abi.nstats(2), device="cuda")
    out = eng.run("global", theta=th, y=yy, n_steps=0, gf=0.5, trace_layout=abi.TRACE_CHAIN_MAJOR, stats=st)   # row 0 only
    assert out.shape == (7, 1, 2) and torch.equal(out[:, 0], th) and float(st.abs().sum()) == 0.0
    aux = torch.zeros(7, abi.AUX_SLOTS, device="cuda")
    out = eng.run("isir", theta=th, y=yy, aux=aux, n_steps=0, gf=0.9, K=5, trace_layout=abi.TRACE_TIME_MAJOR)
    assert out.shape == (1, 7, 2) and torch.equal(out[0], th)
    host = torch.zeros(0, 5, 2)
    eng.run_host("global", theta=torch.zeros(0, 2), y=torch.zeros(0, 2), n_steps=4, gf=0.5, trace=host,
                 trace_layout=abi.TRACE_CHAIN_MAJOR)        # nothing to do, no error


def test_maximum_sizes(objs):
    g, model, lp, gp = objs
    out, st = g.GLMCMC(model, 300, torch.zeros(2), None, lp, None, 0.9, gp, abi.MAX_K, num_chains=257, seed=3, return_stats=True)
    assert out.shape == (257, 300, 2) and 0 < float(st.move_rate.mean()) < 0.5
    out, st = g.AGLMCMC(model, 600, torch.zeros(2), None, lp, gp, None, 1.0, abi.AG_MAX_BLOCK // 16, 16, 0.8, 0.2, num_chains=8,
                        seed=3, return_stats=True)          # one 4096-candidate block per chain, one adaptation
    assert out.shape == (8, 600, 2) and bool(torch.isfinite(out).all())
    with pytest.raises(ValueError):
        g.GLMCMC(model, 10, torch.zeros(2), None, lp, None, 0.9, gp, abi.MAX_K + 1, num_chains=4)
    with pytest.raises(ValueError):
        g.AGLMCMC(model, 10, torch.zeros(2), None, lp, gp, None, 1.0, abi.AG_MAX_BLOCK // 16 + 1, 16, 0.8, 0.2, num_chains=4)


def test_boundary_errors_carry_a_message(objs):
    g, model, lp, gp = objs
    from glabc_b200.engine import Engine
    eng = Engine()                                            # a fresh context: nothing bound
    th, yy = torch.zeros(4, 2, device="cuda"), torch.zeros(4, 2, device="cuda")
    with pytest.raises(abi.GlabcError, match="no model bound") as ei:
        eng.run("global", theta=th, y=yy, n_steps=3, gf=0.5)
    assert ei.value.status == abi.ERR_INVALID
    eng.bind_model(model)
    with pytest.raises(abi.GlabcError, match="LOCAL and GLOBAL"):
        eng.run("global", theta=th, y=yy, n_steps=3, gf=0.5)
    eng.bind_proposal(abi.SLOT_LOCAL, lp)
    eng.bind_proposal(abi.SLOT_GLOBAL, gp)
    eng.bind_proposal(abi.SLOT_IMPORTANCE, gp)
    with pytest.raises(abi.GlabcError, match="block_threads"):
        eng.run("global", theta=th, y=yy, n_steps=3, gf=0.5, block_threads=48)
    with pytest.raises(abi.GlabcError, match="aux"):
        eng.run("isir", theta=th, y=yy, n_steps=3, gf=0.5, K=5)
    with pytest.raises(abi.GlabcError, match="fall outside"):
        eng.run("global", theta=th, y=yy, n_steps=30, gf=0.5, trace=torch.zeros(4, 10, 2, device="cuda"), trace_rows=10)
    with pytest.raises(abi.GlabcError, match="replay mode needs tape32"):
        eng.run("global", theta=th, y=yy, n_steps=3, gf=0.5, rng_mode=abi.RNG_REPLAY)
    with pytest.raises(abi.GlabcError, match="GLABC_TRACE_EVENTS") as ei:
        eng.run("mala", theta=th, y=yy, n_steps=3, gf=0.5, K=5, num_grad=4, tau=0.3, aux=torch.zeros(4, abi.AUX_SLOTS, device="cuda"),
                state64=torch.zeros(4, abi.STATE64_SLOTS, device="cuda", dtype=torch.float64),
                trace_layout=abi.TRACE_EVENTS, trace=torch.zeros(4, 8, 3, device="cuda"), trace_rows=8)
    assert ei.value.status == abi.ERR_UNSUPPORTED
    # the event layout is a device-buffer layout: the host entry points refuse it before anything is allocated or launched
    hth, hy = torch.zeros(4, 2), torch.zeros(4, 2)
    for sampler, kw in (("global", {}), ("isir", dict(K=5, aux=torch.zeros(4, abi.AUX_SLOTS)))):
        ev = torch.full((4, 8, 3), 7.0)
        with pytest.raises(abi.GlabcError, match="device-buffer layout") as ei:
            eng.run_host(sampler, theta=hth, y=hy, n_steps=30, gf=0.5, trace=ev, trace_layout=abi.TRACE_EVENTS, **kw)
        assert ei.value.status == abi.ERR_UNSUPPORTED and bool((ev == 7.0).all())
    # a 5-dimensional model has no fused family
    with pytest.raises(NotImplementedError):
        g.AbsNormalModel(0.05, y_obs=(1.0,) * 5).lower()
    # the failed calls left the context usable
    out = eng.run("global", theta=th, y=yy, n_steps=3, gf=0.5, trace_layout=abi.TRACE_TIME_MAJOR)
    assert out.shape == (4, 4, 2)
