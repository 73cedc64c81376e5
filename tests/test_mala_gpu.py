"""GPU parity tests of K3 (GLMALA step kernel, one warp per chain) through the C-ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import (abi, check_mala_debug, check_mala_free_running, fresh_mala_state, gauss_pod, load_cases,
                     mala_teacher_forced, model_pod)
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from glabc_b200.engine import Engine
    return Engine()


def bind(eng, model, ip):
    eng.ctx.check(eng.lib.glabc_model_set(eng.ctx.handle, C.byref(model), C.sizeof(model)))
    eng.ctx.check(eng.lib.glabc_dist_set(eng.ctx.handle, abi.SLOT_IMPORTANCE, C.byref(ip), C.sizeof(ip)))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def fresh_dev(c):
    aux, s64 = fresh_mala_state(c)
    return dev(aux), dev(s64)


@pytest.mark.parametrize("arith", [abi.ARITH_STRICT, abi.ARITH_FAST])
@pytest.mark.parametrize("ci", range(12))
def test_replay_golden(eng, ci, arith):
    """the reference's own draws, three independent recordings of each of the four cases (seeded torch / numpy /
    secrets: tests/golden/make_golden.py regenerates them bit-identically).  Free-running: decisions bit-exact (until
    finite-difference noise flips one, helpers.check_mala_free_running).  Restarted from the reference's recorded state
    before every step (one launch, one pseudo-chain per (step, chain)): decisions and resample indices bit-exact; theta',
    y', log-densities, log_acc and the gradient's float64 statistics mu+-, Sigma+- within 1e-5 RELATIVE with no absolute
    slack; grad' through helpers.check_mala_debug's conditioning bound (STRICT) / 5e-3 of the terms (FAST: float32 lane
    partials, MUFU sqrt / exp)."""
    case = load_cases("glmala.npz")[ci]
    T, Cn, K, num = int(case["T"]), case["theta0"].shape[0], int(case["K"]), int(case["num_grad"])
    bind(eng, model_pod(case), gauss_pod(case, "ip"))
    kw = dict(gf=float(case["gf"]), rng_mode=abi.RNG_REPLAY, arith=arith, K=K, num_grad=num, tau=float(case["tau"]))
    if arith == abi.ARITH_STRICT:
        theta, y = dev(case["theta0"]), dev(case["y0"])
        aux, s64 = fresh_dev(Cn)
        dbg = torch.zeros(T - 1, abi.DEBUG64_SLOTS, Cn, device="cuda", dtype=torch.float64)
        tr = eng.run("mala", theta=theta, y=y, aux=aux, state64=s64, n_steps=T - 1, trace_layout=abi.TRACE_TIME_MAJOR,
                     tape32=dev(case["tape32"]), tape64=dev(case["tape64"]), tape_grad0=dev(case["tape_grad0"]), debug64=dbg, **kw)
        torch.cuda.synchronize()
        check_mala_free_running(dbg[:, 0].cpu().numpy(), tr.cpu().numpy(), case)

    tf = mala_teacher_forced(case)
    dbg1 = torch.zeros(1, abi.DEBUG64_SLOTS, tf["n"], device="cuda", dtype=torch.float64)
    eng.run("mala", theta=dev(tf["theta"]), y=dev(tf["y"]), aux=dev(tf["aux"]), state64=dev(tf["state64"]), n_steps=1,
            trace_layout=abi.TRACE_NONE, tape32=dev(tf["tape32"]), tape64=dev(tf["tape64"]), tape_grad0=dev(tf["tape_grad0"]),
            debug64=dbg1, **kw)
    torch.cuda.synchronize()
    d1 = dbg1[0].cpu().numpy()
    if arith == abi.ARITH_STRICT:
        check_mala_debug(d1, tf["rec"], K, tol=1e-5, grec=tf["grec"], eps2=tf["eps2"], num_grad=num, kern_c=tf["kern_c"])
    else:
        same = d1[0].astype(np.int64) == tf["rec"][0].astype(np.int64)
        assert same.mean() > 0.999
        loc = ((tf["rec"][0].astype(np.int64) & 1) == 0) & same
        for a, b in ((1, 1), (2, 2), (3, 3), (6, 4), (7, 5), (14, 8), (15, 9), (16, 10)):
            scale = np.maximum(np.abs(tf["rec"][b][loc]), 1e-3)
            if a == 1:   # log_acc is a difference of O(100) float32 log-densities in FAST mode: relative to the terms
                scale = scale + np.abs(tf["rec"][8][loc]) + np.abs(tf["rec"][9][loc]) + np.abs(tf["rec"][10][loc])
            err = np.abs(d1[a][loc] - tf["rec"][b][loc]) / scale
            assert err.max() < 5e-3, (a, float(err.max()))


def synthetic_case(d, Cn, T, K, num, seed):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)  # noqa: E731
    case = dict(y_obs=(1.0 + 0.5 * rng.random(d)).astype(np.float32), noise_loc=0.05 * f(d),
                noise_scale=(0.2 + 0.2 * rng.random(d)).astype(np.float32), prior_loc=0.1 * f(d),
                prior_log_scale=0.2 * f(d), eps_log_scale=np.float32(np.log(0.4)), epsilon=0.4, ip_loc=0.2 * f(d),
                ip_log_scale=0.3 * f(d))
    case["prior_scale"] = np.exp(case["prior_log_scale"])
    case["eps_scale"] = np.exp(case["eps_log_scale"])
    case["ip_scale"] = np.exp(case["ip_log_scale"])
    slots = abi.tape_mala_slots(d, d, K, num)
    tape = f(T, slots, Cn)
    tape[:, 0] = rng.random((T, Cn), dtype=np.float32)
    tape[:, 1 + 2 * K * d] = rng.random((T, Cn), dtype=np.float32)
    return case, tape, rng.random((T, Cn)), f(d * num * d, Cn), f(Cn, d), 1.0 + 0.3 * f(Cn, d)


@pytest.mark.parametrize("d,K,num", [(1, 1, 5), (2, 5, 100), (3, 7, 33), (4, 16, 64), (2, 15, 2)])
@pytest.mark.parametrize("layout", [abi.TRACE_TIME_MAJOR, abi.TRACE_CHAIN_MAJOR])
def test_replay_matches_oracle(eng, d, K, num, layout):
    """seeded synthetic tapes, both model families, every theta_dim, ragged chain counts: the STRICT kernel and
    the oracle take the same decisions and agree on the float64 state (their float64 sums differ only in order)"""
    Cn, T = 70 + d, 90
    family = abi.MODEL_ABS_NORMAL if d % 2 == 0 else abi.MODEL_ID_NORMAL
    case, tape, tape64, tg0, theta0, y0 = synthetic_case(d, Cn, T, K, num, seed=100 * d + K)
    m, ip = model_pod(case, family), gauss_pod(case, "ip")
    kw = dict(n_steps=T, gf=0.5, rng_mode=abi.RNG_REPLAY, K=K, num_grad=num, tau=0.25, trace_layout=layout)
    th_o, y_o = theta0.copy(), y0.copy()
    aux_o, s64_o = fresh_mala_state(Cn)
    st_o = np.zeros((Cn, abi.nstats(d)), np.float32)
    dbg_o = np.zeros((T, abi.DEBUG64_SLOTS, Cn))
    want = oracle.run("mala", m, None, ip, theta=th_o, y=y_o, aux=aux_o, state64=s64_o, tape32=tape, tape64=tape64,
                      tape_grad0=tg0, stats=st_o, debug64=dbg_o, **kw)
    bind(eng, m, ip)
    theta, y = dev(theta0), dev(y0)
    aux, s64 = fresh_dev(Cn)
    stats = torch.zeros(Cn, abi.nstats(d), device="cuda")
    dbg = torch.zeros(T, abi.DEBUG64_SLOTS, Cn, device="cuda", dtype=torch.float64)
    got = eng.run("mala", theta=theta, y=y, aux=aux, state64=s64, arith=abi.ARITH_STRICT, tape32=dev(tape), tape64=dev(tape64),
                  tape_grad0=dev(tg0), stats=stats, debug64=dbg, block_threads=96, **kw)
    torch.cuda.synchronize()
    fl, fl_o = dbg[:, 0].cpu().numpy().astype(np.int64), dbg_o[:, 0].astype(np.int64)
    clean = (fl == fl_o).all(0)
    assert clean.mean() > 0.97
    g = got.cpu().numpy()
    g, w = (g, want) if layout == abi.TRACE_TIME_MAJOR else (g.transpose(1, 0, 2), want.transpose(1, 0, 2))
    assert np.allclose(g[:, clean], w[:, clean], rtol=1e-5, atol=1e-5) and (g[:, clean] == w[:, clean]).mean() > 0.98
    assert np.array_equal(aux.cpu().numpy()[clean, :5], aux_o[clean, :5])
    assert np.allclose(s64.cpu().numpy()[clean], s64_o[clean], rtol=1e-6, atol=1e-6)
    st = stats.cpu().numpy()
    assert np.array_equal(st[clean, :4], st_o[clean, :4])


def readme_pods():
    case = load_cases("glmala.npz")[0]
    return case, model_pod(case), gauss_pod(case, "ip")


def test_native_draws_replayed_by_oracle(eng):
    """native Philox mode: the kernel dumps every draw it used (incl. the gradient normals) and the oracle replays
    them to the same chain"""
    case, m, ip = readme_pods()
    Cn, T, d, K, num = 101, 120, 2, 5, 100
    bind(eng, m, ip)
    theta0 = np.zeros((Cn, d), np.float32)
    y0 = (np.random.default_rng(1).standard_normal((Cn, d)) * 0.2236).astype(np.float32)
    theta, y = dev(theta0), dev(y0)
    aux, s64 = fresh_dev(Cn)
    slots = abi.tape_mala_slots(d, d, K, num)
    dump = torch.zeros(T, slots, Cn, device="cuda")
    dump64 = torch.zeros(T, Cn, device="cuda", dtype=torch.float64)
    dump_g0 = torch.zeros(d * num * d, Cn, device="cuda")
    kw = dict(n_steps=T, gf=0.8, K=K, num_grad=num, tau=0.3)
    got = eng.run("mala", theta=theta, y=y, aux=aux, state64=s64, seed=21, chain_id_base=3, arith=abi.ARITH_STRICT,
                  trace_layout=abi.TRACE_TIME_MAJOR, tape_dump=dump, tape64_dump=dump64, tape_grad0_dump=dump_g0, **kw)
    torch.cuda.synchronize()
    th_o, y_o = theta0.copy(), y0.copy()
    aux_o, s64_o = fresh_mala_state(Cn)
    want = oracle.run("mala", m, None, ip, theta=th_o, y=y_o, aux=aux_o, state64=s64_o, rng_mode=abi.RNG_REPLAY,
                      tape32=dump.cpu().numpy(), tape64=dump64.cpu().numpy(), tape_grad0=dump_g0.cpu().numpy(), **kw)
    g = got.cpu().numpy()
    moved_k, moved_o = (g[1:] != g[:-1]).any(-1), (want[1:] != want[:-1]).any(-1)
    clean = (moved_k == moved_o).all(0)
    assert clean.mean() > 0.97
    assert np.allclose(g[:, clean], want[:, clean], rtol=1e-5, atol=1e-5)
    # and the undumped launch is the same chain
    theta2, y2 = dev(theta0), dev(y0)
    aux2, s642 = fresh_dev(Cn)
    again = eng.run("mala", theta=theta2, y=y2, aux=aux2, state64=s642, seed=21, chain_id_base=3, arith=abi.ARITH_STRICT,
                    trace_layout=abi.TRACE_TIME_MAJOR, **kw)
    assert torch.equal(again, got)


@pytest.mark.parametrize("d,K,num", [(2, 5, 100), (1, 3, 7), (3, 4, 33), (4, 6, 50)])
def test_thread_per_chain_kernel_equals_warp_per_chain_kernel(eng, d, K, num):
    """FAST native mode runs k_mala_fast (a thread per chain, the gradient dealt over the warp); block_threads = 96 keeps
    the warp-per-chain kernel.  Same Philox layout, same formulas, float32 instead of float64 state: the two take the
    same decisions until float32 finite-difference noise flips one (as between kernel and oracle), and agree on the chain
    values up to then."""
    case, tape, tape64, tg0, theta0, y0 = synthetic_case(d, 333, 1, K, num, seed=7 * d + K)
    family = abi.MODEL_ABS_NORMAL if d % 2 == 0 else abi.MODEL_ID_NORMAL
    bind(eng, model_pod(case, family), gauss_pod(case, "ip"))
    Cn, T = 333, 60
    out = {}
    for name, bt in (("thread", 0), ("warp", 96)):
        theta, y = dev(theta0), dev(y0)
        aux, s64 = fresh_dev(Cn)
        st = torch.zeros(Cn, abi.nstats(d), device="cuda")
        tr = eng.run("mala", theta=theta, y=y, aux=aux, state64=s64, n_steps=T, gf=0.6, seed=5, chain_id_base=11, K=K, num_grad=num,
                     tau=0.25, stats=st, trace_layout=abi.TRACE_TIME_MAJOR, block_threads=bt)
        torch.cuda.synchronize()
        out[name] = (tr.cpu().numpy(), st.cpu().numpy(), aux.cpu().numpy(), s64.cpu().numpy())
    a, b = out["thread"][0], out["warp"][0]
    moved_a, moved_b = (a[1:] != a[:-1]).any(-1), (b[1:] != b[:-1]).any(-1)
    clean = (moved_a == moved_b).all(0)
    assert clean.mean() > 0.9, clean.mean()
    assert moved_a.mean() > 0.02                                  # the chains do move
    # values: identical until a chain's first accepted MALA move, then apart by the amplified float32 finite-difference noise
    # of the prior gradient (1e-2 in the drift, helpers.check_mala_free_running uses the same 2e-2 against the reference)
    diff = np.abs(a[:, clean] - b[:, clean])
    assert diff.max() < 2e-2, float(diff.max())
    assert (diff < 1e-5).mean() > 0.5, float((diff < 1e-5).mean())
    assert np.array_equal(out["thread"][1][clean, :4], out["warp"][1][clean, :4])     # step / global / accept counters
    assert np.array_equal(out["thread"][2][clean, :5], out["warp"][2][clean, :5])     # local / wide / have_grad flags


def test_native_invariances(eng):
    """chunked / resumed runs (float64 state carried in state64), sharding by chain_id_base and both trace layouts
    give bit-identical chains"""
    case, m, ip = readme_pods()
    Cn, T, d, K, num = 96, 130, 2, 5, 40
    bind(eng, m, ip)
    theta0 = torch.zeros(Cn, d, device="cuda")
    y0 = (torch.randn(Cn, d, generator=torch.Generator().manual_seed(3)) * 0.2236).cuda()
    run = lambda **kw: eng.run("mala", gf=0.7, seed=7, K=K, num_grad=num, tau=0.3, **kw)  # noqa: E731
    th, yy = theta0.clone(), y0.clone()
    ax, s64 = fresh_dev(Cn)
    full = run(theta=th, y=yy, aux=ax, state64=s64, n_steps=T - 1, trace_layout=abi.TRACE_TIME_MAJOR)
    th2, yy2 = theta0.clone(), y0.clone()
    ax2, s642 = fresh_dev(Cn)
    buf = torch.zeros(Cn, T, d, device="cuda")
    base = 0
    for n in (33, 64, T - 1 - 97):
        run(theta=th2, y=yy2, aux=ax2, state64=s642, n_steps=n, step_base=base, trace=buf, trace_rows=T,
            trace_layout=abi.TRACE_CHAIN_MAJOR, write_row0=(base == 0))
        base += n
    assert torch.equal(buf.permute(1, 0, 2), full) and torch.equal(th2, th) and torch.equal(s642, s64)
    parts = []
    for lo, hi in ((0, 37), (37, Cn)):
        t, yv = theta0[lo:hi].clone(), y0[lo:hi].clone()
        a, s = fresh_dev(hi - lo)
        parts.append(run(theta=t, y=yv, aux=a, state64=s, n_steps=T - 1, chain_id_base=lo, trace_layout=abi.TRACE_TIME_MAJOR))
    assert torch.equal(torch.cat(parts, dim=1), full)
    host = torch.zeros(T, Cn, d).pin_memory()
    hth, hy = theta0.cpu().clone(), y0.cpu().clone()
    hax, hs64 = (torch.from_numpy(a) for a in fresh_mala_state(Cn))
    eng.run_host("mala", theta=hth, y=hy, aux=hax, state64=hs64, n_steps=T - 1, gf=0.7, seed=7, K=K, num_grad=num, tau=0.3,
                 trace=host, trace_layout=abi.TRACE_TIME_MAJOR, chunk_steps=32)
    assert torch.equal(host, full.cpu()) and torch.equal(hth, th.cpu()) and torch.equal(hs64, s64.cpu())


def test_native_posterior_and_reference_bands(eng):
    """README model via run_glmala settings (gf=0.8, K=5, tau=0.3, num_grad=100): closed-form posterior (SURVEY.md
    App. D) and the reference's measured move rate 1.18 % and ESJD 0.0273 +- 0.0037."""
    from scipy import stats as sst
    from glabc_b200.engine import RunStats
    case, m, ip = readme_pods()
    Cn, T, d, K = 8192, 2500, 2, 5
    bind(eng, m, ip)
    theta = torch.zeros(Cn, d, device="cuda")
    y = torch.randn(Cn, d, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.2236
    aux, s64 = fresh_dev(Cn)
    kw = dict(theta=theta, y=y, aux=aux, state64=s64, gf=0.8, seed=11, K=K, num_grad=100, tau=0.3, trace_layout=abi.TRACE_NONE)
    eng.run("mala", n_steps=T, **kw)
    st = torch.zeros(Cn, abi.nstats(d), device="cuda")
    eng.run("mala", n_steps=T, step_base=T, stats=st, **kw)
    torch.cuda.synchronize()
    a = theta.abs().cpu().numpy().astype(np.float64)
    for i in range(d):
        assert sst.kstest(a[:, i], sst.norm(1.42518, np.sqrt(0.049881)).cdf).statistic < 0.03
    quad = ((theta[:, 0] > 0).long() * 2 + (theta[:, 1] > 0).long()).bincount(minlength=4).cpu().numpy() / Cn
    assert np.abs(quad - 0.25).max() < 0.03
    rs = RunStats(st, d)
    move = float(rs.move_rate.mean())
    assert 0.009 < move < 0.0145, move
    e = float(rs.esjd().mean())
    assert 0.020 < e < 0.036, e


def test_public_api_glmala(eng, tmp_path):
    """examples/Mixture.py:77: run_glmala(num_ite, theta0, y0, 0.8, ip, 5, 0.3, 100)"""
    import glabc_b200 as g
    torch.manual_seed(0)
    model = g.Mixture_set(epsilon=0.05)
    theta0 = torch.tensor([0.0, 0.0])
    y0 = model.generate_samples(theta0)
    ip = g.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    runner = g.MCMCRunner(model, output_dir=str(tmp_path))
    chain = runner.run_glmala(500, theta0, y0, 0.8, ip, 5, 0.3, 100, output_file="glmala_results.csv", verbose=False)
    assert chain.shape == (500, 2) and chain.dtype == torch.float32 and torch.equal(chain[0], theta0)
    assert (tmp_path / "glmala_results.csv").exists()
    out, st = runner.run_glmala(200, theta0, None, 0.8, ip, 5, 0.3, 100, output_file=None, num_chains=64, seed=4,
                                return_stats=True)
    assert out.shape == (64, 200, 2)
    assert np.allclose(g.esjd(out), st.esjd().cpu().numpy(), rtol=1e-4, atol=1e-7)
