"""GPU parity tests of K1 (GlobalMCMC step kernel) through the C-ABI, against the golden vectors
recorded from the reference and against the CPU oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from helpers import abi, gauss_pod, load_cases, model_pod, rel_err
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from glabc_b200.engine import Engine
    return Engine()


def bind(eng, model, lp, gp):
    import ctypes as C
    eng.ctx.check(eng.lib.glabc_model_set(eng.ctx.handle, C.byref(model), C.sizeof(model)))
    eng.ctx.check(eng.lib.glabc_dist_set(eng.ctx.handle, abi.SLOT_LOCAL, C.byref(lp), C.sizeof(lp)))
    eng.ctx.check(eng.lib.glabc_dist_set(eng.ctx.handle, abi.SLOT_GLOBAL, C.byref(gp), C.sizeof(gp)))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- Philox known answers ------------------------------------------------------------------------
KAT = [  # Random123 kat_vectors (philox4x32-10) as recalled; independently pinned by test_philox_vs_curand
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_philox_known_answers(eng):
    ctr = [k[0] for k in KAT]
    key = [k[1] for k in KAT]
    out = eng.philox(ctr, key).cpu().tolist()
    for (c, k, want), got in zip(KAT, out):
        assert got == want
        assert oracle.philox(c, k) == want
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, size=(4096, 4), dtype=np.int64)
    key = rng.integers(0, 2**32, size=(4096, 2), dtype=np.int64)
    out = eng.philox(ctr, key).cpu().numpy()
    for i in range(0, 4096, 97):
        assert out[i].tolist() == oracle.philox(ctr[i], key[i])


def test_philox_vs_curand(eng):
    """the product's Philox4x32-10 against NVIDIA cuRAND's device implementation (tests/cuda/curand_kat.cu):
    counter = (chain_lo, chain_hi, block, slot) <-> curand_init(seed, subsequence = block | slot << 32,
    offset = 4 * chain)."""
    import ctypes as C
    import os
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda", "libcurand_kat.so")
    if not os.path.exists(so):
        pytest.skip("tests/cuda/libcurand_kat.so not built (run __graft_entry__.build())")
    lib = C.CDLL(so)
    rng = np.random.default_rng(5)
    n = 2048
    chain = rng.integers(0, 2**40, size=n, dtype=np.int64)
    block = rng.integers(0, 2**32, size=n, dtype=np.int64)
    slot = rng.integers(0, 2**32, size=n, dtype=np.int64)
    seed = 0x9E3779B97F4A7C15
    sub = torch.from_numpy(block | (slot << 32)).cuda()
    off = torch.from_numpy(chain).cuda()
    out = torch.zeros(n, 4, dtype=torch.int32, device="cuda")
    lib.curand_philox_blocks.argtypes = [C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    assert lib.curand_philox_blocks(seed, sub.data_ptr(), off.data_ptr(), n, out.data_ptr()) == 0
    ctr = np.stack([chain & 0xFFFFFFFF, chain >> 32, block, slot], axis=1)
    key = np.tile(np.array([[seed & 0xFFFFFFFF, seed >> 32]], dtype=np.int64), (n, 1))
    mine = eng.philox(ctr, key).cpu().numpy()
    assert np.array_equal(mine, out.cpu().numpy().astype(np.int64) & 0xFFFFFFFF)


# ---- replay of the reference's own draws ----------------------------------------------------------
@pytest.mark.parametrize("arith", [abi.ARITH_STRICT, abi.ARITH_FAST])
@pytest.mark.parametrize("ci", range(4))
def test_replay_golden(eng, ci, arith):
    case = load_cases("global_mcmc.npz")[ci]
    T, C = int(case["T"]), case["theta0"].shape[0]
    bind(eng, model_pod(case), gauss_pod(case, "lp"), gauss_pod(case, "gp"))
    theta, y = dev(case["theta0"]), dev(case["y0"])
    debug = torch.zeros(T - 1, abi.DEBUG_SLOTS, C, device="cuda")
    stats = torch.zeros(C, abi.nstats(2), device="cuda")
    trace = eng.run("global", theta=theta, y=y, n_steps=T - 1, gf=float(case["gf"]), rng_mode=abi.RNG_REPLAY,
                    arith=arith, trace_layout=abi.TRACE_TIME_MAJOR, tape32=dev(case["tape32"]), debug=debug, stats=stats)
    torch.cuda.synchronize()
    dbg, rec = debug.cpu().numpy(), case["rec"]
    # accept/branch decisions bit-exact (north_star), hence the whole trace
    assert np.array_equal(dbg[:, 0].astype(np.int32), rec[:, 0].astype(np.int32))
    if arith == abi.ARITH_STRICT:
        assert np.array_equal(trace.cpu().numpy(), case["trace"])
    else:  # FMA contraction: same decisions, values to float32 rounding
        assert np.allclose(trace.cpu().numpy(), case["trace"], rtol=2e-6, atol=1e-6)
    # log-densities and the kernel within 1e-5 relative (north_star); log_acc is a difference of
    # those (|kernel| runs to ~800 while log_acc may be O(1)), so its error is measured against the
    # magnitude of the terms it is made of
    # (the kernel log-density crosses zero — 2.0768 - 0.5 (dis/eps)^2 — so values within a few ulp of
    # zero get an absolute floor of 3e-6 = 1.5 ulp of the O(2) constants they are made of)
    for k in (1, 2):
        assert np.allclose(dbg[:, k], rec[:, k], rtol=1e-5, atol=3e-6)
        if arith == abi.ARITH_STRICT:
            assert rel_err(dbg[:, k], rec[:, k]).max() <= 1e-5
    scale = np.abs(rec[:, 1]) + np.abs(rec[:, 2]) + np.abs(rec[:, 3])
    assert (np.abs(dbg[:, 3].astype(np.float64) - rec[:, 3]) / scale).max() <= 1e-5
    if arith == abi.ARITH_STRICT:
        assert np.array_equal(dbg[:, 1], rec[:, 1])  # prior: +,-,*,/ only -> identical bits
    flags = rec[:, 0].astype(np.int32)
    st = stats.cpu().numpy()
    assert np.array_equal(st[:, abi.STAT_STEPS], np.full(C, T - 1, np.float32))
    assert np.array_equal(st[:, abi.STAT_GLOBAL_STEPS], (flags & 1).sum(0))
    assert np.array_equal(st[:, abi.STAT_ACC_GLOBAL], ((flags & 3) == 3).sum(0))
    assert np.array_equal(st[:, abi.STAT_ACC_LOCAL], ((flags & 3) == 2).sum(0))


def synthetic_case(d, C, T, seed, family=abi.MODEL_ABS_NORMAL):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)  # noqa: E731
    case = dict(y_obs=(1.0 + 0.5 * rng.random(d)).astype(np.float32), noise_loc=0.05 * f(d),
                noise_scale=(0.2 + 0.2 * rng.random(d)).astype(np.float32), prior_loc=0.1 * f(d),
                prior_log_scale=0.2 * f(d), eps_log_scale=np.float32(np.log(0.3)), lp_loc=0.01 * f(d),
                lp_log_scale=np.log(0.2 + 0.3 * rng.random(d)).astype(np.float32), gp_loc=0.2 * f(d),
                gp_log_scale=0.3 * f(d))
    case["prior_scale"] = np.exp(case["prior_log_scale"])
    case["eps_scale"] = np.exp(case["eps_log_scale"])
    case["lp_scale"], case["gp_scale"] = np.exp(case["lp_log_scale"]), np.exp(case["gp_log_scale"])
    tape = f(T, 2 + 2 * d, C)
    tape[:, 0] = rng.random((T, C), dtype=np.float32)
    tape[:, 1 + 2 * d] = rng.random((T, C), dtype=np.float32)
    tape[0, 1 + 2 * d, 0] = 0.0  # log(0) = -inf must accept (B-16)
    return case, tape, f(C, d), 1.0 + 0.3 * f(C, d)


@pytest.mark.parametrize("d", [1, 2, 3, 4])
@pytest.mark.parametrize("layout", [abi.TRACE_TIME_MAJOR, abi.TRACE_CHAIN_MAJOR])
def test_replay_matches_oracle_all_dims(eng, d, layout):
    """ragged chain count (tail warp), every fused dimension, both trace layouts, both families"""
    C, T = 1000 + 37, 150
    family = abi.MODEL_ABS_NORMAL if d % 2 == 0 else abi.MODEL_ID_NORMAL
    case, tape, theta0, y0 = synthetic_case(d, C, T, seed=d, family=family)
    m, lp, gp = model_pod(case, family), gauss_pod(case, "lp"), gauss_pod(case, "gp")
    th_o, y_o = theta0.copy(), y0.copy()
    st_o = np.zeros((C, abi.nstats(d)), np.float32)
    want = oracle.run("global", m, lp, gp, theta=th_o, y=y_o, n_steps=T, gf=0.4, rng_mode=abi.RNG_REPLAY,
                      tape32=tape, trace_layout=layout, stats=st_o)
    bind(eng, m, lp, gp)
    theta, y = dev(theta0), dev(y0)
    stats = torch.zeros(C, abi.nstats(d), device="cuda")
    got = eng.run("global", theta=theta, y=y, n_steps=T, gf=0.4, rng_mode=abi.RNG_REPLAY, arith=abi.ARITH_STRICT,
                  trace_layout=layout, tape32=dev(tape), stats=stats, block_threads=96)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(theta.cpu().numpy(), th_o) and np.array_equal(y.cpu().numpy(), y_o)
    st = stats.cpu().numpy()
    assert np.array_equal(st[:, :4], st_o[:, :4])
    assert np.allclose(st[:, 4:], st_o[:, 4:], rtol=1e-5, atol=1e-5)


# ---- native RNG -----------------------------------------------------------------------------------
def readme_pods():
    case = load_cases("global_mcmc.npz")[0]
    return case, model_pod(case), gauss_pod(case, "lp"), gauss_pod(case, "gp")


def test_native_draws_replayed_by_oracle(eng):
    """the kernel dumps the Philox/Box-Muller draws it used; the CPU oracle replays them and must land
    on the same chain (strict arithmetic) — ties the native path to the same restatement."""
    case, m, lp, gp = readme_pods()
    C, T, d = 777, 400, 2
    bind(eng, m, lp, gp)
    theta0 = np.zeros((C, d), np.float32)
    y0 = (np.random.default_rng(1).standard_normal((C, d)) * 0.2236).astype(np.float32)
    theta, y = dev(theta0), dev(y0)
    dump = torch.zeros(T, 6, C, device="cuda")
    got = eng.run("global", theta=theta, y=y, n_steps=T, gf=0.5, seed=11, chain_id_base=5, arith=abi.ARITH_STRICT,
                  trace_layout=abi.TRACE_TIME_MAJOR, tape_dump=dump)
    torch.cuda.synchronize()
    th_o, y_o = theta0.copy(), y0.copy()
    want = oracle.run("global", m, lp, gp, theta=th_o, y=y_o, n_steps=T, gf=0.5, rng_mode=abi.RNG_REPLAY,
                      tape32=dump.cpu().numpy())
    assert np.array_equal(got.cpu().numpy(), want)
    # and the fast-arithmetic kernel takes the same decisions on the same draws
    theta, y = dev(theta0), dev(y0)
    fast = eng.run("global", theta=theta, y=y, n_steps=T, gf=0.5, seed=11, chain_id_base=5, arith=abi.ARITH_FAST,
                   trace_layout=abi.TRACE_TIME_MAJOR)
    moved_f = (fast[1:] != fast[:-1]).any(-1).cpu().numpy()
    moved_s = (want[1:] != want[:-1]).any(-1)
    assert (moved_f != moved_s).mean() < 1e-4
    # the oracle's own native mode draws the same Philox streams (normals differ in the last bits only:
    # libm vs MUFU), so its uniforms agree exactly with the dump
    th_n, y_n = theta0.copy(), y0.copy()
    nat = oracle.run("global", m, lp, gp, theta=th_n, y=y_n, n_steps=T, gf=0.5, seed=11, chain_id_base=5)
    dn = dump.cpu().numpy()
    agree = ((nat[1:] != nat[:-1]).any(-1) == moved_s).mean()
    assert agree > 0.999
    assert np.abs(dn[:, 1:5]).max() < 6.5 and abs(dn[:, 1:5].mean()) < 0.01 and abs(dn[:, 1:5].std() - 1) < 0.01


def test_native_invariances(eng):
    """time chunking / resume, sharding by chain_id_base, trace layouts and the host-buffer entry
    point all give the same chains."""
    case, m, lp, gp = readme_pods()
    C, T, d = 300, 257, 2
    bind(eng, m, lp, gp)
    theta0 = torch.zeros(C, d, device="cuda")
    y0 = (torch.randn(C, d, generator=torch.Generator().manual_seed(3)) * 0.2236).cuda()
    run = lambda **kw: eng.run("global", gf=0.5, seed=99, **kw)  # noqa: E731
    th, yy = theta0.clone(), y0.clone()
    st_full = torch.zeros(C, abi.nstats(d), device="cuda")
    full = run(theta=th, y=yy, n_steps=T - 1, trace_layout=abi.TRACE_TIME_MAJOR, stats=st_full)
    # chain-major layout
    th2, yy2 = theta0.clone(), y0.clone()
    cm = run(theta=th2, y=yy2, n_steps=T - 1, trace_layout=abi.TRACE_CHAIN_MAJOR)
    assert torch.equal(cm.permute(1, 0, 2), full)
    # three time chunks (odd boundaries) into one buffer, chain-major, plus accumulated stats
    th3, yy3 = theta0.clone(), y0.clone()
    buf = torch.zeros(C, T, d, device="cuda")
    st = torch.zeros(C, abi.nstats(d), device="cuda")
    base = 0
    for n in (37, 100, T - 1 - 137):
        run(theta=th3, y=yy3, n_steps=n, step_base=base, trace=buf, trace_rows=T, trace_layout=abi.TRACE_CHAIN_MAJOR,
            write_row0=(base == 0), stats=st)
        base += n
    assert torch.equal(buf, cm) and torch.equal(th3, th) and torch.equal(yy3, yy)
    assert torch.equal(st[:, :4], st_full[:, :4]) and torch.allclose(st, st_full, rtol=1e-4, atol=1e-3)  # float32 sums, different chunking
    # two shards keyed by global chain id
    h = 128
    parts = []
    for lo, hi in ((0, h), (h, C)):
        t, yv = theta0[lo:hi].clone(), y0[lo:hi].clone()
        parts.append(run(theta=t, y=yv, n_steps=T - 1, chain_id_base=lo, trace_layout=abi.TRACE_TIME_MAJOR))
    assert torch.equal(torch.cat(parts, dim=1), full)
    # host-buffer entry point (pinned), both layouts, small chunks so the double buffering cycles
    for layout, ref in ((abi.TRACE_TIME_MAJOR, full), (abi.TRACE_CHAIN_MAJOR, cm)):
        shape = (T, C, d) if layout == abi.TRACE_TIME_MAJOR else (C, T, d)
        host = torch.zeros(shape).pin_memory()
        hth, hy = theta0.cpu().clone(), y0.cpu().clone()
        hst = torch.zeros(C, abi.nstats(d))
        eng.run_host("global", theta=hth, y=hy, n_steps=T - 1, gf=0.5, seed=99, trace=host, trace_layout=layout,
                     stats=hst, chunk_steps=64)
        assert torch.equal(host, ref.cpu()) and torch.equal(hth, th.cpu())
        assert torch.equal(hst[:, :4], st_full.cpu()[:, :4]) and torch.allclose(hst, st_full.cpu(), rtol=1e-4, atol=1e-3)


def test_native_posterior_matches_closed_form(eng):
    """README model: |theta_i| ~ N(1.42518, 0.049881), four equal modes."""
    from scipy import stats as sst
    case, m, lp, gp = readme_pods()
    C, T, d = 32768, 6000, 2
    bind(eng, m, lp, gp)
    theta = torch.zeros(C, d, device="cuda")
    y = torch.randn(C, d, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.2236
    eng.run("global", theta=theta, y=y, n_steps=T, gf=0.5, seed=2024, trace_layout=abi.TRACE_NONE)  # burn-in
    st = torch.zeros(C, abi.nstats(d), device="cuda")
    eng.run("global", theta=theta, y=y, n_steps=T, step_base=T, gf=0.5, seed=2024, trace_layout=abi.TRACE_NONE, stats=st)
    torch.cuda.synchronize()
    a = theta.abs().cpu().numpy().astype(np.float64)
    for i in range(d):  # chains are independent: the final states are an i.i.d. sample of the posterior
        ks = sst.kstest(a[:, i], sst.norm(1.42518, np.sqrt(0.049881)).cdf).statistic
        assert ks < 0.015, ks
        assert abs(a[:, i].mean() - 1.42518) < 0.006 and abs(a[:, i].var() - 0.049881) < 0.003
    quad = ((theta[:, 0] > 0).long() * 2 + (theta[:, 1] > 0).long()).bincount(minlength=4).cpu().numpy() / C
    assert np.abs(quad - 0.25).max() < 0.02
    s = st.cpu().numpy().astype(np.float64)
    move = (s[:, abi.STAT_ACC_LOCAL] + s[:, abi.STAT_ACC_GLOBAL]).sum() / s[:, abi.STAT_STEPS].sum()
    assert 0.009 < move < 0.014   # reference: 1.14 % +- 0.04
    assert abs(s[:, abi.STAT_GLOBAL_STEPS].sum() / s[:, abi.STAT_STEPS].sum() - 0.5) < 1e-3


def test_native_matches_oracle_native_distribution(eng):
    """same sampler, same Philox streams, CPU libm vs GPU MUFU normals: two-sample KS on |theta|."""
    from scipy import stats as sst
    case, m, lp, gp = readme_pods()
    C, T, d = 4096, 3000, 2
    bind(eng, m, lp, gp)
    theta0 = np.zeros((C, d), np.float32)
    y0 = (np.random.default_rng(8).standard_normal((C, d)) * 0.2236).astype(np.float32)
    theta, y = dev(theta0), dev(y0)
    st = torch.zeros(C, abi.nstats(d), device="cuda")
    eng.run("global", theta=theta, y=y, n_steps=T, gf=0.5, seed=77, trace_layout=abi.TRACE_NONE, stats=st)
    th_o, y_o = theta0.copy(), y0.copy()
    st_o = np.zeros((C, abi.nstats(d)), np.float32)
    oracle.run("global", m, lp, gp, theta=th_o, y=y_o, n_steps=T, gf=0.5, seed=77, trace_layout=abi.TRACE_NONE, stats=st_o)
    g = theta.abs().cpu().numpy()
    for i in range(d):
        assert sst.ks_2samp(g[:, i], np.abs(th_o[:, i])).statistic < 0.03
    sg = st.cpu().numpy().astype(np.float64)
    acc_g = (sg[:, 2] + sg[:, 3]).sum() / (C * T)
    acc_o = (st_o[:, 2].astype(np.float64) + st_o[:, 3]).sum() / (C * T)
    assert abs(acc_g - acc_o) < 5e-4
    assert np.array_equal(sg[:, abi.STAT_GLOBAL_STEPS], st_o[:, abi.STAT_GLOBAL_STEPS])  # same uniforms exactly


# ---- esjd ------------------------------------------------------------------------------------------
def test_esjd_kernel(eng):
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "misc.npz"))
    chains = torch.from_numpy(z["esjd/chains"]).cuda()          # [6, 400, 2]
    got = eng.esjd(chains.contiguous(), abi.TRACE_CHAIN_MAJOR).cpu().numpy()
    assert np.allclose(got, z["esjd/values"], rtol=2e-5)         # vs the reference's esjd()
    got_t = eng.esjd(chains.permute(1, 0, 2).contiguous(), abi.TRACE_TIME_MAJOR).cpu().numpy()
    assert np.allclose(got_t, z["esjd/values"], rtol=2e-5)
    assert np.allclose(oracle.esjd(z["esjd/chains"], abi.TRACE_CHAIN_MAJOR), z["esjd/values"], rtol=2e-5)


def test_public_api_single_chain(eng, tmp_path, capsys):
    """the reference-shaped call: one chain, CPU tensor [num_ite, d] back, CSV written, summary printed"""
    import glabc_b200 as g
    torch.manual_seed(0)
    model = g.Mixture_set(epsilon=0.05)
    lp = g.DiagGaussian(2, loc=torch.zeros(1, 2), log_scale=torch.log(torch.tensor([0.35, 0.35])))
    gp = g.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    theta0 = torch.tensor([0.0, 0.0])
    y0 = model.generate_samples(theta0)
    runner = g.MCMCRunner(model, output_dir=str(tmp_path))
    chain = runner.run_global_mcmc(2000, theta0, y0, 0.5, lp, gp, output_file="global.csv")
    assert chain.shape == (2000, 2) and chain.dtype == torch.float32 and not chain.is_cuda
    assert torch.equal(chain[0], theta0)
    assert "Theta_Re 1:" in capsys.readouterr().out
    rows = np.loadtxt(tmp_path / "global.csv", delimiter=",", dtype=np.float32)
    assert rows.shape == (2000, 2) and np.array_equal(rows, chain.numpy())
    e = g.esjd(chain)
    assert isinstance(e, np.ndarray) and e.shape == () and e.dtype == np.float32
    # many chains: device tensor [C, num_ite, d] + stats; esjd from stats == esjd of the trace
    out, st = runner.run_global_mcmc(500, theta0, None, 0.5, lp, gp, output_file=None, num_chains=256, seed=1, return_stats=True)
    assert out.shape == (256, 500, 2) and out.is_cuda
    assert np.allclose(g.esjd(out), st.esjd().cpu().numpy(), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("which", ["global", "glmcmc", "glmala"])
def test_checkpoint_resume_continues_bit_identically(tmp_path, which):
    """run 151 iterations, checkpoint, resume to 301 == one 301-iteration run (Philox is keyed by
    (seed, global chain id, step): the checkpoint holds no generator state)"""
    import glabc_b200 as g
    model = g.Mixture_set(0.05)
    lp = g.DiagGaussian(2, torch.zeros(1, 2), torch.log(torch.tensor([0.35, 0.35])))
    gp = g.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    y0 = torch.randn(300, 2, generator=torch.Generator().manual_seed(2)) * 0.2236
    kw = dict(num_chains=300, seed=17, chain_id_base=5, trace="time", return_stats=True)

    def run(n, **extra):
        if which == "global":
            return g.GlobalMCMC(model, n, torch.zeros(2), y0, gp, None, 0.5, lp, **kw, **extra)
        if which == "glmcmc":
            return g.GLMCMC(model, n, torch.zeros(2), y0, lp, None, 0.9, gp, 5, **kw, **extra)
        return g.GLMALA(model, n, torch.zeros(2), y0, 0.3, 20, None, 0.7, gp, 5, **kw, **extra)

    full, st_full = run(301)
    ck = str(tmp_path / "ck.pt")
    head, _ = run(151, checkpoint=ck)
    tail, st = run(301, resume=ck)
    assert head.shape[0] == 151 and tail.shape[0] == 150
    assert torch.equal(torch.cat([head, tail]), full)
    assert torch.equal(st.steps, st_full.steps) and torch.equal(st.global_steps, st_full.global_steps)
    assert torch.allclose(st.raw, st_full.raw, rtol=1e-3, atol=5e-3)   # the sums are re-associated at the cut
    with pytest.raises(ValueError, match="sampler"):
        g.GLMCMC(model, 400, torch.zeros(2), y0, lp, None, 0.9, gp, 5, resume=ck) if which != "glmcmc" else \
            g.GlobalMCMC(model, 400, torch.zeros(2), y0, gp, None, 0.5, lp, resume=ck)


def test_host_entry_event_transport(monkeypatch):
    """glabc_run_global_host with a chain-major host trace: part of the chains travels as move events and is expanded by
    the host cores while the rest comes densely over PCIe — the host buffer must equal the device-resident dense trace
    bit for bit, for every split, and a chain with more moves than the event capacity must fall back to the dense path."""
    import glabc_b200 as g
    from glabc_b200.engine import get_engine
    eng = get_engine()
    lp = g.DiagGaussian(2, torch.zeros(1, 2), torch.log(torch.tensor([0.35, 0.35])))
    gp = g.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    Cn, T, d = 1000, 1500, 2
    y0 = torch.randn(Cn, d, generator=torch.Generator().manual_seed(4)) * 0.2236
    for eps, frac in ((0.05, "0.6"), (0.05, "1.0"), (0.05, "0.13"), (3.0, "0.6")):   # eps = 3: most proposals accepted -> overflow
        monkeypatch.setenv("GLABC_HOST_EVENT_FRACTION", frac)
        eng.bind_model(g.Mixture_set(eps))
        eng.bind_proposal(abi.SLOT_LOCAL, lp)
        eng.bind_proposal(abi.SLOT_GLOBAL, gp)
        th, yy = torch.zeros(Cn, d, device="cuda"), y0.cuda()
        st = torch.zeros(Cn, abi.nstats(d), device="cuda")
        want = eng.run("global", theta=th, y=yy, n_steps=T - 1, gf=0.5, seed=31, chain_id_base=7, stats=st,
                       trace_layout=abi.TRACE_CHAIN_MAJOR)
        torch.cuda.synchronize()
        host = torch.full((Cn, T, d), float("nan")).pin_memory()
        hth, hy, hst = torch.zeros(Cn, d), y0.clone(), torch.zeros(Cn, abi.nstats(d))
        eng.run_host("global", theta=hth, y=hy, n_steps=T - 1, gf=0.5, seed=31, chain_id_base=7, trace=host, stats=hst,
                     trace_layout=abi.TRACE_CHAIN_MAJOR)
        assert torch.equal(host, want.cpu()), (eps, frac)
        assert torch.equal(hth, th.cpu()) and torch.equal(hy, yy.cpu()) and torch.equal(hst, st.cpu())
        if eps > 1:
            assert float(st[:, abi.STAT_ACC_LOCAL].mean() + st[:, abi.STAT_ACC_GLOBAL].mean()) > (T - 1) / 32   # really overflowed
    # the events layout itself, device buffers: entry 0 = count, entries = (row, theta) of every move
    monkeypatch.delenv("GLABC_HOST_EVENT_FRACTION")
    eng.bind_model(g.Mixture_set(0.05))
    th, yy = torch.zeros(64, d, device="cuda"), y0[:64].cuda()
    dense = eng.run("global", theta=th, y=yy, n_steps=600, gf=0.5, seed=5, trace_layout=abi.TRACE_CHAIN_MAJOR).cpu()
    th, yy = torch.zeros(64, d, device="cuda"), y0[:64].cuda()
    ev = torch.zeros(64, 128, 3, device="cuda")
    eng.run("global", theta=th, y=yy, n_steps=600, gf=0.5, seed=5, trace_layout=abi.TRACE_EVENTS, trace=ev, trace_rows=128)
    ev = ev.cpu()
    for c in (0, 17, 63):
        m = int(ev[c, 0, 0].view(torch.int32))
        rows = ev[c, 1:m + 1, 0].contiguous().view(torch.int32).tolist()
        moved = [0] + [i for i in range(1, 601) if not torch.equal(dense[c, i], dense[c, i - 1])]
        assert rows == moved and torch.equal(ev[c, 1:m + 1, 1:], dense[c, moved])


def test_summarize_kernel_matches_torch(eng):
    """glabc_summarize (one launch) == the torch restatement in sharding.summarize, d = 1..4"""
    from glabc_b200 import sharding
    from glabc_b200.engine import RunStats
    g = torch.Generator(device="cuda").manual_seed(3)
    for d in (1, 2, 3, 4):
        Cn = 1000 + d
        raw = torch.zeros(Cn, abi.nstats(d), device="cuda")
        raw[:, 0] = 500.0
        raw[:, 1:4] = torch.randint(0, 200, (Cn, 3), device="cuda", generator=g).float()
        th = torch.randn(Cn, 500, d, device="cuda", generator=g) * 0.1
        raw[:, 4:4 + d] = th.sum(1)
        raw[:, 4 + d:4 + 2 * d] = (th * th).sum(1)
        dl = th[:, 1:] - th[:, :-1]
        t = 0
        for i in range(d):
            for j in range(i, d):
                raw[:, 4 + 2 * d + t] = (dl[:, :, i] * dl[:, :, j]).sum(1)
                t += 1
        rs = RunStats(raw, d)
        got = sharding.summarize(rs)
        want = sharding.summarize(rs, esjd_per_chain=rs.esjd())       # the torch path
        assert got.shape == want.shape == (6 + 2 * d,)
        assert torch.allclose(got, want, rtol=1e-9, atol=1e-9), (d, got, want)
