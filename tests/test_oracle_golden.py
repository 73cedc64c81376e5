"""Pins the CPU oracle (oracle/glabc_oracle.c) against the golden vectors recorded from the
reference's own Python (tests/golden/make_golden.py).  Bar (BASELINE.json north_star): accept /
branch decisions and resample indices bit-exact; log-densities within 1e-5 relative."""
import numpy as np
import pytest

from helpers import abi, gauss_pod, load_cases, model_pod, rel_err
from oracle import oracle


@pytest.mark.parametrize("ci", range(4))
def test_global_mcmc_replay(ci):
    case = load_cases("global_mcmc.npz")[ci]
    T, C = int(case["T"]), case["theta0"].shape[0]
    theta, y = case["theta0"].copy(), case["y0"].copy()
    debug = np.zeros((T - 1, abi.DEBUG_SLOTS, C), np.float32)
    stats = np.zeros((C, abi.nstats(2)), np.float32)
    trace = oracle.run("global", model_pod(case), gauss_pod(case, "lp"), gauss_pod(case, "gp"), theta=theta, y=y,
                       n_steps=T - 1, gf=float(case["gf"]), rng_mode=abi.RNG_REPLAY,
                       tape32=np.ascontiguousarray(case["tape32"]), debug=debug, stats=stats)
    rec = case["rec"]
    # decisions bit-exact, traces bit-exact (everything on the path but log(U) is +,-,*,/,sqrt)
    assert np.array_equal(debug[:, 0].astype(np.int32), rec[:, 0].astype(np.int32))
    assert np.array_equal(trace, case["trace"])
    # log-densities: 1e-5 relative.  The prior is +,-,*,/ only and comes out bit-exact; the kernel
    # goes through torch.sqrt, which on the reference's MKL-VML CPU path is not correctly rounded
    # (e.g. sqrt(0.66828066) -> 0.81748432, IEEE gives 0.81748438), so it agrees to ~4e-7.
    for k, name in ((1, "prior"), (2, "kernel"), (3, "log_acc")):
        assert rel_err(debug[:, k], rec[:, k]).max() <= 1e-5, name
    assert np.array_equal(debug[:, 1], rec[:, 1])
    assert np.array_equal(theta, case["trace"][-1])
    flags = rec[:, 0].astype(np.int32)
    assert np.array_equal(stats[:, abi.STAT_STEPS], np.full(C, T - 1, np.float32))
    assert np.array_equal(stats[:, abi.STAT_GLOBAL_STEPS], (flags & 1).sum(0))
    assert np.array_equal(stats[:, abi.STAT_ACC_GLOBAL], ((flags & 3) == 3).sum(0))
    assert np.array_equal(stats[:, abi.STAT_ACC_LOCAL], ((flags & 3) == 2).sum(0))


@pytest.mark.parametrize("ci", range(4))
def test_glmcmc_replay(ci):
    case = load_cases("glmcmc.npz")[ci]
    T, C, K = int(case["T"]), case["theta0"].shape[0], int(case["K"])
    theta, y = case["theta0"].copy(), case["y0"].copy()
    aux = np.zeros((C, abi.AUX_SLOTS), np.float32)
    aux[:, abi.AUX_LOCAL] = 1.0
    debug = np.zeros((T - 1, abi.DEBUG_SLOTS, C), np.float32)
    trace = oracle.run("isir", model_pod(case), gauss_pod(case, "lp"), gauss_pod(case, "ip"), theta=theta, y=y,
                       aux=aux, n_steps=T - 1, gf=float(case["gf"]), rng_mode=abi.RNG_REPLAY, K=K,
                       tape32=np.ascontiguousarray(case["tape32"]), tape64=np.ascontiguousarray(case["tape64"]),
                       debug=debug)
    rec = case["rec"]
    assert np.array_equal(debug[:, 0].astype(np.int32), rec[:, 0].astype(np.int32))  # branch, move, index
    assert np.array_equal(trace, case["trace"])
    n = rec.shape[1]
    both = np.isfinite(rec[:, 1:n]) & np.isfinite(debug[:, 1:n])
    assert np.array_equal(np.isfinite(rec[:, 1:n]), np.isfinite(debug[:, 1:n]))
    assert rel_err(debug[:, 1:n][both], rec[:, 1:n][both]).max() <= 1e-5


def test_golden_has_the_edge_cases():
    """the fixtures exercise: never/always global, `None` resample (all weights underflow, B-1)."""
    g = load_cases("global_mcmc.npz")
    assert (g[1]["rec"][:, 0].astype(int) & 1).sum() == 0 and (g[2]["rec"][:, 0].astype(int) & 1).all()
    i = load_cases("glmcmc.npz")
    fl = np.concatenate([c["rec"][:, 0].astype(int).ravel() for c in i])
    assert (((fl & 1) == 1) & ((fl >> 8) == 0)).sum() > 0
    assert len({int(c["K"]) for c in i}) == 4


@pytest.mark.parametrize("ci", range(12))
def test_glmala_replay(ci):
    """GLMALA.py:150-200 on the reference's own draws.  Free-running: every branch / accept / resample
    decision bit-exact, traces equal up to the amplification of float32 finite-difference noise
    (helpers.mala_teacher_forced).  Restarted from the reference's recorded state before each step:
    theta', y', gradient, log-densities and log_acc within 1e-5 relative."""
    from helpers import check_mala_debug, check_mala_free_running, fresh_mala_state, mala_teacher_forced
    case = load_cases("glmala.npz")[ci]
    T, Cn, K, num = int(case["T"]), case["theta0"].shape[0], int(case["K"]), int(case["num_grad"])
    theta, y = case["theta0"].copy(), case["y0"].copy()
    aux, s64 = fresh_mala_state(Cn)
    dbg = np.zeros((T - 1, abi.DEBUG64_SLOTS, Cn))
    tr = oracle.run("mala", model_pod(case), None, gauss_pod(case, "ip"), theta=theta, y=y, n_steps=T - 1,
                    gf=float(case["gf"]), rng_mode=abi.RNG_REPLAY, tape32=case["tape32"], tape64=case["tape64"],
                    tape_grad0=case["tape_grad0"], aux=aux, state64=s64, debug64=dbg, K=K, num_grad=num,
                    tau=float(case["tau"]))
    clean = check_mala_free_running(dbg[:, 0], tr, case, strict_all=False)
    final = case["state"][-1]
    assert np.array_equal(aux[clean, abi.AUX_WIDE], ((final[7].astype(np.int64) >> 1) & 1).astype(np.float32)[clean])

    tf = mala_teacher_forced(case)
    dbg1 = np.zeros((1, abi.DEBUG64_SLOTS, tf["n"]))
    oracle.run("mala", model_pod(case), None, gauss_pod(case, "ip"), theta=tf["theta"], y=tf["y"], n_steps=1, gf=tf["gf"],
               rng_mode=abi.RNG_REPLAY, tape32=tf["tape32"], tape64=tf["tape64"], tape_grad0=tf["tape_grad0"], aux=tf["aux"],
               state64=tf["state64"], debug64=dbg1, K=K, num_grad=num, tau=tf["tau"], trace_layout=abi.TRACE_NONE)
    worst = check_mala_debug(dbg1[0], tf["rec"], K, grec=tf["grec"], eps2=tf["eps2"], num_grad=num, kern_c=tf["kern_c"])
    print(worst)


def test_kde_golden():
    """kernel_density.py:22-128 — fit (weights, Silverman / Scott bandwidth from the weighted std) and log_prob,
    weighted and unweighted, d = 1..3, including queries so far out that every kernel underflows without the shift"""
    import os
    from helpers import GOLDEN, rel_max
    z = np.load(os.path.join(GOLDEN, "kde.npz"))
    for i in range(int(z["n_cases"])):
        X, x, w = z[f"kde{i}/X"], z[f"kde{i}/x"], z[f"kde{i}/w"]
        weights, bw = oracle.kde_fit(X, w if w.size else None, int(z[f"kde{i}/rule"]))
        assert rel_max(weights, z[f"kde{i}/weights"]) < 1e-6 and rel_max(bw, z[f"kde{i}/bw"]) < 1e-6
        lp = oracle.kde_log_prob(X, weights, bw, x)
        assert np.isfinite(lp).all() and rel_max(lp, z[f"kde{i}/log_prob"], 1.0) < 1e-5


def run_aglmcmc_oracle(case, **over):
    T, Cn, K, S = int(case["T"]), case["theta0"].shape[0], int(case["K"]), int(case["S"])
    B, R = K * S, case["ad_idx"].shape[0]
    theta, y = case["theta0"].copy(), case["y0"].copy()
    out = dict(dbg=np.zeros((T - 1, abi.DEBUG_SLOTS, Cn), np.float32), ad_rec=np.zeros((R, abi.AG_REC_SLOTS, Cn), np.float32),
               ad_blk=np.zeros((R, B, case["theta0"].shape[1] + 3, Cn), np.float32), init_w=np.zeros((B, Cn), np.float32))
    ag = oracle.aglmcmc_params(S=S, alpha=float(case["alpha"]), hat_eps_T=float(case["hat_eps_T"]), init_p=case["init_p"],
                               init_s=case["init_s"], ad_idx=case["ad_idx"], ad_noise=case["ad_noise"], ad_sim=case["ad_sim"],
                               ad_rec=out["ad_rec"], ad_blk=out["ad_blk"], init_w=out["init_w"])
    out["trace"] = oracle.run("aglmcmc", model_pod(case), gauss_pod(case, "lp"), gauss_pod(case, "ip"), theta=theta, y=y,
                              n_steps=T - 1, gf=float(case["gf"]), rng_mode=abi.RNG_REPLAY, tape32=case["tape32"],
                              tape64=case["tape64"], debug=out["dbg"], K=K, ag=ag, **over)
    return out


@pytest.mark.parametrize("ci", range(3))
def test_aglmcmc_replay(ci):
    """AGLMCMC.py:84-272 on the reference's own draws (incl. its torch.multinomial indices): every branch / move /
    resample index bit-exact through 16-18 adaptations per chain; eps-hat, KDE bandwidths, block log-densities and
    weights to 1e-5"""
    from helpers import check_aglmcmc
    case = load_cases("aglmcmc.npz")[ci]
    o = run_aglmcmc_oracle(case)
    check_aglmcmc(case, o["trace"], o["dbg"], o["ad_rec"], o["ad_blk"], o["init_w"])


def test_resample_golden():
    """systematic resampling GLMCMC_NFs.py:29-40 (host logic of the flow training step): the vectorised searchsorted
    form returns the reference's exact index lists, including the short return when the cumulative sum tops out below 1"""
    import os
    import torch
    from helpers import GOLDEN
    from glabc_b200.GLMCMC_NFs import resample
    z = np.load(os.path.join(GOLDEN, "resample.npz"))
    for i in range(int(z["n_cases"])):
        torch.manual_seed(int(z[f"rs{i}/seed"]))
        got = resample(torch.from_numpy(z[f"rs{i}/W"]), int(z[f"rs{i}/N"]))
        assert np.array_equal(got.numpy(), z[f"rs{i}/idx"]), i
    assert len(z["rs2/idx"]) < int(z["rs2/N"])


def test_generic_proposal_oracle_reproduces_the_reference_recordings():
    """GlobalMCMC with GaussianMixture / Uniform / Gamma proposals: the numpy restatement (oracle/generic_oracle.py), fed
    the reference's recorded draws, reproduces its decisions and its float32 chains bit for bit (tests/golden/global_generic.npz)"""
    import os
    from helpers import GOLDEN
    from oracle import generic_oracle as go
    z = np.load(os.path.join(GOLDEN, "global_generic.npz"))
    kinds = ["gauss", "uniform", "gamma", "mixture"]
    for ci in range(int(z["n_cases"])):
        p = lambda k: z[f"case{ci}/{k}"]  # noqa: E731

        def spec(prefix):
            out = {"kind": kinds[int(p(prefix + "_kind"))]}
            for key in z.files:
                if key.startswith(f"case{ci}/{prefix}_") and not key.endswith("_kind"):
                    out[key.split(f"{prefix}_", 1)[1]] = z[key]
            return out
        lp, gp = spec("lp"), spec("gp")
        model = dict(y_obs=p("y_obs"), noise_scale=p("noise_scale"), eps_scale=p("eps_scale"), eps_log_scale=p("eps_log_scale"))
        for c in range(min(3, p("theta0").shape[0])):
            trace, rec = go.replay_chain(model, lp, gp, p("theta0")[c], p("y0")[c], float(p("gf")), p("tape32")[:, :, c], p("tape64")[:, :, c])
            assert np.array_equal(rec[:, 0], p("rec")[:, 0, c]), (ci, c)
            assert np.array_equal(trace, p("trace")[:, c]), (ci, c)
            want = p("rec")[:, 1:4, c]
            fin = np.isfinite(want)
            assert np.array_equal(np.isneginf(rec[:, 1:4]), np.isneginf(want))
            assert np.max(np.abs(rec[:, 1:4][fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1.0)) < 1e-4   # the reference is float32 before promotion


def test_generic_isir_oracle_reproduces_the_reference_recordings():
    """GLMCMC with Uniform / GaussianMixture / Gamma proposals: the numpy restatement (oracle/generic_oracle.py
    replay_isir_chain), fed the reference's recorded draws, reproduces every branch, resample index (incl. None), weight dtype
    and its float32 chains bit for bit (tests/golden/glmcmc_generic.npz)"""
    from oracle import generic_oracle as go
    kinds = ["gauss", "uniform", "gamma", "mixture"]

    def spec(case, pre):
        d = {"kind": kinds[int(case[pre + "_kind"])]}
        d.update({k[len(pre) + 1:]: case[k] for k in case if k.startswith(pre + "_") and k != pre + "_kind"})
        return d
    seen64 = seen_none = 0
    for case in load_cases("glmcmc_generic.npz"):
        lp, ip = spec(case, "lp"), spec(case, "ip")
        model = dict(y_obs=case["y_obs"], noise_scale=case["noise_scale"], eps_scale=case["eps_scale"], eps_log_scale=case["eps_log_scale"])
        K = int(case["K"])
        for c in range(case["theta0"].shape[0]):
            tr, rec = go.replay_isir_chain(model, lp, ip, case["theta0"][c], case["y0"][c], float(case["gf"]), K,
                                           case["tape32"][:, :, c], case["tape64"][:, :, c])
            assert np.array_equal(tr, case["trace"][:, c])
            assert np.array_equal(rec[:, 0], case["rec"][:, 0, c])
            fl = rec[:, 0].astype(int)
            seen64 += int((((fl >> 16) & 1) == 1).sum())
            seen_none += int((((fl & 1) == 1) & (((fl >> 8) & 0xff) == 0)).sum())
    assert seen64 > 100 and seen_none > 0      # the recordings exercise float64 weights and `None` resamples
