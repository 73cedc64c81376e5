"""Shared test helpers: golden loading, POD construction from golden params, oracle access."""
import os

import numpy as np

import glabc_b200  # noqa: F401  (alias import registers the package)
from glabc_b200 import _abi as abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases(name):
    z = np.load(os.path.join(GOLDEN, name))
    n = int(z["n_cases"])
    cases = []
    for i in range(n):
        pre = f"case{i}/"
        cases.append({k[len(pre):]: z[k] for k in z.files if k.startswith(pre)})
    return cases


def model_pod(case, family=abi.MODEL_ABS_NORMAL):
    d = len(case["y_obs"])
    m = abi.ModelPOD(family=family, theta_dim=d, y_dim=d)
    abi.fill(m.y_obs, case["y_obs"])
    abi.fill(m.noise_loc, case["noise_loc"])
    abi.fill(m.noise_scale, case["noise_scale"])
    abi.fill(m.prior_loc, case["prior_loc"])
    abi.fill(m.prior_log_scale, case["prior_log_scale"])
    abi.fill(m.prior_scale, case["prior_scale"])
    m.eps_log_scale = float(case["eps_log_scale"])
    m.eps_scale = float(case["eps_scale"])
    m.epsilon = float(case["epsilon"]) if "epsilon" in case else float(case["eps_scale"])
    return m


def gauss_pod(case, prefix):
    loc = case[prefix + "_loc"]
    p = abi.DistPOD(kind=abi.DIST_DIAG_GAUSSIAN, dim=len(loc))
    abi.fill(p.a, loc)
    abi.fill(p.b, case[prefix + "_log_scale"])
    abi.fill(p.c, case[prefix + "_scale"])
    return p


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


def fresh_mala_state(c):
    aux = np.zeros((c, abi.AUX_SLOTS), np.float32)
    aux[:, abi.AUX_LOCAL] = 1.0
    return aux, np.zeros((c, abi.STATE64_SLOTS), np.float64)


def mala_teacher_forced(case):
    """Every (step, chain) of a GLMALA golden case as an independent one-step pseudo-chain started from the
    reference's own recorded state before that step.  The reference's float32 finite-difference prior
    gradient (GLMALA.py:84-85, h = 1e-5) turns 1e-8 differences in theta into 1e-2 differences in the
    gradient, so free-running chains agree on decisions but not to 1e-5 on values; one-step restarts do."""
    T, Cn, K, num = int(case["T"]), case["theta0"].shape[0], int(case["K"]), int(case["num_grad"])
    S = T - 1
    n = S * Cn
    st = case["state"][:S]                                       # [S, 8, C] state before step s
    flat = lambda a: np.ascontiguousarray(a.reshape(a.shape[0] * a.shape[1], *a.shape[2:]))  # noqa: E731
    theta64 = flat(np.moveaxis(st[:, 0:2], 1, 2))                # [S*C, 2]
    y64 = flat(np.moveaxis(st[:, 2:4], 1, 2))
    grad = flat(np.moveaxis(st[:, 4:6], 1, 2))
    lw = st[:, 6].reshape(-1)
    bits = st[:, 7].reshape(-1).astype(np.int64)
    aux = np.zeros((n, abi.AUX_SLOTS), np.float32)
    aux[:, abi.AUX_LOCAL] = bits & 1
    aux[:, abi.AUX_WIDE] = (bits >> 1) & 1
    aux[:, abi.AUX_LW_WIDE] = (bits >> 2) & 1
    aux[:, abi.AUX_HAVE_GRAD] = (bits >> 3) & 1
    s64 = np.zeros((n, abi.STATE64_SLOTS), np.float64)
    s64[:, abi.S64_THETA:abi.S64_THETA + 2] = theta64
    s64[:, abi.S64_Y:abi.S64_Y + 2] = y64
    s64[:, abi.S64_GRAD:abi.S64_GRAD + 2] = grad
    s64[:, abi.S64_LOGW] = lw
    slots = case["tape32"].shape[1]
    tape32 = np.ascontiguousarray(np.moveaxis(case["tape32"], 1, 0).reshape(1, slots, n))   # [1, slots, S*C]
    tape64 = np.ascontiguousarray(case["tape64"].reshape(1, n))
    grad0 = np.ascontiguousarray(np.broadcast_to(case["tape_grad0"][:, None, :], (case["tape_grad0"].shape[0], S, Cn)).reshape(-1, n))
    rec = np.ascontiguousarray(np.moveaxis(case["rec"], 1, 0).reshape(case["rec"].shape[1], n))  # [slots, S*C]
    grec = np.ascontiguousarray(np.moveaxis(case["grec"], 1, 0).reshape(case["grec"].shape[1], n))
    eps2 = float(case["epsilon"]) ** 2 if "epsilon" in case else float(case["eps_scale"]) ** 2
    kern_c = -0.5 * np.log(2 * np.pi) - float(case["eps_log_scale"])
    return dict(n=n, grec=grec, eps2=eps2, kern_c=kern_c, theta=theta64.astype(np.float32), y=y64.astype(np.float32), aux=aux, state64=s64, tape32=tape32,
                tape64=tape64, tape_grad0=grad0, rec=rec, K=K, num_grad=num, tau=float(case["tau"]), gf=float(case["gf"]))


def mala_grad_parts(mu_p, mu_m, s_p, s_m, eps2):
    """the likelihood part of numberical_gradient_logABC from its float64 statistics (GLMALA.py:90-94)"""
    lp = -0.5 * np.log(s_p + eps2) - 0.5 * mu_p ** 2 / (s_p + eps2)
    lm = -0.5 * np.log(s_m + eps2) - 0.5 * mu_m ** 2 / (s_m + eps2)
    return (lp - lm) / (2 * 1e-1)


def mala_grad_bound(mu, sig, eps2, n, u=2.0 ** -23):
    """First-order bound on |d logp| (GLMALA.py:90-93) when each of the n float32 discrepancies d_j moves by at most one
    unit in the last place, |delta_j| <= u d_j.  Why one ulp: the reference takes torch.sqrt, which on the CPU build that
    made the fixtures (MKL VML) returns the float32 BELOW the correctly rounded root for 0.64 % of its inputs
    (tests/golden/README in DESIGN.md section 2); IEEE sqrt (oracle, kernels) differs from it by exactly that ulp.
      |d mu| <= u mu,   |d Sigma| <= 2/(n-1) sum |d_j - mu| |delta_j| <= 2 u sqrt(Sigma) sqrt(Sigma + n mu^2 / (n-1))
      d logp = -(mu / V) d mu + (-1/(2V) + mu^2 / (2 V^2)) d Sigma,   V = Sigma + eps^2"""
    v = sig + eps2
    d_mu = u * np.abs(mu)
    d_sig = 2.0 * u * np.sqrt(sig) * np.sqrt(sig + n * mu ** 2 / (n - 1.0))
    return np.abs(mu / v) * d_mu + np.abs(-0.5 / v + 0.5 * mu ** 2 / v ** 2) * d_sig


def check_mala_debug(dbg, rec, K, tol=1e-5, grec=None, eps2=None, num_grad=None, kern_c=2.08):
    """dbg [DEBUG64_SLOTS, n] (kernel / oracle) vs rec [12+K, n] (reference): flags exact, values within `tol`
    RELATIVE with no absolute slack.  Two quantities are not judged by a bare relative error, with the reason:
      * log K(y') crosses zero (log K(0) = +2.08, decreasing in the distance), so its error is taken relative to the
        larger of its two terms, |c| and (dis/eps)^2 / 2, not to their difference;
      * grad' is a finite difference of synthetic log-likelihoods divided by 0.2 (GLMALA.py:94): it is pinned through
        its ingredients — mu+-, Sigma+- (the reference's own torch.mean / torch.var outputs, `grec`) within `tol` relative,
        the combination step by recomputing it from the implementation's own statistics, and the value itself within
        `tol` relative plus the first-order effect of a one-ulp perturbation of the float32 discrepancies
        (mala_grad_bound: the reference's sqrt is not correctly rounded)."""
    fl_o, fl_r = dbg[0].astype(np.int64), rec[0].astype(np.int64)
    assert np.array_equal(fl_o, fl_r), f"{(fl_o != fl_r).sum()} decisions differ"
    loc = (fl_r & 1) == 0
    worst = {}

    def cmp(name, a, b, m, scale=None, slack=None, rtol=tol):
        a, b = a[m], b[m]
        if a.size:
            den = np.maximum(np.abs(b), 1e-300) if scale is None else np.maximum(scale[m], 1e-300)
            excess = np.abs(a - b) - (0.0 if slack is None else slack[m])
            err = np.maximum(excess, 0.0) / den
            worst[name] = float(err.max())
            assert err.max() <= rtol, (name, float(err.max()))

    # log_acc = prior' + kern' + lq_rev - prior - kern - lq_fwd (GLMALA.py:190-193): relative to its largest term
    acc_scale = np.maximum.reduce([np.abs(rec[1]), np.abs(rec[8]), np.abs(rec[9]), np.abs(rec[10]), np.abs(rec[11])])
    cmp("log_acc", dbg[1], rec[1], loc, scale=acc_scale)
    for k in range(2):
        cmp(f"theta'{k}", dbg[2 + k], rec[2 + k], loc)
        cmp(f"y'{k}", dbg[6 + k], rec[4 + k], loc)
    cmp("prior'", dbg[14], rec[8], loc)
    cmp("kern'", dbg[15], rec[9], loc, scale=np.maximum(np.abs(rec[9]), abs(kern_c)))   # kern_c = -(1/2) log 2 pi - log eps
    cmp("lq_rev", dbg[16], rec[10], loc)
    cmp("lq_fwd", dbg[17], rec[11], loc)
    if grec is not None:
        d = 2
        mu_p, mu_m, s_p, s_m = grec[0:d], grec[d:2 * d], grec[2 * d:3 * d], grec[3 * d:4 * d]
        for k in range(d):
            cmp(f"mu+{k}", dbg[20 + k], mu_p[k], loc)
            cmp(f"mu-{k}", dbg[24 + k], mu_m[k], loc)
            cmp(f"Sigma+{k}", dbg[28 + k], s_p[k], loc)
            cmp(f"Sigma-{k}", dbg[32 + k], s_m[k], loc)
            like_ref = mala_grad_parts(mu_p[k], mu_m[k], s_p[k], s_m[k], eps2)
            gprior = rec[6 + k] - like_ref                       # the float32 finite-difference prior gradient (GLMALA.py:84-85)
            like_own = mala_grad_parts(dbg[20 + k], dbg[24 + k], dbg[28 + k], dbg[32 + k], eps2)
            # the combination step, from the implementation's own statistics: exact up to float64 rounding
            cmp(f"grad'{k} (own statistics)", dbg[10 + k], like_own + gprior, loc, rtol=1e-9,
                scale=np.maximum(np.abs(like_own), np.abs(gprior)))
            bound = (mala_grad_bound(mu_p[k], s_p[k], eps2, num_grad) + mala_grad_bound(mu_m[k], s_m[k], eps2, num_grad)) / 0.2
            cmp(f"grad'{k}", dbg[10 + k], rec[6 + k], loc, slack=bound)
            worst[f"grad'{k} abs / bound"] = float((np.abs(dbg[10 + k] - rec[6 + k])[loc] / bound[loc]).max()) if loc.any() else 0.0
    else:
        for k in range(2):
            cmp(f"grad'{k}", dbg[10 + k], rec[6 + k], loc, scale=np.maximum(np.abs(rec[6 + k]), 1.0) * 100)
    g = ~loc
    cmp("lw_old", dbg[1], rec[1], g & np.isfinite(rec[1]))
    pos = g & (rec[2] > 1e-300)
    cmp("S", dbg[2], rec[2], pos)
    cmp("w0", dbg[3], rec[3], pos)
    for j in range(K):
        cmp(f"lw{j}", dbg[4 + j], rec[4 + j], g & np.isfinite(rec[4 + j]))
    return worst


def check_mala_free_running(flags, trace, case, strict_all=False):
    """free-running GLMALA replay vs the reference: decisions are bit-exact until float32 finite-difference
    noise (see mala_teacher_forced) flips one — rare; traces of the chains without a flip agree to 2e-2."""
    ref_fl = case["rec"][:, 0].astype(np.int64)
    ne = flags.astype(np.int64) != ref_fl                       # [S, C]
    first = np.where(ne.any(0), ne.argmax(0), ne.shape[0])       # first mismatching step per chain
    clean = first == ne.shape[0]
    assert clean.mean() >= 0.75, f"decisions diverged in {(~clean).sum()} of {clean.size} chains"
    if strict_all:
        assert clean.all()
    for c in np.flatnonzero(clean):
        assert np.allclose(trace[:, c], case["trace"][:, c], rtol=0, atol=2e-2), c
    return clean


def rel_max(a, b, floor=1e-30):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max()) if a.size else 0.0


def check_aglmcmc(case, trace, dbg, ad_rec, ad_blk, init_w, tol=1e-5):
    """AGLMCMC replay outputs (oracle or kernel) vs the reference's records: decisions and resample indices
    bit-exact; eps-hat, bandwidths, block log-densities / log-weights, per-step densities within `tol` relative;
    traces to float32 rounding (the KDE bandwidth carries ~1e-7 relative summation-order noise into theta0)."""
    rec = case["rec"]
    fl_o, fl_r = dbg[:, 0].astype(np.int64), rec[:, 0].astype(np.int64)
    assert np.array_equal(fl_o, fl_r), f"{(fl_o != fl_r).sum()} decisions differ"
    assert np.allclose(trace[1:], case["trace"][1:], rtol=2e-6, atol=2e-6)   # row 0: the reference leaves zeros (B-10)
    assert np.array_equal(trace[0], case["theta0"])
    assert rel_max(init_w, case["init_w"], 1e-30) < 20 * tol                   # exp of O(100) log-weights
    n = int(case["n_adapt"].min())
    d = case["theta0"].shape[1]
    assert rel_max(ad_rec[:n, 0], case["ad_rec"][:n, 0]) < tol                 # eps-hat (torch.quantile)
    assert np.array_equal(ad_rec[:n, 1], case["ad_rec"][:n, 1].astype(np.float32))  # KDE training points kept
    assert rel_max(ad_rec[:n, 2:2 + d], case["ad_rec"][:n, 2:2 + d]) < tol     # bandwidth
    assert np.allclose(ad_blk[:n, :, 0:d], case["ad_theta"][:n], rtol=2e-6, atol=2e-6)
    assert rel_max(ad_blk[:n, :, d], case["ad_lq"][:n], 1.0) < tol             # KDE.log_prob of the block
    assert np.allclose(ad_blk[:n, :, d + 2], case["ad_dis"][:n], rtol=1e-5, atol=2e-6)
    wr, wo = case["ad_w"][:n], ad_blk[:n, :, d + 1]
    pos = wr > 1e-30
    assert np.array_equal(wo > 0, wr > 0)
    lw_err = np.abs(np.log(wo[pos].astype(np.float64)) - np.log(wr[pos].astype(np.float64)))
    assert lw_err.max() < 3e-4, lw_err.max()                                    # log-weights are O(10..100): 1e-5 relative
    g = (fl_r & 1) == 1
    assert rel_max(dbg[:, 1][g], rec[:, 1][g], 1.0) < tol                      # proposal log-density of the current state
    m = g & (rec[:, 2] > 1e-30)
    assert np.abs(np.log(dbg[:, 2][m].astype(np.float64)) - np.log(rec[:, 2][m])).max() < 3e-4
    loc = ~g
    for k in (1, 2, 3):
        assert rel_max(dbg[:, k][loc], rec[:, k][loc], 1e-2) < 20 * tol
