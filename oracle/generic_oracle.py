"""TEST INFRASTRUCTURE — not part of the product (only tests/ may import this).

numpy (float64) restatement of the reference's GlobalMCMC loop body (GlobalMCMC.py:37-68) for ANY of its proposal classes
in the Local / Global slots — Uniform (distribution.py:50-86), Gamma (:90-137), DiagGaussian (:143-203), GaussianMixture
(:206-293) — driven by recorded draws (replay).  Pinned by tests/golden/global_generic.npz, which was recorded from the
reference itself (tests/golden/make_golden.py `golden_global_generic`); the CUDA kernel `k_global_generic<REPLAY>` is
checked against the same recordings.  Pure-Python loop: small cases only."""
import math

import numpy as np
from scipy.special import gammaln, logsumexp

HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)


def log_prob(spec, z):
    """log density of one point z[d] under `spec` (dict: kind + parameters as in the golden file)"""
    kind = spec["kind"]
    z = np.asarray(z, np.float64)
    if kind == "gauss":       # distribution.py:176-181
        ls = np.log(np.asarray(spec["sigma"], np.float64))
        r = (z - spec["loc"]) / np.exp(ls)
        return -len(z) * HALF_LOG_2PI - np.sum(ls + 0.5 * r * r)
    if kind == "uniform":     # :79-85
        if np.any(z < spec["low"]) or np.any(z > spec["high"]):
            return -np.inf
        return -math.log(np.prod(np.asarray(spec["high"]) - np.asarray(spec["low"])))
    if kind == "gamma":       # :122-137: log pdf, -inf where the pdf is 0
        a, b = np.asarray(spec["shape"], np.float64), np.asarray(spec["rate"], np.float64)
        if np.any(z < 0) or np.any((z == 0) & (a != 1)):
            return -np.inf
        with np.errstate(divide="ignore"):
            return float(np.sum(a * np.log(b) - gammaln(a) + (a - 1) * np.log(z) - b * z))
    if kind == "mixture":     # :270-291
        loc, sc = np.asarray(spec["loc"], np.float64), np.asarray(spec["scale"], np.float64)
        w = np.asarray(spec["weights"], np.float64)
        w = w / w.sum()
        eps = (z[None, :] - loc) / sc
        return float(logsumexp(-len(z) * HALF_LOG_2PI + np.log(w) - 0.5 * np.sum(eps * eps, 1) - np.sum(np.log(sc), 1)))
    raise ValueError(kind)


def is_float64_kind(spec):
    return spec["kind"] in ("gamma", "mixture")   # scipy / nn.Parameter(float64): distribution.py:118,238-240


def replay_chain(model, lp, gp, theta0, y0, gf, tape32, tape64):
    """One chain.  model: dict(y_obs, noise_scale, eps_scale, eps_log_scale); tape32 [steps][4] = U_b, eps_sim[2], U_a;
    tape64 [steps][2] = the proposal's draw.  Returns (trace float32 [steps + 1][2], rec [steps][4] = flags, prior', kernel',
    log_acc).  Mixture.py:13-53 for the model, prior N(0, I)."""
    y_obs, ns = np.asarray(model["y_obs"], np.float64), np.asarray(model["noise_scale"], np.float64)
    eps, leps = float(model["eps_scale"]), float(model["eps_log_scale"])

    def prior(th):
        return -len(th) * HALF_LOG_2PI - 0.5 * float(np.sum(th * th))

    def kernel(y):
        dis = math.sqrt(float(np.sum((y - y_obs) ** 2)))
        return -HALF_LOG_2PI - (leps + 0.5 * (dis / eps) ** 2)

    steps = tape32.shape[0]
    theta, y, wide = np.asarray(theta0, np.float64), np.asarray(y0, np.float64), False
    trace = np.zeros((steps + 1, len(theta)), np.float32)
    trace[0] = theta
    rec = np.zeros((steps, 4))
    for s in range(steps):
        u_b, u_a, e_s, draw = tape32[s, 0], tape32[s, 3], tape32[s, 1:3].astype(np.float64), tape64[s]
        is_global = bool(np.float32(u_b) < np.float32(gf))            # GlobalMCMC.py:39
        if is_global:
            th_p, p_wide = draw.copy(), is_float64_kind(gp)
            corr = log_prob(gp, theta) - log_prob(gp, th_p)           # :45-46
        else:
            p_wide = wide or is_float64_kind(lp)
            th_p = draw + theta if p_wide else (draw.astype(np.float32) + theta.astype(np.float32)).astype(np.float64)   # :56
            corr = 0.0
        y_p = np.abs(th_p) + ns * e_s                                 # :41,57 (Mixture.py:13-26)
        pr_p, k_p = prior(th_p), kernel(y_p)
        log_acc = pr_p + k_p + corr - prior(theta) - kernel(y)        # :44-46 / :60-61
        with np.errstate(divide="ignore"):
            accept = bool(np.log(np.float32(u_a)) < log_acc)          # :47-49,62-64
        if accept:
            theta, y, wide = th_p, y_p, p_wide
        trace[s + 1] = theta
        rec[s] = [int(is_global) | (int(accept) << 1), pr_p, k_p, log_acc]
    return trace, rec


def torch_sum(v):
    """torch.sum over a short contiguous vector in ATen's order (SumKernel row_sum, four interleaved partials for n < 16; the
    16-lane vector path, tail first, for 16 <= n < 32) in the vector's own dtype — SURVEY.md B-3"""
    v = np.asarray(v)
    t = v.dtype.type
    n = len(v)
    if n >= 16:
        acc = t(0)
        for i in range(16, n):
            acc = t(acc + v[i])
        for i in range(16):
            acc = t(acc + v[i])
        return acc
    p = [t(0)] * 4
    rows = n // 4
    for r in range(rows):
        for k in range(4):
            p[k] = t(p[k] + v[4 * r + k])
    for i in range(4 * rows, n):
        p[0] = t(p[0] + v[i])
    for k in range(1, 4):
        p[0] = t(p[0] + p[k])
    return p[0]


def replay_isir_chain(model, lp, ip, theta0, y0, gf, K, tape32, tape64):
    """One GLMCMC chain (GLMCMC.py:48-104, weight_sampling :7-22) with any proposal classes, driven by recorded draws.
    tape32 [steps][2 + 2K] = U_b, eps_sim[K][2] (local: eps_sim[2] first), U_a (last); tape64 [steps][1 + 2K] = the numpy
    resampling uniform, then the proposal's draws.  dtype rules of the reference: a Gamma / GaussianMixture draw is float64, so
    the state (torch.cat / the local add promote) and from then on the log-weights are float64 tensors — whose exp does not
    underflow near -104 as the float32 one does (B-1).  Returns (trace float32 [steps + 1][2], rec [steps][4 + K])."""
    y_obs, ns = np.asarray(model["y_obs"], np.float64), np.asarray(model["noise_scale"], np.float64)
    eps, leps = float(model["eps_scale"]), float(model["eps_log_scale"])

    def prior(th):
        return -len(th) * HALF_LOG_2PI - 0.5 * float(np.sum(th * th))

    def kernel(y):
        dis = math.sqrt(float(np.sum((y - y_obs) ** 2)))
        return -HALF_LOG_2PI - (leps + 0.5 * (dis / eps) ** 2)

    steps = tape32.shape[0]
    theta, y, wide = np.asarray(theta0, np.float64), np.asarray(y0, np.float64), False
    ip64, lp64 = is_float64_kind(ip), is_float64_kind(lp)
    lw_old, lw_wide = prior(theta) + kernel(y) - log_prob(ip, theta), False       # :52-55
    local = True
    trace = np.zeros((steps + 1, len(theta)), np.float32)
    trace[0] = theta
    rec = np.zeros((steps, 4 + K))
    for s in range(steps):
        u_b = tape32[s, 0]
        is_global = bool(np.float32(u_b) < np.float32(gf))                         # :59
        prev = theta.astype(np.float32)
        if is_global:
            if local:                                                              # :60-64
                lw_old, lw_wide = prior(theta) + kernel(y) - log_prob(ip, theta), wide
            local = False
            u64 = float(tape64[s, 0])
            th = tape64[s, 1:1 + 2 * K].reshape(K, 2)
            x = np.abs(th) + ns * tape32[s, 1:1 + 2 * K].reshape(K, 2).astype(np.float64)          # :71
            lw = np.array([prior(th[j]) + kernel(x[j]) - log_prob(ip, th[j]) for j in range(K)])    # :72-74
            w64 = lw_wide or ip64                                                  # dtype of torch.cat((log_weight_old, log_weight0))
            allw = np.concatenate(([lw_old], lw)).astype(np.float64 if w64 else np.float32)
            with np.errstate(over="ignore", under="ignore", invalid="ignore"):
                w = np.exp(allw)                                                   # :78, no max-shift
                w[np.isnan(w)] = 0                                                 # :80-81
                S = torch_sum(w)
                wn = w / S                                                         # :82 (all-zero -> NaN row)
            ind, run = None, 0.0
            for j in range(K + 1):                                                 # weight_sampling, :7-22: float64 running sum
                run += float(wn[j])
                if u64 < run:
                    ind = j
                    break
            if ind is not None and ind != 0:                                       # :84-88
                theta, y, lw_old = th[ind - 1].copy(), x[ind - 1].copy(), float(allw[ind])
                wide, lw_wide = wide or ip64, w64
            changed = bool(np.any(theta.astype(np.float32) != prev))
            rec[s, 0] = 1 | (int(changed) << 1) | ((0 if ind is None else ind + 1) << 8) | (int(w64) << 16)
            rec[s, 1], rec[s, 2], rec[s, 3] = float(allw[0]), float(S), float(wn[0])
            rec[s, 4:] = lw
        else:
            z, e_s, u_a = tape64[s, 1:3], tape32[s, 1:3].astype(np.float64), tape32[s, 1 + 2 * K]
            p_wide = wide or lp64
            th_p = z + theta if p_wide else (z.astype(np.float32) + theta.astype(np.float32)).astype(np.float64)   # :91
            y_p = np.abs(th_p) + ns * e_s
            pr_p, k_p = prior(th_p), kernel(y_p)
            log_acc = pr_p + k_p - prior(theta) - kernel(y)                        # :96-97
            with np.errstate(divide="ignore"):
                accept = bool(np.log(np.float32(u_a)) < log_acc)                   # :98-99
            if accept:
                theta, y, wide, local = th_p, y_p, p_wide, True                    # :100-103
            rec[s, 0] = int(accept) << 1
            rec[s, 1], rec[s, 2], rec[s, 3] = pr_p, k_p, log_acc
        trace[s + 1] = theta
    return trace, rec
