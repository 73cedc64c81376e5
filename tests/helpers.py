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
