#!/usr/bin/env python
"""Times the UNMODIFIED reference package on the host cores — the number BASELINE.md / north_star ask for beside the GPU
figure ("the reference's CPU path timed on the same box's host cores in the same run, with the core count stated").

The package is the pip install of /root/reference under baseline/_ref/ (git-ignored; made by __graft_entry__.build() in the
build container, it travels to the GPU box with the snapshot), plus the reference's own example model
examples/Mixture.py copied next to it.  Nothing of this repo is on that path: the reference's MCMCRunner.run_global_mcmc
(MCMCRunner.py:17-33 -> GlobalMCMC.py:6-98) runs its stock Python loop, one chain per process, torch pinned to one thread per
process, exactly as examples/Mixture_hyper.py:32-37 times it.  `normflows` / `matplotlib` are absent from the image and
unused on this path; empty stub modules let `import glabcmcmc` succeed (as tests/golden/make_golden.py does).

    python baseline/ref_python.py --worker ITS SEED      one chain, prints {"its", "sec", "esjd"}
    measure(its, procs)                                   `procs` workers at once -> aggregate chain-steps/s
"""
import json
import os
import subprocess
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "glabcmcmc", "GlobalMCMC.py")) and os.path.exists(os.path.join(REF, "examples", "Mixture.py"))


def worker(its, seed, sampler="global"):
    for name in ("normflows", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path[:0] = [REF, os.path.join(REF, "examples")]
    import numpy as np
    import torch
    torch.set_num_threads(1)
    import tqdm
    tqdm.tqdm = lambda it, *a, **k: it                      # progress bars off (GlobalMCMC.py:37 wraps range in tqdm)
    import glabcmcmc.GlobalMCMC
    import glabcmcmc.GLMCMC
    for m in (glabcmcmc.GlobalMCMC, glabcmcmc.GLMCMC):
        m.tqdm = lambda it, *a, **k: it
    import glabcmcmc.distribution as distribution
    from glabcmcmc.ESJD import esjd
    from glabcmcmc.MCMCRunner import MCMCRunner
    from Mixture import Mixture_set
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = Mixture_set(epsilon=0.05)                        # README.md:108-122 / BASELINE configs[0]
    theta0 = torch.tensor([0.0, 0.0])
    y0 = model.generate_samples(theta0)
    lp = distribution.DiagGaussian(2, loc=torch.zeros(1, 2), log_scale=torch.log(torch.tensor([0.35, 0.35])))
    gp = distribution.DiagGaussian(2, torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]))
    runner = MCMCRunner(model, output_dir=os.environ.get("TMPDIR", "/tmp"))
    devnull = open(os.devnull, "w")
    out, sys.stdout = sys.stdout, devnull                    # the end-of-run summary print (GlobalMCMC.py:77-97)
    try:
        t0 = time.perf_counter()
        if sampler == "glmcmc":
            chain = runner.run_glmcmc(its, theta0, y0, 0.9, lp, gp, 5, output_file=None)
        else:
            chain = runner.run_global_mcmc(its, theta0, y0, 0.5, lp, gp, output_file=None)
        dt = time.perf_counter() - t0
    finally:
        sys.stdout = out
    print(json.dumps({"its": its - 1, "sec": dt, "esjd": float(esjd(chain))}), flush=True)


def measure(its=20001, procs=None, sampler="global"):
    """One reference chain of `its` iterations per process, `procs` processes at once (default: every host core).
    Returns a dict for cpu_baseline.reference_python, or {"unavailable": why}."""
    if not available():
        return {"unavailable": "baseline/_ref is not installed (run __graft_entry__.build() where /root/reference exists)"}
    procs = procs or os.cpu_count() or 1
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(its), str(1 + i), sampler],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for i in range(procs)]
    rows, errs = [], []
    for p in ps:
        o, e = p.communicate()
        try:
            rows.append(json.loads(o.strip().splitlines()[-1]))
        except (ValueError, IndexError):
            errs.append(e[-300:])
    wall = time.perf_counter() - t0
    if not rows:
        return {"unavailable": "reference workers failed: " + (errs[0] if errs else "no output")}
    per_core = [r["its"] / r["sec"] for r in rows]
    value = sum(per_core)
    esjd = sum(r["esjd"] for r in rows) / len(rows)
    return {"value": value, "unit": "chain-steps/s", "cores": len(rows), "kind": "reference",
            "per_core": sum(per_core) / len(per_core), "esjd_mean_per_chain": esjd, "esjd_per_sec_per_chain": esjd * sum(per_core) / len(per_core),
            "sample": f"the unmodified reference (baseline/_ref, pip install of glabcmcmc 1.0.1): MCMCRunner.run_{'glmcmc' if sampler == 'glmcmc' else 'global_mcmc'}, "
                      f"{its} iterations x {len(rows)} processes (one chain and one torch thread each), only the run_* call timed "
                      f"(Mixture_hyper.py:32-37); {wall:.1f} s wall incl. interpreter start-up"}


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else "global")
    else:
        print(json.dumps(measure(int(sys.argv[1]) if len(sys.argv) > 1 else 2001)))
