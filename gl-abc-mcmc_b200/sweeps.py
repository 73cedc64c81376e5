"""ESJD-per-second hyper-parameter sweep — the tuning loop of reference examples/Mixture_hyper.py:12-41: for every
`global_frequency` in a grid and every seed, run a sampler for `num_ite` iterations, score esjd(chain) / seconds-per-iteration,
report the best grid point.  Here every grid point runs `num_chains` chains at once on the GPU (the seeds of the reference's
inner loop become chains), ESJD comes from the in-kernel Gram accumulators, and with several GPUs the grid points are dealt
round-robin to the ranks and the score table is all-reduced (the only collective)."""
import time

import torch

from .engine import get_engine


def esjd_sweep(run, grid=None, num_chains=1024, seed=0, rank=0, world=1, device=None):
    """`run(global_frequency, num_chains=..., seed=..., return_stats=True, trace="none")` -> (_, RunStats) — e.g.
    `lambda gf, **kw: runner.run_glmcmc(1000, theta0, None, gf, lp, ip, 5, output_file=None, **kw)`.
    Returns (best_gf, table) with table[i] = (gf, mean esjd, chain-steps/s, esjd * chain-steps/s) — Mixture_hyper.py:36-40."""
    grid = [i / 10 for i in range(11)] if grid is None else list(grid)          # Mixture_hyper.py:23
    eng = get_engine(device)
    table = torch.zeros(len(grid), 4, dtype=torch.float64, device=eng.device)
    for i, gf in enumerate(grid):
        if i % world != rank:
            continue
        run(gf, num_chains=num_chains, seed=seed, return_stats=True, trace="none")   # warm-up (context, JIT-free but cold caches)
        torch.cuda.synchronize(eng.device)
        t0 = time.perf_counter()
        _, st = run(gf, num_chains=num_chains, seed=seed + 1, return_stats=True, trace="none")
        torch.cuda.synchronize(eng.device)
        dt = time.perf_counter() - t0
        steps = float(st.steps.sum())
        e = float(st.esjd().mean())
        table[i] = torch.tensor([gf, e, steps / dt, e * steps / dt], dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(table)
    best = int(torch.argmax(table[:, 3]))
    return float(table[best, 0]), table.cpu()
