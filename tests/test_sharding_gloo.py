"""Host-side multi-rank logic on CPU: shard ranges, the summary all-reduce over gloo (world_size 2),
and that the sharded result equals the single-rank one.  The per-rank compute here is the CPU oracle
standing in for the kernel (tests only; the product path has no CPU route)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import abi, gauss_pod, load_cases, model_pod


def test_shard_ranges_cover_and_balance():
    from glabc_b200.sharding import shard_range
    for c in (0, 1, 7, 64, 65536, 65537, 262144):
        for w in (1, 2, 3, 4, 8):
            rng = [shard_range(c, r, w) for r in range(w)]
            assert rng[0][0] == 0 and rng[-1][1] == c
            assert all(rng[i][1] == rng[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rng]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, C, T, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glabc_b200 import sharding
    from glabc_b200.engine import RunStats
    from oracle import oracle
    case = load_cases("global_mcmc.npz")[0]
    lo, hi = sharding.shard_range(C)
    n = hi - lo
    theta = np.zeros((n, 2), np.float32)
    y = np.full((n, 2), 0.1, np.float32)
    stats = np.zeros((n, abi.nstats(2)), np.float32)
    trace = oracle.run("global", model_pod(case), gauss_pod(case, "lp"), gauss_pod(case, "gp"), theta=theta, y=y,
                       n_steps=T, gf=0.5, seed=5, chain_id_base=lo, stats=stats, threads=1)
    rs = RunStats(torch.from_numpy(stats), 2)
    summary = sharding.allreduce_summary(sharding.summarize(rs))
    np.save(os.path.join(out_dir, f"trace{rank}.npy"), trace)
    np.save(os.path.join(out_dir, f"summary{rank}.npy"), summary.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path):
    C, T = 37, 200
    mp.spawn(_worker, args=(2, _free_port(), C, T, str(tmp_path)), nprocs=2, join=True)
    s0, s1 = np.load(tmp_path / "summary0.npy"), np.load(tmp_path / "summary1.npy")
    assert np.array_equal(s0, s1)                      # every rank holds the reduced summary
    parts = np.concatenate([np.load(tmp_path / "trace0.npy"), np.load(tmp_path / "trace1.npy")], axis=1)
    # single rank, same global chain ids
    from glabc_b200 import sharding
    from glabc_b200.engine import RunStats
    from oracle import oracle
    case = load_cases("global_mcmc.npz")[0]
    theta, y = np.zeros((C, 2), np.float32), np.full((C, 2), 0.1, np.float32)
    stats = np.zeros((C, abi.nstats(2)), np.float32)
    whole = oracle.run("global", model_pod(case), gauss_pod(case, "lp"), gauss_pod(case, "gp"), theta=theta, y=y,
                       n_steps=T, gf=0.5, seed=5, stats=stats, threads=1)
    assert np.array_equal(parts, whole)                # traces do not depend on the world size
    ref = sharding.summarize(RunStats(torch.from_numpy(stats), 2)).numpy()
    assert np.allclose(s0, ref, rtol=1e-12, atol=1e-9)
    d = sharding.describe(torch.from_numpy(s0), 2)
    assert d["chains"] == C and d["chain_steps"] == C * T and 0.0 <= d["move_rate"] <= 1.0
