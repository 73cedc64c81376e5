"""Host logic of the shared-proposal samplers on CPU over gloo (world_size 2): pooled quantile, tolerance update,
training-draw all-gather, gradient averaging, RoundSync — glabc_b200/pooled.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_single_process_quantile_and_draws():
    from glabc_b200 import pooled
    g = torch.Generator().manual_seed(0)
    x = torch.rand(10001, generator=g)
    for q in (0.0, 0.013, 0.5, 0.8, 1.0):
        v = float(pooled.global_quantile(x, q))
        assert v == float(torch.quantile(x, q, interpolation="higher")) or abs(v - float(torch.quantile(x, q))) < 2e-4
    # tolerance rule, AGLMCMC.py:174-199
    dis = torch.rand(5000, generator=g) * 3
    e1 = pooled.update_hat_eps(dis, 1000000.0, 0.8, 0.2)
    assert abs(e1 - float(torch.quantile(dis, 0.8))) < 2e-3
    e2 = pooled.update_hat_eps(dis, e1, 0.8, 0.2)
    assert abs(e2 - float(torch.quantile(dis, 0.8 * float((dis < e1).sum()) / 5000))) < 2e-3
    assert pooled.update_hat_eps(dis, 0.2, 0.8, 0.2) == 0.2 and pooled.update_hat_eps(dis * 0.01, 0.5, 0.8, 0.2) == 0.2
    # systematic draws: counts proportional to the weights, total weight preserved
    th = torch.arange(8, dtype=torch.float32).reshape(8, 1).repeat(1, 2)
    w = torch.tensor([0, 1, 0, 3, 0, 0, 4, 0], dtype=torch.float32)
    X, ww = pooled.gather_training_draws(th, w, 800, 0.37)
    assert X.shape == (800, 2) and abs(float(ww.sum()) - 8.0) < 1e-4
    cnt = torch.bincount(X[:, 0].long(), minlength=8)
    assert cnt.tolist() == [0, 100, 0, 300, 0, 0, 400, 0]
    X0, w0 = pooled.gather_training_draws(th, torch.zeros(8), 16, 0.5)
    assert X0.shape[0] == 0 and w0.shape[0] == 0


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glabc_b200 import pooled
    g = torch.Generator().manual_seed(7)
    full = torch.rand(6001, generator=g) * 2
    mine = full[rank::world].clone()                       # ragged shards (3001 / 3000)
    qs = [float(pooled.global_quantile(mine, q)) for q in (0.05, 0.5, 0.93)]
    eps = pooled.update_hat_eps(mine, 1000000.0, 0.8, 0.2)
    # training draws: rank 0 holds 3x the weight mass of rank 1
    th = torch.arange(10, dtype=torch.float32).reshape(10, 1).repeat(1, 2) + 100 * rank
    w = torch.ones(10) * (3.0 if rank == 0 else 1.0)
    X, ww = pooled.gather_training_draws(th, w, 50, 0.25)
    # gradient averaging of a shared model
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        for p in lin.parameters():
            p.fill_(0.5)
    lin(torch.full((4, 3), float(rank + 1))).sum().backward()
    pooled.average_gradients(list(lin.parameters()))
    # the flat gradient buffer of the native flow training step (glabc_flow_grad) and its loss: mean over the ranks, in place
    flat = torch.arange(7, dtype=torch.float32) * (rank + 1)
    loss = torch.tensor([2.0 + rank])
    pooled.average_flat(flat, loss)
    # RoundSync: rank 1 finishes after 2 rounds, rank 0 after 4 — both must leave the loop at round 4
    sync = pooled.RoundSync("cpu")
    mine_done_at = 4 if rank == 0 else 2
    left_at = None
    for r in range(1, 10):
        if sync.round_end(r >= mine_done_at):
            left_at = r
            break
    np.save(os.path.join(out_dir, f"r{rank}.npy"),
            np.array(qs + [eps, float(X.shape[0]), float(ww.sum()), float((X[:, 0] >= 100).sum()), float(ww[0]), float(ww[-1]),
                           float(lin.weight.grad[0, 0]), float(left_at), float(flat[3]), float(flat.sum()), float(loss)], dtype=np.float64))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1)                          # every rank holds the same pooled result
    g = torch.Generator().manual_seed(7)
    full = torch.rand(6001, generator=g) * 2
    for v, q in zip(r0[:3], (0.05, 0.5, 0.93)):
        assert abs(v - float(torch.quantile(full, q))) < 1e-3
    assert abs(r0[3] - float(torch.quantile(full, 0.8))) < 1e-3
    assert r0[4] == 100 and abs(r0[5] - 40.0) < 1e-4 and r0[6] == 50     # 50 draws per rank, total weight 30 + 10
    assert abs(r0[7] - 0.6) < 1e-6 and abs(r0[8] - 0.2) < 1e-6          # per-draw weight = rank mass / m
    assert abs(r0[9] - 4 * 1.5) < 1e-6                                   # mean of 4*1 and 4*2
    assert r0[10] == 4
    assert abs(r0[11] - 3 * 1.5) < 1e-6 and abs(r0[12] - 21 * 1.5) < 1e-6 and abs(r0[13] - 2.5) < 1e-6   # average_flat
