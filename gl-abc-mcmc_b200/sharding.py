"""Multi-GPU layout: chains are independent units, so ranks own contiguous chain ranges and no
collective sits on the data path.  Philox streams are keyed by the GLOBAL chain
id, so a chain's trace is identical for any world size.  The only collective is the end-of-run
all-reduce of the summary statistics."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_chains, rank=None, world_size=None):
    """[begin, end) of the chains rank `rank` owns: floor split, remainder to the low ranks."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n_chains), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_summary(local_sums):
    """Sum a small float64 tensor of additive run statistics (steps, accepts, sum theta, sum theta^2,
    sum of per-chain ESJD, ...) over all ranks.  NCCL for CUDA tensors, gloo for CPU tensors."""
    t = local_sums.to(torch.float64).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def summarize(stats, esjd_per_chain=None):
    """Additive summary vector of a shard's RunStats: [chains, steps, global_steps, acc_local,
    acc_global, sum esjd, sum theta (d), sum theta^2 (d)] — ready for `allreduce_summary`."""
    d = stats.dim
    if stats.raw.is_cuda and esjd_per_chain is None and stats.raw.is_contiguous() and d <= 4:
        from .engine import get_engine              # one fused kernel instead of ~25 small torch launches
        return get_engine(stats.raw.device).summarize(stats.raw, d)
    raw = stats.raw.double()
    e = stats.esjd() if esjd_per_chain is None else esjd_per_chain.double()
    head = torch.stack([torch.tensor(float(raw.shape[0]), dtype=torch.float64, device=raw.device),
                        raw[:, 0].sum(), raw[:, 1].sum(), raw[:, 2].sum(), raw[:, 3].sum(), e.sum()])
    return torch.cat([head, raw[:, 4:4 + 2 * d].sum(0)])


def describe(summary, d):
    """Human-readable dict from a (reduced) summary vector."""
    chains, steps, gsteps, accl, accg, esjd = [float(x) for x in summary[:6]]
    s1 = summary[6:6 + d] / max(steps, 1.0)
    s2 = summary[6 + d:6 + 2 * d] / max(steps, 1.0)
    return dict(chains=int(chains), chain_steps=steps, global_fraction=gsteps / max(steps, 1.0),
                move_rate=(accl + accg) / max(steps, 1.0), mean_esjd=esjd / max(chains, 1.0),
                mean=s1.tolist(), var=(s2 - s1 * s1).tolist())
