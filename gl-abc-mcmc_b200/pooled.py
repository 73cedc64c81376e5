"""Cross-chain / cross-GPU pieces of the samplers whose importance proposal is SHARED by all chains:

  * AGLMCMC with one pooled KernelDensity (BASELINE config 5: a KDE over 1e5 accepted draws): the tolerance rule of
    AGLMCMC.py:174-199 on the pooled discrepancies (`global_quantile`, counts all-reduced), and the all-gather of every
    rank's weighted training draws before `fit` (`gather_training_draws`: fixed count per rank, so no length header);
  * GLMCMC-NFs with one shared flow: `average_gradients` (all-reduce of the 548,932 gradient floats per Adam step,
    GLMCMC_NFs.py:112-124) and `RoundSync`, which keeps ranks that finished early in the collective sequence.

Everything here is host logic over torch tensors + torch.distributed: NCCL for CUDA tensors, gloo for the CPU tests
(tests/test_pooled_gloo.py).  No sampler arithmetic lives here — candidates, weights, KDE and flow evaluations are the
kernels behind the C-ABI."""
import torch
import torch.distributed as dist


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_(t, op=None):
    """in-place all-reduce (SUM by default) when a process group with more than one rank is up; returns t"""
    if _world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op is None else op)
    return t


def global_count(mask):
    """number of True entries over all ranks (float64 scalar tensor on mask's device)"""
    return allreduce_(mask.sum().to(torch.float64).reshape(1))[0]


def global_quantile(x, q, iters=48):
    """q-quantile of the union of every rank's `x` without moving the data: bisection on the value, one all-reduce of a
    count per iteration.  Returns the smallest float32 v (to bisection resolution) with #{x <= v} >= q * n — the 'higher'
    interpolation of torch.quantile; with the millions of pooled discrepancies this replaces, the interpolation rule of
    AGLMCMC.py:193 (linear) differs by less than one order statistic."""
    x = x.reshape(-1)
    n = allreduce_(torch.tensor([float(x.numel())], dtype=torch.float64, device=x.device))[0]
    if n == 0:
        raise ValueError("global_quantile of an empty set")
    big = torch.finfo(torch.float64).max
    lo = torch.tensor([float(x.min()) if x.numel() else big], dtype=torch.float64, device=x.device)
    hi = torch.tensor([float(x.max()) if x.numel() else -big], dtype=torch.float64, device=x.device)
    allreduce_(lo, dist.ReduceOp.MIN)
    allreduce_(hi, dist.ReduceOp.MAX)
    lo, hi = lo[0], hi[0]
    need = torch.clamp(torch.ceil(torch.as_tensor(float(q), dtype=torch.float64, device=x.device) * n), min=1.0)
    xd = x.double()
    for _ in range(iters):          # invariant: count(x <= hi) >= need, count(x <= lo') < need for lo' < lo
        mid = 0.5 * (lo + hi)
        ok = global_count(xd <= mid) >= need          # tensor predicate: no host synchronisation per iteration
        hi, lo = torch.where(ok, mid, hi), torch.where(ok, lo, mid)
    return hi.to(torch.float32)


def update_hat_eps(dis, hat_eps, alpha, hat_eps_T):
    """AGLMCMC.py:174-199 on the pooled block: q = alpha * #{dis < hat_eps} / n, hat_eps <- max(quantile_q(dis), hat_eps_T)."""
    if hat_eps <= hat_eps_T:
        return float(hat_eps)
    valid = dis[~torch.isnan(dis)]
    n = global_count(torch.ones_like(valid, dtype=torch.bool))
    if n > 0:
        num_a = global_count(dis < hat_eps)
        q = float(torch.clamp(alpha * num_a / n, 0.0, 1.0))
        hat_eps = float(global_quantile(valid, q))
    return max(float(hat_eps), float(hat_eps_T))


def systematic_indices(w, m, u):
    """m systematic-resampling indices from unnormalised weights w (GLMCMC_NFs.py:29-40 with the cumulative sums
    normalised, so exactly m indices come back); u in [0, 1) is the single uniform."""
    cs = torch.cumsum(w.double(), 0)
    pos = (u + torch.arange(m, device=w.device, dtype=torch.float64)) / m * cs[-1]
    return torch.searchsorted(cs, pos, right=True).clamp_(max=w.numel() - 1)


def gather_training_draws(theta, w, m, u):
    """This rank's candidates theta [n, d] with unnormalised training weights w [n] -> the pooled training set of
    world * m draws: m systematic draws per rank, each carrying weight (rank's total weight) / m, all-gathered.
    A rank whose weights are all zero contributes m zero-weight rows (dropped by the caller)."""
    w = torch.nan_to_num(w.reshape(-1).double(), nan=0.0, posinf=0.0).clamp_(min=0.0)
    total = w.sum()
    d = theta.shape[-1]
    if total > 0:
        idx = systematic_indices(w, m, u)
        loc = torch.cat([theta.reshape(-1, d)[idx].float(), (total / m).float().expand(m, 1)], dim=1).contiguous()
    else:
        loc = torch.zeros(m, d + 1, dtype=torch.float32, device=theta.device)
    _, world = _world()
    if world > 1:
        out = torch.empty(world * m, d + 1, dtype=torch.float32, device=theta.device)
        dist.all_gather_into_tensor(out, loc) if loc.is_cuda else dist.all_gather(list(out.chunk(world)), loc)
    else:
        out = loc
    keep = out[:, d] > 0
    return out[keep, :d].contiguous(), out[keep, d].contiguous()


def average_gradients(params):
    """all-reduce (mean) of the gradients of a shared model: one flat buffer, one collective"""
    _, world = _world()
    if world == 1:
        return
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat /= world
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = flat[off:off + n].view_as(p).clone()
        off += n


def average_flat(grad, loss=None):
    """mean over the ranks of a flat gradient buffer (glabc_flow_grad's output) and of the loss: one collective each"""
    _, world = _world()
    if world == 1:
        return
    dist.all_reduce(grad)
    grad /= world
    if loss is not None:
        dist.all_reduce(loss)
        loss /= world


class RoundSync:
    """Keeps the collective sequence of a shared-proposal sampler aligned across ranks.  Every rank calls
    `round_end(local_done)` each time ALL of its chains have consumed their block (or finished); the call returns True
    when every rank has finished.  A rank that finishes early keeps calling it (its step kernel is a no-op by then), so
    the per-round collectives (gradient / training-draw exchange) always have every participant."""

    def __init__(self, device):
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.rounds = 0

    def round_end(self, local_done):
        self.rounds += 1
        self.flag.fill_(1 if local_done else 0)
        allreduce_(self.flag, dist.ReduceOp.MIN)
        return bool(self.flag.item())
