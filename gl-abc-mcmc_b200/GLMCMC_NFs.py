"""GLMCMC_NF — reference glabcmcmc/GLMCMC_NFs.py:43-186: iSIR whose importance proposal is a 32-block RealNVP trained
online by forward KL on importance-resampled candidates.  Per block of `step_size` global moves: candidates
theta0 ~ flow (GLMCMC_NFs.py:72,127), weights prior * kernel / q (:80-85); each global move resamples among the current
state and the next `batch_size` candidates (:92-110); after the block, at most `Train_step` times, one Adam step on
-mean log q(theta0[resample(w0)]) (:112-124), then a new block from the updated flow.

B200 mapping (many chains, ONE shared flow — the reference's single chain has one flow):
  * flow sample / log_prob for all chains' candidates: tcgen05 tensor-core kernel (csrc/flow.cuh);
  * the chain step against the block: fused kernel `k_ag_step<EXT>` (csrc/step_aglmcmc.cuh) — a chain pauses when its
    block is consumed or when the proposal log-density of a changed state is needed; this host loop then refreshes those
    log-densities in one batched log_prob launch, and at the end of a block trains / rebinds the flow and refills;
  * the <= Train_step Adam steps use torch autograd on the fp32 module (flows.RealNVP) over the pooled candidates.
Parity note: `normflows` cannot be installed here, so this path is unpinned against it."""
import ctypes as C

import torch

from . import _abi
from .engine import RunStats, get_engine
from .flows import RealNVP
from .samplers import _ARITH, _LAYOUT, default_seed, initial_state, print_summary, write_csv


def resample(W, N):
    """Systematic resampling, GLMCMC_NFs.py:29-40: u_i = (U + i) / N, index j repeated once per u_i in
    [Psum[j-1], Psum[j]); u_i beyond the last cumulative weight are dropped (the reference returns fewer than N then)."""
    W = torch.as_tensor(W)
    u = ((torch.rand(1).item() + torch.arange(N)) / N).to(W.device, W.dtype)
    psum = torch.cumsum(W, dim=0)
    idx = torch.searchsorted(psum, u, right=True)      # first j with Psum[j] > u_i
    return idx[idx < W.shape[0]]


def _base_params(base):
    if base is None:
        return None, None
    loc, ls = getattr(base, "loc", None), getattr(base, "log_scale", None)
    if loc is None or ls is None:
        raise NotImplementedError("base must be a DiagGaussian(2) with loc / log_scale (Mixture.py:70)")
    return loc.detach().reshape(1, 2).float(), ls.detach().reshape(1, 2).float()


def GLMCMC_NF(ABCset, num_ite, Initial_theta, Initial_y, Local_Proposal, filelocation, global_frequency, step_size, batch_size,
              base, Train_step, *, num_chains=None, seed=None, chain_id_base=0, arith="fast", trace="chain", return_stats=False,
              verbose=None, device=None, n_blocks=32, train_batch=65536, lr=5e-4, weight_decay=1e-5, flow=None,
              return_flow=False):
    """Same positional signature and return value as the reference for one chain; keyword extensions as in `GlobalMCMC`,
    plus `flow` (continue with a given RealNVP), `train_batch` (pooled resample size) and `return_flow`."""
    if num_ite < 1:
        raise ValueError("num_ite must be at least 1")
    K, S = int(batch_size), int(step_size)
    if not 1 <= K <= _abi.MAX_K or S < 1 or K * S > _abi.AG_MAX_BLOCK:
        raise ValueError(f"batch_size in 1..{_abi.MAX_K} and batch_size * step_size <= {_abi.AG_MAX_BLOCK}")
    eng = get_engine(device)
    pod = eng.bind_model(ABCset)
    if pod.theta_dim != 2:
        raise NotImplementedError("the reference's flow is MLP([1,128,128,2]) couplings: theta_dim must be 2 (GLMCMC_NFs.py:56)")
    eng.bind_proposal(_abi.SLOT_LOCAL, Local_Proposal)
    seed = default_seed() if seed is None else int(seed)
    theta, y, c = initial_state(eng, pod, Initial_theta, Initial_y, num_chains, seed)
    single = num_chains is None and c == 1
    d, B, dev = 2, K * S, eng.device
    if flow is None:
        loc, ls = _base_params(base)
        with torch.random.fork_rng(devices=[]):          # the flow's initial weights: a function of the seed only
            torch.manual_seed(seed & 0x7FFFFFFF)
            flow = RealNVP(n_blocks=n_blocks, base_loc=loc, base_log_scale=ls)
    flow.to(dev)
    opt = torch.optim.Adam(flow.parameters(), lr=lr, weight_decay=weight_decay)     # GLMCMC_NFs.py:63
    flow.bind(eng)
    gen = torch.Generator(device=dev).manual_seed((seed * 0x9E3779B1 + chain_id_base + 0x5F) & 0x7FFFFFFFFFFFFFFF)

    blk_theta = torch.empty(c, B, d, device=dev)
    blk_x, blk_w, blk_lq = torch.empty(c, B, d, device=dev), torch.empty(c, B, device=dev), torch.empty(c, B, device=dev)
    kk = torch.zeros(c, dtype=torch.int32, device=dev)
    pending, lq_valid = torch.zeros_like(kk), torch.zeros_like(kk)
    next_step = torch.ones(c, dtype=torch.int32, device=dev)
    lq_cur = torch.zeros(c, device=dev)
    blk = _abi.BlockIsirPOD(step_size=S, block=B, blk_theta=blk_theta.data_ptr(), blk_x=blk_x.data_ptr(), blk_w=blk_w.data_ptr(),
                            blk_lq=blk_lq.data_ptr(), kk=kk.data_ptr(), pending=pending.data_ptr(), next_step=next_step.data_ptr(),
                            lq_cur=lq_cur.data_ptr(), lq_valid=lq_valid.data_ptr())
    common = dict(theta=theta, y=y, gf=global_frequency, seed=seed, chain_id_base=chain_id_base, K=K, blk=blk)

    def refill(rnd):    # GLMCMC_NFs.py:70-85 / :125-140
        eps = torch.randn(c * B, d, generator=gen, device=dev)
        flow.fused_sample_from(eps, eng, theta=blk_theta, log_q=blk_lq)
        eng.run("block_weights", n_steps=0, step_base=rnd, trace_layout=_abi.TRACE_NONE, **common)

    def refresh_lq(idx=None):   # NF_model.log_prob(Theta_old), GLMCMC_NFs.py:96-98, only where the state changed
        if idx is None:
            lq_cur.copy_(flow.fused_log_prob(theta, eng))
            lq_valid.fill_(1)
        else:
            lq_cur[idx] = flow.fused_log_prob(theta[idx], eng)
            lq_valid[idx] = 1

    refill(0)
    refresh_lq()
    layout = _LAYOUT[trace]
    n_steps = num_ite - 1
    out = None
    if layout != _abi.TRACE_NONE:
        out = torch.empty((num_ite, c, d) if layout == _abi.TRACE_TIME_MAJOR else (c, num_ite, d), device=dev)
    stats = torch.zeros(c, _abi.nstats(d), device=dev)
    num_train, rnd, first, losses = 0, 0, True, []
    while True:
        eng.run("block_isir", n_steps=n_steps, arith=_ARITH[arith], trace_layout=layout, trace=out, trace_rows=num_ite,
                write_row0=first, stats=stats, **common)
        first = False
        need = (pending & 2) != 0
        settled = (next_step > n_steps) | ((pending & 1) != 0)
        n_need, n_done, n_settled = torch.stack([need.sum(), (next_step > n_steps).sum(), settled.sum()]).tolist()
        if n_need:
            refresh_lq(need.nonzero().squeeze(1))
        if n_done == c:
            break
        if n_need == 0 and n_settled == c:      # every chain consumed its block (GLMCMC_NFs.py:111)
            if num_train < Train_step:           # :113-124, pooled over the chains
                opt.zero_grad()
                w = blk_w.reshape(-1)
                idx = resample(w / torch.sum(w), min(c * B, int(train_batch)))
                loss = flow.forward_kld(blk_theta.reshape(-1, d)[idx].detach().float())
                if not (torch.isnan(loss) | torch.isinf(loss)):
                    loss.backward()
                opt.step()                       # runs even when backward was skipped
                num_train += 1
                losses.append(float(loss))
                flow.bind(eng)
            rnd += 1
            refill(rnd)
            refresh_lq()
            kk.zero_()
            pending.zero_()
    rs = RunStats(stats, d)
    if single:
        chain = (out[0] if layout == _abi.TRACE_CHAIN_MAJOR else out[:, 0]).cpu() if out is not None else None
        if filelocation is not None and chain is not None:
            write_csv(filelocation, chain)
        if verbose is not False and chain is not None:
            print_summary(chain)
        result = chain
    else:
        if filelocation is not None and out is not None:
            import numpy as np
            np.save(filelocation if str(filelocation).endswith(".npy") else str(filelocation) + ".npy", out.cpu().numpy())
        result = out
    extra = ((rs,) if return_stats else ()) + ((flow, losses) if return_flow else ())
    return (result,) + extra if extra else result
