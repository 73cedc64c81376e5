"""AGLMCMC — reference glabcmcmc/AGLMCMC.py:44-288: iSIR against a block of batch_size * step_size pre-generated
importance candidates; after step_size global moves the auxiliary tolerance eps-hat shrinks by a quantile rule, a
weighted KernelDensity is refit on the block and the next block is sampled from it.  The loop body and the adaptation
run in the kernels of csrc/step_aglmcmc.cuh + csrc/kde.cuh for all chains at once (every chain owns its block and KDE,
as the reference's single chain does).

Deviations from the reference, on purpose: the chain is returned (the reference returns None and
crashes past 10,000 iterations) and row 0 holds the initial theta."""
from . import _abi
from .engine import get_engine
from .samplers import run_chains


def AGLMCMC(ABCset, num_ite, Initial_theta, Initial_y, Local_Proposal, Initial_ISIR_prop, filelocation, global_frequency,
            step_size, batch_size, alpha, hat_eps_T, device=None, *, num_chains=None, seed=None, chain_id_base=0, arith="fast",
            trace="chain", return_stats=False, verbose=None, block_threads=0):
    """Same positional signature as the reference; keyword extensions as in `GlobalMCMC`."""
    if not 1 <= int(batch_size) <= _abi.MAX_K:
        raise ValueError(f"batch_size must be in 1..{_abi.MAX_K}")
    if int(step_size) < 1 or int(step_size) * int(batch_size) > _abi.AG_MAX_BLOCK:
        raise ValueError(f"batch_size * step_size must be in 1..{_abi.AG_MAX_BLOCK}")
    eng = get_engine(device)
    pod = eng.bind_model(ABCset)
    eng.bind_proposal(_abi.SLOT_LOCAL, Local_Proposal)
    eng.bind_proposal(_abi.SLOT_IMPORTANCE, Initial_ISIR_prop)
    ag = eng.aglmcmc_params(step_size=step_size, alpha=alpha, hat_eps_T=hat_eps_T)
    return run_chains("aglmcmc", eng, pod, num_ite=num_ite, Initial_theta=Initial_theta, Initial_y=Initial_y,
                      global_frequency=global_frequency, filelocation=filelocation, num_chains=num_chains, seed=seed,
                      chain_id_base=chain_id_base, arith=arith, trace=trace, return_stats=return_stats, verbose=verbose,
                      K=int(batch_size), block_threads=block_threads, ag=ag)
