"""GlobalMCMC — reference glabcmcmc/GlobalMCMC.py:6-98: each iteration is, with probability
`global_frequency`, an independence Metropolis–Hastings move from `Global_Proposal`, otherwise a
random-walk move from `Local_Proposal`; the ABC likelihood is a Gaussian kernel on the discrepancy
of one simulator draw.  The loop body (GlobalMCMC.py:37-68) runs in the fused kernel
`k_global_mcmc` (csrc/step_global.cu) for all chains at once."""
from . import _abi
from .engine import get_engine
from .samplers import run_chains


def GlobalMCMC(ABCset, num_ite, Initial_theta, Initial_y, Global_Proposal, filelocation, global_frequency,
               Local_Proposal=None, *, num_chains=None, seed=None, chain_id_base=0, arith="fast", trace="chain",
               return_stats=False, verbose=None, device=None, block_threads=0):
    """Same positional signature and return value as the reference (a float32 CPU tensor
    `[num_ite, theta_dim]`, row 0 = `Initial_theta`) when called for one chain.

    Keyword extensions: `num_chains=C` runs C independent chains and returns a device tensor
    `[C, num_ite, d]` (`trace="chain"`), `[num_ite, C, d]` (`trace="time"`) or None
    (`trace="none"`, statistics only); `Initial_theta` / `Initial_y` may then be one row (shared) or
    C rows, and `Initial_y=None` draws y0 from the simulator per chain.  `seed` defaults to
    `torch.initial_seed()`; chain c draws from the Philox stream of global id `chain_id_base + c`.
    `arith="strict"` evaluates in the reference's float32 operation order.  `return_stats=True`
    also returns the in-kernel `RunStats` (accept counts, moments, ESJD numerators)."""
    if Local_Proposal is None:
        raise ValueError("Local_Proposal is required (the reference dereferences it on every local move, GlobalMCMC.py:56)")
    eng = get_engine(device)
    pod = eng.bind_model(ABCset)
    eng.bind_proposal(_abi.SLOT_LOCAL, Local_Proposal)
    eng.bind_proposal(_abi.SLOT_GLOBAL, Global_Proposal)
    return run_chains("global", eng, pod, num_ite=num_ite, Initial_theta=Initial_theta, Initial_y=Initial_y,
                      global_frequency=global_frequency, filelocation=filelocation, num_chains=num_chains, seed=seed,
                      chain_id_base=chain_id_base, arith=arith, trace=trace, return_stats=return_stats, verbose=verbose,
                      block_threads=block_threads)
