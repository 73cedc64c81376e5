"""esjd() — reference ESJD.py:2-25: det(D^T D / (N-1))^(1/d) of a chain's successive differences."""
import numpy as np
import torch

from . import _abi


def esjd(data):
    """`data` [N, d] -> 0-d numpy float32 array (the reference's return type); `data` [C, N, d] ->
    numpy array of C values.  Runs in the CUDA kernel behind `glabc_esjd`."""
    from .engine import get_engine
    eng = get_engine()
    t = torch.as_tensor(data, dtype=torch.float32)
    single = t.dim() == 2
    if single:
        t = t[None]
    t = t.to(eng.device).contiguous()
    out = eng.esjd(t, _abi.TRACE_CHAIN_MAJOR).cpu().numpy()
    return np.asarray(out[0]) if single else out
