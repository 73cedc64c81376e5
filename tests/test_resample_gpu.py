"""`resample` (systematic resampling, GLMCMC_NFs.py:29-40 / AGLMCMC.py:30-41) on the device — csrc/resample.cu through
glabc_resample — against the reference's own recorded index lists (tests/golden/resample.npz) and, at pooled size, against
a float64 numpy restatement."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def test_resample_golden_on_device():
    """the reference's exact index lists, including the short return when the cumulative sum tops out below 1"""
    from glabc_b200.GLMCMC_NFs import resample
    z = np.load(os.path.join(GOLDEN, "resample.npz"))
    for i in range(int(z["n_cases"])):
        torch.manual_seed(int(z[f"rs{i}/seed"]))           # resample draws its one uniform with torch.rand(1), as the reference does
        got = resample(torch.from_numpy(z[f"rs{i}/W"]).cuda(), int(z[f"rs{i}/N"]))
        assert got.is_cuda and got.dtype == torch.int64
        assert np.array_equal(got.cpu().numpy(), z[f"rs{i}/idx"]), i
    assert len(z["rs2/idx"]) < int(z["rs2/N"])


@pytest.mark.parametrize("n,N", [(1, 5), (4095, 100), (4097, 7), (3_000_000, 65536), (40_000_000, 65536)])
def test_resample_matches_float64_restatement(n, N):
    """any length (several 4096-weight blocks, ragged tails): idx[i] = first j with float32(cumsum64(W)[j]) > float32 u_i"""
    from glabc_b200.GLMCMC_NFs import resample
    g = torch.Generator(device="cuda").manual_seed(n % 1000 + N)
    W = torch.rand(n, device="cuda", generator=g) ** 4
    W = W / W.sum()
    torch.manual_seed(n + N)
    got = resample(W, N).cpu().numpy()
    torch.manual_seed(n + N)
    u0 = np.float32(torch.rand(1).item())
    u = ((u0 + np.arange(N, dtype=np.float32)).astype(np.float32) / np.float32(N)).astype(np.float32)
    psum = np.cumsum(W.cpu().numpy().astype(np.float64)).astype(np.float32)
    want = np.searchsorted(psum, u, side="right")
    want = want[want < n]
    assert got.shape == want.shape
    # the parallel float64 sum differs from numpy's sequential one in the last ulp of float64: an index may move by one only
    # where two float32 prefixes tie to within that — never on these sizes, but allowed for
    assert (got == want).mean() > 0.9999 and np.abs(got - want).max() <= 1
    assert np.all(np.diff(got) >= 0)


def test_resample_empty_and_degenerate():
    from glabc_b200.GLMCMC_NFs import resample
    assert resample(torch.zeros(10, device="cuda"), 8).numel() == 0            # no mass: nothing is emitted
    one = resample(torch.tensor([0.0, 1.0, 0.0], device="cuda"), 6)
    assert one.tolist() == [1] * 6
    assert resample(torch.ones(4, device="cuda") / 4, 0).numel() == 0
