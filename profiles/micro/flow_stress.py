"""Stress of the flow pipeline kernels: random batch sizes (1 .. 1.3 M, incl. many chunks per CTA and ragged tails), both
precisions, sample / native sample / log_prob; every launch repeated and compared bit for bit (the kernels are deterministic),
outputs finite, FAST against PRECISE within the FAST tolerance.  A protocol bug in the mbarrier pipeline shows up as a hang
(run under `timeout`) or as a mismatch."""
import sys
import time

import torch

sys.path.insert(0, '/root/repo')
import glabc_b200  # noqa: E402,F401
from glabc_b200.flows import RealNVP  # noqa: E402

torch.manual_seed(1)
f = RealNVP(device="cuda")
with torch.no_grad():
    f.w3.copy_(0.05 * torch.randn_like(f.w3))
eng = f.train_init()
g = torch.Generator().manual_seed(5)
t0 = time.time()
n_launch = 0
sizes = [1, 2, 127, 128, 129, 255, 256, 511, 512, 513, 4096, 148 * 128, 148 * 128 + 1, 148 * 256 - 1, 606208, 606209, 1212416 + 77]
sizes += [int(torch.randint(1, 1_300_000, (1,), generator=g)) for _ in range(120)]
sizes += [int(torch.randint(1, 3000, (1,), generator=g)) for _ in range(150)]
for it, n in enumerate(sizes):
    eps = torch.randn(n, 2, device="cuda")
    out = {}
    for mode in ("fast", "precise"):
        th1, lq1 = f.fused_sample_from(eps, eng, precision=mode)
        th2, lq2 = f.fused_sample_from(eps, eng, precision=mode)
        assert torch.equal(th1, th2) and torch.equal(lq1, lq2), (mode, n, "sample not deterministic")
        assert torch.isfinite(th1).all() and torch.isfinite(lq1).all(), (mode, n)
        lp1 = f.fused_log_prob(th1, eng, precision=mode)
        lp2 = f.fused_log_prob(th1, eng, precision=mode)
        assert torch.equal(lp1, lp2), (mode, n, "log_prob not deterministic")
        tn1, ln1 = f.fused_sample(n, seed=it, eng=eng, precision=mode)
        tn2, ln2 = f.fused_sample(n, seed=it, eng=eng, precision=mode)
        assert torch.equal(tn1, tn2) and torch.equal(ln1, ln2), (mode, n, "native sample not deterministic")
        out[mode] = (th1, lq1, lp1)
        n_launch += 6
    e = ((out["fast"][0] - out["precise"][0]).abs() / (1 + out["precise"][0].abs())).max(1).values
    assert float(e.median()) < 3e-3, (n, float(e.median()))
    d = (out["precise"][2] - out["precise"][1]).abs()     # PRECISE: log_prob(sample) reproduces the sample's log q
    assert float(d.median()) < 1e-4, (n, float(d.median()))
torch.cuda.synchronize()
print(f"flow stress ok: {len(sizes)} sizes, {n_launch} launches, {time.time() - t0:.1f} s")
