"""ncu target: one launch of every flow kernel at the size the bench quotes (2,097,152 samples; training: 65,536)."""
import sys

import torch

sys.path.insert(0, '/root/repo')
import glabc_b200  # noqa: E402,F401
from glabc_b200.flows import RealNVP  # noqa: E402

torch.manual_seed(0)
f = RealNVP(device="cuda")
with torch.no_grad():
    f.w3.copy_(0.05 * torch.randn_like(f.w3))
eng = f.train_init()
eps = torch.randn(1 << 21, 2, device="cuda")
mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
if mode == "train":
    x, _ = f.fused_sample_from(eps[:65536], eng, precision="precise")
    f.grad(x, eng)
else:
    f.fused_sample_from(eps, eng, precision=mode)
torch.cuda.synchronize()
