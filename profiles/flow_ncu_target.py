import torch, sys
sys.path.insert(0, '/root/repo')
import glabc_b200
from glabc_b200.flows import RealNVP
torch.manual_seed(0)
f = RealNVP(device="cuda")
with torch.no_grad():
    f.w3.copy_(0.05 * torch.randn_like(f.w3))
f.bind()
eps = torch.randn(1 << 21, 2, device="cuda")
f.fused_sample_from(eps)
torch.cuda.synchronize()
