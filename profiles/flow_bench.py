import torch, time, sys
sys.path.insert(0, '/root/repo')
import glabc_b200
from glabc_b200.flows import RealNVP
torch.manual_seed(0)
f = RealNVP(device="cuda")
with torch.no_grad():
    f.w3.copy_(0.05 * torch.randn_like(f.w3))
f.bind()
for n in (1 << 20, 1 << 23):
    eps = torch.randn(n, 2, device="cuda")
    for _ in range(2): f.fused_sample_from(eps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): th, lq = f.fused_sample_from(eps)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"sample n={n}: {ms:.3f} ms  {n/ms*1e3:.3e} samples/s  {n*1.073e6/ms*1e3/1e12:.1f} TFLOP/s (dense 1.049 MFLOP/sample: {n*1.049e6/ms*1e3/1e12:.1f})")
    e0.record()
    for _ in range(5): lp = f.fused_log_prob(th)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"log_prob n={n}: {ms:.3f} ms  {n/ms*1e3:.3e} samples/s")
    with torch.no_grad():
        t0=time.perf_counter(); th_r, lq_r = f.sample_from(eps[:1<<20]); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(f"torch fp32 eager sample 1M: {dt*1e3:.1f} ms")
