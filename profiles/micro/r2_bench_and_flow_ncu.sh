o=gpurun_out
python bench.py > $o/r2_bench.json 2> $o/r2_bench.err; tail -2 $o/r2_bench.err
N="ncu --set full --clock-control none --import-source on -c 1 -f"
$N -k k_flow -o $o/r2_k4_flow_fast python profiles/flow_ncu_target.py fast > $o/r2_ncu_k4a.log 2>&1
$N -k k_flow -o $o/r2_k4_flow_precise python profiles/flow_ncu_target.py precise > $o/r2_ncu_k4b.log 2>&1
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import glabc_b200
from glabc_b200.flows import RealNVP
torch.manual_seed(0)
f = RealNVP(device="cuda")
with torch.no_grad():
    f.w3.copy_(0.05 * torch.randn_like(f.w3))
eng = f.train_init()
x, _ = f.fused_sample_from(torch.randn(65536, 2, device="cuda"), eng, precision="precise")
ref = RealNVP(device="cuda"); ref.load_state_dict(f.state_dict())
opt = torch.optim.Adam(ref.parameters(), lr=5e-4, weight_decay=1e-5)
def native():
    g, l = f.grad(x, eng); f.adam_step(g, l, eng, sync_module=False)
def torch_step():
    opt.zero_grad(); l = ref.forward_kld(x); l.backward(); opt.step()
for name, fn in (("native", native), ("torch", torch_step)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print("train step 65536 samples", name, e0.elapsed_time(e1) / 10, "ms")
PY
