import sys, time, torch
sys.path.insert(0, '/root/repo')
import glabc_b200 as g
from glabc_b200 import block_isir, GLMCMC_NFs
model = g.Mixture_set(0.05)
lp = g.DiagGaussian(2, torch.zeros(1, 2), torch.log(torch.tensor([0.35, 0.35])))
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0; T[name + "#"] = T.get(name + "#", 0) + 1
        return r
    return w
FP = GLMCMC_NFs.FlowProposal
FP.fill = timed("fill", FP.fill); FP.log_prob = timed("log_prob", FP.log_prob); FP.adapt = timed("adapt", FP.adapt)
from glabc_b200.engine import Engine
Engine.run = timed("engine.run", Engine.run)
for rep in range(2):
    T.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g.GLMCMC_NF(model, 1001, torch.zeros(2), None, lp, None, 0.5, 200, 5, None, 50, num_chains=131072, seed=rep, trace="none", return_stats=True, verbose=False)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print("total %.3f s" % tot, {k: (round(v, 3) if not k.endswith('#') else v) for k, v in T.items()})
