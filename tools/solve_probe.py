"""Device CG on the cfg-2 system: time per iteration (SpMV + 2 dot products + 2 vector updates)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from petiga_b200.cases import Case
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = Case(3, p=3, N=N, bcv=[(d, s, 0, 1.0) for d in range(3) for s in range(2)]).product()
g.SetForm("SYSTEM", "POISSON")
A, B, X = g.CreateMat(), g.CreateVec(), g.CreateVec()
g.ComputeSystem(A, B)
g.Solve(A, B, X, rtol=1e-30, maxits=20)
torch.cuda.synchronize(); t0 = time.perf_counter()
its, rel = g.Solve(A, B, X, rtol=1e-30, maxits=100)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({"mesh": N, "iterations": its, "ms_per_iteration": dt / its * 1e3, "spmv_GBps": 12.0 * A.nnz / (dt / its) / 1e9}))
