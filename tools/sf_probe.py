"""Sum-factorised kernel probe at cfg 2 (and the mapped variant): full kernel vs integration only (scatter skipped)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from petiga_b200.cases import Case
dall = lambda dim, v=1.0: [(d, s, 0, v) for d in range(dim) for s in range(2)]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
IMPL = int(sys.argv[2]) if len(sys.argv) > 2 else -1      # -1: the library's choice (quad_sf3 at p = 3), 0: quad_sf, 1: pair loop
for name, geo in (("identity", None), ("mapped", ("perturbed", 0.05))):
    case = Case(3, p=3, N=N, bcv=dall(3), geometry=geo)
    g = case.product()
    g.SetOption("path", 1); g.SetOption("quad_impl", IMPL)
    g.SetForm("SYSTEM", "POISSON")
    A, B = g.CreateMat(), g.CreateVec()
    for sc in (0, 99):
        g.SetOption("scatter", sc)
        for _ in range(2): g.ComputeSystem(A, B)
        ms = 0.0
        for _ in range(3):
            g.ComputeSystem(A, B); ms += g.GetStat("last_kernel_ms") / 3
        print(json.dumps({"geometry": name, "mesh": N, "scatter": "on" if sc == 0 else "skipped", "kernel_ms": ms, "impl": int(g.GetStat("last_impl")), "static": int(g.GetStat("last_sf3_static"))}), flush=True)
    A.destroy(); B.destroy(); g.Destroy()
