"""Scalar separable kernel: 3 CTAs of 80 registers vs 4 CTAs of 64 registers per SM, by pencil length (option kron_minb_rows)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petiga_b200.cases import Case
dall = lambda dim, v=1.0: [(d, s, 0, v) for d in range(dim) for s in range(2)]
for N in (48, 64, 80, 96, 112, 128):
    out = {"mesh": N}
    for name, thr in (("cta4_64regs", 100000), ("cta3_80regs", 0)):
        g = Case(3, p=3, N=N, bcv=dall(3)).product()
        g.SetOption("kron_minb_rows", thr)
        g.SetForm("SYSTEM", "POISSON")
        A, B = g.CreateMat(), g.CreateVec()
        for _ in range(3): g.ComputeSystem(A, B)
        ms = []
        for _ in range(20):
            g.ComputeSystem(A, B); ms.append(g.GetStat("last_kernel_ms"))
        out[name] = sorted(ms)[10]
        A.destroy(); B.destroy(); g.Destroy()
    print(json.dumps(out), flush=True)
