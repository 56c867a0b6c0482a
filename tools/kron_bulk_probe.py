"""A/B of the separable kernel's fast passes: direct warp stores vs cp.async.bulk (TMA) stores of rows staged in shared memory."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petiga_b200.cases import Case
dall = lambda dim, v=1.0: [(d, s, 0, v) for d in range(dim) for s in range(2)]
for N in ([int(a) for a in sys.argv[1:]] or [64, 128]):
    ref = None
    for bulk in (0, 1):
        case = Case(3, p=3, N=N, bcv=dall(3))
        g = case.product()
        g.SetOption("kron_bulk", bulk)
        g.SetForm("SYSTEM", "POISSON")
        A, B = g.CreateMat(), g.CreateVec()
        for _ in range(3): g.ComputeSystem(A, B)
        ms = []
        for _ in range(10):
            g.ComputeSystem(A, B); ms.append(g.GetStat("last_kernel_ms"))
        chk = None
        if N <= 64:
            v = A.values(); r = B.get()
            if ref is None: ref = (v, r)
            chk = [float(np.abs(v - ref[0]).max()), float(np.abs(r - ref[1]).max())]
        print(json.dumps({"mesh": N, "bulk": bulk, "kernel_ms_min": min(ms), "kernel_ms_med": sorted(ms)[5], "maxdiff_vs_direct": chk}), flush=True)
        A.destroy(); B.destroy(); g.Destroy()
