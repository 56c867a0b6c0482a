"""Multi-rank parity against committed golden vectors (no oracle code at run time).

`tests/golden/multirank_cases.npz` holds the CPU oracle's assembled systems of a few small cases in NATURAL node numbering
(generator: tests/golden/make_multirank_golden.py).  `check_cases` assembles the same cases on the live communicator -- one
rank per GPU, ghost rows and state halo over NCCL -- and compares every rank's owned rows entry by entry, whatever the
rank count (the PETSc global numbering depends on it; the natural numbering does not).  bench.py reports the result in its
JSON line ("parity"), so that the driver's 1/2/4/8-GPU scaling records carry it."""
import os

import numpy as np

from .cases import Case, state_vectors

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "multirank_cases.npz")


def _dall(dim, v=1.0):
    return [(d, s, 0, v) for d in range(dim) for s in range(2)]


def golden_cases():
    """name -> (Case, slot, form, params, needs_state, shift).  Shared by the generator and the checker."""
    return {
        "poisson3d_p3": (Case(3, p=3, N=8, bcv=_dall(3)), "SYSTEM", "POISSON", [], False, 0.0),
        "elasticity3d_baij": (Case(3, dof=3, p=2, N=4, bcv=[(0, 0, 0, 0.0), (0, 0, 1, 0.0), (0, 0, 2, 0.0), (0, 1, 0, 1.0)]),
                              "SYSTEM", "ELASTICITY3D", [1.0, 1.0], False, 0.0),
        "cahnhilliard2d_ijac": (Case(2, p=2, N=32, C=1, periodic=True), "IJACOBIAN", "CAHNHILLIARD2D", [1.5, 3000.0], True, 1.0e3),
        "mapped_poisson3d": (Case(3, p=2, N=6, geometry=("perturbed", 0.05), bcv=_dall(3)), "SYSTEM", "POISSON", [], False, 0.0),
    }


def _run(g, case, slot, form, params, U, V, shift):
    g.SetForm(slot, form, params)
    A = g.CreateMat() if slot in ("MATRIX", "SYSTEM", "JACOBIAN", "IJACOBIAN") else None
    B = g.CreateVec() if slot in ("VECTOR", "SYSTEM", "FUNCTION", "IFUNCTION") else None
    vU = vV = None
    if U is not None:
        vU = g.CreateVec(); vU.set(U)
    if V is not None:
        vV = g.CreateVec(); vV.set(V)
    if slot == "SYSTEM": g.ComputeSystem(A, B)
    elif slot == "IJACOBIAN": g.ComputeIJacobian(shift, vV, 0.0, vU, A)
    elif slot == "IFUNCTION": g.ComputeIFunction(shift, vV, 0.0, vU, B)
    else: raise ValueError(slot)
    out = {}
    if A is not None:
        out["rowptr"], out["colidx"] = A.pattern()
        out["values"] = A.values().reshape(len(out["colidx"]), -1)
    if B is not None:
        out["rhs"] = B.get()
    for x in (A, B, vU, vV):
        if x is not None:
            x.destroy()
    return out


def check_cases(rank, world, nccl, device, allgather, allreduce_sum, paths=("auto", "quadrature"), names=None):
    """Returns {"ranks", "cases", "max_relerr", "detail"}.  allgather(np.int32 array) -> list of arrays (rank order);
    allreduce_sum(np.float64 array) -> summed array.  With world == 1 both may be identity-like lambdas."""
    gold = np.load(GOLDEN)
    detail, worst, ncases = {}, 0.0, 0
    for name, (case, slot, form, params, state, shift) in golden_cases().items():
        if names and name not in names:
            continue
        rp, ci = gold[name + "/rowptr"], gold[name + "/colidx"]
        K = gold[name + "/K"] if (name + "/K") in gold.files else None
        F = gold[name + "/F"] if (name + "/F") in gold.files else None
        n, dof = len(rp) - 1, case.dof
        for path in paths:
            if state and path == "auto":
                continue                       # state-dependent forms always take the quadrature path
            g = case.product(rank=rank, size=world, nccl=nccl, device=device)
            g.SetOption("path", {"auto": 0, "quadrature": 1}[path])
            nat_own = g.GetOwnedNaturalIndices().astype(np.int64)
            perm = np.concatenate([np.asarray(a, dtype=np.int64) for a in allgather(nat_own.astype(np.int32))])   # global id -> natural id
            U = V = None
            if state:
                Un, Vn = state_vectors(n * dof)
                U, V = Un.reshape(n, dof)[nat_own].reshape(-1), Vn.reshape(n, dof)[nat_own].reshape(-1)
            res = _run(g, case, slot, form, params, U, V, shift)
            sums = np.zeros(4)
            ok_pattern = 1.0
            if K is not None:
                lrp, lci = res["rowptr"].astype(np.int64), res["colidx"].astype(np.int64)
                rows_nat = np.repeat(nat_own, np.diff(lrp))
                cols_nat = perm[lci]
                order = np.lexsort((cols_nat, rows_nat))
                srows = np.sort(nat_own)
                gidx = np.concatenate([np.arange(rp[r], rp[r + 1]) for r in srows]) if len(srows) else np.zeros(0, dtype=np.int64)
                if len(gidx) != len(order) or not np.array_equal(ci[gidx], cols_nat[order]):
                    ok_pattern = 0.0
                else:
                    d = res["values"][order] - K[gidx].reshape(len(gidx), -1)
                    sums[0], sums[1] = float((d * d).sum()), float((K[gidx] ** 2).sum())
            if F is not None:
                d = res["rhs"].reshape(-1, dof) - F[nat_own]
                sums[2], sums[3] = float((d * d).sum()), float((F[nat_own] ** 2).sum())
            tot = allreduce_sum(np.concatenate([sums, [1.0 - ok_pattern]]))
            eK = float(np.sqrt(tot[0] / tot[1])) if tot[1] > 0 else 0.0
            eF = float(np.sqrt(tot[2] / tot[3])) if tot[3] > 0 else 0.0
            bad = tot[4] > 0
            err = float("inf") if bad else max(eK, eF)
            detail["%s/%s" % (name, path)] = {"K": eK, "F": eF, "pattern_ok": not bad, "path_used": int(g.GetStat("last_path"))}
            worst = max(worst, err)
            ncases += 1
            g.Destroy()
    return {"ranks": world, "cases": ncases, "max_relerr": worst, "tol": 1e-12, "pass": bool(worst <= 1e-12),
            "against": "tests/golden/multirank_cases.npz (CPU oracle outputs in natural numbering)", "detail": detail}
