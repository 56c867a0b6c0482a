"""Multi-rank parity check, run under torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multirank_check.py
Every rank assembles its part on its GPU (ghost rows exchanged over NCCL, state halo over NCCL) and compares its
owned rows with the oracle's global assembly emulating the same number of ranks."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import petiga_b200 as pb
    from tests.common import Case, oracle_to_layout, rel_frobenius, state_vectors
    from tests.gpu_common import run_product, MAT_SLOTS, VEC_SLOTS

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = pb.load_cuda()
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        assert L.petiga_cuda_comm_unique_id(raw) == 0
        idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
    comm = C.c_void_p()
    assert L.petiga_cuda_comm_init(C.byref(comm), world, rank, raw, local) == 0, L.petiga_cuda_last_error()

    dall = lambda dim, v=1.0: [(d, s, 0, v) for d in range(dim) for s in range(2)]
    cases = [
        ("poisson3d p2", Case(3, p=2, N=8, bcv=dall(3)), "SYSTEM", "POISSON", [], False),
        ("poisson3d p3", Case(3, p=3, N=(9, 8, 10), bcv=dall(3)), "SYSTEM", "POISSON", [], False),
        ("poisson3d p3 big", Case(3, p=3, N=(24, 20, 22), bcv=dall(3)), "SYSTEM", "POISSON", [], False),
        ("poisson2d p2", Case(2, p=2, N=(16, 12), bcv=dall(2, 0.5)), "SYSTEM", "POISSON", [], False),
        ("elasticity3d", Case(3, dof=3, p=2, N=6, bcv=[(0, 0, 0, 0.0), (0, 0, 1, 0.0), (0, 0, 2, 0.0), (0, 1, 0, 1.0)]), "SYSTEM", "ELASTICITY3D", [1.0, 1.0], False),
        ("elasticity3d aij", Case(3, dof=3, p=2, N=6, mattype="aij", bcv=[(0, 0, 0, 0.0), (0, 1, 0, 1.0)]), "SYSTEM", "ELASTICITY3D", [1.0, 1.0], False),
        ("mass periodic dof2", Case(2, dof=2, p=2, N=(24, 20), periodic=(True, False)), "SYSTEM", "MASS", [], False),
        ("mapped poisson", Case(3, p=2, N=8, geometry=("perturbed", 0.05), bcv=dall(3)), "SYSTEM", "POISSON", [], False),
        ("boundary integral", Case(3, p=2, N=6, bcv=[(0, 0, 0, 1.0)], bcf=[(d, s) for d in range(3) for s in range(2)]), "SYSTEM", "BOUNDARYINTEGRAL", [], False),
        ("boundary int mapped", Case(2, p=3, N=(9, 8), geometry=("perturbed", 0.05), bcv=[(1, 0, 0, 1.0)], bcf=[(0, 0), (0, 1), (1, 1)]), "SYSTEM", "BOUNDARYINTEGRAL", [], False),
        ("neumann demo", Case(2, p=2, N=(12, 10), bcl=[(d, s, 0, (2 * s - 1) * 6.283185307179586) for d in range(2) for s in range(2)]), "SYSTEM", "NEUMANN", [], False),
        ("mapped loads", Case(3, dof=3, p=2, N=(6, 5, 6), order=1, geometry=("perturbed", 0.05), bcv=[(0, 0, c, 0.0) for c in range(3)],
                              bcl=[(0, 1, 0, 1.0), (1, 0, 2, -0.5), (2, 1, 0, 0.25)]), "SYSTEM", "ELASTICITY", [1.0, 1.0], False),
        ("cahnhilliard3d IJ", Case(3, p=2, N=8, C=1, periodic=True, order=2), "IJACOBIAN", "CAHNHILLIARD3D", [1.5, 1.0, 0.0117], True),
        ("cahnhilliard IJ", Case(2, p=2, N=32, C=1, periodic=True), "IJACOBIAN", "CAHNHILLIARD2D", [1.5, 3000.0], True),
        ("cahnhilliard IF", Case(2, p=2, N=32, C=1, periodic=True), "IFUNCTION", "CAHNHILLIARD2D", [1.5, 3000.0], True),
        ("bratu F", Case(3, p=2, N=8, bcv=dall(3, 0.0)), "FUNCTION", "BRATU", [6.8], True),
        ("bratu J", Case(3, p=2, N=8, bcv=dall(3, 0.0)), "JACOBIAN", "BRATU", [6.8], True),
        # round 2: third-generation kernel on mapped geometry, generic kernel (mixed degrees, order-2 on mapped geometry),
        # boundary-integral matrix terms, IE / I2 / RHS drivers
        ("mapped poisson p3", Case(3, p=3, N=(8, 6, 8), geometry=("perturbed", 0.05), bcv=dall(3)), "SYSTEM", "POISSON", [], False),
        ("mixed degree mass", Case(2, dof=3, p=(2, 3), N=(10, 12), periodic=(False, True)), "SYSTEM", "MASS", [], False),
        ("dof 5 mass", Case(2, dof=5, p=(4, 3), N=(8, 9)), "SYSTEM", "MASS", [], False),
        ("nitsche 2d", Case(2, p=2, N=(12, 10), bcf=[(d, s) for d in range(2) for s in range(2)]), "SYSTEM", "NITSCHE", [], False),
        ("nitsche 3d mapped", Case(3, p=2, N=6, geometry=("perturbed", 0.05), bcf=[(d, s) for d in range(3) for s in range(2)]), "SYSTEM", "NITSCHE", [], False),
        ("ch2d mapped IJ", Case(2, p=2, N=(12, 10), order=2, geometry=("perturbed", 0.05)), "IJACOBIAN", "CAHNHILLIARD2D", [1.5, 3000.0], True),
        ("patternform IEJ", Case(2, dof=2, p=2, N=(16, 12), limits=(-1.0, 1.0), periodic=True), "IEJACOBIAN", "PATTERNFORMATION", [1.0, 0.0045, 0.5, 1.0, 0.899, -0.910, -0.899, 0.020, 0.200], True),
        ("patternform IEF", Case(2, dof=2, p=2, N=(16, 12), limits=(-1.0, 1.0), periodic=True), "IEFUNCTION", "PATTERNFORMATION", [0.0, 0.0045, 0.5, 1.0, 0.899, -0.910, -0.899, 0.020, 0.200], True),
        ("elasticrod I2F", Case(1, p=2, N=32, bcv=[(0, 0, 0, 0.0), (0, 1, 0, 0.0)]), "I2FUNCTION", "ELASTICROD", [1.3, 0.7], True),
        ("bratu RHSJ", Case(2, p=2, N=(12, 10), bcv=dall(2, 0.0)), "RHSJACOBIAN", "BRATU", [2.0], True),
    ]
    nfail = 0
    for name, case, slot, form, prm, state in cases:
        o = case.oracle()
        o.setup()
        rp, ci, rs = o.pattern(world)
        n = len(rp) - 1
        U = V = W = None
        if state:
            U, V = state_vectors(n * case.dof)
            if slot.startswith("IE") or slot.startswith("I2"):
                W = np.random.default_rng(5).random(n * case.dof)
        Ko, Fo = o.assemble(slot, form, prm, size=world, shift=7.0, V=V, U=U, W=W, shift2=0.5, t0=0.1)
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        for path in (["quadrature", "auto"] if not state else ["quadrature"]):
            g = case.product(rank=rank, size=world, nccl=comm.value, device=local)
            Ul = None if U is None else U[r0 * case.dof:r1 * case.dof]
            Vl = None if V is None else V[r0 * case.dof:r1 * case.dof]
            Wl = None if W is None else W[r0 * case.dof:r1 * case.dof]
            needs_v = slot in ("IFUNCTION", "IJACOBIAN") or W is not None
            res = run_product(case, slot, form, prm, U=Ul, V=Vl if needs_v else None, W=Wl, shift=7.0, shift2=0.5, t0=0.1, path=path, g=g)
            ok, msg = True, ""
            if slot in MAT_SLOTS:
                rpl = rp[r0:r1 + 1] - rp[r0]
                cil = ci[rp[r0]:rp[r1]]
                if res["baij"] or case.dof == 1:
                    pat_ok = bool(np.array_equal(res["rowptr"], rpl) and np.array_equal(res["colidx"], cil))
                    if not pat_ok:
                        msg += " PATTERN(rank %d: rowptr %s, colidx mismatches %d of %d)" % (
                            rank, np.array_equal(res["rowptr"], rpl), int(np.sum(res["colidx"] != cil)) if len(res["colidx"]) == len(cil) else -1, len(cil))
                    ok &= pat_ok
                exp = oracle_to_layout(Ko[rp[r0]:rp[r1]], rpl, case.dof, res["baij"])
                e = rel_frobenius(res["values"], exp)
                ok &= e <= 1e-12
                msg += " K=%.1e" % e
            if slot in VEC_SLOTS:
                e = rel_frobenius(res["rhs"], Fo[r0:r1].reshape(-1))
                ok &= e <= 1e-12
                msg += " F=%.1e" % e
            flag = torch.tensor([0 if ok else 1], device="cuda")
            dist.all_reduce(flag)
            if not ok and rank != 0:
                print("rank %d: %-20s %-10s FAIL%s" % (rank, name, path, msg), flush=True)
            if rank == 0:
                print("%-20s %-10s path=%d %s%s" % (name, path, res["path"], "ok " if flag.item() == 0 else "FAIL", msg), flush=True)
            nfail += int(flag.item() != 0)
            g.Destroy()
    # ---- IGASetFixTable on a distributed vector: G2L of the table over NCCL (test/IGAFixTable.c) ----
    for name, case in [("fixtable 2d", Case(2, p=2, N=(12, 10), bcv=dall(2))), ("fixtable 3d dof2", Case(3, dof=2, p=2, N=6, bcv=[(0, 0, 0, 0.0), (1, 1, 1, 0.0), (2, 0, 0, 0.0)]))]:
        o = case.oracle()
        o.setup()
        rp, ci, rs = o.pattern(world)
        nn = len(rp) - 1
        table = np.random.default_rng(21).standard_normal((nn, case.dof))
        o.fixtable(table)
        form = "POISSON" if case.dof == 1 else "MASS"
        Ko, Fo = o.assemble("SYSTEM", form, [], size=world)
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        g = case.product(rank=rank, size=world, nccl=comm.value, device=local)
        res = run_product(case, "SYSTEM", form, [], fixtable=table[r0:r1].reshape(-1), path="auto", g=g)
        rpl = rp[r0:r1 + 1] - rp[r0]
        eK = rel_frobenius(res["values"], oracle_to_layout(Ko[rp[r0]:rp[r1]], rpl, case.dof, res["baij"]))
        eF = rel_frobenius(res["rhs"], Fo[r0:r1].reshape(-1))
        flag = torch.tensor([0 if (eK <= 1e-12 and eF <= 1e-12) else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print("%-20s path=%d %s K=%.1e F=%.1e" % (name, res["path"], "ok " if flag.item() == 0 else "FAIL", eK, eF), flush=True)
        nfail += int(flag.item() != 0)
        g.Destroy()
    # ---- IGAComputeScalar / IGAComputeErrorNorm: state halo + ncclAllReduce (src/petigacomp.c:35-186) ----
    scalars = [
        ("errnorm 3d k=1", Case(3, dof=4, p=2, N=6, order=2), "ERRNORM", [1, 1, 0], 4),
        ("errnorm mapped k=2", Case(2, dof=4, p=3, N=8, order=2, geometry=("perturbed", 0.05)), "ERRNORM", [2, 1, 0], 4),
        ("ch stats", Case(2, p=2, N=32, C=1, periodic=True, order=2), "CH_STATS", [1.5, 3000.0, 0.63], 3),
    ]
    for name, case, sid, prm, n in scalars:
        o = case.oracle()
        o.setup()
        rp, ci, rs = o.pattern(world)
        nn = len(rp) - 1
        if sid == "CH_STATS":
            U, _ = state_vectors(nn)
        else:
            U = np.random.default_rng(9).standard_normal((nn, case.dof))
        exp = o.compute_scalar(sid, prm, n, U=U, size=world)
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        g = case.product(rank=rank, size=world, nccl=comm.value, device=local)
        vU = g.CreateVec()
        vU.set(np.asarray(U).reshape(nn, -1)[r0:r1].reshape(-1))
        if sid == "CH_STATS":
            got = g.ComputeScalar(vU, 3, "CahnHilliard2D_Stats", prm)
            err = max(abs(got[0] - exp[0]) / abs(exp[0]), abs(got[1] - exp[1]) / abs(exp[1]), abs(got[2] - exp[2]) / abs(exp[1]))
        else:
            got = g.ComputeErrorNorm(int(prm[0]), vU, "ErrNormTest") ** 2
            err = float(np.max(np.abs(got - exp) / np.abs(exp)))
        ok = err <= 1e-11
        allg = [None] * world
        dist.all_gather_object(allg, [float(x) for x in got])
        ok &= all(a == allg[0] for a in allg)            # every rank holds the same sums (MPI_Allreduce semantics)
        flag = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print("%-20s scalar     %s err=%.1e" % (name, "ok " if flag.item() == 0 else "FAIL", err), flush=True)
        nfail += int(flag.item() != 0)
        vU.destroy()
        g.Destroy()
    # ---- the step after the path on several ranks: MatMult with the operand gathered over NCCL, CG + Jacobi (pc_solve.cu) ----
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    for name, case, form, prm in [("solve poisson3d", Case(3, p=2, N=(8, 7, 9), bcv=dall(3)), "POISSON", []),
                                  ("solve elasticity3d", Case(3, dof=3, p=2, N=5, bcv=[(0, 0, c, 0.0) for c in range(3)] + [(0, 1, 0, 1.0)]), "ELASTICITY", [1.0, 1.0])]:
        o = case.oracle()
        o.setup()
        rp, ci, rs = o.pattern(world)
        Ko, Fo = o.assemble("SYSTEM", form, prm, size=world)
        dof = case.dof
        n = (len(rp) - 1) * dof
        M = (sp.bsr_matrix((Ko.reshape(-1, dof, dof), ci, rp), shape=(n, n)) if dof > 1 else sp.csr_matrix((Ko.reshape(-1), ci, rp), shape=(n, n))).tocsc()
        xs = spla.spsolve(M, Fo.reshape(-1))
        xr = np.random.default_rng(4).standard_normal(n)
        r0, r1 = int(rs[rank]) * dof, int(rs[rank + 1]) * dof
        g = case.product(rank=rank, size=world, nccl=comm.value, device=local)
        g.SetForm("SYSTEM", form, [prm[1], prm[0]] if form == "ELASTICITY" else prm)
        A, B, X, Y = g.CreateMat(), g.CreateVec(), g.CreateVec(), g.CreateVec()
        g.ComputeSystem(A, B)
        X.set(xr[r0:r1])
        g.MatMult(A, X, Y)
        e_mv = np.linalg.norm(Y.get() - (M @ xr)[r0:r1]) / np.linalg.norm(M @ xr)
        its, rel = g.Solve(A, B, X, rtol=1e-12, maxits=3000)
        e_x = np.linalg.norm(X.get() - xs[r0:r1]) / np.linalg.norm(xs)
        ok = e_mv <= 1e-12 and e_x <= 1e-8 and rel <= 1e-12
        flag = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print("%-20s solve      %s matmult=%.1e x=%.1e its=%d rel=%.1e" % (name, "ok " if flag.item() == 0 else "FAIL", e_mv, e_x, its, rel), flush=True)
        nfail += int(flag.item() != 0)
        for v in (A, B, X, Y):
            v.destroy()
        g.Destroy()
    dist.barrier()
    if rank == 0:
        print("MULTIRANK %s: %d failures on %d ranks" % ("PASS" if nfail == 0 else "FAIL", nfail, world), flush=True)
    L.petiga_cuda_comm_destroy(comm)
    dist.destroy_process_group()
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
