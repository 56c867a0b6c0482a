#!/usr/bin/env python
"""Time the five BASELINE.json configurations on one GPU (secondary rows of SURVEY 8d).  Prints one JSON line per
(config, path).  Timing: CUDA events on the plan's stream (library stat last_kernel_ms) averaged over `steps` calls."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from petiga_b200.cases import Case, state_vectors

FP64_TF = 37.2


def run(name, case, slot, form, prm, w_e, path, steps=5, state=False, shift=1.0e3, quad_impl=-1):
    g = case.product()
    g.SetOption("path", {"auto": 0, "quadrature": 1}[path])
    g.SetOption("quad_impl", quad_impl)
    g.SetForm(slot, form, prm)
    A = g.CreateMat() if slot in ("SYSTEM", "IJACOBIAN", "MATRIX") else None
    B = g.CreateVec() if slot in ("SYSTEM", "IFUNCTION") else None
    U = V = None
    if state:
        n = g.CreateVec(); nn = n.size; n.destroy()
        u, v = state_vectors(nn)
        U, V = g.CreateVec(), g.CreateVec()
        U.set(u); V.set(v)

    def call():
        if slot == "SYSTEM": g.ComputeSystem(A, B)
        elif slot == "IJACOBIAN": g.ComputeIJacobian(shift, V, 0.0, U, A)
        elif slot == "IFUNCTION": g.ComputeIFunction(shift, V, 0.0, U, B)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ms, t0 = 0.0, time.perf_counter()
    for _ in range(steps):
        call()
        ms += g.GetStat("last_kernel_ms")
    wall = (time.perf_counter() - t0) / steps * 1e3
    ms /= steps
    inf = g.info()
    nel = int(np.prod(inf["nel"]))
    nnz = (A.nnz * (case.dof ** 2 if A.baij else 1)) if A is not None else 0
    out = dict(config=name, path={1: "quadrature", 2: "kronecker"}[int(g.GetStat("last_path"))], quad_impl=int(g.GetStat("last_impl")) if int(g.GetStat("last_path")) == 1 else -1, slot=slot, form=form,
               elements=nel, nnz_scalar=nnz, kernel_ms=ms, wall_ms_per_call=wall, elements_per_s=nel / (ms * 1e-3),
               mnnz_per_s=nnz / (ms * 1e-3) / 1e6 if nnz else None,
               fp64_frac_of_nominal=(w_e * nel / (ms * 1e-3) / 1e12 / FP64_TF) if w_e else None,
               hbm_GBps=(8.0 * nnz / (ms * 1e-3) / 1e9) if nnz else None)
    print(json.dumps(out), flush=True)
    for x in (A, B, U, V):
        if x is not None:
            x.destroy()
    g.Destroy()


def main():
    dall = lambda dim, v=1.0: [(d, s, 0, v) for d in range(dim) for s in range(2)]
    which = sys.argv[1:] or ["1", "2", "3", "4", "5"]
    if "1" in which:
        c = Case(2, p=2, N=64, bcv=dall(2))
        for path in ("auto", "quadrature"):
            run("cfg1 Poisson2D p=2 64^2", c, "SYSTEM", "POISSON", [], 3888, path, steps=20)
    if "2" in which:
        c = Case(3, p=3, N=128, bcv=dall(3))
        for path in ("auto", "quadrature"):
            run("cfg2 Poisson3D p=3 128^3", c, "SYSTEM", "POISSON", [], 1847296, path)
        run("cfg2 Poisson3D p=3 128^3 (pair-loop kernel)", c, "SYSTEM", "POISSON", [], 1847296, "quadrature", steps=2, quad_impl=1)
        c = Case(3, p=3, N=128, bcv=dall(3), geometry=("perturbed", 0.05))
        run("cfg2g Poisson3D p=3 128^3 mapped", c, "SYSTEM", "POISSON", [], 1847296, "auto")
    if "3" in which:
        c = Case(3, p=4, N=64, limits=(-1.0, 1.0))
        run("cfg3 L2Projection 3D p=4 64^3", c, "SYSTEM", "L2PROJECTION", [0], 5906250, "auto", steps=3)
        run("cfg3 mass 3D p=4 64^3 (IGACreate form)", c, "SYSTEM", "MASS", [], 5906250, "auto", steps=3)
    if "4" in which:
        bcv = [(0, 0, 0, 0.0), (0, 0, 1, 0.0), (0, 0, 2, 0.0), (0, 1, 0, 1.0)]
        c = Case(3, dof=3, p=2, N=96, bcv=bcv)
        for path in ("auto", "quadrature"):
            run("cfg4 Elasticity3D p=2 96^3 BAIJ", c, "SYSTEM", "ELASTICITY3D", [1.0, 1.0], 1495908, path, steps=3)
    if "5" in which:
        c = Case(2, p=2, N=512, C=1, periodic=True)
        run("cfg5 CahnHilliard2D p=2 512^2 IJacobian", c, "IJACOBIAN", "CAHNHILLIARD2D", [1.5, 3000.0], 19200, "auto", steps=20, state=True)
        run("cfg5 CahnHilliard2D p=2 512^2 IFunction", c, "IFUNCTION", "CAHNHILLIARD2D", [1.5, 3000.0], 972, "auto", steps=20, state=True)
        run("cfg5 CahnHilliard2D p=2 512^2 IJacobian (sum-factorised kernel)", c, "IJACOBIAN", "CAHNHILLIARD2D", [1.5, 3000.0], 19200, "auto", steps=20, state=True, quad_impl=0)


if __name__ == "__main__":
    main()
