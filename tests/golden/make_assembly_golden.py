#!/usr/bin/env python
"""Golden assembly vectors: pattern, values and right-hand sides of a few small cases, produced by the CPU oracle
(oracle/petiga_oracle.c, itself pinned by the reference's known answers -- the reference cannot run here) and committed as
tests/golden/assembly_cases.npz.  They freeze the oracle (tests/test_golden_assembly.py checks it reproduces them on CPU) and
give the GPU parity tests a fixed target that does not depend on the oracle being rebuilt on the GPU box.
Run from the repo root:  python tests/golden/make_assembly_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    from tests.common import Case, state_vectors
    dall = lambda dim, v=1.0: [(d, s, 0, v) for d in range(dim) for s in range(2)]
    out = {
        # BASELINE cfg 1 shrunk: demo/Poisson2D p=2
        "poisson2d_p2": (Case(2, p=2, N=6, bcv=dall(2)), "SYSTEM", "POISSON", [], None),
        # BASELINE cfg 2 shrunk: demo/Poisson3D p=3 C2
        "poisson3d_p3": (Case(3, p=3, N=4, bcv=dall(3)), "SYSTEM", "POISSON", [], None),
        # cfg 2g: mapped geometry
        "poisson3d_p2_mapped": (Case(3, p=2, N=3, bcv=dall(3), geometry=("perturbed", 0.05)), "SYSTEM", "POISSON", [], None),
        # BASELINE cfg 3 shrunk: demo/L2Projection p=4, -function linear
        "l2projection3d_p4": (Case(3, p=4, N=2, limits=(-1.0, 1.0)), "SYSTEM", "L2PROJECTION", [0], None),
        # BASELINE cfg 4 shrunk: demo/Elasticity3D p=2, BAIJ blocks
        "elasticity3d_p2": (Case(3, dof=3, p=2, N=3, bcv=[(0, 0, 0, 0.0), (0, 0, 1, 0.0), (0, 0, 2, 0.0), (0, 1, 0, 1.0)]), "SYSTEM", "ELASTICITY3D", [1.0, 1.0], None),
        # BASELINE cfg 5 shrunk: demo/CahnHilliard2D p=2 C1 periodic, IJacobian + IFunction on the seeded state
        "cahnhilliard2d_ijac": (Case(2, p=2, N=8, C=1, periodic=True), "IJACOBIAN", "CAHNHILLIARD2D", [1.5, 3000.0], "state"),
        "cahnhilliard2d_ifun": (Case(2, p=2, N=8, C=1, periodic=True), "IFUNCTION", "CAHNHILLIARD2D", [1.5, 3000.0], "state"),
        # boundary-integral pass
        "boundary_integral2d": (Case(2, p=2, N=5, bcv=[(0, 0, 0, 1.0)], bcf=[(0, 1)]), "SYSTEM", "BOUNDARYINTEGRAL", [], None),
    }
    return out, state_vectors


def compute(case, slot, form, prm, state, state_vectors):
    o = case.oracle()
    o.setup()
    rp, ci, _ = o.pattern()
    U = V = None
    if state:
        U, V = state_vectors((len(rp) - 1) * case.dof)
    K, F = o.assemble(slot, form, prm, shift=1e3, V=V, U=U)
    return rp, ci, K, F


def main():
    cs, sv = cases()
    blob = {}
    for name, (case, slot, form, prm, state) in cs.items():
        rp, ci, K, F = compute(case, slot, form, prm, state, sv)
        blob[name + "/rowptr"] = rp.astype(np.int32)
        blob[name + "/colidx"] = ci.astype(np.int32)
        if K is not None:
            blob[name + "/values"] = K
        if F is not None:
            blob[name + "/rhs"] = F
    np.savez_compressed(os.path.join(HERE, "assembly_cases.npz"), **blob)
    print("wrote", len(blob), "arrays")


if __name__ == "__main__":
    main()
