"""Writes tests/golden/multirank_cases.npz: the CPU oracle's assembled systems of petiga_b200.parity.golden_cases() on ONE
rank, whose PETSc numbering is the natural one (i fastest).  Run from the repo root: python tests/golden/make_multirank_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from petiga_b200.cases import state_vectors  # noqa: E402
from petiga_b200.parity import GOLDEN, golden_cases  # noqa: E402
from tests.common import Case  # noqa: E402


def main():
    out = {}
    for name, (pc, slot, form, params, state, shift) in golden_cases().items():
        case = Case.__new__(Case)
        case.__dict__.update(pc.__dict__)
        o = case.oracle()
        o.setup()
        rp, ci, _ = o.pattern(1)
        n = len(rp) - 1
        U = V = None
        if state:
            U, V = state_vectors(n * case.dof)
        K, F = o.assemble(slot, form, params, size=1, shift=shift, U=U, V=V)
        out[name + "/rowptr"], out[name + "/colidx"] = rp.astype(np.int32), ci.astype(np.int32)
        if K is not None:
            out[name + "/K"] = K
        if F is not None:
            out[name + "/F"] = F
        print(name, "rows", n, "nnz blocks", len(ci))
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e6, "MB")


if __name__ == "__main__":
    main()
