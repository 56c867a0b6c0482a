import sys, numpy as np
sys.path.insert(0, ".")
from tests.common import Case, rel_frobenius
from tests.gpu_common import run_product
import scipy.sparse as sp

def dbg(case, slot, form, params=()):
    o = case.oracle(); o.setup()
    Ko, Fo = o.assemble(slot, form, params)
    res = run_product(case, slot, form, params, path="quadrature")
    rp, ci, _ = o.pattern()
    if Ko is not None:
        A = sp.csr_matrix((res["values"].reshape(-1), ci, rp)).toarray()
        B = sp.csr_matrix((Ko.reshape(-1), ci, rp)).toarray()
        D = np.abs(A - B)
        print(case.name, "K err", rel_frobenius(A, B), "sumA", A.sum(), "sumB", B.sum(), "max", D.max())
        bad = np.argwhere(D > 1e-10)
        print(" bad entries", len(bad), "of", np.count_nonzero(B), bad[:8].tolist())
        if len(bad):
            i, j = bad[0]; print("  A", A[i, j], "B", B[i, j])
            print("  rowA", A[i][A[i] != 0][:8], "\n  rowB", B[i][B[i] != 0][:8])
    if Fo is not None:
        print(case.name, "F err", rel_frobenius(res["rhs"], Fo.reshape(-1)), res["rhs"][:6], Fo.reshape(-1)[:6])

dbg(Case(1, p=1, N=4, name="1d p1"), "SYSTEM", "MASS")
dbg(Case(1, p=2, N=4, name="1d p2"), "SYSTEM", "MASS")
dbg(Case(2, p=1, N=3, name="2d p1"), "SYSTEM", "MASS")
dbg(Case(2, p=2, N=3, name="2d p2"), "SYSTEM", "POISSON")
dbg(Case(3, p=1, N=3, name="3d p1"), "SYSTEM", "POISSON")
dbg(Case(3, p=2, N=3, name="3d p2"), "SYSTEM", "POISSON")
