"""Matrix ENTRIES pinned independently of the oracle (CPU): closed-form uniform B-spline stencils -> scipy/numpy 1-D
matrices -> Kronecker sums -> the oracle's assembled CSR.  The GPU twin is tests/test_gpu_fullsize.py::test_entry_pins_gpu."""
import numpy as np
import pytest

from tests.common import Case
from tests.independent_ref import CLOSED_FORM, csr_to_dense, mass_matrix, matrices_1d, poisson_matrix


@pytest.mark.parametrize("p", [1, 2, 3])
def test_independent_1d_matrices_match_closed_forms(p):
    N = 12
    h = 1.0 / N
    M, K = matrices_1d(p, N)
    mst, kst = CLOSED_FORM[p]
    for r in range(2 * p, M.shape[0] - 2 * p):       # rows whose whole support is interior (uniform knots)
        assert np.allclose(M[r, r - p:r + p + 1], mst * h, rtol=1e-13, atol=1e-16)
        assert np.allclose(K[r, r - p:r + p + 1], kst / h, rtol=1e-12, atol=1e-13)
    assert abs(M.sum() - 1.0) < 1e-13                 # partition of unity: sum of the mass matrix = |domain|
    assert np.abs(K.sum(axis=1)).max() < 1e-11        # constants are in the kernel of the stiffness matrix


@pytest.mark.parametrize("dim,p,N", [(1, 1, 9), (1, 2, 8), (1, 3, 9), (2, 1, 6), (2, 2, 6), (2, 3, 5), (3, 1, 4), (3, 2, 4)])
def test_oracle_entries_match_independent_reference(dim, p, N):
    case = Case(dim, p=p, N=N)
    o = case.oracle()
    o.setup()
    rp, ci, _ = o.pattern()
    n = len(rp) - 1
    K, _ = o.assemble("MATRIX", "POISSON")
    M, _ = o.assemble("MATRIX", "MASS")
    Kd, Md = csr_to_dense(rp, ci, K.reshape(-1), n), csr_to_dense(rp, ci, M.reshape(-1), n)
    Ki, Mi = poisson_matrix(dim, p, N), mass_matrix(dim, p, N)
    assert np.linalg.norm(Kd - Ki) <= 1e-12 * np.linalg.norm(Ki)
    assert np.linalg.norm(Md - Mi) <= 1e-12 * np.linalg.norm(Mi)
    # the CSR pattern holds exactly the structurally nonzero entries of the tensor-product space
    dense_pattern = np.zeros((n, n), dtype=bool)
    for r in range(n):
        dense_pattern[r, ci[rp[r]:rp[r + 1]]] = True
    assert not np.any((np.abs(Mi) > 1e-14) & ~dense_pattern)
