"""The reference's demos end to end with the assembly and the functionals on the device (the Krylov solve, which is not on
the path, is scipy on the host): each test follows the demo's main() and checks the demo's own acceptance criterion."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from tests.common import Case
from tests.geomutil import greville, uniform_knots

pytestmark = pytest.mark.gpu


def solve(A, b, vec):
    rp, ci = A.pattern()
    x = spla.spsolve(sp.csr_matrix((A.values(), ci, rp)).tocsc(), b)
    vec.set(x)
    return x


@pytest.mark.parametrize("dim,axis,side", [(2, 0, 1), (3, 1, 0)])
def test_demo_boundary_integral(dim, axis, side):
    """demo/BoundaryIntegral.c:138-200: Dirichlet 1 on one end of `axis`, unit flux through the other via the boundary form;
    -check_error: L2 error of u = x + 1 (or 2 - x) below 1e-3 -- here it is reproduced to rounding."""
    N = 8 if dim == 2 else 5
    g = Case(dim, p=2, N=N, bcv=[(axis, 1 - side, 0, 1.0)], bcf=[(axis, side)]).product()
    g.SetForm("SYSTEM", "BOUNDARYINTEGRAL")
    A, b, x = g.CreateMat(), g.CreateVec(), g.CreateVec()
    g.ComputeSystem(A, b)
    u = solve(A, b.get(), x)
    gr = greville(uniform_knots(2, N), 2)
    xa = np.meshgrid(*[gr] * dim, indexing="ij")[axis]
    exact = ((2 - xa) if side == 0 else (xa + 1)).transpose(*range(dim)[::-1]).reshape(-1)
    assert np.abs(u - exact).max() < 1e-10
    # ||u_h||_L2 through IGAComputeErrorNorm(iga,0,x,NULL,...): int (x+1)^2 = 7/3
    assert abs(g.ComputeErrorNorm(0, x, None)[0] ** 2 - 7.0 / 3.0) < 1e-10


def test_demo_neumann():
    """demo/Neumann.c:88-176: flux loads on every face, forcing f, mean removed with the mass vector, -check_error < 1e-3."""
    dim, N = 2, 32
    bcl = [(d, s, 0, (+1 if s else -1) * 2 * np.pi) for d in range(dim) for s in range(2)]
    g = Case(dim, p=2, N=N, bcl=bcl).product()
    g.SetForm("SYSTEM", "NEUMANN")
    g.SetForm("VECTOR", "MASS")
    A, b, x, Q = g.CreateMat(), g.CreateVec(), g.CreateVec(), g.CreateVec()
    g.ComputeSystem(A, b)
    g.ComputeVector(Q)
    rp, ci = A.pattern()
    M = sp.csr_matrix((A.values(), ci, rp))
    q = Q.get()
    aug = sp.bmat([[M, sp.csr_matrix(q.reshape(-1, 1))], [sp.csr_matrix(q.reshape(1, -1)), None]]).tocsc()
    u = spla.spsolve(aug, np.concatenate([b.get(), [0.0]]))[:-1]
    x.set(u)
    err = g.ComputeErrorNorm(0, x, "Neumann")[0]
    assert err < 1e-3, err


@pytest.mark.parametrize("p", [1, 2, 3])
def test_demo_convtest_rates(p):
    """test/ConvTest.c + ConvTest.py in 2-D: L2 rate p+1 and H1 rate p within 0.075, assembly and error norms on the device."""
    dim, Ns = 2, (12, 16)
    eL2, eH1 = [], []
    for N in Ns:
        bcv = [(d, s, 0, 0.0) for d in range(dim) for s in range(2)]
        g = Case(dim, p=p, N=N, order=1, bcv=bcv).product()
        g.SetForm("SYSTEM", "CONVTEST", [1.0, 1.0])
        A, b = g.CreateMat(), g.CreateVec()
        g.ComputeSystem(A, b)
        rp, ci = A.pattern()
        u = spla.spsolve(sp.csr_matrix((A.values(), ci, rp)).tocsc(), b.get())
        e = Case(dim, p=p, N=N, order=1, q=10).product()       # 10-point rule for the norms (test/ConvTest.c:186-189)
        x = e.CreateVec(); x.set(u)
        l2 = e.ComputeErrorNorm(0, x, "ConvTest")[0]
        h1 = e.ComputeErrorNorm(1, x, "ConvTest")[0]
        eL2.append(l2); eH1.append(np.hypot(l2, h1))
    h = 1.0 / np.array(Ns, float)
    rL2 = np.polyfit(np.log10(h), np.log10(eL2), 1)[0]
    rH1 = np.polyfit(np.log10(h), np.log10(eH1), 1)[0]
    assert (p + 1) - rL2 < 0.075 and p - rH1 < 0.075, (rL2, rH1)
