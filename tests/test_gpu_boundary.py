"""GPU parity of the boundary-integral pass (IGASetBoundaryForm, src/petigaelem.c:427-447,813-868,1012-1029)
and of the demo/Neumann.c system against the CPU oracle; tolerance 1e-12 relative Frobenius."""
import numpy as np
import pytest

from tests.common import Case
from tests.gpu_common import check_against_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("dim,axis,side", [(1, 0, 1), (1, 0, 0), (2, 0, 1), (2, 1, 0), (3, 0, 0), (3, 1, 1), (3, 2, 0)])
@pytest.mark.parametrize("impl", [0, 1])
def test_boundary_integral_identity(dim, axis, side, impl):
    """demo/BoundaryIntegral.c:166-172: Dirichlet 1.0 on one end of `axis`, boundary form on the other."""
    case = Case(dim, p=2, N=5 if dim < 3 else 4, bcv=[(axis, 1 - side, 0, 1.0)], bcf=[(axis, side)])
    res, _ = check_against_oracle(case, "SYSTEM", "BOUNDARYINTEGRAL", path="quadrature", tol=TOL, quad_impl=impl)
    assert res["path"] == 1
    rhs = res["rhs"]
    # total face load of the free dofs <= face area 1 (unit square / cube faces), and > 0
    assert 0 < rhs.sum()


@pytest.mark.parametrize("dim", [2, 3])
def test_boundary_integral_all_faces_mixed_with_dirichlet(dim):
    """Every face visited, Dirichlet values on two of them: the face terms of fixed dofs are discarded (FixSystem)."""
    bcf = [(d, s) for d in range(dim) for s in range(2)]
    case = Case(dim, p=3, N=4 if dim == 2 else 3, bcv=[(0, 0, 0, 2.0), (1, 1, 0, -1.0)], bcf=bcf)
    check_against_oracle(case, "SYSTEM", "BOUNDARYINTEGRAL", tol=TOL)


@pytest.mark.parametrize("dim", [2, 3])
def test_boundary_integral_mapped_geometry(dim):
    bcf = [(d, s) for d in range(dim) for s in range(2)]
    case = Case(dim, p=2, N=4, geometry=("perturbed", 0.05), bcv=[(0, 0, 0, 1.0)], bcf=bcf)
    check_against_oracle(case, "SYSTEM", "BOUNDARYINTEGRAL", tol=TOL)


def test_boundary_integral_nurbs_annulus():
    """Face integrals on the quarter annulus: the outer arc has length pi, so the face term sums to pi (test/IGAGeometryMap.c geometry)."""
    import petiga_b200 as pb
    from oracle.oracle import OracleIGA
    from tests.common import rel_frobenius
    from tests.geomutil import refine_annulus
    o, X, W = refine_annulus(OracleIGA, N=(4, 5))
    g, _, _ = refine_annulus(pb.IGA, N=(4, 5))
    o.boundary_form(0, 1, True)
    g.SetBoundaryForm(0, 1, True)
    o.setup()
    Ko, Fo = o.assemble("SYSTEM", "BOUNDARYINTEGRAL")
    g.SetUp()
    g.SetForm("SYSTEM", "BOUNDARYINTEGRAL")
    A, B = g.CreateMat(), g.CreateVec()
    g.ComputeSystem(A, B)
    assert rel_frobenius(A.values(), Ko.reshape(-1)) <= TOL and rel_frobenius(B.get(), Fo.reshape(-1)) <= TOL
    assert abs(B.get().sum() - np.pi) < 1e-6      # sum_a int N_a dS = arc length of r = 2 over a quarter circle


def test_boundary_form_needs_a_form_with_a_face_term():
    import petiga_b200 as pb
    g = Case(2, p=2, N=4, bcf=[(0, 1)]).product()
    g.SetForm("SYSTEM", "POISSON")
    A, B = g.CreateMat(), g.CreateVec()
    with pytest.raises(pb.IGAError) as e:
        g.ComputeSystem(A, B)
    assert e.value.code == 56       # PETSC_ERR_SUP


@pytest.mark.parametrize("dim,N", [(1, 16), (2, 12), (3, 5)])
def test_neumann_demo_system(dim, N):
    """demo/Neumann.c: flux loads on every face (AddFlux) + the forcing f(x); separable matrix, point-wise vector."""
    bcl = [(d, s, 0, (+1 if s else -1) * 2 * np.pi) for d in range(dim) for s in range(2)]
    case = Case(dim, p=2, N=N, bcl=bcl)
    for path in ("quadrature", "auto"):
        check_against_oracle(case, "SYSTEM", "NEUMANN", path=path, tol=TOL)


def test_neumann_demo_error_norm_on_device():
    """IGAComputeErrorNorm with the demo's Exact (demo/Neumann.c:80-86) against the oracle on a seeded field."""
    case = Case(2, p=2, N=8)
    o = case.oracle()
    inf = o.setup()
    n = int(np.prod(inf["nnp"][:2]))
    U = np.random.default_rng(4).standard_normal(n)
    exp = o.compute_scalar("ERRNORM", [0, 3, 0], 1, U=U)
    g = case.product()
    vU = g.CreateVec(); vU.set(U)
    got = g.ComputeErrorNorm(0, vU, "Neumann") ** 2
    assert abs(got[0] - exp[0]) <= 1e-12 * abs(exp[0])


@pytest.mark.parametrize("dim,p,N", [(1, 3, 12), (2, 2, 9), (3, 3, 4), (3, 1, 6)])
def test_convtest_system_and_error_norms(dim, p, N):
    """test/ConvTest.c: reaction-diffusion system and the L2 / H1 error functionals of its exact solution."""
    case = Case(dim, p=p, N=N, order=1, bcv=[(d, s, 0, 0.0) for d in range(dim) for s in range(2)])
    check_against_oracle(case, "SYSTEM", "CONVTEST", [1.5, 0.5], tol=TOL)
    o = case.oracle()
    inf = o.setup()
    n = int(np.prod(inf["nnp"][:dim]))
    U = np.random.default_rng(8).standard_normal(n)
    g = case.product()
    vU = g.CreateVec(); vU.set(U)
    for k in (0, 1):
        exp = o.compute_scalar("ERRNORM", [k, 4, 0], 1, U=U)
        got = g.ComputeErrorNorm(k, vU, "ConvTest") ** 2
        assert abs(got[0] - exp[0]) <= 1e-12 * abs(exp[0]), (k, got, exp)
