"""GPU parity of the generic runtime-degree kernel (pc_quadg.cuh) and of what only it provides (VERDICT r1 missing 1-4):
mixed degrees per axis, degree > 4, dof 4..8, second derivatives on mapped / NURBS geometry, the IE / RHS / I2 drivers and
boundary-integral MATRIX terms (Nitsche).  Everything against the CPU oracle at 1e-12; the Nitsche case also through the
demo's own check (L2 error <= 1e-6, demo/makefile:218-219)."""
import numpy as np
import pytest

from tests.common import Case, state_vectors
from tests.geomutil import refine_annulus
from tests.gpu_common import check_against_oracle, run_product

pytestmark = pytest.mark.gpu
TOL = 1e-12
PF = [1.0, 0.0045, 0.5, 1.0, 0.899, -0.910, -0.899, 0.020, 0.200]


def dall(dim, v=1.0, field=0):
    return [(d, s, field, v) for d in range(dim) for s in range(2)]


# ---- the generic kernel forced on cases the tuned kernels also run: same answers --------------------------------------------
@pytest.mark.parametrize("dim,p,N", [(1, 2, 9), (2, 3, 6), (3, 2, 5), (3, 4, 3)])
def test_generic_kernel_poisson(dim, p, N):
    res, _ = check_against_oracle(Case(dim, p=p, N=N, bcv=dall(dim)), "SYSTEM", "POISSON", path="quadrature", quad_impl=2, tol=TOL)
    assert res["impl"] == 2


def test_generic_kernel_elasticity_and_loads():
    bcv = [(0, 0, c, 0.0) for c in range(3)]
    bcl = [(0, 1, 0, 1.0), (1, 0, 2, -0.5)]
    check_against_oracle(Case(3, dof=3, p=2, N=(4, 3, 4), order=1, bcv=bcv, bcl=bcl), "SYSTEM", "ELASTICITY", [1.0, 1.0], path="quadrature", quad_impl=2, tol=TOL)
    check_against_oracle(Case(3, dof=3, p=2, N=(4, 3, 4), order=1, bcv=bcv, bcl=bcl, geometry=("perturbed", 0.05)), "SYSTEM", "ELASTICITY", [1.0, 1.0], path="quadrature", quad_impl=2, tol=TOL)
    check_against_oracle(Case(3, dof=3, p=2, N=4, mattype="aij", bcv=bcv), "SYSTEM", "ELASTICITY3D", [1.0, 1.0], path="quadrature", quad_impl=2, tol=TOL)


def test_generic_kernel_cahnhilliard_identity():
    case = Case(2, p=2, N=16, C=1, periodic=True)
    U, V = state_vectors(16 * 16)
    for slot in ("IJACOBIAN", "IFUNCTION"):
        check_against_oracle(case, slot, "CAHNHILLIARD2D", [1.5, 3000.0], U=U, V=V, shift=1e3, quad_impl=2, tol=TOL)


# ---- mixed degrees per axis, degree > 4, dof up to 8: the IGACreate sweep (test/makefile:23-40) ---------------------------------
@pytest.mark.parametrize("dim,dof,p,periodic,N", [
    (1, 4, 2, False, 8), (1, 8, 3, False, 7), (2, 4, 2, False, 5),
    (2, 3, (2, 3), (False, True), (6, 9)), (2, 3, (2, 3), (True, False), (9, 5)), (2, 3, (2, 3), (True, True), (9, 10)),
    (2, 5, (4, 3), (False, False), (4, 5)), (2, 5, (4, 3), (False, True), (4, 9)),
    (3, 2, 2, False, 4), (3, 1, (1, 2, 1), (False, True, True), (4, 6, 5)), (3, 1, (3, 2, 1), False, (4, 3, 5)),
    (2, 1, 5, False, 4), (2, 1, (6, 2), False, (3, 5)), (1, 1, 8, False, 4), (3, 1, (5, 1, 2), False, (2, 4, 3)),
])
def test_igacreate_style_sweep(dim, dof, p, periodic, N):
    case = Case(dim, dof=dof, p=p, N=N, periodic=periodic)
    for slot in ("SYSTEM", "MATRIX", "VECTOR"):
        res, _ = check_against_oracle(case, slot, "MASS", path="quadrature", tol=TOL)
    pmax = max(p) if isinstance(p, tuple) else p
    if isinstance(p, tuple) or dof > 3 or pmax > 4:
        assert res["impl"] == 2                   # none of the tuned kernels instantiates these


def test_mixed_degree_poisson_with_bcs_and_geometry():
    case = Case(2, p=(2, 3), N=(6, 5), bcv=dall(2, 0.5))
    check_against_oracle(case, "SYSTEM", "POISSON", path="quadrature", tol=TOL)
    check_against_oracle(case, "SYSTEM", "POISSON", path="auto", tol=TOL)


# ---- second derivatives on mapped and NURBS geometry (src/petigamapinv.f90.in:47-67, petigamapshf.f90.in:44-61, petigarat.f90.in:36-46)
def _annulus_case(N, dof=1):
    from oracle.oracle import OracleIGA

    class Holder:
        pass
    o, X, W = refine_annulus(OracleIGA, N=N, p=2, dof=dof)
    return o, X, W


def test_cahnhilliard2d_on_refined_quarter_annulus():
    """VERDICT r1 done-criterion: CahnHilliard2D (order-2 shape functions) on the exact NURBS quarter annulus of
    test/IGAGeometryMap.c:18-32, knot-refined to 8 x 8 elements."""
    import petiga_b200 as pb
    from tests.common import rel_frobenius
    o, X, W = _annulus_case((8, 8))
    o.order(2)
    o.setup()
    rp, ci, _ = o.pattern()
    n = len(rp) - 1
    U, V = state_vectors(n)
    g = pb.IGA(2, 1)
    inf = o.info()
    tabs = [o.tables(d) for d in range(2)]
    for d in range(2):
        g.AxisSetKnots(d, 2, tabs[d]["U"])
    g.SetOrder(2)
    g.SetGeometryArrays(X, W)
    g.SetUp()
    for slot in ("IJACOBIAN", "IFUNCTION"):
        Ko, Fo = o.assemble(slot, "CAHNHILLIARD2D", [1.5, 3000.0], shift=1e3, V=V, U=U)
        case = Case(2)          # only used for its dof
        res = run_product(case, slot, "CAHNHILLIARD2D", [1.5, 3000.0], U=U, V=V, shift=1e3, g=g)
        assert res["impl"] == 2
        if Ko is not None:
            assert np.array_equal(res["rowptr"], rp) and np.array_equal(res["colidx"], ci)
            assert rel_frobenius(res["values"], Ko.reshape(-1)) <= TOL
        else:
            assert rel_frobenius(res["rhs"], Fo.reshape(-1)) <= TOL


@pytest.mark.parametrize("form,dim,prm", [("CAHNHILLIARD2D", 2, [1.5, 3000.0]), ("CAHNHILLIARD3D", 3, [1.5, 1.0, 0.0117])])
def test_cahnhilliard_on_perturbed_geometry(form, dim, prm):
    case = Case(dim, p=2, N=6 if dim == 2 else 4, order=2, geometry=("perturbed", 0.05))
    n = (6 + 2) ** 2 if dim == 2 else (4 + 2) ** 3
    U, V = state_vectors(n)
    for slot in ("IJACOBIAN", "IFUNCTION"):
        res, _ = check_against_oracle(case, slot, form, prm, U=U, V=V, shift=7.0, tol=TOL)
        assert res["impl"] == 2


# ---- dof 4 with state: test/Test_SNES_2D.c (Function + Jacobian, its boundary values :166-186) ------------------------------------
def test_snes2d_function_jacobian():
    bcv = [(d, s, 1, 1.0) for d in range(2) for s in range(2)] + [(d, s, 2, 0.0) for d in range(2) for s in range(2)] + \
          [(d, s, 3, 0.0) for d in range(2) for s in range(2)]
    case = Case(2, dof=4, p=2, N=8, limits=(-1.0, 1.0), bcv=bcv)
    U = 0.3 + 0.2 * np.random.default_rng(5).random(10 * 10 * 4)
    for slot in ("FUNCTION", "JACOBIAN"):
        res, _ = check_against_oracle(case, slot, "SNES2D", U=U, tol=TOL)
        assert res["impl"] == 2


# ---- the IE / RHS / I2 drivers (src/petigats.c:182-477, src/petigats2.c:23-175) -----------------------------------------------
@pytest.mark.parametrize("implicit", [1.0, 0.0])
def test_patternformation_ie_drivers(implicit):
    case = Case(2, dof=2, p=2, N=12, limits=(-1.0, 1.0), periodic=True)
    n = 12 * 12 * 2
    rng = np.random.default_rng(11)
    U, V, U0 = rng.random(n), 2 * rng.random(n) - 1, rng.random(n)
    prm = [implicit] + PF[1:]
    for slot in ("IEFUNCTION", "IEJACOBIAN"):
        res, _ = check_against_oracle(case, slot, "PATTERNFORMATION", prm, U=U, V=V, W=U0, shift=2.5, t=0.3, t0=0.2, tol=TOL)
        assert res["impl"] == 2


def test_elasticrod_i2_drivers():
    case = Case(1, p=2, N=32, bcv=[(0, 0, 0, 0.0), (0, 1, 0, 0.0)])       # demo/ElasticRod.c:43-58 (64 elements there)
    n = 34
    rng = np.random.default_rng(12)
    U, V, A = rng.random(n), rng.random(n), rng.random(n)
    for slot in ("I2FUNCTION", "I2JACOBIAN"):
        check_against_oracle(case, slot, "ELASTICROD", [1.3, 0.7], U=U, V=V, W=A, shift=4.0, shift2=2.0, t=0.1, tol=TOL)


def test_rhs_drivers():
    case = Case(2, p=2, N=9, bcv=dall(2, 0.0))
    U = 0.1 * np.random.default_rng(13).random(11 * 11)
    for slot in ("RHSFUNCTION", "RHSJACOBIAN"):
        check_against_oracle(case, slot, "BRATU", [2.0], U=U, t=0.5, tol=TOL)


# ---- boundary-integral MATRIX terms: demo/NitscheMethod.c ----------------------------------------------------------------------
@pytest.mark.parametrize("dim,N,geometry", [(1, 16, None), (2, 16, None), (3, 5, None), (2, 9, ("perturbed", 0.05)), (3, 4, ("perturbed", 0.05))])
def test_nitsche_system(dim, N, geometry):
    case = Case(dim, p=2, N=N, bcf=[(d, s) for d in range(dim) for s in range(2)], geometry=geometry)
    res, _ = check_against_oracle(case, "SYSTEM", "NITSCHE", path="auto", tol=TOL)
    assert res["path"] == 1        # a visited face keeps the assembly on the quadrature path


def test_nitsche_partial_faces_with_dirichlet():
    """Nitsche terms on two faces, strong Dirichlet values on a third: the face pass must respect the element's fix-up."""
    case = Case(2, p=3, N=7, bcf=[(0, 1), (1, 0)], bcv=[(0, 0, 0, 0.25)])
    check_against_oracle(case, "SYSTEM", "NITSCHE", tol=TOL)


@pytest.mark.parametrize("dim", [1, 2])
def test_nitsche_demo_check_on_device(dim):
    """./NitscheMethod -check_error 1e-6 -iga_dim {1,2} -iga_degree 2 (demo/makefile:218-219) with the DEVICE matrix."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    case = Case(dim, p=2, N=16, bcf=[(d, s) for d in range(dim) for s in range(2)])
    res = run_product(case, "SYSTEM", "NITSCHE")
    n = len(res["rowptr"]) - 1
    A = sp.csr_matrix((res["values"], res["colidx"], res["rowptr"]), shape=(n, n))
    x = spla.spsolve(A.tocsc(), res["rhs"])
    g = case.product()
    vU = g.CreateVec(); vU.set(x)
    err = g.ComputeErrorNorm(0, vU, "L2Projection", [1.0])       # exact solution sum x_i^2 = demo/L2Projection.c "quadratic"
    assert err[0] <= 1e-6, err
