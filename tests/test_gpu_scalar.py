"""GPU parity of IGAComputeScalar / IGAComputeErrorNorm (src/petigacomp.c:35-186): the device path through
the host mirror against the CPU oracle on the same seeded inputs, and against the closed forms of test/IGAErrNorm.c.
Tolerance: 1e-12 relative on the accumulated scalars (the squared norms)."""
import numpy as np
import pytest

from tests.common import Case, state_vectors

pytestmark = pytest.mark.gpu

TOL = 1e-12
S = np.sqrt
EXPECTED_L2 = [[1, 1 / S(3), 1 / S(5), 1 / S(3)], [1, S(7) / S(6), S(28) / S(45), 1 / S(9)], [1, S(5) / S(2), S(19) / S(15), 1 / S(27)]]
EXPECTED_H1 = [[0, 1, 2 / S(3), 1], [0, S(2), S(8) / S(3), S(2) / S(3)], [0, S(3), 2, 1 / S(3)]]
EXPECTED_H2 = [[0, 0, 2, 0], [0, 0, S(8), S(2)], [0, 0, S(12), S(2)]]


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def nodes_of(case):
    o = case.oracle()
    inf = o.setup()
    return o, int(np.prod(inf["nnp"][:case.dim]))


@pytest.mark.parametrize("dim,N", [(1, 8), (2, 8), (3, 4)])
def test_errnorm_closed_forms(dim, N):
    """test/IGAErrNorm.c:101-121 on the device: norms of 1, sum x, sum x^2, prod x (tolerance of the reference test)."""
    case = Case(dim, dof=4, p=2, N=N, q=3, order=2)
    g = case.product()
    for k, exp in ((0, EXPECTED_L2), (1, EXPECTED_H1), (2, EXPECTED_H2)):
        got = g.ComputeErrorNorm(k, None, "ErrNormTest")
        assert np.all(np.abs(got - np.array(exp[dim - 1], float)) < 1.4901161193847656e-08), (k, got)


@pytest.mark.parametrize("dim,p,N", [(1, 3, 9), (2, 2, 7), (2, (3, 2), (5, 6)), (3, 2, 4), (3, 3, 3), (3, (1, 2, 4), (3, 4, 2))])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_errnorm_random_state_vs_oracle(dim, p, N, k):
    """|D^k u_exact - D^k u_h|^2 with a random discrete field, dof 4, mixed degrees included."""
    pmin = min(p) if isinstance(p, tuple) else p
    if k > pmin:
        pytest.skip("derivative order above the degree")
    case = Case(dim, dof=4, p=p, N=N, order=2)
    o, n = nodes_of(case)
    U = np.random.default_rng(7).standard_normal((n, 4))
    exp = o.compute_scalar("ERRNORM", [k, 1, 0], 4, U=U)
    g = case.product()
    vU = g.CreateVec(); vU.set(U)
    got = g.ComputeErrorNorm(k, vU, "ErrNormTest") ** 2
    assert rel(got, exp) <= TOL, (got, exp)
    got0 = g.ComputeErrorNorm(k, vU, None) ** 2          # Exact == NULL: seminorm of the discrete field
    assert rel(got0, o.compute_scalar("ERRNORM", [k, 0, 0], 4, U=U)) <= TOL


@pytest.mark.parametrize("choice", [0, 3, 4, 6])
def test_errnorm_l2projection_exact(choice):
    case = Case(2, p=2, N=6, limits=(-1.0, 1.0))
    o, n = nodes_of(case)
    U = np.random.default_rng(3).standard_normal(n)
    exp = o.compute_scalar("ERRNORM", [0, 2, choice], 1, U=U)
    g = case.product()
    vU = g.CreateVec(); vU.set(U)
    got = g.ComputeErrorNorm(0, vU, "L2Projection", ctx=[float(choice)]) ** 2
    assert rel(got, exp) <= TOL


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_errnorm_mapped_geometry(dim, k):
    """K5-K7 on the field side: smooth non-rational map (cfg 2g geometry), derivative order up to 2."""
    case = Case(dim, dof=4, p=3 if dim == 2 else 2, N=4, order=2, geometry=("perturbed", 0.05))
    o, n = nodes_of(case)
    U = np.random.default_rng(11).standard_normal((n, 4))
    exp = o.compute_scalar("ERRNORM", [k, 1, 0], 4, U=U)
    g = case.product()
    vU = g.CreateVec(); vU.set(U)
    got = g.ComputeErrorNorm(k, vU, "ErrNormTest") ** 2
    assert rel(got, exp) <= 1e-11, (got, exp)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_errnorm_nurbs_annulus(k):
    """Rational geometry: the quarter annulus of test/IGAGeometryMap.c:18-32 refined to 4x4 elements."""
    import petiga_b200 as pb
    from oracle.oracle import OracleIGA
    from tests.geomutil import refine_annulus

    o, X, W = refine_annulus(OracleIGA, N=(4, 4))
    o.order(2)
    inf = o.setup()
    n = int(np.prod(inf["nnp"][:2]))
    U = np.random.default_rng(5).standard_normal(n)
    exp = o.compute_scalar("ERRNORM", [k, 0, 0], 1, U=U)

    g, _, _ = refine_annulus(pb.IGA, N=(4, 4))
    g.SetOrder(2)
    g.SetUp()
    vU = g.CreateVec(); vU.set(U)
    got = g.ComputeErrorNorm(k, vU, None) ** 2
    assert rel(got, exp) <= 1e-11, (got, exp)
    if k == 0:   # ||1||^2 = area, tolerance of test/IGAGeometryMap.c
        vU.set(np.ones(n))
        assert abs(g.ComputeErrorNorm(0, vU, None)[0] ** 2 - 3 * np.pi / 4) < 1e-6


def test_cahnhilliard_stats_vs_oracle():
    """demo/CahnHilliard2D.c:43-58 monitor functional on the cfg-5 synthetic state (64^2 here)."""
    case = Case(2, p=2, N=64, C=1, periodic=True, order=2)
    o, n = nodes_of(case)
    U, _ = state_vectors(n)
    prm = [1.5, 3000.0, 0.63]
    exp = o.compute_scalar("CH_STATS", prm, 3, U=U)
    g = case.product()
    vU = g.CreateVec(); vU.set(U)
    got = g.ComputeScalar(vU, 3, "CahnHilliard2D_Stats", prm)
    assert rel(got[:2], exp[:2]) <= TOL, (got, exp)
    assert abs(got[2] - exp[2]) <= 1e-12 * abs(exp[1])     # third moment: signed sum that nearly cancels
    again = g.ComputeScalar(vU, 3, "CahnHilliard2D_Stats", prm)
    assert np.array_equal(got, again)                      # deterministic reduction


def test_scalar_error_behaviour():
    import petiga_b200 as pb
    g = Case(2, dof=1, p=2, N=4).product()
    with pytest.raises(pb.IGAError) as e:
        g.ComputeErrorNorm(-1)                 # PETSC_ERR_ARG_OUTOFRANGE (petigacomp.c:170)
    assert e.value.code == 63
    with pytest.raises(pb.IGAError):
        g.ComputeErrorNorm(0, None, "ErrNormTest")   # that Exact is for dof == 4
