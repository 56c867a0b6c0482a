"""Pin the oracle's IGAComputeScalar / IGAComputeErrorNorm (src/petigacomp.c:35-186) against the reference's own
known answers: test/IGAErrNorm.c:101-147 (closed-form norms of 1, sum x, sum x^2, prod x on [0,1]^d and the zero error
of their L2 projection) and the CahnHilliard monitor functional (demo/CahnHilliard2D.c:36-58).  CPU only."""
import numpy as np
import pytest

from oracle.oracle import OracleIGA

S = np.sqrt
# test/IGAErrNorm.c:107-118, rows = dim 1..3, columns = the four fields
EXPECTED_L2 = [[1, 1 / S(3), 1 / S(5), 1 / S(3)], [1, S(7) / S(6), S(28) / S(45), 1 / S(9)], [1, S(5) / S(2), S(19) / S(15), 1 / S(27)]]
EXPECTED_H1 = [[0, 1, 2 / S(3), 1], [0, S(2), S(8) / S(3), S(2) / S(3)], [0, S(3), 2, 1 / S(3)]]
EXPECTED_H2 = [[0, 0, 2, 0], [0, 0, S(8), S(2)], [0, 0, S(12), S(2)]]
TOL = 1.4901161193847656e-08     # PETSC_SQRT_MACHINE_EPSILON, the reference's AssertEQUAL tolerance (:78-84)


def errnorm_iga(dim, N=8, p=2):
    """The IGA of test/IGAErrNorm.c:93-104: dof 4, Gauss-Legendre rule of size 3, order 2, [0,1]^dim."""
    o = OracleIGA(dim, 4)
    for d in range(dim):
        o.axis_uniform(d, p, N)
        o.rule_size(d, 3)
    o.order(2)
    return o


def exact_fields(x):
    """Exact() of test/IGAErrNorm.c:26-52, order 0."""
    x = np.atleast_2d(x)
    return np.stack([np.ones(len(x)), x.sum(1), (x * x).sum(1), x.prod(1)], axis=1)


def project_exact(o, dim):
    """Coefficients of the four exact fields in the spline space (they are quadratics, so p=2 holds them exactly):
    least squares over all quadrature points, using the oracle's own tabulation."""
    inf = o.info()
    nnp, nel = inf["nnp"][:dim], inf["nel"][:dim]
    offs = [o.tables(d)["offset"] for d in range(dim)]
    nen1 = [inf["nen"][d] for d in range(dim)]
    rows, rhs = [], []
    nn = int(np.prod(nnp))
    for e in np.ndindex(*nel[::-1]):
        ID = list(e[::-1])
        t = o.tabulate(ID)
        nodes = []
        for a in np.ndindex(*nen1[::-1]):
            ia = a[::-1]
            g, stride = 0, 1
            for d in range(dim):
                g += (offs[d][ID[d]] + ia[d]) * stride
                stride *= nnp[d]
            nodes.append(g)
        R = np.zeros((len(t["weight"]), nn))
        R[:, nodes] = t["shape0"]
        rows.append(R)
        rhs.append(exact_fields(t["X0"]))
    A, b = np.vstack(rows), np.vstack(rhs)
    U, *_ = np.linalg.lstsq(A, b, rcond=None)
    return U          # [nodes][4], natural = PETSc numbering on one rank


@pytest.mark.parametrize("dim,size", [(1, 1), (2, 1), (3, 1), (2, 3), (3, 2)])
def test_errnorm_of_exact_fields(dim, size):
    """IGAComputeErrorNorm(iga,k,NULL,Exact,...) = norms of the exact fields (test/IGAErrNorm.c:101-121)."""
    o = errnorm_iga(dim, N=8 if dim < 3 else 4)
    o.setup()
    for k, exp in ((0, EXPECTED_L2), (1, EXPECTED_H1), (2, EXPECTED_H2)):
        got = o.error_norm(k, U=None, exact=1, size=size)
        assert np.all(np.abs(got - np.array(exp[dim - 1], dtype=float)) < TOL), (k, got)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_errnorm_of_projection_is_zero(dim):
    """IGAComputeErrorNorm(iga,k,x,Exact,...) == 0 for the projected fields (test/IGAErrNorm.c:135-143)."""
    o = errnorm_iga(dim, N=4 if dim == 3 else 6)
    o.setup()
    U = project_exact(o, dim)
    for k in (0, 1, 2):
        got = o.error_norm(k, U=U, exact=1)
        assert np.all(got < TOL), (k, got)
    # and with Exact == NULL the same call returns the norms of the discrete field = those of the exact one
    got = o.error_norm(0, U=U, exact=0)
    assert np.all(np.abs(got - np.array(EXPECTED_L2[dim - 1], dtype=float)) < TOL)


def test_errnorm_mapped_geometry_area():
    """On the quarter annulus of test/IGAGeometryMap.c:18-32 the L2 norm of the constant field 1 is sqrt(area)."""
    from tests.geomutil import refine_annulus
    o, X, W = refine_annulus(OracleIGA, N=(4, 4))
    o.setup()
    one = np.ones(int(np.prod(o.info()["nnp"][:2])))
    got = o.compute_scalar("ERRNORM", [0, 0, 0], 1, U=one)
    assert abs(got[0] - np.pi * 3 / 4) < 1e-6    # area of the quarter annulus; tolerance of test/IGAGeometryMap.c (the integrand is rational)


def test_cahnhilliard_stats_of_constant_state():
    """demo/CahnHilliard2D.c:36-58 on c == cbar: moments vanish, energy = area * (c log c + (1-c) log(1-c) + 2 theta c (1-c))."""
    o = OracleIGA(2, 1)
    for d in range(2):
        o.axis_uniform(d, 2, 8, 0.0, 1.0, 1, True)
    o.order(2)
    o.setup()
    theta, alpha, c = 1.5, 3000.0, 0.63
    n = int(np.prod(o.info()["nnp"][:2]))
    got = o.compute_scalar("CH_STATS", [theta, alpha, c], 3, U=np.full(n, c), size=1)
    e = c * np.log(c) + (1 - c) * np.log(1 - c) + 2 * theta * c * (1 - c)
    assert abs(got[0] - e) < 1e-13 and abs(got[1]) < 1e-28 and abs(got[2]) < 1e-40
    got4 = o.compute_scalar("CH_STATS", [theta, alpha, c], 3, U=np.full(n, c), size=4)
    assert np.allclose(got4, got, rtol=1e-13, atol=1e-30)


# test/ConvTest.c + test/ConvTest.py: reaction-diffusion with u = prod sin(pi x_i); rates p+1 (L2) and p (H1), tolerance 0.075
@pytest.mark.parametrize("dim,Ns", [(1, (48, 64)), (2, (12, 16))])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_convtest_rates(dim, Ns, p):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    eL2, eH1 = [], []
    for N in Ns:
        o = OracleIGA(dim, 1)
        for d in range(dim):
            o.axis_uniform(d, p, N)
            o.boundary_value(d, 0, 0, 0.0)
            o.boundary_value(d, 1, 0, 0.0)
        o.order(1)                                   # test/ConvTest.c:140-141
        o.setup()
        K, F = o.assemble("SYSTEM", "CONVTEST", [1.0, 1.0])
        rp, ci, _ = o.pattern()
        x = spla.spsolve(sp.csr_matrix((K.reshape(-1), ci, rp)).tocsc(), F.reshape(-1))
        e = OracleIGA(dim, 1)                        # error norms with the 10-point rule (:186-189)
        for d in range(dim):
            e.axis_uniform(d, p, N)
            e.rule_size(d, 10)
        e.order(1)
        e.setup()
        l2 = e.error_norm(0, U=x, exact=4)[0]
        h1s = e.error_norm(1, U=x, exact=4)[0]
        eL2.append(l2)
        eH1.append(np.sqrt(l2 ** 2 + h1s ** 2))
    h = 1.0 / np.array(Ns, dtype=float)
    rL2 = np.polyfit(np.log10(h), np.log10(eL2), 1)[0]
    rH1 = np.polyfit(np.log10(h), np.log10(eH1), 1)[0]
    assert (p + 1) - rL2 < 0.075 and p - rH1 < 0.075, (rL2, rH1)      # checkrate() of test/ConvTest.py:72-73
