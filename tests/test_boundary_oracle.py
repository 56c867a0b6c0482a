"""Pin the oracle's boundary-integral pass (src/petigaelem.c:427-447,813-868,1012-1029, IGA_GetNormal
src/petigaval.F90:45-99) against the reference's known answers: normals / surface Jacobians of the quarter annulus
(test/IGAGeometryMap.c:275-389), the exact linear solution of demo/BoundaryIntegral.c and the error bound of demo/Neumann.c."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle.oracle import OracleIGA
from tests.geomutil import refine_annulus

TOL = 1e-6      # AssertEQUAL of test/IGAGeometryMap.c:7-12


@pytest.mark.parametrize("dim", [2, 3])
def test_annulus_boundary_normals_and_detS(dim):
    o, X, W = refine_annulus(OracleIGA, N=(3, 4), height=None if dim == 2 else (2, 2.0))
    inf = o.setup()
    nel = inf["nel"][:dim]
    for axis in range(dim):
        for side in range(2):
            for e in np.ndindex(*[nel[d] if d != axis else 1 for d in range(dim)][::-1]):
                ID = list(e[::-1])
                ID[axis] = nel[axis] - 1 if side else 0
                t = o.tabulate_boundary(ID, axis, side)
                x, y = t["X0"][:, 0], t["X0"][:, 1]
                r = np.hypot(x, y)
                n = t["normal"]
                if axis == 0:      # Boundary_00 / Boundary_01: inner (R=1) and outer (R=2) arcs, radial normals
                    R, sg = (2.0, +1.0) if side else (1.0, -1.0)
                    assert np.allclose(r, R, atol=TOL)
                    assert np.allclose(n[:, 0], sg * x / r, atol=TOL) and np.allclose(n[:, 1], sg * y / r, atol=TOL)
                    if dim == 3:
                        assert np.allclose(n[:, 2], 0.0, atol=TOL)
                elif axis == 1:    # Boundary_10 / Boundary_11: straight edges, dS = 1 (2-D) or 2 (3-D)
                    assert np.allclose(t["detS"], 1.0 if dim == 2 else 2.0, atol=TOL)
                    exp = (0.0, -1.0) if side == 0 else (-1.0, 0.0)
                    assert np.allclose(n[:, 0], exp[0], atol=TOL) and np.allclose(n[:, 1], exp[1], atol=TOL)
                else:              # Boundary_20 / Boundary_21: dS = dV/(d-1), normal -/+ e_z
                    ti = o.tabulate(ID)      # interior tabulation of the same element for detX at matching (u,v)
                    assert np.allclose(n[:, 2], +1.0 if side else -1.0, atol=TOL) and np.allclose(n[:, :2], 0.0, atol=TOL)
                    nq2 = len(t["detS"])
                    assert np.allclose(t["detS"], ti["detX"][:nq2] / (dim - 1), atol=TOL)   # z = 2w: detX does not depend on w


def csr(o, vals, size=1):
    rp, ci, _ = o.pattern(size)
    return sp.csr_matrix((vals.reshape(-1), ci, rp))


@pytest.mark.parametrize("dim,axis,side,size", [(1, 0, 1, 1), (2, 0, 1, 1), (2, 1, 0, 1), (3, 2, 1, 1), (2, 0, 0, 4), (3, 1, 1, 2)])
def test_boundary_integral_demo_exact_solution(dim, axis, side, size):
    """demo/BoundaryIntegral.c:166-172: u = 1 on one end of `axis`, unit flux through the other via the boundary form;
    the solution x_axis + 1 (or 2 - x_axis) is reproduced exactly (:121-131)."""
    o = OracleIGA(dim, 1)
    for d in range(dim):
        o.axis_uniform(d, 2, 5 if dim < 3 else 3)
    o.boundary_value(axis, 1 - side, 0, 1.0)
    o.boundary_form(axis, side, True)
    inf = o.setup()
    K, F = o.assemble("SYSTEM", "BOUNDARYINTEGRAL", size=size)
    u = spla.spsolve(csr(o, K, size).tocsc(), F.reshape(-1))
    # exact coefficients: the Greville abscissae of the axis (+1), in PETSc numbering -> compare through the L2 error
    err = o.compute_scalar("ERRNORM", [0, 0, 0], 1, U=u, size=size)      # ||u_h||^2, cross-check below with the exact field
    nn = inf["nnp"][:dim]
    from tests.geomutil import greville, uniform_knots
    g = greville(uniform_knots(2, 5 if dim < 3 else 3), 2)
    grids = np.meshgrid(*[g] * dim, indexing="ij")      # grids[d][i0,i1,..]
    xa = grids[axis]
    exact_nat = ((2 - xa) if side == 0 else (xa + 1)).transpose(*range(dim)[::-1]).reshape(-1)   # natural: i fastest
    if size == 1:
        assert np.allclose(u, exact_nat, atol=1e-10)
    else:    # global (rank-major) numbering: compare norms and extrema instead of entries
        assert abs(np.sort(u) - np.sort(exact_nat)).max() < 1e-10
    assert err[0] > 0


def test_neumann_demo_error_bound():
    """demo/Neumann.c: -lap u = f with flux loads on all faces, mean removed through the mass vector; -check_error: < 1e-3."""
    dim, N = 2, 32
    o = OracleIGA(dim, 1)
    for d in range(dim):
        o.axis_uniform(d, 2, N)
        for s in range(2):
            o.boundary_load(d, s, 0, (+1 if s else -1) * 2 * np.pi)       # Flux(dir,side), :15-18,111-116
    o.setup()
    K, F = o.assemble("SYSTEM", "NEUMANN")
    A = csr(o, K)
    b = F.reshape(-1)
    n = A.shape[0]
    # singular (pure Neumann): constrain with the mass vector Q as the demo does after the solve (:150-158)
    _, Q = o.assemble("VECTOR", "MASS")
    Q = Q.reshape(-1)
    Aug = sp.bmat([[A, sp.csr_matrix(Q.reshape(-1, 1))], [sp.csr_matrix(Q.reshape(1, -1)), None]]).tocsc()
    x = spla.spsolve(Aug, np.concatenate([b, [0.0]]))[:n]
    err = np.sqrt(o.compute_scalar("ERRNORM", [0, 3, 0], 1, U=x))[0]
    assert err < 1e-3, err
