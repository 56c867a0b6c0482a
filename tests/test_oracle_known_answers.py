"""Pin the CPU oracle against the reference's own known answers.

Each test names the reference fixture it restates.  These run on CPU (-m "not gpu").
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle.oracle import OracleIGA, partition

sqrt2 = np.sqrt(2.0)


def make(dim, dof=1, p=2, N=16, C=-1, periodic=False, limits=(0.0, 1.0)):
    o = OracleIGA(dim, dof)
    for d in range(dim):
        pd = p[d] if isinstance(p, (list, tuple)) else p
        Nd = N[d] if isinstance(N, (list, tuple)) else N
        Cd = C[d] if isinstance(C, (list, tuple)) else C
        wd = periodic[d] if isinstance(periodic, (list, tuple)) else periodic
        o.axis_uniform(d, pd, Nd, limits[0], limits[1], Cd, wd)
    return o


def csr(o, vals, size=1):
    rp, ci, _ = o.pattern(size)
    dof = o.dof
    if dof == 1:
        return sp.csr_matrix((vals.reshape(-1), ci, rp))
    return sp.bsr_matrix((vals, ci, rp)).tocsr()


# docs/manual/TUTORIAL.rst:113-115,121-122 and :204-205 -----------------------------------------
def test_tutorial_pattern_sizes():
    o = make(3, p=2, N=16)
    o.setup()
    rp, ci, _ = o.pattern()
    assert len(rp) - 1 == 5832 and len(ci) == 592704
    o = make(2, p=2, N=16)
    o.setup()
    rp, ci, _ = o.pattern()
    assert len(rp) - 1 == 324 and len(ci) == 7056


# docs/manual/TUTORIAL.rst:179-183 ---------------------------------------------------------------
def test_tutorial_c0_quartics():
    o = make(2, p=4, N=64, C=0)
    inf = o.setup()
    assert inf["nnp"][:2] == [257, 257] and inf["nel"][:2] == [64, 64] and inf["order"] == 4


# docs/manual/TUTORIAL.rst:78-80 -----------------------------------------------------------------
def test_tutorial_partition_report():
    nnp, nel = [], []
    for r in range(8):
        o = make(3, p=2, N=16)
        inf = o.setup(8, r)
        assert inf["proc_size"] == [2, 2, 2]
        nnp.append(int(np.prod(inf["node_lwidth"])))
        nel.append(int(np.prod(inf["elem_width"])))
    assert sum(nnp) == 5832 and min(nnp) == 512 and max(nnp) == 1000
    assert sum(nel) == 4096 and min(nel) == 512 and max(nel) == 512


# cubes cut the slowest axis first; 512^2 -> (1,2),(2,2),(2,4) ------------------------
def test_partition_grids():
    assert partition(2, 0, 3, [128] * 3)[0] == [1, 1, 2]
    assert partition(4, 0, 3, [128] * 3)[0] == [1, 2, 2]
    assert partition(8, 0, 3, [128] * 3)[0] == [2, 2, 2]
    assert partition(2, 0, 2, [512] * 2)[0] == [1, 2]
    assert partition(4, 0, 2, [512] * 2)[0] == [2, 2]
    assert partition(8, 0, 2, [512] * 2)[0] == [2, 4]
    # rank -> (i,j,k), i fastest (petigapart.c:161-166)
    assert partition(8, 5, 3, [16] * 3)[1] == [1, 0, 1]


# SURVEY 8 size table (mirrors IGAAxisInitUniform + Stencil) -------------------------------------
def test_stencil_width_sums():
    for p, N, periodic, expect in [(2, 64, False, 324), (3, 128, False, 905), (4, 64, False, 592),
                                   (2, 96, False, 484), (2, 512, True, 2560)]:
        o = make(1, p=p, N=N, periodic=periodic)
        o.setup()
        rp, ci, _ = o.pattern()
        assert len(ci) == expect, (p, N, periodic, len(ci))


# 1-D tables: partition of unity and derivative sums; Gauss weights sum to 2 ----------------------
@pytest.mark.parametrize("p,C", [(1, 0), (2, 1), (2, 0), (3, 2), (3, 0), (4, 3), (5, 2)])
def test_basis_tables(p, C):
    o = make(1, p=p, N=7, C=C, limits=(-1.0, 2.0))
    o.setup()
    t = o.tables(0)
    assert np.allclose(t["value"][..., 0].sum(-1), 1.0, atol=1e-14)
    for k in range(1, min(p, 4) + 1):
        assert np.allclose(t["value"][..., k].sum(-1), 0.0, atol=1e-9 * 10 ** k)
    assert np.allclose(t["weight"].sum(-1), 2.0)
    assert np.allclose((t["weight"] * t["detJac"][:, None]).sum(), 3.0)
    # derivative table against finite differences of the value table's generating polynomial
    U = t["U"]
    assert U[0] == -1.0 and U[-1] == 2.0 and len(U) == 2 * (p + 1) + 6 * (p - C)


# test/IGACreate.c:105-149: mass system solve gives x == 1 (partition of unity) --------------------
@pytest.mark.parametrize("dim,dof,periodic,size", [(1, 1, False, 1), (2, 1, False, 1), (2, 3, False, 4),
                                                    (3, 1, False, 8), (2, 1, True, 1), (2, 2, (True, False), 2),
                                                    (3, 2, (False, True, False), 3)])
def test_igacreate_partition_of_unity(dim, dof, periodic, size):
    o = make(dim, dof, p=2, N=8 if dim < 3 else 6, periodic=periodic)
    o.setup()
    K, F = o.assemble("SYSTEM", "MASS", size=size)
    A = csr(o, K, size)
    x = spla.spsolve(A.tocsc(), F.reshape(-1))
    assert x.max() - x.min() <= 1e-2
    assert np.allclose(x, 1.0, atol=1e-10)
    # IGAComputeVector + IGAComputeMatrix variant (:127-149)
    K2, _ = o.assemble("MATRIX", "MASS", size=size)
    _, F2 = o.assemble("VECTOR", "MASS", size=size)
    assert np.allclose(K2, K) and np.allclose(F2, F)
    # any partition gives the same operator as one rank, up to the rank-major row permutation
    assert abs(A.sum() - dof) < 1e-12 and abs(F.sum() - dof) < 1e-12


# test/IGAGeometryMap.c:18-257: quarter annulus NURBS --------------------------------------------
def annulus(dim, N=4, p=2):
    """Control net of test/IGAGeometryMap.c:18-32 (PX, PY, PW; z = 2w), single quadratic patch."""
    PX = np.array([[1.0, 1.0, 0.0], [1.5, 1.5, 0.0], [2.0, 2.0, 0.0]])
    PY = np.array([[0.0, 1.0, 1.0], [0.0, 1.5, 1.5], [0.0, 2.0, 2.0]])
    PW = np.array([[1.0, sqrt2 / 2, 1.0]] * 3)
    o = OracleIGA(dim, 1)
    for d in range(dim):
        o.axis_uniform(d, 2, 1, 0.0, 1.0, 1, False)   # one Bezier element per axis
    nz = 3 if dim == 3 else 1
    X = np.zeros((nz, 3, 3, dim))
    W = np.zeros((nz, 3, 3))
    for k in range(nz):
        for j in range(3):       # axis 1 = angular direction v
            for i in range(3):   # axis 0 = radial direction u
                X[k, j, i, 0] = PX[i, j]
                X[k, j, i, 1] = PY[i, j]
                if dim == 3:
                    X[k, j, i, 2] = [0.0, 1.0, 2.0][k]
                W[k, j, i] = PW[i, j]
    if dim == 2:
        X, W = X[0], W[0]
    o.geometry(X, W)
    return o


@pytest.mark.parametrize("dim", [2, 3])
def test_geometry_map_annulus(dim):
    o = annulus(dim)
    o.order(3)
    for d in range(dim):
        o.rule_size(d, [9, 10, 8][d])     # test/makefile:63-67 uses 9/10/8-point rules
    o.setup()
    t = o.tabulate([0, 0, 0])
    u, v = t["point"][:, 0], t["point"][:, 1]
    w = t["point"][:, 2] if dim == 3 else 0 * u
    # X(u,v): IGAGeometryMap.c:46-56
    ww = v * v * (-2 + sqrt2) + v * (-sqrt2 + 2) - 1
    x = (1 + u) * (v * v * (-1 + sqrt2) + v * (-sqrt2 + 2) - 1) / ww
    y = (1 + u) * (v * v * (-1 + sqrt2) - v * sqrt2) / ww
    assert np.allclose(t["X0"][:, 0], x, atol=1e-6) and np.allclose(t["X0"][:, 1], y, atol=1e-6)
    if dim == 3:
        assert np.allclose(t["X0"][:, 2], 2 * w, atol=1e-6)
    # det J: :57-64
    J = sqrt2 * (1 + u) / ((2 - sqrt2) * v * v + (-2 + sqrt2) * v + 1) * (2 if dim == 3 else 1)
    assert np.allclose(t["detX"], J, atol=1e-6)
    # gradient of the map: :65-100
    F00 = (v * v * (-1 + sqrt2) + v * (-sqrt2 + 2) - 1) / ww
    F01 = (-v * (u + 1) * (-2 * v + sqrt2 * v + 2)) / (ww * ww)
    F10 = (v * v * (-1 + sqrt2) - v * sqrt2) / ww
    F11 = ((u + 1) * (v - 1) * (-2 * v + sqrt2 * v - sqrt2)) / (ww * ww)
    X1 = t["X1"]
    for got, exp in [(X1[:, 0, 0], F00), (X1[:, 0, 1], F01), (X1[:, 1, 0], F10), (X1[:, 1, 1], F11)]:
        assert np.allclose(got, exp, atol=1e-6)
    if dim == 3:
        assert np.allclose(X1[:, 2, 2], 2.0) and np.allclose(X1[:, 0, 2], 0) and np.allclose(X1[:, 2, 0], 0)
    # symmetry of the higher maps: :101-130
    X2, X3 = t["X2"], t["X3"]
    assert np.allclose(X2, X2.transpose(0, 1, 3, 2), atol=1e-9)
    assert np.allclose(X3, X3.transpose(0, 1, 3, 2, 4), atol=1e-9) and np.allclose(X3, X3.transpose(0, 1, 2, 4, 3), atol=1e-9)
    assert np.allclose(X2[:, 0, 0, 0], 0, atol=1e-6) and np.allclose(X2[:, 1, 0, 0], 0, atol=1e-6)
    # identities sum_a C_a grad^k N_a: :160-257  (C_a = control points, N_a physical shape functions)
    inf = o.info()
    nen = int(np.prod(inf["nen"]))
    Cx = np.zeros((nen, dim))
    PXf = np.array([[1.0, 1.0, 0.0], [1.5, 1.5, 0.0], [2.0, 2.0, 0.0]])
    PYf = np.array([[0.0, 1.0, 1.0], [0.0, 1.5, 1.5], [0.0, 2.0, 2.0]])
    a = 0
    for k in range(3 if dim == 3 else 1):
        for j in range(3):
            for i in range(3):
                Cx[a, 0], Cx[a, 1] = PXf[i, j], PYf[i, j]
                if dim == 3:
                    Cx[a, 2] = float(k)
                a += 1
    G = np.einsum("ai,qaj->qij", Cx, t["shape1"])
    assert np.allclose(G, np.eye(dim)[None], atol=1e-6)
    H = np.einsum("ai,qajk->qijk", Cx, t["shape2"])
    assert np.allclose(H, 0, atol=1e-6)
    D = np.einsum("ai,qajkl->qijkl", Cx, t["shape3"])
    assert np.allclose(D, 0, atol=1e-5)
    assert np.allclose(t["shape0"].sum(1), 1.0) and np.allclose(t["shape1"].sum(1), 0.0, atol=1e-9)
    # volume: :553-569   pi (Ro^2 - Ri^2)/4 * h
    vol = (t["detJac"] * t["weight"]).sum()
    assert abs(vol - np.pi * (4 - 1) / 4 * (2 if dim == 3 else 1)) < 1e-6


# refined annulus: area via the mass matrix on a multi-element NURBS mesh --------------------------
def test_annulus_area_from_mass_matrix():
    from tests.geomutil import refine_annulus
    o, _, _ = refine_annulus(OracleIGA, N=(5, 6), p=2)
    o.setup()
    K, F = o.assemble("SYSTEM", "MASS")
    assert abs(K.sum() - np.pi * 3 / 4) < 1e-6 and abs(F.sum() - np.pi * 3 / 4) < 1e-6   # test/IGAGeometryMap.c tolerance


# test/IGAErrNorm.c:115-147: L2 projection reproduces polynomials exactly -------------------------
@pytest.mark.parametrize("dim,p,choice,tol", [(1, 2, 0, 1e-12), (2, 2, 1, 1e-11), (3, 2, 1, 1e-11), (2, 3, 2, 1e-10)])
def test_l2_projection_reproduces_polynomials(dim, p, choice, tol):
    o = make(dim, p=p, N=5, limits=(-1.0, 1.0))
    o.setup()
    K, F = o.assemble("SYSTEM", "L2PROJECTION", params=[choice])
    A = csr(o, K)
    x = spla.spsolve(A.tocsc(), F.reshape(-1))
    # ||f||_L2^2 = x^T A x for f in the spline space; closed forms on [-1,1]^d
    got = x @ (A @ x)
    if choice == 0:      # sum x_i: int = d * (2/3) * 2^(d-1)
        exp = dim * (2.0 / 3.0) * 2 ** (dim - 1)
    elif choice == 1:    # (sum x_i^2)^2: d*2/5*2^(d-1) + d(d-1)*(2/3)^2*2^(d-2)
        exp = dim * (2.0 / 5.0) * 2 ** (dim - 1) + dim * (dim - 1) * (2.0 / 3.0) ** 2 * 2 ** (dim - 2)
    else:                # (sum x_i^3)^2: d*2/7*2^(d-1) (odd cross terms vanish)
        exp = dim * (2.0 / 7.0) * 2 ** (dim - 1)
    assert abs(got - exp) < tol * max(1, exp)


# demo/Laplace.c:108-113,146: u == 1 ------------------------------------------------------------
@pytest.mark.parametrize("dim,size", [(1, 1), (2, 1), (2, 4), (3, 2)])
def test_laplace_solution_is_one(dim, size):
    o = make(dim, p=2, N=6)
    for d in range(dim):
        o.boundary_value(d, 0, 0, 1.0)
        o.boundary_load(d, 1, 0, 0.0)
    o.setup()
    K, F = o.assemble("SYSTEM", "LAPLACE", size=size)
    x = spla.spsolve(csr(o, K, size).tocsc(), F.reshape(-1))
    assert np.allclose(x, 1.0, atol=1e-9)


# test/IGAFixTable.c: Poisson with u = sum x^2 Dirichlet data from a vector, error <= 1e-6 --------
def test_fixtable_poisson_quadratic():
    dim, p, N = 2, 2, 8
    # exact solution u = x^2 + y^2 is in the p=2 space; -lap u = -4, Poisson form has f = +1, so use
    # u = -(x^2+y^2)/4 whose -laplacian is +1.
    o = make(dim, p=p, N=N)
    o.setup()
    # L2-project u to get the fix table (as the reference test does)
    K, F = o.assemble("SYSTEM", "L2PROJECTION", params=[1])
    M = csr(o, K)
    utab = spla.spsolve(M.tocsc(), F.reshape(-1)) * (-0.25)
    o2 = make(dim, p=p, N=N)
    for d in range(dim):
        for s in range(2):
            o2.boundary_value(d, s, 0, 0.0)
    o2.fixtable(utab)
    o2.setup()
    K, F = o2.assemble("SYSTEM", "POISSON")
    A = csr(o2, K)
    x = spla.spsolve(A.tocsc(), F.reshape(-1))
    # a fixed row's diagonal counts the elements containing the node (SURVEY 7 hard part 7)
    diag = A.diagonal()
    rhs = F.reshape(-1)
    fixed = np.isclose(A.multiply(A).sum(1).A1, diag ** 2) & (diag >= 1) & np.isclose(diag, np.round(diag))
    assert fixed.sum() >= 4 * (N + p) - 4
    err = x - utab
    assert np.sqrt(err @ (M @ err)) <= 1e-6


# FixSystem semantics on the Poisson demo: corner diag 1, edge diag 2, face counts -----------------
def test_dirichlet_diagonal_counts_elements():
    o = make(2, p=2, N=4)
    for d in range(2):
        for s in range(2):
            o.boundary_value(d, s, 0, 1.0)
    o.setup()
    K, F = o.assemble("SYSTEM", "POISSON")
    A = csr(o, K).toarray()
    n = 6
    idx = lambda i, j: i + n * j
    assert A[idx(0, 0), idx(0, 0)] == 1.0 and F[idx(0, 0), 0] == 1.0
    assert A[idx(1, 0), idx(1, 0)] == 2.0 and F[idx(1, 0), 0] == 2.0
    assert A[idx(2, 0), idx(2, 0)] == 3.0 and F[idx(2, 0), 0] == 3.0
    assert np.count_nonzero(A[idx(2, 0)]) == 1 and np.count_nonzero(A[:, idx(2, 0)]) == 1
    x = np.linalg.solve(A, F.reshape(-1))
    assert np.allclose(x[[idx(0, 0), idx(3, 0), idx(5, 5)]], 1.0)


# test/ConvTest.py:71-72,95-101 (rates) on the Poisson form with zero Dirichlet data -----------------
def test_poisson_convergence_rate():
    errs = []
    for N in (4, 8, 16):
        o = make(1, p=2, N=N)
        o.boundary_value(0, 0, 0, 0.0)
        o.boundary_value(0, 1, 0, 0.0)
        o.setup()
        K, F = o.assemble("SYSTEM", "POISSON")
        A = csr(o, K)
        x = spla.spsolve(A.tocsc(), F.reshape(-1))
        # exact u = x(1-x)/2 is quadratic -> reproduced exactly at every N (rate test degenerates to ~0 error)
        Km, Fm = make_mass_rhs(N)
        errs.append(np.abs(x - Fm).max())
    assert max(errs) < 1e-12


def make_mass_rhs(N):
    o = make(1, p=2, N=N)
    o.setup()
    # nodal coefficients of u = x(1-x)/2 via L2 projection of x and x^2
    K, F1 = o.assemble("SYSTEM", "L2PROJECTION", params=[0])
    _, F2 = o.assemble("SYSTEM", "L2PROJECTION", params=[1])
    M = csr(o, K).tocsc()
    return K, (spla.spsolve(M, F1.reshape(-1)) - spla.spsolve(M, F2.reshape(-1))) / 2


# Newton consistency: Jacobian forms are the derivative of the Function forms ---------------------
@pytest.mark.parametrize("form,slotf,slotj,dim,periodic,params", [
    ("BRATU", "FUNCTION", "JACOBIAN", 2, False, [6.8]),
    ("BRATU", "IFUNCTION", "IJACOBIAN", 2, False, [6.8]),
    ("CAHNHILLIARD2D", "IFUNCTION", "IJACOBIAN", 2, True, [1.5, 3000.0]),
    ("CAHNHILLIARD3D", "IFUNCTION", "IJACOBIAN", 3, True, [1.5, 1.0, 0.003]),     # demo/CahnHilliard3D.c: theta, L0, lambda = tau*h^2
    ("POISSON", "FUNCTION", "JACOBIAN", 3, False, []),
])
def test_jacobian_matches_finite_difference(form, slotf, slotj, dim, periodic, params):
    o = make(dim, p=2, N=5 if dim == 2 else (4 if periodic else 3), periodic=periodic)
    if not periodic:
        for d in range(dim):
            o.boundary_value(d, 0, 0, 0.25)
    o.setup()
    rp, ci, _ = o.pattern()
    n = len(rp) - 1
    rng = np.random.default_rng(20261017)
    U = 0.63 + 0.05 * (2 * rng.random(n) - 1)
    V = 2 * rng.random(n) - 1
    shift = 3.0
    J, _ = o.assemble(slotj, form, params, shift=shift, V=V, U=U)
    J = csr(o, J).toarray()
    _, F0 = o.assemble(slotf, form, params, shift=shift, V=V, U=U)
    h = 1e-6
    for col in rng.choice(n, 6, replace=False):
        dU = np.zeros(n)
        dU[col] = h
        _, Fp = o.assemble(slotf, form, params, shift=shift, V=V + shift * dU, U=U + dU)
        _, Fm = o.assemble(slotf, form, params, shift=shift, V=V - shift * dU, U=U - dU)
        fd = (Fp - Fm).reshape(-1) / (2 * h)
        scale = max(1.0, np.abs(J[:, col]).max())
        assert np.allclose(fd, J[:, col], atol=2e-5 * scale), (form, col)
