"""GPU parity tests: libpetiga_cuda (through the host mirror of the reference API) against the CPU oracle on the
same inputs.  Pattern and numbering bit-exact; values within 1e-12 relative Frobenius error (BASELINE north_star)."""
import numpy as np
import pytest

from tests.common import Case, state_vectors
from tests.gpu_common import check_against_oracle, run_product

pytestmark = pytest.mark.gpu

TOL = 1e-12


def dirichlet_all(dim, value=1.0, field=0):
    return [(d, s, field, value) for d in range(dim) for s in range(2)]


# ---- F1 Poisson (BASELINE cfg 1 and 2 at reduced mesh) --------------------------------------------------
@pytest.mark.parametrize("dim,p,N", [(1, 1, 9), (1, 2, 8), (1, 3, 7), (1, 4, 6), (2, 1, 7), (2, 2, 9), (2, 3, 6), (2, 4, 5),
                                     (3, 1, 5), (3, 2, 6), (3, 3, 6), (3, 4, 4)])
@pytest.mark.parametrize("path", ["quadrature", "auto"])
def test_poisson_system(dim, p, N, path):
    case = Case(dim, p=p, N=N, bcv=dirichlet_all(dim))
    if path == "quadrature":     # both quadrature kernels explicitly (the default picks one by element size)
        for impl in (0, 1):
            check_against_oracle(case, "SYSTEM", "POISSON", path=path, tol=TOL, quad_impl=impl)
    res, _ = check_against_oracle(case, "SYSTEM", "POISSON", path=path, tol=TOL)
    assert res["path"] == (2 if path == "auto" else 1)      # identity geometry + constant form -> separable path


def test_poisson2d_cfg1_full_size():
    """BASELINE configs[0]: Poisson2D p=2 64x64, IGAComputeSystem."""
    case = Case(2, p=2, N=64, bcv=dirichlet_all(2))
    for path in ("quadrature", "auto"):
        check_against_oracle(case, "SYSTEM", "POISSON", path=path, tol=TOL)


@pytest.mark.parametrize("C", [0, 1])
def test_poisson_lower_continuity_and_rules(C):
    case = Case(2, p=3, N=5, C=C, q=(5, 4), bcv=dirichlet_all(2, 0.5))
    check_against_oracle(case, "SYSTEM", "POISSON", path="quadrature", tol=TOL)
    check_against_oracle(case, "SYSTEM", "POISSON", path="auto", tol=TOL)


# ---- F2 Laplace: value on side 0, zero load on side 1 (demo/Laplace.c:108-113) ---------------------------
@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("path", ["quadrature", "auto"])
def test_laplace_system(dim, path):
    bcv = [(d, 0, 0, 1.0) for d in range(dim)]
    bcl = [(d, 1, 0, 0.0) for d in range(dim)]
    check_against_oracle(Case(dim, p=2, N=6, bcv=bcv, bcl=bcl), "SYSTEM", "LAPLACE", path=path, tol=TOL)


# ---- F3 L2 projection, all eight right-hand sides (demo/L2Projection.c:3-61), cfg 3 at reduced mesh -------
@pytest.mark.parametrize("choice", range(8))
def test_l2projection_functions(choice):
    check_against_oracle(Case(2, p=2, N=7, limits=(-1.0, 1.0)), "SYSTEM", "L2PROJECTION", params=[choice], tol=TOL)


@pytest.mark.parametrize("path", ["quadrature", "auto"])
def test_l2projection_3d_p4(path):
    check_against_oracle(Case(3, p=4, N=4, limits=(-1.0, 1.0)), "SYSTEM", "L2PROJECTION", params=[0], path=path, tol=TOL)


# ---- mass form of test/IGACreate.c: all three linear drivers, dof 1-3, AIJ and BAIJ, periodic mixes --------
@pytest.mark.parametrize("dim,dof,periodic,N", [(1, 1, False, 8), (2, 2, False, 6), (2, 3, (True, False), (10, 5)), (3, 1, True, 10),
                                                (3, 2, False, 4), (3, 3, (False, True, False), (4, 10, 3))])
@pytest.mark.parametrize("mattype", ["aij", "baij"])
def test_mass_matrix_vector_system(dim, dof, periodic, N, mattype):
    case = Case(dim, dof=dof, p=2, N=N, periodic=periodic, mattype=mattype)
    for slot in ("SYSTEM", "MATRIX", "VECTOR"):
        for path in ("quadrature", "auto"):
            check_against_oracle(case, slot, "MASS", path=path, tol=TOL)


# ---- F4 Elasticity3D (cfg 4 at reduced mesh), BAIJ default and AIJ --------------------------------------
@pytest.mark.parametrize("mattype", [None, "aij"])
@pytest.mark.parametrize("path", ["quadrature", "auto"])
def test_elasticity3d(mattype, path):
    bcv = [(0, 0, 0, 0.0), (0, 0, 1, 0.0), (0, 0, 2, 0.0), (0, 1, 0, 1.0)]      # demo/Elasticity3D.c:67-70
    case = Case(3, dof=3, p=2, N=(5, 4, 4), bcv=bcv, mattype=mattype)
    res, _ = check_against_oracle(case, "SYSTEM", "ELASTICITY3D", params=[1.0, 1.0], path=path, tol=TOL)
    check_against_oracle(case, "SYSTEM", "ELASTICITY3D", params=[2.5, 0.7], path=path, tol=TOL)   # exercises the mu*mu term
    if path == "quadrature":
        check_against_oracle(case, "SYSTEM", "ELASTICITY3D", params=[2.5, 0.7], path=path, tol=TOL, quad_impl=1)


# ---- F5 Elasticity (dim-generic) with a Neumann load: AddFlux / BoundaryArea ------------------------------
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("path", ["quadrature", "auto"])
def test_elasticity_generic_with_load(dim, path):
    bcv = [(0, 0, i, 0.0) for i in range(dim)]
    bcl = [(0, 1, i, [1.0, 0.5, -0.25][i]) for i in range(dim)]
    case = Case(dim, dof=dim, p=2, N=5, order=1, bcv=bcv, bcl=bcl)
    check_against_oracle(case, "SYSTEM", "ELASTICITY", params=[1.0, 1.0], path=path, tol=TOL)
    check_against_oracle(case, "SYSTEM", "ELASTICITY", params=[2.0, 0.5], path=path, tol=TOL)   # lambda != mu: C is not symmetric in (al, be)


# ---- F6 CahnHilliard2D (cfg 5 at reduced mesh): IFunction + IJacobian with state ---------------------------
@pytest.mark.parametrize("N", [8, 32])
def test_cahnhilliard2d(N):
    case = Case(2, p=2, N=N, C=1, periodic=True)
    n = N * N
    U, V = state_vectors(n)
    prm = [1.5, 3000.0]
    for impl in (0, 1):
        check_against_oracle(case, "IFUNCTION", "CAHNHILLIARD2D", prm, U=U, V=V, shift=1.0e3, tol=TOL, quad_impl=impl)
        check_against_oracle(case, "IJACOBIAN", "CAHNHILLIARD2D", prm, U=U, V=V, shift=1.0e3, tol=TOL, quad_impl=impl)


# ---- SNES / TS drivers with Dirichlet data: Bratu and the Poisson residual --------------------------------
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_bratu_function_jacobian(dim):
    case = Case(dim, p=2, N=5, bcv=dirichlet_all(dim, 0.0))
    o = case.oracle(); o.setup()
    n = len(o.pattern()[0]) - 1
    U, V = state_vectors(n)
    check_against_oracle(case, "FUNCTION", "BRATU", [6.8], U=U, tol=TOL)
    check_against_oracle(case, "JACOBIAN", "BRATU", [6.8], U=U, tol=TOL)
    check_against_oracle(case, "IFUNCTION", "BRATU", [6.8], U=U, V=V, shift=2.0, tol=TOL)
    check_against_oracle(case, "IJACOBIAN", "BRATU", [6.8], U=U, V=V, shift=2.0, tol=TOL)


def test_poisson_function_jacobian_nonzero_dirichlet():
    case = Case(2, p=3, N=5, bcv=dirichlet_all(2, 0.75))
    o = case.oracle(); o.setup()
    n = len(o.pattern()[0]) - 1
    U, _ = state_vectors(n)
    check_against_oracle(case, "FUNCTION", "POISSON", U=U, tol=TOL)
    check_against_oracle(case, "JACOBIAN", "POISSON", U=U, tol=TOL)


# ---- Dirichlet data from a vector (IGASetFixTable, test/IGAFixTable.c) --------------------------------------
def test_fixtable():
    case = Case(2, p=2, N=6, bcv=dirichlet_all(2, 0.0))
    o = case.oracle(); o.setup()
    n = len(o.pattern()[0]) - 1
    table = np.cos(np.arange(n) * 0.37)
    check_against_oracle(case, "SYSTEM", "POISSON", fixtable=table, path="quadrature", tol=TOL)
    check_against_oracle(case, "SYSTEM", "POISSON", fixtable=table, path="auto", tol=TOL)


# ---- mapped geometry (K5-K7) and NURBS (K4) ------------------------------------------------------------
@pytest.mark.parametrize("dim,p,N", [(2, 2, 6), (2, 3, 5), (3, 2, 4), (3, 3, 4)])
def test_mapped_geometry_poisson(dim, p, N):
    case = Case(dim, p=p, N=N, geometry=("perturbed", 0.05), bcv=dirichlet_all(dim))
    for impl in (0, 1):
        check_against_oracle(case, "SYSTEM", "POISSON", tol=TOL, quad_impl=impl)
    res, _ = check_against_oracle(case, "SYSTEM", "POISSON", tol=TOL)
    assert res["path"] == 1     # mapped geometry can only take the quadrature path


def test_mapped_geometry_elasticity_and_mass():
    case = Case(3, dof=3, p=2, N=3, geometry=("perturbed", 0.04), bcv=[(0, 0, i, 0.0) for i in range(3)])
    check_against_oracle(case, "SYSTEM", "ELASTICITY3D", params=[1.0, 1.0], tol=TOL)
    case = Case(2, dof=2, p=2, N=5, geometry=("perturbed", 0.05))
    check_against_oracle(case, "SYSTEM", "MASS", tol=TOL)


@pytest.mark.parametrize("form,params", [("POISSON", []), ("L2PROJECTION", [6]), ("MASS", [])])
def test_nurbs_quarter_annulus(form, params):
    """Refined quarter annulus of test/IGAGeometryMap.c:18-32 (rational weights => Rationalize path)."""
    import petiga_b200 as pb
    from oracle.oracle import OracleIGA
    from tests.geomutil import refine_annulus
    from tests.common import rel_frobenius
    o, X, W = refine_annulus(OracleIGA, N=(5, 6))
    for d in range(2):
        o.boundary_value(d, 0, 0, 1.0)
    o.setup()
    Ko, Fo = o.assemble("SYSTEM", form, params)
    g, _, _ = refine_annulus(pb.IGA, N=(5, 6))
    for d in range(2):
        g.SetBoundaryValue(d, 0, 0, 1.0)
    g.SetUp()
    g.SetForm("SYSTEM", form, params)
    A, B = g.CreateMat(), g.CreateVec()
    g.ComputeSystem(A, B)
    rp, ci = A.pattern()
    rpo, cio, _ = o.pattern()
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    assert rel_frobenius(A.values(), Ko.reshape(-1)) <= TOL
    assert rel_frobenius(B.get(), Fo.reshape(-1)) <= TOL
    if form == "MASS":   # without the Dirichlet rows the mass matrix sums to the area of the quarter annulus
        g.SetForm("MATRIX", "MASS")
        g.ComputeMatrix(A)
        assert abs(A.values().sum() - np.pi * 3 / 4) < 1e-6


# ---- graded (non-uniform) meshes through IGAAxisInitBreaks: the separable path's 1-D matrices are per node, not per mesh ----
@pytest.mark.parametrize("dim,p", [(2, 2), (3, 3)])
def test_graded_mesh_init_breaks(dim, p):
    import petiga_b200 as pb
    from oracle.oracle import OracleIGA
    from tests.common import rel_frobenius
    brk = [np.linspace(0.0, 1.0, 9 + d) ** (1.3 + 0.4 * d) for d in range(dim)]
    g = pb.IGA(dim, 1)
    o = OracleIGA(dim, 1)
    for d in range(dim):
        g.AxisInitBreaks(d, p, brk[d])
        o.axis_knots(d, p, g.AxisGetKnots(d))
        for s in range(2):
            g.SetBoundaryValue(d, s, 0, 0.5 + d)
            o.boundary_value(d, s, 0, 0.5 + d)
    g.SetUp()
    o.setup()
    Ko, Fo = o.assemble("SYSTEM", "POISSON")
    rpo, cio, _ = o.pattern()
    for path in (0, 1):
        g.SetOption("path", path)
        g.SetForm("SYSTEM", "POISSON")
        A, B = g.CreateMat(), g.CreateVec()
        g.ComputeSystem(A, B)
        rp, ci = A.pattern()
        assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
        assert rel_frobenius(A.values(), Ko.reshape(-1)) <= TOL and rel_frobenius(B.get(), Fo.reshape(-1)) <= TOL
        assert int(g.GetStat("last_path")) == (2 if path == 0 else 1)


# ---- CahnHilliard3D (demo/CahnHilliard3D.c): order-2 form with state in 3-D, periodic, both quadrature kernels ----------
@pytest.mark.parametrize("p,N", [(2, 6), (3, 8)])
def test_cahnhilliard3d(p, N):
    case = Case(3, p=p, N=N, C=p - 1, periodic=True, order=2)
    o = case.oracle(); o.setup()
    n = len(o.pattern()[0]) - 1
    U, V = state_vectors(n)
    prm = [1.5, 1.0, 0.75 / N ** 2]
    for impl in ((0, 1) if p == 2 else (None,)):     # p = 3: 49 tensor pairs do not fit the sum-factorised kernel's slot; auto falls to the pair loop
        check_against_oracle(case, "IFUNCTION", "CAHNHILLIARD3D", prm, U=U, V=V, shift=1e3, tol=TOL, quad_impl=impl)
        check_against_oracle(case, "IJACOBIAN", "CAHNHILLIARD3D", prm, U=U, V=V, shift=1e3, tol=TOL, quad_impl=impl)


# ---- large elements: 3-D p=3 dof=3 and p=4 dof 2-3 (only the sum-factorised kernel instantiates them) -----------------
@pytest.mark.parametrize("p,dof,form,prm", [(3, 3, "ELASTICITY3D", [1.0, 1.0]), (4, 2, "MASS", []), (4, 3, "ELASTICITY", [2.0, 0.5])])
def test_large_element_block_forms(p, dof, form, prm):
    bcv = [(0, 0, c, 0.0) for c in range(dof)] + [(0, 1, 0, 1.0)]
    case = Case(3, dof=dof, p=p, N=3, order=1, bcv=bcv)
    check_against_oracle(case, "SYSTEM", form, prm, path="quadrature", tol=TOL)
    case = Case(3, dof=dof, p=p, N=3, order=1, bcv=bcv, geometry=("perturbed", 0.05))
    check_against_oracle(case, "SYSTEM", form, prm, path="quadrature", tol=TOL)


# ---- Neumann loads on a mapped geometry: BoundaryArea integrates the surface Jacobian (petigaelem.c:1132-1162) ------
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("impl", [0, 1])
def test_boundary_loads_on_mapped_geometry(dim, impl):
    bcv = [(0, 0, c, 0.0) for c in range(dim)]
    bcl = [(0, 1, 0, 1.0), (1, 0, dim - 1, -0.5), (dim - 1, 1, 0, 0.25)]
    case = Case(dim, dof=dim, p=2, N=5, order=1, geometry=("perturbed", 0.05), bcv=bcv, bcl=bcl)
    check_against_oracle(case, "SYSTEM", "ELASTICITY", [1.0, 1.0], path="quadrature", tol=TOL, quad_impl=impl)


def test_boundary_loads_on_nurbs_annulus():
    import petiga_b200 as pb
    from oracle.oracle import OracleIGA
    from tests.geomutil import refine_annulus
    from tests.common import rel_frobenius
    o, X, W = refine_annulus(OracleIGA, N=(5, 4))
    g, _, _ = refine_annulus(pb.IGA, N=(5, 4))
    for obj, bv, bl in ((o, o.boundary_value, o.boundary_load), (g, g.SetBoundaryValue, g.SetBoundaryLoad)):
        bv(0, 0, 0, 1.0)
        bl(0, 1, 0, 2.0)      # outer arc r = 2: total load = 2 * (pi/2 * 2)
        bl(1, 0, 0, -1.0)
    o.setup()
    Ko, Fo = o.assemble("SYSTEM", "POISSON")
    g.SetUp()
    g.SetForm("SYSTEM", "POISSON")
    A, B = g.CreateMat(), g.CreateVec()
    g.ComputeSystem(A, B)
    assert rel_frobenius(A.values(), Ko.reshape(-1)) <= TOL and rel_frobenius(B.get(), Fo.reshape(-1)) <= TOL


# ---- separable path, multi-row interior passes with masked Dirichlet columns (needs >= 2p+1+4 nodes per axis) ---------
def test_poisson3d_p3_separable_interior_passes_vs_oracle():
    """Different Dirichlet values per face (precedence k over j over i) and one free face; mesh large enough for the
    4-row passes of kron_rows_kernel, small enough for the oracle."""
    bcv = [(0, 0, 0, 0.25), (0, 1, 0, -1.5), (1, 0, 0, 2.0), (2, 0, 0, 0.5), (2, 1, 0, 3.0)]
    case = Case(3, p=3, N=(14, 12, 13), bcv=bcv, bcl=[(1, 1, 0, 0.75)])
    res, _ = check_against_oracle(case, "SYSTEM", "POISSON", path="auto", tol=TOL)
    assert res["path"] == 2


@pytest.mark.parametrize("p,N", [(3, 24), (2, 20), (4, 18)])
def test_poisson3d_separable_equals_quadrature_midsize(p, N):
    bcv = [(d, s, 0, 1.0 + d - 0.5 * s) for d in range(3) for s in range(2)]
    case = Case(3, p=p, N=N, bcv=bcv)
    a = run_product(case, "SYSTEM", "POISSON", path="auto")
    b = run_product(case, "SYSTEM", "POISSON", path="quadrature")
    assert a["path"] == 2 and b["path"] == 1
    from tests.common import rel_frobenius
    assert rel_frobenius(a["values"], b["values"]) <= TOL and rel_frobenius(a["rhs"], b["rhs"]) <= TOL


# (full-size and near-full-size parity: tests/test_gpu_fullsize.py)


# ---- multi-GPU: ghost-row exchange + state halo over NCCL (skipped on a single-GPU box) ------------------------
def test_multirank_nccl_parity():
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "multirank_check.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(out.stdout[-4000:])
    assert "MULTIRANK PASS" in out.stdout
