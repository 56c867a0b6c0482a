"""Known answers for the forms and drivers added in round 2 (CPU, oracle only):
   * demo/NitscheMethod.c: `-check_error 1e-6` for dim 1 and 2 at degree 2 (demo/makefile:218-219) -- boundary-integral MATRIX terms;
   * the tangents of test/Test_SNES_2D.c (dof 4), demo/PatternFormation.c (IEFunction/IEJacobian, -implicit), demo/ElasticRod
     (I2Function/I2Jacobian) and the RHS form against finite differences of their residuals -- the IE/RHS/I2 drivers
     (src/petigats.c:182-477, src/petigats2.c:23-175)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle.oracle import OracleIGA

PF = [1.0, 0.0045, 0.5, 1.0, 0.899, -0.910, -0.899, 0.020, 0.200]     # demo/PatternFormation.c:150-158, IMPLICIT = true


def csr(o, vals):
    rp, ci, _ = o.pattern()
    n, dof = len(rp) - 1, o.dof
    if dof == 1:
        return sp.csr_matrix((vals.reshape(-1), ci, rp), shape=(n, n))
    return sp.bsr_matrix((vals, ci, rp), shape=(n * dof, n * dof)).tocsr()


@pytest.mark.parametrize("dim", [1, 2])
def test_nitsche_demo_check_error(dim):
    """./NitscheMethod -check_error 1e-6 -iga_dim {1,2} -iga_degree 2: defaults are 16 elements per axis on [0,1]."""
    o = OracleIGA(dim, 1)
    for d in range(dim):
        o.axis_uniform(d, 2, 16)
        for s in range(2):
            o.boundary_form(d, s, True)
    o.setup()
    K, F = o.assemble("SYSTEM", "NITSCHE")
    A = csr(o, K)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()            # the demo declares the matrix symmetric and SPD
    x = spla.spsolve(A.tocsc(), F.reshape(-1))
    err = o.error_norm(0, U=x.reshape(-1, 1), exact=2, choice=1)  # exact solution sum x_i^2 (demo/NitscheMethod.c:32-38)
    assert err[0] <= 1e-6, err


def _fd_check(o, slotf, slotj, form, params, nvec, shift=3.0, shift2=0.7, cols=5, third_mode=None, atol=2e-5):
    rp, ci, _ = o.pattern()
    n, dof = len(rp) - 1, o.dof
    rng = np.random.default_rng(20261017)
    U = 0.3 + 0.2 * rng.random(n * dof)
    V = 2 * rng.random(n * dof) - 1
    W = 2 * rng.random(n * dof) - 1
    kw = dict(shift=shift, shift2=shift2, t0=0.1, V=V, U=U, W=W if third_mode else None)
    J, _ = o.assemble(slotj, form, params, **kw)
    J = csr(o, J).toarray()
    h = 1e-6
    for col in rng.choice(n * dof, cols, replace=False):
        dU = np.zeros(n * dof)
        dU[col] = h
        # total derivative along U with V = shift*U (+ A = shift*U for I2, where `shift` is shiftA and the form ignores V)
        def F(sgn):
            k2 = dict(kw)
            k2["U"] = U + sgn * dU
            k2["V"] = V + sgn * shift * dU
            if third_mode == "i2":
                k2["W"] = W + sgn * shift * dU      # A moves with shiftA = `shift` (first shift argument of the I2 callbacks)
            return o.assemble(slotf, form, params, **k2)[1].reshape(-1)
        fd = (F(+1) - F(-1)) / (2 * h)
        scale = max(1.0, np.abs(J[:, col]).max())
        assert np.allclose(fd, J[:, col], atol=atol * scale), (form, col, np.abs(fd - J[:, col]).max())


def test_snes2d_tangent_fd():
    o = OracleIGA(2, 4)
    for d in range(2):
        o.axis_uniform(d, 2, 5, -1.0, 1.0)
        for s in range(2):
            o.boundary_value(d, s, 1, 1.0); o.boundary_value(d, s, 2, 0.0); o.boundary_value(d, s, 3, 0.0)   # test/Test_SNES_2D.c:166-186
    o.setup()
    _fd_check(o, "FUNCTION", "JACOBIAN", "SNES2D", [], 1, shift=0.0)


def test_patternformation_ie_tangent_fd():
    o = OracleIGA(2, 2)
    for d in range(2):
        o.axis_uniform(d, 2, 6, -1.0, 1.0, -1, True)
    o.setup()
    _fd_check(o, "IEFUNCTION", "IEJACOBIAN", "PATTERNFORMATION", PF, 3, third_mode="ie")


def test_patternformation_explicit_uses_u0():
    """IMPLICIT = false (the demo's default): the reaction terms read U0, so the residual is affine in U and the tangent
    has no reaction block."""
    o = OracleIGA(2, 2)
    for d in range(2):
        o.axis_uniform(d, 2, 4, -1.0, 1.0, -1, True)
    o.setup()
    rp, ci, _ = o.pattern()
    n = (len(rp) - 1) * 2
    rng = np.random.default_rng(3)
    U, V, U0 = rng.random(n), rng.random(n), rng.random(n)
    prm = [0.0] + PF[1:]
    _, F1 = o.assemble("IEFUNCTION", "PATTERNFORMATION", prm, shift=2.0, V=V, U=U, W=U0)
    _, F2 = o.assemble("IEFUNCTION", "PATTERNFORMATION", prm, shift=2.0, V=V, U=U, W=U0 + 0.1)
    assert np.abs(F1 - F2).max() > 1e-6                      # depends on U0
    J, _ = o.assemble("IEJACOBIAN", "PATTERNFORMATION", prm, shift=2.0, V=V, U=U, W=U0)
    assert np.abs(J[:, 0, 1]).max() == 0 and np.abs(J[:, 1, 0]).max() == 0


def test_elasticrod_i2_tangent_fd():
    o = OracleIGA(1, 1)
    o.axis_uniform(0, 2, 12)
    o.boundary_value(0, 0, 0, 0.0); o.boundary_value(0, 1, 0, 0.0)       # demo/ElasticRod.c:52-58
    o.setup()
    _fd_check(o, "I2FUNCTION", "I2JACOBIAN", "ELASTICROD", [1.3, 0.7], 3, third_mode="i2")


def test_rhs_tangent_fd():
    o = OracleIGA(2, 1)
    for d in range(2):
        o.axis_uniform(d, 2, 5)
        o.boundary_value(d, 0, 0, 0.0)
    o.setup()
    _fd_check(o, "RHSFUNCTION", "RHSJACOBIAN", "BRATU", [2.0], 1, shift=0.0)
