"""GPU-side helpers: run one assembly through the host mirror (C-ABI underneath) and through the oracle."""
import numpy as np

from tests.common import oracle_to_layout, rel_frobenius

MAT_SLOTS = ("MATRIX", "SYSTEM", "JACOBIAN", "IJACOBIAN", "IEJACOBIAN", "RHSJACOBIAN", "I2JACOBIAN")
VEC_SLOTS = ("VECTOR", "SYSTEM", "FUNCTION", "IFUNCTION", "IEFUNCTION", "RHSFUNCTION", "I2FUNCTION")


def run_product(case, slot, form, params=(), U=None, V=None, shift=0.0, t=0.0, path=None, fixtable=None, g=None, quad_impl=None,
                W=None, shift2=0.0, t0=0.0, options=None):
    g = g or case.product()
    for name, value in (options or {}).items():
        g.SetOption(name, value)
    if quad_impl is not None:
        g.SetOption("quad_impl", quad_impl)
    if path is not None:
        g.SetOption("path", {"auto": 0, "quadrature": 1, "kronecker": 2}[path])
    # params are (lambda, mu) as the oracle takes them; the AppCtx of demo/Elasticity.c:10-13 is {mu, lambda}
    g.SetForm(slot, form, [params[1], params[0]] if form == "ELASTICITY" else params)
    A = g.CreateMat() if slot in MAT_SLOTS else None
    B = g.CreateVec() if slot in VEC_SLOTS else None
    vU = vV = vT = None
    if fixtable is not None:
        vT = g.CreateVec(); vT.set(fixtable); g.SetFixTable(vT)
    if U is not None:
        vU = g.CreateVec(); vU.set(U)
    if V is not None:
        vV = g.CreateVec(); vV.set(V)
    vW = None
    if W is not None:
        vW = g.CreateVec(); vW.set(W)
    out_obj = A if slot in MAT_SLOTS else B
    if slot in ("IEFUNCTION", "IEJACOBIAN"): g.ComputeIEFunction(shift, vV, t, vU, t0, vW, out_obj) if slot == "IEFUNCTION" else g.ComputeIEJacobian(shift, vV, t, vU, t0, vW, out_obj)
    elif slot in ("RHSFUNCTION", "RHSJACOBIAN"): g.ComputeRHSFunction(t, vU, out_obj) if slot == "RHSFUNCTION" else g.ComputeRHSJacobian(t, vU, out_obj)
    elif slot in ("I2FUNCTION", "I2JACOBIAN"): g.ComputeI2Function(shift, vW, shift2, vV, t, vU, out_obj) if slot == "I2FUNCTION" else g.ComputeI2Jacobian(shift, vW, shift2, vV, t, vU, out_obj)
    elif slot == "VECTOR": g.ComputeVector(B)
    elif slot == "MATRIX": g.ComputeMatrix(A)
    elif slot == "SYSTEM": g.ComputeSystem(A, B)
    elif slot == "FUNCTION": g.ComputeFunction(vU, B)
    elif slot == "JACOBIAN": g.ComputeJacobian(vU, A)
    elif slot == "IFUNCTION": g.ComputeIFunction(shift, vV, t, vU, B)
    elif slot == "IJACOBIAN": g.ComputeIJacobian(shift, vV, t, vU, A)
    out = dict(path=int(g.GetStat("last_path")), impl=int(g.GetStat("last_impl")), sf3_variant=int(g.GetStat("last_sf3_variant")),
               sf3_static=int(g.GetStat("last_sf3_static")), g=g)
    if A is not None:
        out["rowptr"], out["colidx"] = A.pattern()
        out["values"] = A.values()
        out["baij"] = A.baij
    if B is not None:
        out["rhs"] = B.get()
    for v in (A, B, vU, vV, vT, vW):
        if v is not None:
            v.destroy()
    return out


def check_against_oracle(case, slot, form, params=(), U=None, V=None, shift=0.0, t=0.0, path=None, fixtable=None, tol=1e-12, quad_impl=None,
                         W=None, shift2=0.0, t0=0.0, options=None):
    """Pattern bit-exact; values / vectors within `tol` relative Frobenius error (north_star: 1e-12)."""
    o = case.oracle()
    if fixtable is not None:
        o.fixtable(fixtable)
    o.setup()
    Ko, Fo = o.assemble(slot, form, params, shift=shift, V=V, t=t, U=U, W=W, shift2=shift2, t0=t0)
    res = run_product(case, slot, form, params, U=U, V=V, shift=shift, t=t, path=path, fixtable=fixtable, quad_impl=quad_impl,
                      W=W, shift2=shift2, t0=t0, options=options)
    rp_o, ci_o, _ = o.pattern()
    errs = {}
    if Ko is not None:
        dof = case.dof
        if res["baij"] or dof == 1:
            assert np.array_equal(res["rowptr"], rp_o) and np.array_equal(res["colidx"], ci_o)
        else:
            assert len(res["rowptr"]) - 1 == (len(rp_o) - 1) * dof and len(res["colidx"]) == len(ci_o) * dof * dof
        exp = oracle_to_layout(Ko, rp_o, dof, res["baij"])
        errs["K"] = rel_frobenius(res["values"], exp)
        assert errs["K"] <= tol, ("matrix", errs["K"])
    if Fo is not None:
        errs["F"] = rel_frobenius(res["rhs"], Fo.reshape(-1))
        scale = np.linalg.norm(Fo)
        if scale == 0:
            assert np.abs(res["rhs"]).max() == 0
        else:
            assert errs["F"] <= tol, ("vector", errs["F"])
    return res, errs


def check_against_parallel_oracle(case, slot, form, params=(), U=None, V=None, shift=0.0, t=0.0, path=None, tol=1e-12, quad_impl=None,
                                  T=None):
    """Same as check_against_oracle, with the oracle's one-rank element loop split over host threads (tests/par_oracle.py):
    for the cases at or near BASELINE size, where one thread would take minutes."""
    from tests.par_oracle import assemble_parallel
    rp_o, ci_o, Ko, Fo = assemble_parallel(case, slot, form, params, T=T, shift=shift, V=V, t=t, U=U)
    res = run_product(case, slot, form, params, U=U, V=V, shift=shift, t=t, path=path, quad_impl=quad_impl)
    errs = {}
    if Ko is not None:
        dof = case.dof
        if res["baij"] or dof == 1:
            assert np.array_equal(res["rowptr"], rp_o) and np.array_equal(res["colidx"], ci_o)
        exp = oracle_to_layout(Ko, rp_o, dof, res["baij"])
        errs["K"] = rel_frobenius(res["values"], exp)
        assert errs["K"] <= tol, ("matrix", errs["K"])
    if Fo is not None:
        errs["F"] = rel_frobenius(res["rhs"], Fo.reshape(-1))
        assert errs["F"] <= tol, ("vector", errs["F"])
    return res, errs
