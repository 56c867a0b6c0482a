"""GPU parity of the third-generation quadrature kernels against the CPU oracle: everything they accept, forced with quad_impl = 3.
Two matrix kernels share the stages (DMMA sum factorisation, cp.async.bulk + mbarrier ring): quad_sf3r_kernel (pc_quad3r.cuh) carries
a pencil's rows in the DMMA accumulators and runs where the axis-0 rows advance one per element (variant 0); quad_sf3_kernel
(pc_quad3.cuh) sums them in a shared-memory window and takes every knot vector (variant 1)."""
import numpy as np
import pytest

from tests.common import Case
from tests.gpu_common import check_against_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-12
VARIANTS = [0, 1]


def dall(v=1.0):
    return [(d, s, 0, v) for d in range(3) for s in range(2)]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("N", [4, (5, 3, 6), (9, 2, 3), (1, 1, 1), (2, 7, 1), (3, 2, 2)])
def test_sf3_poisson_identity(N, variant):
    res, _ = check_against_oracle(Case(3, p=3, N=N, bcv=dall()), "SYSTEM", "POISSON", path="quadrature", quad_impl=3, tol=TOL, options={"sf3_variant": variant})
    assert res["impl"] == 3 and res["sf3_variant"] == variant


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("N", [4, (6, 3, 5)])
def test_sf3_poisson_mapped(N, variant):
    res, _ = check_against_oracle(Case(3, p=3, N=N, bcv=dall(0.5), geometry=("perturbed", 0.05)), "SYSTEM", "POISSON", path="quadrature", quad_impl=3, tol=TOL, options={"sf3_variant": variant})
    assert res["impl"] == 3 and res["sf3_variant"] == variant


@pytest.mark.parametrize("C", [0, 1, 2])
def test_sf3_lower_continuity_and_periodic(C):
    """C < p-1: the window advances by p - C rows per element (window kernel); periodic axes wrap the row slots (both kernels)."""
    res, _ = check_against_oracle(Case(3, p=3, N=(4, 3, 3), C=C, bcv=dall()), "SYSTEM", "POISSON", path="quadrature", quad_impl=3, tol=TOL)
    assert res["impl"] == 3 and res["sf3_variant"] == (0 if C == 2 else 1)
    res, _ = check_against_oracle(Case(3, p=3, N=(8, 7, 9), C=C, periodic=(True, False, True)), "SYSTEM", "MASS", path="quadrature", quad_impl=3, tol=TOL)
    assert res["impl"] == 3 and res["sf3_variant"] == (0 if C == 2 else 1)


@pytest.mark.parametrize("variant", VARIANTS)
def test_sf3_other_forms_and_slots(variant):
    from functools import partial
    check_against_oracle = partial(globals()["check_against_oracle"], options={"sf3_variant": variant})
    case = Case(3, p=3, N=(4, 5, 3), limits=(-1.0, 1.0))
    for choice in (0, 4, 6):
        check_against_oracle(case, "SYSTEM", "L2PROJECTION", [choice], path="quadrature", quad_impl=3, tol=TOL)
    for slot in ("SYSTEM", "MATRIX", "VECTOR"):
        check_against_oracle(case, slot, "MASS", path="quadrature", quad_impl=3, tol=TOL)
    bcv = [(d, 0, 0, 1.0) for d in range(3)]
    bcl = [(d, 1, 0, 0.25 * (d + 1)) for d in range(3)]
    check_against_oracle(Case(3, p=3, N=4, bcv=bcv, bcl=bcl), "SYSTEM", "LAPLACE", path="quadrature", quad_impl=3, tol=TOL)
    check_against_oracle(Case(3, p=3, N=4, bcv=bcv, bcl=bcl, geometry=("perturbed", 0.05)), "SYSTEM", "LAPLACE", path="quadrature", quad_impl=3, tol=TOL)
    bcl = [(d, s, 0, (2 * s - 1) * 6.283185307179586) for d in range(3) for s in range(2)]
    check_against_oracle(Case(3, p=3, N=4, bcl=bcl), "SYSTEM", "NEUMANN", path="quadrature", quad_impl=3, tol=TOL)
    check_against_oracle(Case(3, p=3, N=4, bcv=dall(0.0)), "SYSTEM", "CONVTEST", [1.5, 0.75], path="quadrature", quad_impl=3, tol=TOL)
    check_against_oracle(Case(3, p=3, N=4, bcv=dall(0.0), geometry=("perturbed", 0.05)), "SYSTEM", "CONVTEST", [1.5, 0.75], path="quadrature", quad_impl=3, tol=TOL)


@pytest.mark.parametrize("variant", VARIANTS)
def test_sf3_fixtable(variant):
    case = Case(3, p=3, N=4, bcv=dall())
    table = np.random.default_rng(21).standard_normal(7 ** 3)
    check_against_oracle(case, "SYSTEM", "POISSON", fixtable=table, path="quadrature", quad_impl=3, tol=TOL, options={"sf3_variant": variant})


@pytest.mark.parametrize("variant", VARIANTS)
def test_sf3_midsize_many_pencils_per_cta(variant):
    """More pencils than SMs: every CTA walks several pencils (ring and flush-event parity, table reuse across pencils)."""
    from functools import partial
    from tests.gpu_common import run_product
    from tests.common import rel_frobenius
    run_product = partial(run_product, options={"sf3_variant": variant})
    case = Case(3, p=3, N=(20, 16, 14), bcv=[(d, s, 0, 1.0 + d - 0.5 * s) for d in range(3) for s in range(2)])
    a = run_product(case, "SYSTEM", "POISSON", path="quadrature", quad_impl=3)
    b = run_product(case, "SYSTEM", "POISSON", path="quadrature", quad_impl=0)
    assert a["impl"] == 3 and b["impl"] == 0
    assert rel_frobenius(a["values"], b["values"]) <= TOL and rel_frobenius(a["rhs"], b["rhs"]) <= TOL
    case = Case(3, p=3, N=(20, 16, 14), bcv=dall(), geometry=("perturbed", 0.05))
    a = run_product(case, "SYSTEM", "POISSON", path="quadrature", quad_impl=3)
    b = run_product(case, "SYSTEM", "POISSON", path="quadrature", quad_impl=0)
    assert rel_frobenius(a["values"], b["values"]) <= TOL and rel_frobenius(a["rhs"], b["rhs"]) <= TOL


@pytest.mark.parametrize("static", [0, 1])
def test_sf3r_compiled_in_structures(static):
    """The register-carried kernel runs stages A + B either from a program interpreted at run time or from one of two compiled-in
    form structures (diagonal gradient pairs on identity geometry, all nine on mapped geometry), chosen only when the run-time lists
    match the constants exactly: both must give the oracle's matrix, and the expected one must be the one that ran."""
    opts = {"sf3_variant": 0, "sf3_static": static}
    res, _ = check_against_oracle(Case(3, p=3, N=(6, 4, 5), bcv=dall()), "SYSTEM", "POISSON", path="quadrature", quad_impl=3, tol=TOL, options=opts)
    assert res["sf3_variant"] == 0 and res["sf3_static"] == (1 if static else 0)
    res, _ = check_against_oracle(Case(3, p=3, N=(5, 4, 6), bcv=[(0, 0, 0, 1.0), (0, 1, 0, 1.0)]), "SYSTEM", "LAPLACE", path="quadrature", quad_impl=3, tol=TOL, options=opts)
    assert res["sf3_static"] == (1 if static else 0)
    res, _ = check_against_oracle(Case(3, p=3, N=(6, 4, 5), bcv=dall(0.5), geometry=("perturbed", 0.05)), "SYSTEM", "POISSON", path="quadrature", quad_impl=3, tol=TOL, options=opts)
    assert res["sf3_static"] == (2 if static else 0)
    # a structure that is not compiled in (mass + stiffness: ConvTest) always runs the interpreted program
    res, _ = check_against_oracle(Case(3, p=3, N=4, bcv=dall(0.0)), "SYSTEM", "CONVTEST", [1.5, 0.75], path="quadrature", quad_impl=3, tol=TOL, options=opts)
    assert res["sf3_static"] == 0


@pytest.mark.parametrize("variant", VARIANTS)
def test_sf3_long_pencils_are_segmented(variant):
    """Pencils longer than the row tables of a work item (157 elements) are cut into segments; the rows shared by two segments
    receive their partial sums from both (tail flush of the first, head rows of the second)."""
    opts = {"sf3_variant": variant}
    res, _ = check_against_oracle(Case(3, p=3, N=(170, 1, 2), bcv=dall()), "SYSTEM", "POISSON", path="quadrature", quad_impl=3, tol=TOL, options=opts)
    assert res["impl"] == 3 and res["sf3_variant"] == variant
    res, _ = check_against_oracle(Case(3, p=3, N=(330, 1, 1), bcv=dall(0.5), geometry=("perturbed", 0.02)), "SYSTEM", "POISSON", path="quadrature", quad_impl=3, tol=TOL, options=opts)
    assert res["impl"] == 3 and res["sf3_variant"] == variant
