"""GPU parity of the vector-only sum-factorised kernel (pc_quadv.cuh, quad_vec3_kernel<P>) against the CPU oracle:
IGAComputeVector and the hybrid system of BASELINE cfg 3 (matrix by the separable path, load vector by this kernel)."""
import pytest

from tests.common import Case
from tests.gpu_common import check_against_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("p", [2, 3, 4])
@pytest.mark.parametrize("geometry", [None, ("perturbed", 0.05)])
def test_vec3_compute_vector(p, geometry):
    """IGAComputeVector (the only built-in Vector-slot form is the IGACreate test's mass form, test/IGACreate.c:14-22)."""
    case = Case(3, p=p, N=(4, 3, 5), limits=(-1.0, 1.0), geometry=geometry)
    res, _ = check_against_oracle(case, "VECTOR", "MASS", path="quadrature", tol=TOL)
    assert res["impl"] == 4
    if geometry is None:      # (the perturbed-geometry generator of petiga_b200/cases.py is for maximally smooth knot vectors)
        case = Case(3, p=p, N=(3, 6, 2), C=0, periodic=(True, False, False))
    else:
        case = Case(3, p=p, N=(3, 6, 2), geometry=geometry)
    res, _ = check_against_oracle(case, "VECTOR", "MASS", path="quadrature", tol=TOL)
    assert res["impl"] == 4


@pytest.mark.parametrize("p", [2, 3, 4])
def test_vec3_hybrid_system_cfg3_shape(p):
    """demo/L2Projection.c: separable mass matrix + point-wise load; lower continuity and a periodic axis on the way."""
    case = Case(3, p=p, N=(5, 4, 6), limits=(-1.0, 1.0))
    for choice in (0, 4, 6):
        res, _ = check_against_oracle(case, "SYSTEM", "L2PROJECTION", [choice], path="auto", tol=TOL)
        assert res["path"] == 2 and res["impl"] == 4
    case = Case(3, p=p, N=(6, 5, 7), C=0, periodic=(False, True, False), limits=(-1.0, 1.0))
    res, _ = check_against_oracle(case, "SYSTEM", "L2PROJECTION", [4], path="auto", tol=TOL)
    assert res["path"] == 2 and res["impl"] == 4
