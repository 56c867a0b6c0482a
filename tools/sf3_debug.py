"""Small cases of the third-generation kernels against the second-generation kernel (debug helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petiga_b200.cases import Case
from tests.gpu_common import run_product
from tests.common import rel_frobenius
dall = lambda v=1.0: [(d, s, 0, v) for d in range(3) for s in range(2)]
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for N, geo in (((1, 1, 1), None), (4, None), ((9, 2, 3), None), ((20, 16, 14), None), ((6, 3, 5), ("perturbed", 0.05))):
    case = Case(3, p=3, N=N, bcv=dall(), geometry=geo)
    a = run_product(case, "SYSTEM", "POISSON", path="quadrature", quad_impl=3, options={"sf3_variant": variant})
    b = run_product(case, "SYSTEM", "POISSON", path="quadrature", quad_impl=0)
    print(N, geo, "variant", a["sf3_variant"], "K", rel_frobenius(a["values"], b["values"]), "F", rel_frobenius(a["rhs"], b["rhs"]), flush=True)
