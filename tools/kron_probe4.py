"""Separable path at cfg 4 (Elasticity3D p=2 96^3 BAIJ) with and without its Dirichlet faces."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petiga_b200.cases import Case
from tools.bench_configs import run
bcv = [(0, 0, 0, 0.0), (0, 0, 1, 0.0), (0, 0, 2, 0.0), (0, 1, 0, 1.0)]
run("cfg4 Elasticity3D p=2 96^3 BAIJ bc", Case(3, dof=3, p=2, N=96, bcv=bcv), "SYSTEM", "ELASTICITY3D", [1.0, 1.0], 0, "auto", steps=10)
run("cfg4 Elasticity3D p=2 96^3 BAIJ no bc", Case(3, dof=3, p=2, N=96), "SYSTEM", "ELASTICITY3D", [1.0, 1.0], 0, "auto", steps=10)
run("cfg4 Elasticity3D p=2 96^3 AIJ bc", Case(3, dof=3, p=2, N=96, bcv=bcv, mattype="aij"), "SYSTEM", "ELASTICITY3D", [1.0, 1.0], 0, "auto", steps=10)
