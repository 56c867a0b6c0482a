import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import Case
from tools.bench_configs import run
dall = lambda dim, v=1.0: [(d, s, 0, v) for d in range(dim) for s in range(2)]
run("mass p=3 128^3 no bc", Case(3, p=3, N=128), "SYSTEM", "MASS", [], 0, "auto", steps=20)
run("poisson p=3 128^3 no bc", Case(3, p=3, N=128), "SYSTEM", "POISSON", [], 0, "auto", steps=20)
run("poisson p=3 128^3 bc", Case(3, p=3, N=128, bcv=dall(3)), "SYSTEM", "POISSON", [], 0, "auto", steps=20)
run("poisson p=3 128^3 bc MATRIX only", Case(3, p=3, N=128, bcv=dall(3)), "MATRIX", "POISSON", [], 0, "auto", steps=20) if False else None
run("poisson p=2 160^3 bc", Case(3, p=2, N=160, bcv=dall(3)), "SYSTEM", "POISSON", [], 0, "auto", steps=20)
run("poisson p=4 64^3 bc", Case(3, p=4, N=64, bcv=dall(3)), "SYSTEM", "POISSON", [], 0, "auto", steps=20)
