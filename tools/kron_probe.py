"""Separable-path probe: kernel time vs mesh size with and without boundary conditions (fixed cost vs streaming rate)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import Case
from tools.bench_configs import run
dall = lambda dim, v=1.0: [(d, s, 0, v) for d in range(dim) for s in range(2)]
for N in (32, 48, 64, 96, 128):
    run("poisson p=3 %d^3 bc" % N, Case(3, p=3, N=N, bcv=dall(3)), "SYSTEM", "POISSON", [], 0, "auto", steps=30)
    run("poisson p=3 %d^3 no bc" % N, Case(3, p=3, N=N), "SYSTEM", "POISSON", [], 0, "auto", steps=30)
