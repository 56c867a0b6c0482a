"""Shared test helpers: build the oracle and the product from one case description."""
import numpy as np

from oracle.oracle import OracleIGA


def _per_axis(v, d):
    return v[d] if isinstance(v, (list, tuple)) else v


class Case:
    """One discretisation + BC + form configuration, applicable to both the oracle and the product."""

    def __init__(self, dim, dof=1, p=2, N=8, C=-1, periodic=False, limits=(0.0, 1.0), q=None, order=None,
                 bcv=(), bcl=(), bcf=(), geometry=None, mattype=None, name=""):
        self.dim, self.dof, self.p, self.N, self.C, self.periodic = dim, dof, p, N, C, periodic
        self.limits, self.q, self.order, self.bcv, self.bcl = limits, q, order, list(bcv), list(bcl)
        self.bcf = list(bcf)          # faces (axis, side) visited by the boundary-integral pass (IGASetBoundaryForm)
        self.geometry = geometry      # None | ("perturbed", amp) | ("arrays", X, W)
        self.mattype = mattype
        self.name = name

    def geometry_arrays(self):
        if self.geometry is None:
            return None, None
        if self.geometry[0] == "perturbed":
            from tests.geomutil import perturbed_identity
            assert not isinstance(self.p, (list, tuple))
            return perturbed_identity(self.dim, self.p, [_per_axis(self.N, d) for d in range(self.dim)], self.geometry[1]), None
        return self.geometry[1], self.geometry[2]

    def _apply(self, o, uniform, rule, order, bv, bl, geom, bf=None):
        for d in range(self.dim):
            uniform(d, _per_axis(self.p, d), _per_axis(self.N, d), self.limits[0], self.limits[1], _per_axis(self.C, d),
                    bool(_per_axis(self.periodic, d)))
            if self.q is not None:
                rule(d, _per_axis(self.q, d))
        if self.order is not None:
            order(self.order)
        for (a, s, f, v) in self.bcv:
            bv(a, s, f, v)
        for (a, s, f, v) in self.bcl:
            bl(a, s, f, v)
        for (a, s) in self.bcf:
            bf(a, s, True)
        X, W = self.geometry_arrays()
        if X is not None:
            geom(X, W)

    def oracle(self, native=False):
        o = OracleIGA(self.dim, self.dof, native=native)
        self._apply(o, o.axis_uniform, o.rule_size, o.order, o.boundary_value, o.boundary_load, o.geometry, o.boundary_form)
        return o

    def product(self, rank=0, size=1, nccl=None, device=0, setup=True):
        import petiga_b200 as pb
        g = pb.IGA(self.dim, self.dof, rank=rank, size=size, nccl=nccl, device=device)
        self._apply(g, g.AxisInitUniform, g.SetRuleSize, g.SetOrder, g.SetBoundaryValue, g.SetBoundaryLoad, g.SetGeometryArrays,
                    g.SetBoundaryForm)
        if self.mattype:
            g.SetMatType(self.mattype)
        if setup:
            g.SetUp()
        return g


def rel_frobenius(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)


def oracle_to_layout(vals, rowptr_blocks, dof, baij):
    """Oracle values [nnzb, dof, dof] (row-major blocks) -> the product's host view (Mat.values()):
    BAIJ: same [nnzb, dof, dof]; AIJ: scalar CSR order [row a, comp i][col b, comp j]."""
    if baij or dof == 1:
        return vals if dof > 1 else vals.reshape(-1)
    out = []
    for r in range(len(rowptr_blocks) - 1):
        blk = vals[rowptr_blocks[r]:rowptr_blocks[r + 1]]       # [W, i, j]
        out.append(blk.transpose(1, 0, 2).reshape(-1))          # [i][W][j]
    return np.concatenate(out) if out else np.zeros(0)


def state_vectors(n, seed=20261017):
    """SURVEY 8d cfg 5 synthetic state: U = cbar + 0.05(2r-1), V = 2r-1 from one seeded stream."""
    rng = np.random.default_rng(seed)
    r = rng.random(2 * n)
    return 0.63 + 0.05 * (2 * r[:n] - 1), 2 * r[n:] - 1
