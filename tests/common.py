"""Shared test helpers: build the oracle and the product from one case description."""
import numpy as np

from oracle.oracle import OracleIGA
from petiga_b200.cases import Case as _ProductCase, state_vectors  # noqa: F401


class Case(_ProductCase):
    """The product-side case description (petiga_b200/cases.py) plus the oracle construction (tests only)."""

    def oracle(self, native=False):
        o = OracleIGA(self.dim, self.dof, native=native)
        self._apply(o, o.axis_uniform, o.rule_size, o.order, o.boundary_value, o.boundary_load, o.geometry, o.boundary_form)
        return o


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
