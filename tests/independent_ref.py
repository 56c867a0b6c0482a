"""An entry-level reference that shares NO code with the oracle or the product.

The reference's tests hold no matrix-entry vectors ("parity unpinned" at the value-array level), so the only
way to pin entries independently is mathematics: on the identity map the Poisson / mass matrices are Kronecker sums of
1-D B-spline matrices.  Here the 1-D matrices come from scipy's BSpline evaluator and numpy's Gauss-Legendre rule
(different B-spline recurrence, different quadrature table than src/petigabsb.f90.in / src/petigarule.c), and the
uniform-knot closed forms (h/6 [1 4 1], 1/h [-1 2 -1], h/120 [1 26 66 26 1], ...) pin them in turn."""
import numpy as np
from scipy.interpolate import BSpline

from petiga_b200.cases import uniform_knots


def matrices_1d(p, N, lo=0.0, hi=1.0, q=None):
    """(M, K) dense [nnp, nnp]: mass and stiffness of the open uniform C^{p-1} B-spline space with N elements."""
    U = uniform_knots(p, N, lo=lo, hi=hi)
    n = len(U) - p - 1
    q = q or p + 1
    x, w = np.polynomial.legendre.leggauss(q)
    breaks = np.unique(U)
    M, K = np.zeros((n, n)), np.zeros((n, n))
    eye = np.eye(n)
    for e in range(len(breaks) - 1):
        a, b = breaks[e], breaks[e + 1]
        J = (b - a) / 2
        u = (x + 1) * J + a
        B0 = np.stack([BSpline(U, eye[i], p, extrapolate=False)(u) for i in range(n)], axis=1)
        B1 = np.stack([BSpline(U, eye[i], p, extrapolate=False).derivative()(u) for i in range(n)], axis=1)
        B0, B1 = np.nan_to_num(B0), np.nan_to_num(B1)
        M += B0.T @ (B0 * (w * J)[:, None])
        K += B1.T @ (B1 * (w * J)[:, None])
    return M, K


def kron_axes(mats):
    """Kronecker product with axis 0 fastest in the node numbering (node = i + n0*(j + n1*k))."""
    out = np.array([[1.0]])
    for m in mats:          # axis 0 first -> innermost
        out = np.kron(m, out)
    return out


def poisson_matrix(dim, p, N):
    MK = [matrices_1d(p, N) for _ in range(dim)]
    A = 0
    for d in range(dim):
        A = A + kron_axes([MK[e][1] if e == d else MK[e][0] for e in range(dim)])
    return A


def mass_matrix(dim, p, N):
    return kron_axes([matrices_1d(p, N)[0] for _ in range(dim)])


# closed-form interior stencils of uniform B-splines (knot spacing h): mass / h and stiffness * h
CLOSED_FORM = {
    1: (np.array([1, 4, 1]) / 6.0, np.array([-1, 2, -1]) / 1.0),
    2: (np.array([1, 26, 66, 26, 1]) / 120.0, np.array([-1, -2, 6, -2, -1]) / 6.0),
    3: (np.array([1, 120, 1191, 2416, 1191, 120, 1]) / 5040.0, np.array([-1, -24, -15, 80, -15, -24, -1]) / 120.0),
}


def csr_to_dense(rowptr, colidx, vals, n):
    A = np.zeros((n, n))
    for r in range(n):
        A[r, colidx[rowptr[r]:rowptr[r + 1]]] = vals[rowptr[r]:rowptr[r + 1]]
    return A
