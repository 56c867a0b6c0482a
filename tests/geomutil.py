"""Geometry fixtures shared by the oracle and GPU tests (pure numpy; no reference code)."""
import numpy as np


from petiga_b200.cases import greville, uniform_knots, perturbed_identity  # noqa: F401


def _insert_knot_1d(U, p, Pw, u):
    """Boehm knot insertion on homogeneous control points Pw[n, c] (textbook A5.1 for one knot)."""
    k = np.searchsorted(U, u, side="right") - 1
    n = len(Pw)
    Q = np.zeros((n + 1, Pw.shape[1]))
    Q[:k - p + 1] = Pw[:k - p + 1]
    Q[k + 1:] = Pw[k:]
    for i in range(k - p + 1, k + 1):
        a = (u - U[i]) / (U[i + p] - U[i])
        Q[i] = a * Pw[i] + (1 - a) * Pw[i - 1]
    return np.insert(U, k + 1, u), Q


def refine_annulus(cls, N=(4, 4), p=2, dof=1, height=None):
    """Quarter annulus of test/IGAGeometryMap.c:18-32 knot-refined to N elements per axis (exact NURBS).
    Returns (iga object of class `cls` with axes+geometry set, X, W).  dim = 2, or 3 when height is given."""
    s2 = np.sqrt(2.0)
    PX = np.array([[1.0, 1.0, 0.0], [1.5, 1.5, 0.0], [2.0, 2.0, 0.0]])
    PY = np.array([[0.0, 1.0, 1.0], [0.0, 1.5, 1.5], [0.0, 2.0, 2.0]])
    PW = np.array([[1.0, s2 / 2, 1.0]] * 3)
    # homogeneous net [i (radial)][j (angular)][x*w, y*w, w]
    net = np.stack([PX * PW, PY * PW, PW], axis=-1)
    U0 = np.array([0, 0, 0, 1, 1, 1.0])
    U1 = U0.copy()
    for u in np.arange(1, N[0]) / N[0]:
        U0, flat = _insert_knot_1d(U0, 2, net.reshape(net.shape[0], -1), u)
        net = flat.reshape(-1, net.shape[1], 3)
    for v in np.arange(1, N[1]) / N[1]:
        t = net.transpose(1, 0, 2)
        U1, flat = _insert_knot_1d(U1, 2, t.reshape(t.shape[0], -1), v)
        net = flat.reshape(-1, t.shape[1], 3).transpose(1, 0, 2)
    W = net[..., 2]
    XY = net[..., :2] / W[..., None]
    dim = 2 if height is None else 3
    o = cls(dim, dof)
    o.axis_knots(0, 2, U0)
    o.axis_knots(1, 2, U1)
    if dim == 2:
        X = XY.transpose(1, 0, 2).copy()          # natural order [j][i][c]
        Wn = W.T.copy()
    else:
        Nz, hz = height
        U2 = uniform_knots(2, Nz)
        o.axis_knots(2, 2, U2)
        gz = greville(U2, 2) * hz
        X = np.zeros((len(gz),) + XY.transpose(1, 0, 2).shape[:2] + (3,))
        X[..., :2] = XY.transpose(1, 0, 2)[None]
        X[..., 2] = gz[:, None, None]
        Wn = np.broadcast_to(W.T[None], X.shape[:3]).copy()
    o.geometry(X, Wn)
    return o, X, Wn
