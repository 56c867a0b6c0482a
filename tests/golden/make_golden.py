#!/usr/bin/env python
"""Generate the golden geometry/vector files under tests/golden/ (committed; run from the repo root).

The reference (PETSc + MPI + gfortran) cannot run here, so the files are produced from the published on-disk format the
reference reads and writes, stated in src/petigaio.c:
  IGASave   :75-139   int32 IGA_FILE_CLASSID (1211299, include/petiga.h:394) | int32 info (bit0 geometry, bit1 property)
                      | int32 dim | per axis {int32 p, int32 m+1, float64 U[m+1]} | [int32 nsd | Vec]
  VecView (PETSc binary)  int32 VEC_FILE_CLASSID (1211214) | int32 n | float64[n]
  IGASaveGeometry :288-369  natural order (i fastest), per control point (w*x_0..w*x_{nsd-1}, w)
all big-endian (PetscBinaryWrite).  Only numpy is used -- no code of this repository -- so the files pin both
IGARead/IGAWrite of the host mirror and the oracle's reader independently.

  annulus_4x4.dat   quarter annulus of test/IGAGeometryMap.c:18-32, knot-refined to 4x4 elements (rational)
  cube_perturbed.dat  3-D p=2, 3x3x3 elements, Greville points + 0.05 prod sin(2 pi x) (W == 1, non-rational)
  vec_natural.dat   a natural-ordering Vec for the cube's node grid, dof 2, values = index/7
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def be_int(*v):
    return np.asarray(v, dtype=">i4").tobytes()


def be_real(a):
    return np.asarray(a, dtype=">f8").tobytes()


def write_iga(path, axes, X=None, W=None):
    """axes = [(p, U)]; X natural [..][nsd] or None; W natural or None (-> 1)."""
    b = be_int(1211299, 1 if X is not None else 0, len(axes))
    for p, U in axes:
        b += be_int(p, len(U)) + be_real(U)
    if X is not None:
        nsd = X.shape[-1]
        Xf = X.reshape(-1, nsd)
        Wf = np.ones(len(Xf)) if W is None else W.reshape(-1)
        Xw = np.concatenate([Xf * Wf[:, None], Wf[:, None]], axis=1)
        b += be_int(nsd) + be_int(1211214, Xw.size) + be_real(Xw.reshape(-1))
    open(path, "wb").write(b)


def main():
    from tests.geomutil import greville, perturbed_identity, refine_annulus, uniform_knots

    class Rec:   # records what refine_annulus hands to an IGA-like object
        def __init__(self, dim, dof):
            self.axes = {}

        def axis_knots(self, d, p, U):
            self.axes[d] = (p, np.asarray(U, dtype=float))

        def geometry(self, X, W):
            self.X, self.W = X, W

    r, X, W = refine_annulus(Rec, N=(4, 4))
    write_iga(os.path.join(HERE, "annulus_4x4.dat"), [r.axes[0], r.axes[1]], X, W)
    p, N = 2, 3
    U = uniform_knots(p, N)
    Xc = perturbed_identity(3, p, N, 0.05)
    write_iga(os.path.join(HERE, "cube_perturbed.dat"), [(p, U)] * 3, Xc, None)
    nn = len(greville(U, p)) ** 3 * 2
    open(os.path.join(HERE, "vec_natural.dat"), "wb").write(be_int(1211214, nn) + be_real(np.arange(nn) / 7.0))


if __name__ == "__main__":
    main()
