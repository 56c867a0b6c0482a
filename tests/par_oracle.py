"""Multi-threaded use of the CPU oracle for the larger parity cases (test infrastructure).

The oracle's element loop is single-threaded like the reference's.  To compare against it at sizes closer to BASELINE's,
T emulated MPI ranks of the reference's own box partition run in T threads (ctypes releases the GIL), each into its own
zero-initialised arrays (calloc: only the pages a rank touches become resident), and the per-rank results are summed --
what MatAssemblyEnd/VecAssemblyEnd produce.  Summation order differs from a serial run only in the ghost rows, far below
the 1e-12 bar."""
import os
import threading

import numpy as np


def host_threads(limit=32):
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(limit, n))


def assemble_parallel(case, slot, form, params=(), T=None, shift=0.0, V=None, t=0.0, U=None):
    """Returns (rowptr, colidx, vals[nnzb,dof,dof] or None, rhs[n,dof] or None) in the numbering of a ONE-rank run.

    The T emulated ranks only split the element loop; rows are numbered as on one rank (size = 1 pattern), so the result is
    directly comparable with a single-GPU assembly."""
    T = T or host_threads()
    base = case.oracle()
    base.setup()
    rp, ci, _ = base.pattern(1)
    n, nnz, dof = len(rp) - 1, len(ci), case.dof
    want_mat = slot in ("MATRIX", "SYSTEM", "JACOBIAN", "IJACOBIAN")
    want_vec = slot in ("VECTOR", "SYSTEM", "FUNCTION", "IFUNCTION")
    inf = base.info()
    nel = [inf["nel"][d] for d in range(case.dim)]
    # split the slowest axis into T slabs of elements; each thread assembles its slab through the oracle's own per-rank loop
    # by emulating a (1,..,T) processor grid -- which is exactly IGA_Partition's answer only for some T, so instead of the
    # oracle's rank emulation (whose numbering depends on T) every thread gets its own oracle object restricted to a slab
    # via oiga_assemble_range (element index range of the one-rank loop)
    from oracle.oracle import SLOT, FORM, _d
    total = int(np.prod(nel))
    T = max(1, min(T, total))
    cuts = [total * k // T for k in range(T + 1)]
    prm = np.ascontiguousarray(list(params) + [0.0] * 4, dtype=np.float64)
    Uc = None if U is None else np.ascontiguousarray(U, dtype=np.float64)
    Vc = None if V is None else np.ascontiguousarray(V, dtype=np.float64)
    outs, errs = [None] * T, []

    def work(k):
        try:
            o = base if k == 0 else case.oracle()
            if k:
                o.setup()
                o._pat_borrow = base._pat[1]          # the pattern is read-only during assembly: share it
            vals = np.zeros((nnz, dof, dof)) if want_mat else None
            rhs = np.zeros((n, dof)) if want_vec else None
            rc = o.L.oiga_assemble_range(o.h, cuts[k], cuts[k + 1], SLOT[slot], FORM[form], _d(prm), shift, _d(Vc), t, _d(Uc),
                                         base._pat[1], _d(vals), _d(rhs))
            assert rc == 0, rc
            outs[k] = (vals, rhs)
        except Exception as e:       # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(T)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    if errs:
        raise errs[0]
    vals = rhs = None
    if want_mat:
        vals = outs[0][0]
        for k in range(1, T):
            vals += outs[k][0]
    if want_vec:
        rhs = outs[0][1]
        for k in range(1, T):
            rhs += outs[k][1]
    return rp, ci, vals, rhs
