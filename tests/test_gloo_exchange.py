"""world_size-2 (and 4) gloo test of the multi-rank HOST logic on CPU: the exchange lists and the unified
owned+ghost row layout of petiga_layout_* are driven with real messages.  The per-rank element contributions come
from the oracle (one emulated rank each); the product's layout decides where they live locally, what is sent to
whom and where received rows are added -- exactly what pc_comm.cu does with ncclSend/ncclRecv on the GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.common import Case


def _worker(rank, world, port, case_kw, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = Case(**case_kw)
        dof = case.dof
        o = case.oracle()
        o.setup()
        rp, ci, rs = o.pattern(world)
        n = len(rp) - 1
        Kg, Fg = o.assemble("SYSTEM", "MASS" if dof > 1 else "POISSON", size=world)          # the assembled answer
        Kr, Fr = np.zeros((len(ci), dof, dof)), np.zeros((n, dof))
        o.assemble_rank("SYSTEM", "MASS" if dof > 1 else "POISSON", [], world, rank, Kr, Fr)   # this rank's element loop only
        g = case.product(rank=rank, size=world)                                              # host logic only: no GPU touched
        L = g.layout()
        sz = L.sizes()
        lg, lr, rb = L.lgmap(), L.localrow(), L.rowbase()
        # local unified buffers (owned rows first, then ghost rows), filled from this rank's contributions
        vals = np.zeros((sz["nnz_loc"], dof, dof))
        vec = np.zeros((sz["nloc"], dof))
        seen = set()
        for gnode, row in zip(lg, lr):
            if row in seen:
                continue
            seen.add(row)
            assert rb[row + 1] - rb[row] == rp[gnode + 1] - rp[gnode]          # ghost rows are laid out like the owner's row
            vals[rb[row]:rb[row + 1]] = Kr[rp[gnode]:rp[gnode + 1]]
            vec[row] = Fr[gnode]
        # everything this rank integrated must sit in rows it can see
        assert np.isclose(np.abs(vals).sum(), np.abs(Kr).sum())
        send, recv = L.exchange(0), L.exchange(1)
        reqs, bufs = [], []
        for (peer, first, nrows, nblocks) in send:
            b0 = rb[first]
            tm = torch.from_numpy(vals[b0:b0 + nblocks].copy().reshape(-1))
            tv = torch.from_numpy(vec[first:first + nrows].copy().reshape(-1))
            reqs += [dist.isend(tm, int(peer), tag=1), dist.isend(tv, int(peer), tag=2)]
            bufs += [tm, tv]
        incoming = []
        for i, (peer, _, nrows, nblocks) in enumerate(recv):
            tm = torch.zeros(int(nblocks) * dof * dof, dtype=torch.float64)
            tv = torch.zeros(int(nrows) * dof, dtype=torch.float64)
            reqs += [dist.irecv(tm, int(peer), tag=1), dist.irecv(tv, int(peer), tag=2)]
            incoming.append((i, int(nrows), tm, tv))
        for r in reqs:
            r.wait()
        for i, nrows, tm, tv in incoming:
            rows = L.recv_rows(i, nrows)
            m = tm.numpy().reshape(-1, dof, dof)
            v = tv.numpy().reshape(-1, dof)
            off = 0
            for t, row in enumerate(rows):
                w = rb[row + 1] - rb[row]
                vals[rb[row]:rb[row + 1]] += m[off:off + w]
                off += w
                vec[row] += v[t]
            assert off == len(m)
        r0, r1 = rs[rank], rs[rank + 1]
        nown = sz["nown"]
        assert nown == r1 - r0
        ok = np.allclose(vals[:rb[nown]], Kg[rp[r0]:rp[r1]], rtol=1e-13, atol=1e-15) and np.allclose(vec[:nown], Fg[r0:r1], rtol=1e-13, atol=1e-15)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case_kw", [
    (2, dict(dim=3, p=2, N=6, bcv=[(d, s, 0, 1.0) for d in range(3) for s in range(2)])),
    (2, dict(dim=2, dof=2, p=2, N=(12, 10), periodic=(True, False))),
    (4, dict(dim=2, p=3, N=(9, 10), bcv=[(0, 0, 0, 2.0)])),
])
def test_ghost_row_exchange_over_gloo(world, case_kw):
    import random
    port = 29600 + random.randint(0, 300)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, case_kw, out), nprocs=world, join=True)
    assert all(out.get(r, False) for r in range(world)), dict(out)
