"""Geometry/vector files in the reference's PETSc binary format (src/petigaio.c): the host mirror's
IGARead/IGAWrite and the oracle's numpy reader against the golden files of tests/golden/ (written from the format
specification by tests/golden/make_golden.py, independent of both)."""
import os

import numpy as np
import pytest

from oracle.oracle import OracleIGA, oracle_from_file, read_iga_file
from tests.geomutil import perturbed_identity, refine_annulus, uniform_knots

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ANNULUS = os.path.join(GOLD, "annulus_4x4.dat")
CUBE = os.path.join(GOLD, "cube_perturbed.dat")
VEC = os.path.join(GOLD, "vec_natural.dat")


class Rec:
    def __init__(self, dim, dof):
        self.axes = {}

    def axis_knots(self, d, p, U):
        self.axes[d] = (p, np.asarray(U, float))

    def geometry(self, X, W):
        self.X, self.W = X, W


def test_golden_files_match_their_generator(tmp_path):
    """The committed fixtures are what make_golden.py writes today (guards against silent drift)."""
    import shutil
    import subprocess
    import sys
    work = tmp_path / "golden"
    work.mkdir()
    shutil.copy(os.path.join(GOLD, "make_golden.py"), work / "make_golden.py")
    # the generator writes next to itself and imports tests.geomutil from the repo root two levels up
    src = open(work / "make_golden.py").read().replace("ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))",
                                                       "ROOT = %r" % os.path.dirname(os.path.dirname(GOLD)))
    open(work / "make_golden.py", "w").write(src)
    subprocess.check_call([sys.executable, str(work / "make_golden.py")])
    for name in ("annulus_4x4.dat", "cube_perturbed.dat", "vec_natural.dat"):
        assert open(work / name, "rb").read() == open(os.path.join(GOLD, name), "rb").read(), name


def test_oracle_reads_annulus():
    axes, nsd, X, W, rational = read_iga_file(ANNULUS)
    r, Xe, We = refine_annulus(Rec, N=(4, 4))
    assert nsd == 2 and rational and len(axes) == 2
    for d in range(2):
        assert axes[d][0] == r.axes[d][0] and np.array_equal(axes[d][1], r.axes[d][1])
    assert np.array_equal(W, We) and np.allclose(X, Xe, rtol=4e-16, atol=0)
    o = oracle_from_file(ANNULUS)
    o.setup()
    K, F = o.assemble("SYSTEM", "MASS")
    assert abs(K.sum() - 3 * np.pi / 4) < 1e-6        # test/IGAGeometryMap.c: area of the quarter annulus


def test_oracle_reads_cube():
    axes, nsd, X, W, rational = read_iga_file(CUBE)
    assert nsd == 3 and not rational and [a[0] for a in axes] == [2, 2, 2]
    assert np.array_equal(axes[0][1], uniform_knots(2, 3))
    assert np.array_equal(X, perturbed_identity(3, 2, 3, 0.05)) and np.all(W == 1.0)


def test_mirror_reads_what_the_oracle_reads():
    import petiga_b200 as pb
    for path, dim in ((ANNULUS, 2), (CUBE, 3)):
        axes, nsd, X, W, rational = read_iga_file(path)
        g = pb.IGA()
        g.Read(path)
        assert g.dim == dim
        sizes, nsd_g, rat_g, Xg, Wg = g.GetGeometryArrays()
        assert nsd_g == nsd and rat_g == rational and sizes[:dim] == list(X.shape[:dim][::-1])
        assert np.array_equal(Xg, X.reshape(-1, nsd)) and np.array_equal(Wg, W.reshape(-1))    # same de-homogenisation, bit for bit
        g.SetDof(1)
        g.SetUp()
        inf = g.info()
        o = oracle_from_file(path)
        oi = o.setup()
        for key in ("p", "m", "nnp", "nel", "geom_size"):
            assert inf[key] == oi[key], key
        g.Read(path)            # reading twice, as test/IGAInputOutput.c:41-44 does
        g.SetDof(1)
        g.SetUp()


def test_mirror_write_roundtrip(tmp_path):
    import petiga_b200 as pb
    # non-rational file: byte-exact
    g = pb.IGA()
    g.Read(CUBE)
    g.SetDof(1)
    g.SetUp()
    out = str(tmp_path / "cube.dat")
    g.Write(out)
    g.Write(out)                # test/IGAInputOutput.c:36-37 writes twice
    assert open(out, "rb").read() == open(CUBE, "rb").read()
    # rational file: (w x)/w*w may differ from w x in the last bit; the reloaded geometry must be identical
    g = pb.IGA()
    g.Read(ANNULUS)
    g.SetDof(1)
    g.SetUp()
    out = str(tmp_path / "annulus.dat")
    g.Write(out)
    a, b = read_iga_file(out), read_iga_file(ANNULUS)
    assert np.array_equal(a[3], b[3]) and np.allclose(a[2], b[2], rtol=4e-16, atol=0) and a[4] == b[4]
    raw_a, raw_b = open(out, "rb").read(), open(ANNULUS, "rb").read()
    assert len(raw_a) == len(raw_b) and raw_a[:12 + 2 * (8 + 8 * 10)] == raw_b[:12 + 2 * (8 + 8 * 10)]   # header + knots byte-exact
    # an identity-geometry IGA writes a file without the geometry record (info == 0)
    g = pb.IGA(2, 1)
    g.AxisInitUniform(0, 2, 4)
    g.AxisInitUniform(1, 3, 5)
    g.SetUp()
    out = str(tmp_path / "plain.dat")
    g.Write(out)
    axes, nsd, X, W, rational = read_iga_file(out)
    assert nsd == 0 and X is None and [a[0] for a in axes] == [2, 3]
    assert np.array_equal(axes[1][1], uniform_knots(3, 5))


def test_mirror_read_errors(tmp_path):
    import petiga_b200 as pb
    g = pb.IGA()
    with pytest.raises(pb.IGAError) as e:
        g.Read(str(tmp_path / "missing.dat"))
    assert e.value.code == 65                                   # PETSC_ERR_FILE_OPEN
    with pytest.raises(pb.IGAError) as e:
        g.Read(VEC)                                             # a Vec file is "Not an IGA in file" (petigaio.c:32)
    assert e.value.code == 62                                   # PETSC_ERR_ARG_WRONG
    bad = tmp_path / "trunc.dat"
    bad.write_bytes(open(ANNULUS, "rb").read()[:300])
    with pytest.raises(pb.IGAError) as e:
        g.Read(str(bad))
    assert e.value.code == 66                                   # PETSC_ERR_FILE_READ


@pytest.mark.gpu
def test_gpu_assembly_on_loaded_geometry_and_vec_io(tmp_path):
    """IGARead -> device assembly equals the oracle on the same file; IGAReadVec/IGAWriteVec round trip."""
    import petiga_b200 as pb
    from tests.common import rel_frobenius
    o = oracle_from_file(ANNULUS)
    o.setup()
    Ko, Fo = o.assemble("SYSTEM", "L2PROJECTION", [6])
    g = pb.IGA()
    g.Read(ANNULUS)
    g.SetDof(1)
    g.SetUp()
    g.SetForm("SYSTEM", "L2PROJECTION", [6])
    A, B = g.CreateMat(), g.CreateVec()
    g.ComputeSystem(A, B)
    assert rel_frobenius(A.values(), Ko.reshape(-1)) <= 1e-12 and rel_frobenius(B.get(), Fo.reshape(-1)) <= 1e-12
    g = pb.IGA()
    g.Read(CUBE)
    g.SetDof(2)
    g.SetUp()
    v = g.CreateVec()
    g.ReadVec(v, VEC)
    assert np.array_equal(v.get(), np.arange(v.size) / 7.0)     # one rank: natural == global
    out = str(tmp_path / "v.dat")
    g.WriteVec(v, out)
    assert open(out, "rb").read() == open(VEC, "rb").read()


@pytest.mark.parametrize("dim,N,size", [(2, (7, 5), 2), (2, (8, 8), 4), (3, (5, 4, 6), 8), (3, (6, 6, 6), 3)])
def test_natural_to_global_map_of_readvec(dim, N, size):
    """IGAReadVec places entry `natural` of the file into owned slot a: check that map against the rank's own numbering
    (lgmap of the ghost box, itself bit-exact vs the oracle in test_host_layout.py) for every rank of several partitions."""
    import petiga_b200 as pb
    seen = []
    start = 0
    for rank in range(size):
        g = pb.IGA(dim, 1, rank=rank, size=size)
        for d in range(dim):
            g.AxisInitUniform(d, 2, N[d])
        g.SetUp()
        inf = g.info()
        nat = g.GetOwnedNaturalIndices()
        gs, gw = inf["node_gstart"], inf["node_gwidth"]
        ls, lw = inf["node_lstart"], inf["node_lwidth"]
        nnp = inf["nnp"]
        lg = g.lgmap().reshape(gw[2], gw[1], gw[0])
        for k in range(ls[2], ls[2] + lw[2]):
            for j in range(ls[1], ls[1] + lw[1]):
                for i in range(ls[0], ls[0] + lw[0]):
                    glob = lg[k - gs[2], j - gs[1], i - gs[0]]
                    assert nat[glob - start] == i + nnp[0] * (j + nnp[1] * k)
        start += len(nat)
        seen.extend(nat.tolist())
    assert sorted(seen) == list(range(int(np.prod([inf["nnp"][d] for d in range(3)]))))      # every natural entry lands exactly once
