"""Committed golden assembly vectors (tests/golden/assembly_cases.npz, written by tests/golden/make_assembly_golden.py):
the oracle must keep reproducing them (CPU), and the device path must match them through the C ABI (GPU)."""
import os

import numpy as np
import pytest

from tests.common import Case, oracle_to_layout, rel_frobenius

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_assembly_golden", os.path.join(GOLD, "make_assembly_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases, state_vectors = mod.cases()
    return mod, cases, state_vectors, np.load(os.path.join(GOLD, "assembly_cases.npz"))


_MOD, CASES, _SV, BLOB = load()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    case, slot, form, prm, state = CASES[name]
    rp, ci, K, F = _MOD.compute(case, slot, form, prm, state, _SV)
    assert np.array_equal(rp, BLOB[name + "/rowptr"]) and np.array_equal(ci, BLOB[name + "/colidx"])      # bit-exact pattern
    if K is not None:
        assert rel_frobenius(K, BLOB[name + "/values"]) <= 1e-14
    if F is not None:
        assert rel_frobenius(F, BLOB[name + "/rhs"]) <= 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_matches_golden(name):
    from tests.gpu_common import run_product
    case, slot, form, prm, state = CASES[name]
    rp, ci = BLOB[name + "/rowptr"], BLOB[name + "/colidx"]
    U = V = None
    if state:
        U, V = _SV((len(rp) - 1) * case.dof)
    res = run_product(case, slot, form, prm, U=U, V=V if slot in ("IFUNCTION", "IJACOBIAN") else None, shift=1e3)
    if name + "/values" in BLOB.files:
        if res["baij"] or case.dof == 1:
            assert np.array_equal(res["rowptr"], rp) and np.array_equal(res["colidx"], ci)
        exp = oracle_to_layout(BLOB[name + "/values"], rp, case.dof, res["baij"])
        assert rel_frobenius(res["values"], exp) <= 1e-12
    if name + "/rhs" in BLOB.files:
        assert rel_frobenius(res["rhs"], BLOB[name + "/rhs"].reshape(-1)) <= 1e-12


def test_oracle_reproduces_multirank_golden():
    """tests/golden/multirank_cases.npz (what bench.py's "parity" key and the N-rank runs are compared with) is the oracle's
    one-rank output; a changed oracle must fail here before it silently changes what the GPUs are checked against."""
    from petiga_b200.cases import state_vectors
    from petiga_b200.parity import GOLDEN, golden_cases
    gold = np.load(GOLDEN)
    for name, (pc, slot, form, params, state, shift) in golden_cases().items():
        case = Case.__new__(Case)
        case.__dict__.update(pc.__dict__)
        o = case.oracle()
        o.setup()
        rp, ci, _ = o.pattern(1)
        assert np.array_equal(rp, gold[name + "/rowptr"]) and np.array_equal(ci, gold[name + "/colidx"])
        U = V = None
        if state:
            U, V = state_vectors((len(rp) - 1) * case.dof)
        K, F = o.assemble(slot, form, params, size=1, shift=shift, U=U, V=V)
        if K is not None:
            assert rel_frobenius(K, gold[name + "/K"]) <= 1e-14
        if F is not None:
            assert rel_frobenius(F, gold[name + "/F"]) <= 1e-14
