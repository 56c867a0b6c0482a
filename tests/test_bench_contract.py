"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints exactly ONE JSON line with the
contract's keys, also under torchrun with two ranks (rank 0 alone works and prints; the other rank exits 0)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def check(out):
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "assembled_Mnnz_per_s" and d["unit"] == "Mnnz/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mnnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("demo/Poisson3D p=3 C2 128^3")


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--cpu-threads", "2"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    check(r.stdout)


def test_reference_arm_under_torchrun_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29577",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3", "--cpu-threads", "2"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    check(r.stdout)


def test_product_code_does_not_reach_into_oracle_or_tests():
    """The oracle is test infrastructure: the product package must not import it at all, and bench.py may execute it only in the
    cpu_baseline / reference-arm legs (function-local imports of oracle.cpu_reference), never at module level and never `tests`."""
    import ast
    import glob
    for path in glob.glob(os.path.join(ROOT, "petiga_b200", "*.py")):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] in ("oracle", "tests") for n in names), (path, names)
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for node in tree.body:      # module level
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            mod = node.module if isinstance(node, ast.ImportFrom) else node.names[0].name
            assert (mod or "").split(".")[0] not in ("oracle", "tests"), mod
    allowed = {"cpu_reference_run", "main"}       # main(): only the cfg-1 full-size CPU figure inside the cpu_baseline block
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] in ("oracle", "tests"):
                assert node.module.startswith("oracle.cpu_reference") and fn.name in allowed, (fn.name, node.module)
