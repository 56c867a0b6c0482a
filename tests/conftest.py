import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def pytest_sessionstart(session):
    """The tests load the in-tree shared libraries; build them once if a fresh checkout has none (nvcc cross-compiles
    without a GPU, a few minutes).  The oracle builds itself on first use (oracle/oracle.py)."""
    import petiga_b200
    if os.environ.get("PETIGA_NO_AUTOBUILD") != "1":
        try:
            petiga_b200.build(verbose=False)     # make: a no-op when the libraries are newer than every source (ADVICE r1:
        except Exception as e:                   # an existence check let tests run against stale binaries)
            print("petiga_b200 auto-build failed: %s" % e)
