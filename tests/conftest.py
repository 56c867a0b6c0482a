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
    lib = os.path.join(petiga_b200.lib_dir(), "libpetiga_host.so")
    if not os.path.exists(lib) and os.environ.get("PETIGA_NO_AUTOBUILD") != "1":
        try:
            petiga_b200.build(verbose=False)
        except Exception as e:      # leave the failure to the tests that need the library
            print("petiga_b200 auto-build failed: %s" % e)
