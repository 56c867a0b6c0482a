"""Build the in-tree native libraries (nvcc, sm_100a only; cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_dir():
    return os.path.join(_HERE, "lib")


def build(force=False, verbose=False):
    """make -C petiga_b200/csrc: libpetiga_cuda.so + libpetiga_host.so under petiga_b200/lib/."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", src, "clean"])
    out = subprocess.run(["make", "-C", src, "-j8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode:
        print(out.stdout)
    if out.returncode:
        raise RuntimeError("building libpetiga_cuda failed")
    for name in ("libpetiga_cuda.so", "libpetiga_host.so"):
        if not os.path.exists(os.path.join(lib_dir(), name)):
            raise RuntimeError(name + " missing after build")
    return lib_dir()
