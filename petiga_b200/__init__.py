"""petiga_b200 -- B200-native element assembly behind PetIGA's IGA/IGAForm API.

The product is the C-ABI library ``libpetiga_cuda.so`` (CUDA for sm_100a, ``include/petiga_cuda.h``) plus the
PETSc-free host mirror ``libpetiga_host.so`` (``include/petiga_host.h``).  This Python package is only a ctypes
binding over those two libraries for tests and benchmarks; it contains no numerics and no CPU fallback: every
compute call goes through the CUDA library and fails loudly when it (or a GPU) is missing.
"""
from .build import build, lib_dir  # noqa: F401
from .iga import IGA, IGAError, Mat, Vec, FORMS, iga_partition, load_host, load_cuda  # noqa: F401
from .cases import Case, baseline_config, state_vectors  # noqa: F401
