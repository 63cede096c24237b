"""Loader of the CUDA product library.  There is no CPU fallback: a missing extension or a missing GPU is an error."""
from __future__ import annotations

import ctypes as C
import os

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libprestogpu.so")

_backend = None


def library_path() -> str:
    return LIB_PATH


def load_library() -> abi.Backend:
    """dlopen libprestogpu.so and bind every symbol include/presto_gpu.h declares (no CUDA call is made)."""
    global _backend
    if _backend is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). presto_b200 has no CPU fallback.")
        _backend = abi.Backend(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL), "pg_", has_device=True)
    return _backend


def load() -> abi.Backend:
    """The backend every operator uses by default; fails loudly when no B200 is visible."""
    b = load_library()
    n = C.c_int32()
    status = b.device_count(C.byref(n))
    if status != abi.PG_OK or n.value < 1:
        raise abi.PrestoException(abi.PG_ERR_NO_DEVICE, "no CUDA device visible: the presto_b200 operators only run on the GPU (no CPU fallback)")
    return b
