"""ORACLE — TEST INFRASTRUCTURE ONLY.

Python handle on oracle/liboracle.so, the CPU restatement of the reference's Java operators (see oracle/oracle_common.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package; the
product package presto_b200 never does.  It reuses presto_b200's ctypes marshalling (`Backend`) with the prefix ``or_``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from presto_b200 import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_backend = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "presto_gpu.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return LIB_PATH


def backend() -> abi.Backend:
    global _backend
    if _backend is None:
        build()
        lib = C.CDLL(LIB_PATH)
        _backend = abi.Backend(lib, "or_", has_device=False)
        lib.or_group_by_hash_get_capacity.restype = C.c_int32
        lib.or_group_by_hash_get_capacity.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        lib.or_array_size.restype = C.c_int32
        lib.or_array_size.argtypes = [C.c_int64, C.POINTER(C.c_int32)]
    return _backend
