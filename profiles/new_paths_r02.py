#!/usr/bin/env python
"""Rough timings (wall clock around synchronous C-ABI calls, device-resident pages) of the paths added late in round 2: interned VARCHAR
group keys, VARCHAR join keys, VARCHAR predicates, the LZ4 serde.  Prints one JSON object; not a bench line (bench.py's contract is
the headline) — these are the first numbers of new kernels, for DESIGN.md."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from presto_b200 import (AGG_COUNT_STAR, AGG_SUM_DOUBLE, BIGINT, DOUBLE, MEM_DEVICE, VARCHAR, Aggregate, Block, FilterAndProjectOperator, HashAggregationOperator,  # noqa: E402
                         HashBuilderOperator, JoinBridge, LookupJoinOperator, Page, PagesSerde, Program, to_pages)
from presto_b200 import _abi as abi  # noqa: E402
from presto_b200._lib import load  # noqa: E402
from util import page_to_device  # noqa: E402

b = load()
b.check(b.set_device(0))
rng = np.random.default_rng(0)
out = {}


def timed(f, reps=3):
    f()                                  # warm-up (allocator, JIT)
    best = 1e9
    for _ in range(reps):
        b.check(b.device_synchronize()); t = time.perf_counter(); f(); b.check(b.device_synchronize())
        best = min(best, time.perf_counter() - t)
    return best * 1e3


def strings_block(words, pick):
    lens = np.array([len(w) for w in words], dtype=np.int64)
    width = int(lens.max())
    table = np.zeros((len(words), width), dtype=np.uint8)
    for i, w in enumerate(words):
        table[i, :len(w)] = np.frombuffer(w, dtype=np.uint8)
    l = lens[pick]
    offsets = np.zeros(len(pick) + 1, dtype=np.int32); np.cumsum(l, out=offsets[1:])
    mask = np.arange(width)[None, :] < l[:, None]
    data = table[pick][mask]
    return Block(abi.VARCHAR, len(pick), np.ascontiguousarray(data), None, offsets)


keep = []
n = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 20_000_000
for groups in (100, min(100_000, n // 20)):
    words = [b"customer-name-%08d-x" % i for i in range(groups)]            # 23-byte keys
    page = page_to_device(b, Page([strings_block(words, rng.integers(0, groups, n)), Block.from_numpy(DOUBLE, rng.random(n))]), keep)
    def agg():
        op = HashAggregationOperator([VARCHAR], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)], expected_groups=groups, backend=b, output_memory=MEM_DEVICE)
        res = to_pages(op, [page]); assert sum(p.position_count for p in res) == groups
        op.close()
    ms = timed(agg)
    out[f"group_by_varchar23_{'few' if groups == 100 else 'many'}_groups"] = {"rows": n, "groups": groups, "ms": round(ms, 2), "g_rows_s": round(n / ms / 1e6, 2), "bytes_per_row": 23 + 4 + 8}

# VARCHAR join: 1 M unique build keys, 20 M probe rows (70 % hits)
nb = min(1_000_000, n // 8)
words = [b"order-%09d" % i for i in range(nb)]
build = page_to_device(b, Page([strings_block(words, rng.permutation(nb)), Block.from_numpy(BIGINT, np.arange(nb))]), keep)
probe_words = words + [b"missing-%07d" % i for i in range(nb // 2)]
probe = page_to_device(b, Page([strings_block(probe_words, rng.integers(0, len(probe_words), n)), Block.from_numpy(DOUBLE, rng.random(n))]), keep)
def join():
    bridge = JoinBridge(backend=b)
    hb = HashBuilderOperator(bridge, [VARCHAR, BIGINT], [0], [1], backend=b)
    to_pages(hb, [build])
    pr = LookupJoinOperator(bridge, [VARCHAR, DOUBLE], [0], [1], backend=b, output_memory=MEM_DEVICE)
    res = to_pages(pr, [probe]); rows = sum(p.position_count for p in res)
    pr.close(); hb.close(); bridge.close()
    return rows
ms = timed(join)
out["join_varchar15"] = {"build_rows": nb, "probe_rows": n, "ms": round(ms, 2), "probe_g_rows_s": round(n / ms / 1e6, 2), "output_rows": join()}

# VARCHAR predicate: c_mktsegment = 'BUILDING' over 20 M rows, key passes through
segs = [b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"HOUSEHOLD", b"MACHINERY"]
cust = page_to_device(b, Page([strings_block(segs, rng.integers(0, 5, n)), Block.from_numpy(BIGINT, np.arange(n))]), keep)
prog = Program([VARCHAR, BIGINT]); prog.set_filter(prog.input(0) == b"BUILDING"); prog.project(prog.input(1))
def filt():
    fp = FilterAndProjectOperator([VARCHAR, BIGINT], prog, backend=b, output_memory=MEM_DEVICE)       # (to_pages finishes the operator: a fresh one per run)
    fp.add_input(cust); r = fp.get_output(copy=False); fp.close()
ms = timed(filt)
out["filter_mktsegment_eq_building"] = {"rows": n, "ms": round(ms, 2), "g_rows_s": round(n / ms / 1e6, 2)}

# LZ4 serde: a 64 MB compressible page (BIGINT sequence // 3 + small-domain INTEGER) and its decode
m = min(1_000_000, n // 8)       # (the one-warp decoder runs at ~24 MB/s on these 8-byte sequences: keep the page small)
page = page_to_device(b, Page([Block.from_numpy(BIGINT, np.arange(m, dtype=np.int64) // 3), Block.from_numpy(BIGINT, rng.integers(0, 50, m))]), keep)
lz = PagesSerde(compression=abi.COMPRESSION_LZ4, backend=b, output_memory=MEM_DEVICE)
plain = PagesSerde(backend=b, output_memory=MEM_DEVICE)
ms_plain = timed(lambda: plain.serialize(page, copy=False))
ms_lz = timed(lambda: lz.serialize(page, copy=False))
addr, size = lz.serialize(page, copy=False)
wire = np.empty(size, dtype=np.uint8); b.check(b.memcpy_d2h(wire.ctypes.data, addr, size))
import ctypes as C
dptr = C.c_void_p(); b.check(b.device_alloc(size, C.byref(dptr))); b.check(b.memcpy_h2d(dptr, wire.ctypes.data, size))
reader = PagesSerde(backend=b, output_memory=MEM_DEVICE)
ms_dec = timed(lambda: reader.deserialize(dptr.value, [BIGINT, BIGINT], size=int(size), memory=MEM_DEVICE, copy=False))

raw = m * 16
out["serde_lz4"] = {"raw_bytes": raw, "wire_bytes": int(size), "ratio": round(size / raw, 3), "serialize_plain_ms": round(ms_plain, 2), "serialize_lz4_ms": round(ms_lz, 2),
                         "lz4_gb_s": round(raw / ms_lz / 1e6, 2), "deserialize_lz4_ms": round(ms_dec, 2)}
print(json.dumps(out))
