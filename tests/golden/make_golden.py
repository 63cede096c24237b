#!/usr/bin/env python
"""Regenerates the fixtures of this directory with the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference is Java and cannot run in this image, so these vectors are produced by the oracle — AFTER the oracle itself has been
pinned on the reference's own known answers (tests/test_oracle_reference.py, test_oracle_serde.py, test_oracle_semi_join.py: the
expectations of TestGroupByHash, TestHashJoinOperator, TestPagesSerde, TestHashSemiJoinOperator, …).  They freeze byte-level and
bit-level results so that both the oracle and the CUDA path are checked against fixed data, not only against each other."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from presto_b200 import BIGINT, DOUBLE, INTEGER, Block, Page, PagesSerde  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pages():
    """name → (page, types): small pages covering every encoding the serde writes."""
    rng = np.random.default_rng(20260921)
    nulls10 = np.array([1, 0, 0, 0, 0, 0, 0, 0, 0, 1], dtype=np.uint8)
    return {
        "bigint_one_value": (Page([Block.from_numpy(BIGINT, np.array([123]))]), [BIGINT]),
        "bigint_two_values": (Page([Block.from_numpy(BIGINT, np.array([123, 456]))]), [BIGINT]),
        "varchar_alice_bob": (Page([Block.from_strings([b"alice", b"bob"])]), [6]),
        "integer_nulls_first_and_last": (Page([Block.from_numpy(INTEGER, np.arange(10, dtype=np.int32), nulls10)]), [INTEGER]),
        "rle_double_and_flat_bigint": (Page([Block.rle(Block.from_numpy(DOUBLE, np.array([2.5])), 3), Block.from_numpy(BIGINT, np.array([7, 8, 9]))], 3), [DOUBLE, BIGINT]),
        "mixed_33_rows": (Page([Block.from_numpy(BIGINT, rng.integers(-2**40, 2**40, 33), (rng.random(33) < 0.3).astype(np.uint8)),
                                Block.from_numpy(DOUBLE, np.round(rng.random(33) * 100, 3)),
                                Block.from_strings([None if i % 5 == 0 else bytes([97 + i % 26]) * (i % 4) for i in range(33)])], 33), [BIGINT, DOUBLE, 6]),
    }


HASH_INPUTS = [0, 1, -1, 42, 2**31, -2**31, 2**62, -2**63, 123456789012345, 0x7FFFFFFFFFFFFFFF]


def main():
    cpu = oracle.backend()
    out = {}
    for checksum in (False, True):
        s = PagesSerde(checksum_enabled=checksum, backend=cpu)
        for name, (page, types) in pages().items():
            out[f"{name}{'_checksummed' if checksum else ''}"] = {"types": types, "rows": [list(r) for r in _jsonable(page.rows())], "wire_hex": s.serialize(page).tobytes().hex()}
        s.close()
    json.dump(out, open(os.path.join(HERE, "serialized_pages.json"), "w"), indent=1)
    h = {"inputs": HASH_INPUTS,
         "hash_bigint": [int(cpu.hash_bigint(v)) for v in HASH_INPUTS],                      # AbstractLongType.hash
         "murmur3_fmix64": [int(cpu.hash_murmur3_fmix64(v)) for v in HASH_INPUTS],           # HashCommon.murmurHash3 / PagesHash.java:236-252
         "combine_with_31": [int(cpu.hash_combine(v, 31)) for v in HASH_INPUTS],             # CombineHashFunction
         "partition_of_8": [int(cpu.hash_partition(int(cpu.hash_bigint(v)), 8)) for v in HASH_INPUTS],   # HashGenerator.getPartition
         "xxhash64_of_le_bytes": [int(cpu.hash_xxhash64(np.array([v], dtype=np.int64).ctypes.data, 8)) for v in HASH_INPUTS]}
    json.dump(h, open(os.path.join(HERE, "hash_vectors.json"), "w"), indent=1)
    print("wrote", len(out), "serialized pages and", len(HASH_INPUTS), "hash inputs")


def _jsonable(rows):
    return [[x.decode() if isinstance(x, bytes) else x for x in r] for r in rows]


if __name__ == "__main__":
    main()
