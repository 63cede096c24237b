"""Fixed vectors (tests/golden/*.json, made by tests/golden/make_golden.py with the oracle once the oracle was pinned on the
reference's own test expectations): the oracle must keep reproducing them, the numpy restatement of the hash functions and zlib /
xxhash agree with them independently, and on the GPU the CUDA serde and scalar helpers reproduce them byte for byte."""
import json
import os
import struct
import zlib

import numpy as np
import pytest

from presto_b200 import Block, Page, PagesSerde
from presto_b200.hashing import get_partition, hash_long, murmur_hash3

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAGES = json.load(open(os.path.join(HERE, "serialized_pages.json")))
HASHES = json.load(open(os.path.join(HERE, "hash_vectors.json")))


def page_of(case) -> Page:
    """Rebuild the page from the golden rows (flat blocks; the wire of an RLE column differs, so RLE cases are only decoded)."""
    types, rows = case["types"], case["rows"]
    blocks = []
    for c, t in enumerate(types):
        col = [r[c] for r in rows]
        if t == 6:
            blocks.append(Block.from_strings([None if v is None else v.encode() for v in col]))
        else:
            blocks.append(Block.from_values(t, [0 if v is None else v for v in col], [v is None for v in col]) if any(v is None for v in col)
                          else Block.from_values(t, col))
    return Page(blocks, len(rows))


def check_serde(backend):
    for name, case in PAGES.items():
        wire = np.frombuffer(bytes.fromhex(case["wire_hex"]), dtype=np.uint8)
        s = PagesSerde(checksum_enabled=name.endswith("_checksummed"), backend=backend)
        rows = [[x.decode() if isinstance(x, bytes) else x for x in r] for r in s.deserialize(wire, case["types"]).rows()]
        assert rows == case["rows"], name
        if "rle" not in name:
            assert s.serialize(page_of(case)).tobytes().hex() == case["wire_hex"], name
        s.close()


def check_hashes(backend):
    for i, v in enumerate(HASHES["inputs"]):
        assert backend.hash_bigint(v) == HASHES["hash_bigint"][i]
        assert backend.hash_murmur3_fmix64(v) == HASHES["murmur3_fmix64"][i]
        assert backend.hash_combine(v, 31) == HASHES["combine_with_31"][i]
        assert backend.hash_partition(HASHES["hash_bigint"][i], 8) == HASHES["partition_of_8"][i]
        assert backend.hash_xxhash64(np.array([v], dtype=np.int64).ctypes.data, 8) == HASHES["xxhash64_of_le_bytes"][i]


def test_oracle_reproduces_the_golden_vectors(oracle_backend):
    check_serde(oracle_backend)
    check_hashes(oracle_backend)


def test_independent_restatements_agree_with_the_golden_vectors():
    v = np.array(HASHES["inputs"], dtype=np.int64)
    assert hash_long(v).tolist() == HASHES["hash_bigint"]                    # numpy restatement of AbstractLongType.hash
    assert murmur_hash3(v).tolist() == HASHES["murmur3_fmix64"]
    assert get_partition(np.array(HASHES["hash_bigint"], dtype=np.int64), 8).tolist() == HASHES["partition_of_8"]
    for name, case in PAGES.items():                                        # the CRC32 in the header is the JDK's (= zlib's)
        if name.endswith("_checksummed"):
            wire = bytes.fromhex(case["wire_hex"])
            n, markers, uncompressed, size, checksum = struct.unpack("<iBiiq", wire[:21])
            c = zlib.crc32(struct.pack("<i", uncompressed), zlib.crc32(struct.pack("<i", n), zlib.crc32(bytes([markers]), zlib.crc32(wire[21:21 + size]))))
            assert markers == 4 and checksum == c, name
    xxhash = pytest.importorskip("xxhash")
    assert [struct.unpack("<q", struct.pack("<Q", xxhash.xxh64(struct.pack("<q", x)).intdigest()))[0] for x in HASHES["inputs"]] == HASHES["xxhash64_of_le_bytes"]


@pytest.mark.gpu
def test_gpu_reproduces_the_golden_vectors(gpu_backend):
    check_serde(gpu_backend)
    check_hashes(gpu_backend)
