"""CPU-side checks of the drop-in boundary: libprestogpu.so loads and exports every symbol include/presto_gpu.h declares
(no compute calls without a GPU), the ctypes mirrors match the header's struct layouts, and the product fails loudly — it
never falls back to a CPU path — when no CUDA device is visible."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "presto_gpu.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from presto_b200._lib import LIB_PATH, load_library
    assert os.path.exists(LIB_PATH), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"libprestogpu.so does not export {missing}"
    backend = load_library()
    assert backend.version().startswith(b"presto_b200")


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof of every ABI struct, computed by the C compiler from the header, equals the ctypes mirror's."""
    from presto_b200 import _abi as abi
    structs = ["pg_block", "pg_page", "pg_operator_stats", "pg_operator_options", "pg_aggregate", "pg_hash_aggregation_config",
               "pg_group_by_hash_config", "pg_hash_builder_config", "pg_lookup_join_config", "pg_partitioned_output_config",
               "pg_expr_node", "pg_filter_project_config", "pg_fused_aggregation_config", "pg_fused_page_config", "pg_semi_join_config",
               "pg_pages_serde_config", "pg_lookup_outer_config"]
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for s in structs:
        cls = getattr(abi, s)
        src.append(f'printf("{s} %zu", sizeof({s}));')
        for fname, _ in cls._fields_:
            src.append(f'printf(" %zu", offsetof({s}, {fname}));')
        src.append('printf("\\n");')
    src.append("return 0;}")
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(cfile)])
    out = subprocess.check_output([str(exe)]).decode().splitlines()
    for line in out:
        parts = line.split()
        cls = getattr(abi, parts[0])
        expect = [C.sizeof(cls)] + [getattr(cls, f).offset for f, _ in cls._fields_]
        assert [int(x) for x in parts[1:]] == expect, f"layout mismatch for {parts[0]}"


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU every operator constructor must raise NO_DEVICE (this test is skipped where a GPU exists)."""
    from presto_b200 import AGG_COUNT_STAR, BIGINT, Aggregate, HashAggregationOperator, JoinBridge, PrestoException
    from presto_b200._lib import load_library
    b = load_library()
    n = C.c_int32()
    if b.device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(PrestoException) as e:
        HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_COUNT_STAR)])
    assert e.value.code == 8
    with pytest.raises(PrestoException):
        JoinBridge()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "presto_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "liboracle" not in text and "oracle_" not in text, f"{f} references the oracle"


def test_xxhash64_of_slices_of_every_length_matches_the_xxhash_package():
    """pg_hash_xxhash64 is common.cuh's xxHash64Bytes — the function the kernels hash VARCHAR slices with (interned group keys, $hashvalue
    of VARCHAR keys) — compiled for the host: every length 0..100 (stripe loop, 8- / 4- / 1-byte tails) at several alignments."""
    xxhash = pytest.importorskip("xxhash")
    import numpy as np
    from presto_b200._lib import load_library
    b = load_library()
    rng = np.random.default_rng(64)
    data = bytes(rng.integers(0, 256, 300).astype(np.uint8))
    for n in range(0, 101):
        for start in (0, 1, 3, 7):
            s = data[start:start + n]
            want = xxhash.xxh64_intdigest(s)
            assert b.hash_xxhash64(s, n) == (want - (1 << 64) if want >= 1 << 63 else want), (n, start)


def test_host_hash_helpers_match_known_answers():
    """pg_hash_* host helpers (no GPU needed) against SURVEY.md §8c known answers."""
    from presto_b200._lib import load_library
    b = load_library()
    assert b.hash_bigint(1) == 9155312661752487122 and b.hash_bigint(42) == 1279370338401826588
    assert b.hash_murmur3_fmix64(1) == -5451962507482445012 and b.hash_murmur3_fmix64(-1) == 7256831767414464289
    assert [b.hash_partition(b.hash_bigint(v), 8) for v in (0, 1, 42)] == [0, 4, 2]
    assert b.hash_xxhash64(b"A", 1) == 1371800463213966980 and b.hash_xxhash64(b"", 0) == -1205034819632174695
    assert b.hash_combine(3, 4) == 97
    # TypeUtils.doubleHashCode (PC/type/TypeUtils.java:232-238): -0.0 is canonicalised to +0.0 (hash(0) = 0), every NaN to the canonical NaN
    assert b.hash_double(-0.0) == b.hash_double(0.0) == 0
    import struct
    nan_payload = struct.unpack("<d", struct.pack("<Q", 0x7FF0000000000123))[0]
    assert b.hash_double(float("nan")) == b.hash_double(nan_payload) == b.hash_bigint(0x7FF8000000000000)
    assert b.hash_double(1.5) == b.hash_bigint(struct.unpack("<q", struct.pack("<d", 1.5))[0])


def test_local_partition_generator_known_answers(oracle_backend):
    """LocalPartitionGenerator.getPartition (MB/operator/exchange/LocalPartitionGenerator.java:43-63) =
    (int) XxHash64.hash(Long.reverse(rawHash)) & (P - 1): the product helper, the oracle and an independent restatement through the
    python xxhash package (XXH64 of the 8 little-endian bytes) agree; P must be a power of two."""
    import struct

    import numpy as np
    from presto_b200._lib import load_library
    from presto_b200.hashing import hash_long, local_partition
    xxhash = pytest.importorskip("xxhash")
    b = load_library()
    rng = np.random.default_rng(3)
    raw = np.concatenate([np.array([0, 1, -1, 42, 1 << 62, -(1 << 63)], dtype=np.int64), hash_long(rng.integers(-10**12, 10**12, 200))])
    for P in (1, 2, 8, 64, 1024):
        want = []
        for v in raw.tolist():
            rev = int(format(v & 0xFFFFFFFFFFFFFFFF, "064b")[::-1], 2)
            want.append(xxhash.xxh64(struct.pack("<Q", rev)).intdigest() & (P - 1))
        assert [b.hash_local_partition(v, P) for v in raw.tolist()] == want
        assert [oracle_backend.hash_local_partition(v, P) for v in raw.tolist()] == want
        assert local_partition(raw, P).tolist() == want
    assert b.hash_local_partition(5, 6) == -1 and oracle_backend.hash_local_partition(5, 0) == -1


def test_chunked_crc32_combine_equals_zlib():
    """The host half of the SerializedPage checksum (csrc/crc32_host.h): per-chunk CRCs combined with the zero-extension operator
    must equal java.util.zip.CRC32 (= zlib.crc32) of the whole buffer, for chunk counts of 1, 2, many, and ragged last chunks."""
    import zlib

    import numpy as np
    from presto_b200._lib import load_library
    lib = load_library()
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 1_000_003, dtype=np.uint8)
    for n in (0, 1, 4095, 4096, 4097, 8192, 12_289, 100_000, 1_000_003):
        for chunk in (4096, 1000, 65536):
            got = lib.hash_crc32_chunked(data.ctypes.data, n, chunk)
            assert got == zlib.crc32(data[:n].tobytes()), (n, chunk)
