"""-m gpu: the CUDA SerializedPage serde (presto_b200/csrc/pages_serde.cu) against the oracle: wire bytes bit-exact in both
directions (what one writes the other reads), the reference's known sizes, checksums, RLE / DICTIONARY channels, device-resident
pages and wire buffers, rejected inputs."""
import ctypes as C
import struct

import numpy as np
import pytest

from presto_b200 import BIGINT, DOUBLE, INTEGER, MEM_DEVICE, VARCHAR, Block, Page, PagesSerde
from serde_cases import dictionary_id_ranges, dictionary_pages, same_wire_up_to_dictionary_ids, HEADER, header_of, java_checksum, name, random_page, types_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 15, 16, 17, 1000, 4097, 100_003])
@pytest.mark.parametrize("with_nulls", [False, True])
@pytest.mark.parametrize("checksum", [False, True])
def test_wire_bytes_equal_the_oracles_and_round_trip(gpu_backend, oracle_backend, n, with_nulls, checksum):
    page = random_page(np.random.default_rng(7 * n + int(with_nulls)), n, with_nulls)
    g, o = PagesSerde(checksum_enabled=checksum, backend=gpu_backend), PagesSerde(checksum_enabled=checksum, backend=oracle_backend)
    wire_g, wire_o = g.serialize(page), o.serialize(page)
    assert wire_g.size == wire_o.size and (wire_g == wire_o).all()
    if checksum:
        assert header_of(wire_g)[4] == java_checksum(wire_g)
    assert g.deserialize(wire_o, types_of(page)).rows() == page.rows()        # GPU reads what the oracle wrote
    assert o.deserialize(wire_g, types_of(page)).rows() == page.rows()        # and the other way round
    g.close(); o.close()


def test_reference_sizes(gpu_backend):
    g = PagesSerde(backend=gpu_backend)
    assert g.serialize(Page([Block.from_numpy(BIGINT, np.array([123]))])).size == 52          # TestPagesSerde.testBigintSerializedSize: 35 + 17
    assert g.serialize(Page([Block.from_numpy(BIGINT, np.array([123, 456]))])).size == 60
    assert g.serialize(Page([Block.from_strings([])], 0)).size == 52                              # testVarcharSerializedSize
    assert g.serialize(Page([Block.from_strings([b"alice", b"bob"])])).size == 52 + 9 + 7
    g.close()


def test_rle_and_dictionary_channels(gpu_backend, oracle_backend):
    g, o = PagesSerde(backend=gpu_backend), PagesSerde(backend=oracle_backend)
    value = Block.from_numpy(DOUBLE, np.array([2.5]))
    dictionary = Block.from_numpy(BIGINT, np.array([10, 20, 30]), np.array([0, 1, 0], dtype=np.uint8))
    ids = np.array([2, 2, 0, 1, 0], dtype=np.int32)
    page = Page([Block.rle(value, 5), Block.dictionary(dictionary, ids)], 5)
    wire = g.serialize(page)
    assert same_wire_up_to_dictionary_ids(wire, o.serialize(page))
    assert g.deserialize(wire, [DOUBLE, BIGINT]).rows() == page.rows()
    raw = (struct.pack("<i", 1) + name("DICTIONARY") + struct.pack("<i", 5) + name("LONG_ARRAY") + struct.pack("<i", 3) + b"\x01" + bytes([0x40])
           + struct.pack("<qq", 10, 30) + ids.tobytes() + struct.pack("<qqq", 1, 2, 3))
    java_wire = np.frombuffer(struct.pack("<iBiiq", 5, 0, len(raw), len(raw), 0) + raw, dtype=np.uint8)
    assert g.deserialize(java_wire, [BIGINT]).rows() == [(30,), (30,), (10,), (None,), (10,)]
    g.close(); o.close()


@pytest.mark.parametrize("checksum", [False, True])
def test_dictionary_blocks_are_written_compacted(gpu_backend, oracle_backend, checksum):
    """DICTIONARY channels: the GPU writer's bytes equal the oracle's (compacted dictionary in order of first reference, remapped ids) up to
    the dictionary instance ids; both read each other's pages; ids are unique per written dictionary."""
    g, o = PagesSerde(checksum_enabled=checksum, backend=gpu_backend), PagesSerde(checksum_enabled=checksum, backend=oracle_backend)
    seen = set()
    for page in dictionary_pages(np.random.default_rng(11)):
        wg, wo = g.serialize(page), o.serialize(page)
        assert same_wire_up_to_dictionary_ids(wg, wo)
        assert g.deserialize(wo, types_of(page)).rows() == page.rows() == o.deserialize(wg, types_of(page)).rows()
        for lo, hi in dictionary_id_ranges(wg):
            ident = wg[lo:hi].tobytes()
            assert ident not in seen
            seen.add(ident)
    g.close(); o.close()


def test_device_resident_page_and_wire(gpu_backend, oracle_backend):
    """Device page → device wire → device page: nothing but the fixed fields crosses PCIe; checked by bringing both ends back."""
    n = 300_007
    rng = np.random.default_rng(3)
    page = random_page(rng, n, True, varchar=False)
    keep = []
    blocks = []
    for b in page.blocks:
        ptrs = []
        for arr in (b.values, b.is_null):
            if arr is None:
                ptrs.append(0); continue
            p = C.c_void_p()
            gpu_backend.check(gpu_backend.device_alloc(max(arr.nbytes, 16), C.byref(p)))
            gpu_backend.check(gpu_backend.memcpy_h2d(p, arr.ctypes.data, arr.nbytes))
            keep.append(p); ptrs.append(p.value)
        blocks.append(Block.device(b.type, n, ptrs[0], ptrs[1]))
    g = PagesSerde(checksum_enabled=True, backend=gpu_backend, output_memory=MEM_DEVICE)
    addr, size = g.serialize(Page(blocks, n), copy=False)
    wire = np.empty(size, dtype=np.uint8)
    gpu_backend.check(gpu_backend.memcpy_d2h(wire.ctypes.data, C.c_void_p(addr), size))
    o = PagesSerde(checksum_enabled=True, backend=oracle_backend)
    assert (wire == o.serialize(page)).all()
    # read it back from a device copy of the wire with a second serde (the first one's buffer is only valid until its next call)
    g2 = PagesSerde(backend=gpu_backend, output_memory=MEM_DEVICE)
    back = g2.deserialize(addr, types_of(page), size=size, memory=MEM_DEVICE, copy=True)
    assert back.rows() == page.rows()
    for p in keep:
        gpu_backend.device_free(p)
    g.close(); g2.close(); o.close()


def test_rejects_what_it_does_not_speak(gpu_backend):
    g = PagesSerde(checksum_enabled=True, backend=gpu_backend)
    wire = g.serialize(Page([Block.from_numpy(BIGINT, np.arange(100))]))
    bad = wire.copy(); bad[4] |= 2
    with pytest.raises(Exception, match="encrypted"):
        g.deserialize(bad, [BIGINT])
    with pytest.raises(Exception, match="truncated"):
        g.deserialize(wire[:-3], [BIGINT])
    with pytest.raises(Exception, match="encoding"):
        g.deserialize(wire, [INTEGER])
    with pytest.raises(Exception, match="channel count"):
        g.deserialize(wire, [BIGINT, BIGINT])
    flipped = wire.copy(); flipped[HEADER + 60] ^= 4
    with pytest.raises(Exception, match="checksum"):
        g.deserialize(flipped, [BIGINT])
    g.close()


def test_partitioned_output_to_serialized_pages(gpu_backend, oracle_backend):
    """PartitionedOutputOperator → one SerializedPage per partition (the off-box leg): decoded by the oracle, every partition holds
    exactly the rows the oracle's partitioned output puts there, in the same order."""
    from presto_b200 import MEM_DEVICE, PartitionedOutputOperator
    from presto_b200.exchange import serialize_partitions
    rng = np.random.default_rng(8)
    n, parts = 20_011, 5
    page = Page([Block.from_numpy(BIGINT, rng.integers(-1000, 1000, n), (rng.random(n) < 0.1).astype(np.uint8)), Block.from_numpy(DOUBLE, rng.random(n))])
    op = PartitionedOutputOperator([BIGINT, DOUBLE], parts, [0], backend=gpu_backend, output_memory=MEM_DEVICE)
    op.add_input(page)
    out, offs = op.get_output(copy=False), op.get_offsets()
    g = PagesSerde(checksum_enabled=True, backend=gpu_backend)
    wires = serialize_partitions(g, out, offs)
    ref = PartitionedOutputOperator([BIGINT, DOUBLE], parts, [0], backend=oracle_backend)
    ref.add_input(page)
    ref_out = ref.get_output()
    want, woffs = ref_out.rows(), ref.get_offsets()
    o = PagesSerde(checksum_enabled=True, backend=oracle_backend)
    assert list(offs) == list(woffs)
    ref_wires = o.serialize_ranges(ref_out, woffs)          # pg_pages_serde_serialize_ranges on both sides: byte for byte
    for p in range(parts):
        assert o.deserialize(wires[p], [BIGINT, DOUBLE]).rows() == want[woffs[p]:woffs[p + 1]]
        assert (wires[p] == ref_wires[p]).all()
    op.close(); ref.close(); g.close(); o.close()


def test_serialize_ranges_with_varchar_and_empty_ranges(gpu_backend, oracle_backend):
    """Row ranges of a host page (VARCHAR + nullable BIGINT), some of them empty: every range is the SerializedPage of that slice."""
    rng = np.random.default_rng(9)
    n = 5000
    strs = [None if rng.random() < 0.1 else bytes(rng.integers(97, 123, rng.integers(0, 12)).astype(np.uint8)) for _ in range(n)]
    page = Page([Block.from_strings(strs), Block.from_numpy(BIGINT, rng.integers(-9, 9, n), (rng.random(n) < 0.2).astype(np.uint8))])
    cuts = [0, 0, 1234, 1234, 4999, 5000]
    g, o = PagesSerde(backend=gpu_backend), PagesSerde(backend=oracle_backend)
    got, want = g.serialize_ranges(page, cuts), o.serialize_ranges(page, cuts)
    rows = page.rows()
    for r in range(len(cuts) - 1):
        assert (got[r] == want[r]).all()
        assert o.deserialize(got[r], [VARCHAR, BIGINT]).rows() == rows[cuts[r]:cuts[r + 1]]
    with pytest.raises(Exception):
        g.serialize_ranges(page, [0, 6000])
    g.close(); o.close()


# ---- LZ4 (PagesSerde.wrapSlice :118-143 with the PageCompressor of PagesSerdeFactory.java:75; csrc/lz4_codec.cuh) -------------------------
def _lz4_pages():
    rng = np.random.default_rng(12)
    n = 60_000
    yield "sequence", Page([Block.from_numpy(BIGINT, np.arange(n, dtype=np.int64) // 3), Block.from_numpy(INTEGER, (np.arange(n) % 11).astype(np.int32))]), True
    yield "strings+nulls", Page([Block.from_strings([None if i % 13 == 0 else [b"BUILDING", b"AUTOMOBILE", b"", b"MACHINERY"][i % 4] for i in range(n)]),
                                 Block.from_numpy(DOUBLE, np.round(rng.random(n) * 4) / 4, (rng.random(n) < 0.2).astype(np.uint8))]), True
    yield "noise", Page([Block.from_numpy(DOUBLE, rng.random(n)), Block.from_numpy(BIGINT, rng.integers(-2**62, 2**62, n))]), False
    yield "tiny", Page([Block.from_numpy(BIGINT, np.array([1, 2, 3]))]), False
    yield "big", random_page(np.random.default_rng(5), 400_000, True), None          # several hundred 8 KB segments; whether it shrinks is the data's business


@pytest.mark.parametrize("checksum", [False, True])
@pytest.mark.parametrize("device_output", [False, True])
def test_lz4_pages_cross_with_the_oracle_and_liblz4(gpu_backend, oracle_backend, checksum, device_output):
    import pyarrow as pa
    from presto_b200 import _abi as abi
    kw = dict(output_memory=MEM_DEVICE) if device_output else {}
    g = PagesSerde(checksum_enabled=checksum, compression=abi.COMPRESSION_LZ4, backend=gpu_backend, **kw)
    g_plain = PagesSerde(checksum_enabled=checksum, backend=gpu_backend)
    o = PagesSerde(checksum_enabled=checksum, compression=abi.COMPRESSION_LZ4, backend=oracle_backend)
    codec = pa.Codec("lz4_raw")
    for label, page, shrinks in _lz4_pages():
        types = types_of(page)
        wire = g.serialize(page)               # (copied back from the serde's output memory, host or device)
        raw = g_plain.serialize(page)
        position_count, markers, uncompressed, size, crc = header_of(wire)
        compressed = bool(markers & abi.PAGE_MARKER_COMPRESSED)
        if shrinks is not None:
            assert compressed == shrinks, label
        assert position_count == page.position_count and uncompressed == raw.size - HEADER and size == wire.size - HEADER, label
        if compressed:
            assert size <= 0.8 * uncompressed
            assert codec.decompress(bytes(wire[HEADER:]), decompressed_size=uncompressed, asbytes=True) == bytes(raw[HEADER:]), label    # liblz4 reads the block
        else:
            assert (wire[HEADER:] == raw[HEADER:]).all(), label
        if checksum:
            assert crc == java_checksum(wire), label
        assert o.deserialize(wire, types).rows() == page.rows(), label             # the reference-side reader
        assert g_plain.deserialize(wire, types).rows() == page.rows(), label       # a serde without a compressor still reads compressed pages
        wire_o = o.serialize(page)                                                  # and what the oracle's encoder writes
        assert g_plain.deserialize(wire_o, types).rows() == page.rows(), label
    g.close(); g_plain.close(); o.close()


def test_lz4_block_written_by_liblz4_is_read_and_malformed_blocks_are_refused(gpu_backend):
    import pyarrow as pa
    from presto_b200 import PrestoException
    from presto_b200 import _abi as abi
    g = PagesSerde(backend=gpu_backend)
    page = Page([Block.from_numpy(BIGINT, np.arange(30_000, dtype=np.int64) % 97), Block.from_strings([b"x" * (i % 5) for i in range(30_000)])])
    raw = g.serialize(page)
    z = pa.Codec("lz4_raw").compress(bytes(raw[HEADER:]), asbytes=True)
    wire = np.frombuffer(struct.pack("<iBiiq", page.position_count, abi.PAGE_MARKER_COMPRESSED, raw.size - HEADER, len(z), 0) + z, dtype=np.uint8)
    assert g.deserialize(wire, types_of(page)).rows() == page.rows()
    bad = wire.copy(); bad[HEADER + 40:HEADER + 60] = 0xff
    with pytest.raises(PrestoException):
        g.deserialize(bad, types_of(page))
    short = np.frombuffer(struct.pack("<iBiiq", page.position_count, abi.PAGE_MARKER_COMPRESSED, raw.size - HEADER - 1, len(z), 0) + z, dtype=np.uint8)
    with pytest.raises(PrestoException):
        g.deserialize(short, types_of(page))
    g.close()


def test_lz4_decoder_on_blocks_of_every_shape(gpu_backend):
    """The device decoder on blocks liblz4 wrote: a ~1 MB match at offset 1, matches at the largest offsets, literal runs of tens of
    kilobytes, hundreds of thousands of tiny sequences — payloads are handed over as VARCHAR slices of one row so that the bytes are
    exactly these."""
    import pyarrow as pa
    from presto_b200 import _abi as abi
    rng = np.random.default_rng(77)
    noise = bytes(rng.integers(0, 256, 70_000).astype(np.uint8))
    payloads = [
        b"\x00" * 1_000_000,                                                  # one match of ~1 MB at offset 1
        noise[:20_000] + b"ab" * 200_000 + noise[20_000:40_000],              # offset 2, long; literals longer than the stage on both sides
        noise[:65_535] + noise[:65_535] + noise[:100],                        # matches at offset 65535
        noise[:60_000] + b"x" * 10 + noise[:60_000],                          # offset 60010
        b"".join(bytes([65 + i % 7]) * (1 + i % 9) for i in range(60_000)),   # hundreds of thousands of short sequences
        noise,                                                                # incompressible: one literal run of 70 000 bytes
    ]
    g = PagesSerde(backend=gpu_backend)
    codec = pa.Codec("lz4_raw")
    for payload in payloads:
        page = Page([Block.from_strings([payload])])
        raw = g.serialize(page)
        z = codec.compress(bytes(raw[HEADER:]), asbytes=True)
        wire = np.frombuffer(struct.pack("<iBiiq", 1, abi.PAGE_MARKER_COMPRESSED, raw.size - HEADER, len(z), 0) + z, dtype=np.uint8)
        got = g.deserialize(wire, [VARCHAR]).rows()
        assert len(got) == 1 and got[0][0] == payload
    g.close()


from conftest import unverified_on_gpu


@unverified_on_gpu
def test_reference_sizes_hold_with_the_lz4_codec(gpu_backend):
    """TestPagesSerde :82-127 runs its size expectations for every CompressionCodec, LZ4 included: pages this small stay uncompressed."""
    from presto_b200 import _abi as abi
    g = PagesSerde(compression=abi.COMPRESSION_LZ4, backend=gpu_backend)
    assert g.serialize(Page([Block.from_numpy(BIGINT, np.array([123]))])).size == 52
    assert g.serialize(Page([Block.from_numpy(BIGINT, np.array([123, 456]))])).size == 60
    assert g.serialize(Page([Block.from_strings([])], 0)).size == 52
    assert g.serialize(Page([Block.from_strings([b"alice", b"bob"])])).size == 52 + 9 + 7
    names = Block.from_strings([b"alice", b"bob", b"charlie", b"dave"])
    page = Page([names, names, names])
    assert g.deserialize(g.serialize(page), [VARCHAR, VARCHAR, VARCHAR]).rows() == page.rows()
    g.close()
