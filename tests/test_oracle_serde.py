"""CPU: the oracle's SerializedPage writer / reader against the reference's own known answers — serialized sizes asserted by
TestPagesSerde (presto-main-base/src/test/java/com/facebook/presto/execution/buffer/TestPagesSerde.java:83-127), the header layout
of PagesSerdeUtil.writeSerializedPageMetadata (:70-77), EncoderUtil's MSB-first null bits (:31-63), java.util.zip.CRC32 (= zlib) for
the checksum (:109-120) — and round trips of random pages in every supported encoding."""
import struct

import numpy as np
import pytest

from presto_b200 import BIGINT, DOUBLE, INTEGER, VARCHAR, Block, Page, PagesSerde
from serde_cases import dictionary_id_ranges, dictionary_pages, same_wire_up_to_dictionary_ids, HEADER, header_of, java_checksum, name, random_page, types_of


@pytest.fixture
def serde(oracle_backend):
    s = PagesSerde(backend=oracle_backend)
    yield s
    s.close()


def test_bigint_serialized_sizes_of_the_reference_test(serde):
    # testBigintSerializedSize: page overhead 35, first value 17 (8 + name/positions/null flag overhead), second value 8
    one = serde.serialize(Page([Block.from_numpy(BIGINT, np.array([123]))]))
    two = serde.serialize(Page([Block.from_numpy(BIGINT, np.array([123, 456]))]))
    assert one.size == 35 + 17 and two.size == 35 + 17 + 8
    assert header_of(one) == (1, 0, one.size - HEADER, one.size - HEADER, 0)
    assert one[HEADER:].tobytes() == struct.pack("<i", 1) + name("LONG_ARRAY") + struct.pack("<i", 1) + b"\x00" + struct.pack("<q", 123)


def test_varchar_serialized_sizes_of_the_reference_test(serde):
    # testVarcharSerializedSize: empty page 52, + 4 + len("alice"), + 4 + len("bob")
    empty = serde.serialize(Page([Block.from_strings([])], 0))
    alice = serde.serialize(Page([Block.from_strings([b"alice"])]))
    both = serde.serialize(Page([Block.from_strings([b"alice", b"bob"])]))
    assert (empty.size, alice.size, both.size) == (52, 52 + 9, 52 + 9 + 7)
    assert both[HEADER:].tobytes() == (struct.pack("<i", 1) + name("VARIABLE_WIDTH") + struct.pack("<i", 2) + struct.pack("<ii", 5, 8) + b"\x00"
                                       + struct.pack("<i", 8) + b"alicebob")


def test_null_bits_are_msb_first_and_null_values_are_not_written(serde):
    nulls = np.array([1, 0, 0, 0, 0, 0, 0, 0, 0, 1], dtype=np.uint8)
    wire = serde.serialize(Page([Block.from_numpy(INTEGER, np.arange(10, dtype=np.int32), nulls)]))
    body = wire[HEADER + 4 + 4 + len("INT_ARRAY"):].tobytes()
    assert body[:4] == struct.pack("<i", 10) and body[4] == 1 and body[5:7] == bytes([0x80, 0x40])
    assert body[7:] == np.arange(1, 9, dtype=np.int32).tobytes()


def test_checksum_is_the_jdk_crc32_over_data_markers_position_count_and_size(oracle_backend):
    s = PagesSerde(checksum_enabled=True, backend=oracle_backend)
    page = random_page(np.random.default_rng(1), 1000)
    wire = s.serialize(page)
    n, markers, uncompressed, size, checksum = header_of(wire)
    assert (n, markers, uncompressed, size) == (1000, 4, wire.size - HEADER, wire.size - HEADER)      # CHECKSUMMED = bit 3
    assert checksum == java_checksum(wire)
    assert s.deserialize(wire, types_of(page)).rows() == page.rows()
    wire[HEADER + 40] ^= 1
    with pytest.raises(Exception, match="checksum"):
        s.deserialize(wire, types_of(page))
    s.close()


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 1000, 4097])
@pytest.mark.parametrize("with_nulls", [False, True])
def test_round_trip(serde, n, with_nulls):
    page = random_page(np.random.default_rng(n + int(with_nulls)), n, with_nulls)
    back = serde.deserialize(serde.serialize(page), types_of(page))
    assert back.position_count == n and back.rows() == page.rows()


def test_rle_and_dictionary_blocks(serde):
    value = Block.from_numpy(DOUBLE, np.array([2.5]))
    dictionary = Block.from_numpy(BIGINT, np.array([10, 20, 30]), np.array([0, 1, 0], dtype=np.uint8))
    ids = np.array([2, 2, 0, 1, 0], dtype=np.int32)
    page = Page([Block.rle(value, 5), Block.dictionary(dictionary, ids)], 5)
    wire = serde.serialize(page)
    assert name("RLE") in wire.tobytes() and name("DICTIONARY") in wire.tobytes()
    back = serde.deserialize(wire, [DOUBLE, BIGINT])
    assert back.rows() == page.rows() == [(2.5, 30), (2.5, 30), (2.5, 10), (2.5, None), (2.5, 10)]
    # a DICTIONARY block as the Java writer lays it out (DictionaryBlockEncoding.writeBlock :30-52)
    raw = (struct.pack("<i", 1) + name("DICTIONARY") + struct.pack("<i", 5) + name("LONG_ARRAY") + struct.pack("<i", 3) + b"\x01" + bytes([0x40])
           + struct.pack("<qq", 10, 30) + ids.tobytes() + struct.pack("<qqq", 1, 2, 3))
    wire = np.frombuffer(struct.pack("<iBiiq", 5, 0, len(raw), len(raw), 0) + raw, dtype=np.uint8)
    assert serde.deserialize(wire, [BIGINT]).rows() == [(30,), (30,), (10,), (None,), (10,)]
    # and that is, up to the dictionary instance id, what the writer produces for the same block (every entry referenced: no compaction)
    assert same_wire_up_to_dictionary_ids(serde.serialize(Page([Block.dictionary(dictionary, ids)], 5)), wire)


def test_dictionary_blocks_are_written_compacted(serde):
    """DictionaryBlockEncoding.writeBlock :34-54 writes dictionaryBlock.compact() (DictionaryBlock.java:584-628): only referenced entries,
    renumbered in order of first reference; every written dictionary carries its own instance id."""
    d = Block.from_numpy(BIGINT, np.array([10, 20, 30]), np.array([0, 1, 0], dtype=np.uint8))
    wire = serde.serialize(Page([Block.dictionary(d, np.array([2, 2, 2, 1], dtype=np.int32))], 4))
    raw = (struct.pack("<i", 1) + name("DICTIONARY") + struct.pack("<i", 4) + name("LONG_ARRAY") + struct.pack("<i", 2) + b"\x01" + bytes([0x40])
           + struct.pack("<q", 30) + np.array([0, 0, 0, 1], dtype=np.int32).tobytes() + struct.pack("<qqq", 0, 0, 0))
    expected = np.frombuffer(struct.pack("<iBiiq", 4, 0, len(raw), len(raw), 0) + raw, dtype=np.uint8)
    assert same_wire_up_to_dictionary_ids(wire, expected)
    rng = np.random.default_rng(11)
    seen = set()
    for page in dictionary_pages(rng):
        wire = serde.serialize(page)
        back = serde.deserialize(wire, types_of(page))
        assert back.rows() == page.rows()
        for lo, hi in dictionary_id_ranges(wire):
            ident = wire[lo:hi].tobytes()
            assert ident not in seen
            seen.add(ident)


def test_rejects_what_it_does_not_speak(serde):
    wire = serde.serialize(Page([Block.from_numpy(BIGINT, np.array([1]))]))
    encrypted = wire.copy(); encrypted[4] = 2
    with pytest.raises(Exception, match="encrypted"):
        serde.deserialize(encrypted, [BIGINT])
    compressed = wire.copy(); compressed[4] = 1                     # a raw page flagged COMPRESSED is not an LZ4 block of that size
    with pytest.raises(Exception, match="decompression"):
        serde.deserialize(compressed, [BIGINT])
    with pytest.raises(Exception, match="truncated"):
        serde.deserialize(wire[:-3], [BIGINT])
    with pytest.raises(Exception, match="encoding"):
        serde.deserialize(wire, [INTEGER])
    with pytest.raises(Exception, match="channel count"):
        serde.deserialize(wire, [BIGINT, BIGINT])


def test_serialize_ranges_equals_serializing_the_slices(serde):
    rng = np.random.default_rng(5)
    page = random_page(rng, 1000, True)
    cuts = [0, 10, 10, 600, 1000]
    wires = serde.serialize_ranges(page, cuts)
    rows = page.rows()
    for r in range(len(cuts) - 1):
        assert serde.deserialize(wires[r], types_of(page)).rows() == rows[cuts[r]:cuts[r + 1]]
        assert header_of(wires[r])[0] == cuts[r + 1] - cuts[r]


def test_reference_sizes_and_round_trip_hold_with_the_lz4_codec(oracle_backend):
    """TestPagesSerde runs testBigintSerializedSize / testVarcharSerializedSize / testRoundTrip for every CompressionCodec, LZ4 included, and
    expects the SAME sizes (:82-127): pages this small do not reach MINIMUM_COMPRESSION_RATIO and travel uncompressed."""
    from presto_b200 import _abi as abi
    s = PagesSerde(compression=abi.COMPRESSION_LZ4, backend=oracle_backend)
    assert s.serialize(Page([Block.from_numpy(BIGINT, np.array([123]))])).size == 35 + 17
    assert s.serialize(Page([Block.from_numpy(BIGINT, np.array([123, 456]))])).size == 35 + 17 + 8
    assert s.serialize(Page([Block.from_strings([])], 0)).size == 52
    assert s.serialize(Page([Block.from_strings([b"alice"])])).size == 52 + 9
    assert s.serialize(Page([Block.from_strings([b"alice", b"bob"])])).size == 52 + 9 + 7
    names = Block.from_strings([b"alice", b"bob", b"charlie", b"dave"])                  # testRoundTrip :58-80
    page = Page([names, names, names])
    wire = s.serialize(page)
    assert s.deserialize(wire, [6, 6, 6]).rows() == page.rows()
    s.close()
