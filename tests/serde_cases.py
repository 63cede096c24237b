"""SerializedPage cases shared by the oracle (CPU) and CUDA (GPU) serde tests."""
import struct
import zlib

import numpy as np

from presto_b200 import BIGINT, BOOLEAN, DATE, DOUBLE, INTEGER, VARCHAR, Block, Page

HEADER = 21


def header_of(wire: np.ndarray):
    """(positionCount, markers, uncompressedSize, size, checksum) — PagesSerdeUtil.writeSerializedPageMetadata :70-77."""
    return struct.unpack("<iBiiq", wire[:HEADER].tobytes())


def java_checksum(wire: np.ndarray) -> int:
    """PagesSerdeUtil.computeSerializedPageChecksum :109-120 with java.util.zip.CRC32 == zlib.crc32."""
    n, markers, uncompressed, size, _ = header_of(wire)
    c = zlib.crc32(wire[HEADER:HEADER + size].tobytes())
    c = zlib.crc32(bytes([markers]), c)
    c = zlib.crc32(struct.pack("<i", n), c)
    return zlib.crc32(struct.pack("<i", uncompressed), c)


def name(s: str) -> bytes:
    return struct.pack("<i", len(s)) + s.encode()


def random_page(rng, n, with_nulls=True, varchar=True) -> Page:
    def nulls(p):
        return (rng.random(n) < p).astype(np.uint8) if with_nulls else None
    blocks = [Block.from_numpy(BIGINT, rng.integers(-2**62, 2**62, n), nulls(0.2)),
              Block.from_numpy(DOUBLE, rng.random(n), None),
              Block.from_numpy(INTEGER, rng.integers(-2**31, 2**31, n).astype(np.int32), nulls(0.5)),
              Block.from_numpy(DATE, rng.integers(0, 20000, n).astype(np.int32), nulls(0.01)),
              Block.from_numpy(BOOLEAN, rng.integers(0, 2, n).astype(np.uint8), nulls(0.3)),
              Block.from_numpy(DOUBLE, rng.random(n), nulls(1.0) if with_nulls else None)]         # all-NULL column: no values on the wire
    if varchar:
        words = [None if (with_nulls and rng.random() < 0.2) else bytes(rng.integers(97, 123, rng.integers(0, 12)).astype(np.uint8)) for _ in range(n)]
        blocks.append(Block.from_strings(words))
    return Page(blocks, n)


def types_of(page: Page):
    return [b.type for b in page.blocks]


def dictionary_id_ranges(wire: np.ndarray):
    """Byte ranges [begin, end) of the 24-byte dictionary instance ids (DictionaryBlockEncoding.writeBlock :50-53: node UUID + sequence) of a
    SerializedPage: the only bytes that differ between two writers of the same page."""
    buf = wire.tobytes()
    at = 21
    (channels,) = struct.unpack_from("<i", buf, at); at += 4
    ranges = []

    def block(at):
        (ln,) = struct.unpack_from("<i", buf, at); at += 4
        enc = buf[at:at + ln].decode(); at += ln
        (n,) = struct.unpack_from("<i", buf, at); at += 4
        if enc == "RLE":
            return block(at)
        if enc == "DICTIONARY":
            at = block(at)
            at += 4 * n
            ranges.append((at, at + 24))
            return at + 24
        if enc == "VARIABLE_WIDTH":
            at += 4 * n
            flag = buf[at]; at += 1
            if flag:
                at += (n + 7) // 8
            (total,) = struct.unpack_from("<i", buf, at)
            return at + 4 + total
        width = {"LONG_ARRAY": 8, "INT_ARRAY": 4, "BYTE_ARRAY": 1}[enc]
        flag = buf[at]; at += 1
        non_null = n
        if flag:
            bits = np.unpackbits(np.frombuffer(buf, dtype=np.uint8, count=(n + 7) // 8, offset=at))[:n]
            non_null = n - int(bits.sum())
            at += (n + 7) // 8
        return at + non_null * width

    for _ in range(channels):
        at = block(at)
    assert at == len(buf)
    return ranges


def same_wire_up_to_dictionary_ids(a: np.ndarray, b: np.ndarray) -> bool:
    if a.size != b.size:
        return False
    ra, rb = dictionary_id_ranges(a), dictionary_id_ranges(b)
    if ra != rb:
        return False
    a, b = a.copy(), b.copy()
    for lo, hi in ra:
        a[lo:hi] = 0; b[lo:hi] = 0
    a[13:21] = 0; b[13:21] = 0          # the checksum covers the ids
    return bool((a == b).all())


def dictionary_pages(rng):
    """DICTIONARY blocks the way scans and joins produce them: fully referenced dictionaries (written as they are), sparsely referenced
    ones (compacted: referenced entries in order of first reference), NULL entries, VARCHAR dictionaries, an empty page."""
    pages = []
    d = Block.from_numpy(BIGINT, np.array([10, 20, 30]), np.array([0, 1, 0], dtype=np.uint8))
    pages.append(Page([Block.dictionary(d, np.array([2, 2, 0, 1, 0], dtype=np.int32))], 5))                       # all referenced
    pages.append(Page([Block.dictionary(d, np.array([2, 2, 2, 1], dtype=np.int32))], 4))                          # entry 0 unused: [30, NULL]
    big = Block.from_numpy(DOUBLE, rng.random(1000), (rng.random(1000) < 0.2).astype(np.uint8))
    ids = rng.integers(0, 1000, 300).astype(np.int32)
    strs = Block.from_strings([None if i % 7 == 3 else (b"s%d" % i) * (i % 4) for i in range(50)])
    pages.append(Page([Block.dictionary(big, ids), Block.from_numpy(INTEGER, rng.integers(0, 9, 300).astype(np.int32)),
                       Block.dictionary(strs, rng.integers(0, 50, 300).astype(np.int32)),
                       Block.dictionary(strs, (np.arange(300) % 5 + 40).astype(np.int32))], 300))
    pages.append(Page([Block.dictionary(d, np.zeros(0, dtype=np.int32))], 0))
    return pages
