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
