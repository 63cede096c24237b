"""The LZ4 block codec of the SerializedPage serde (presto_b200/csrc/lz4_codec.cuh) run on the host through pg_lz4_*_host — the very
functions the kernels compile — against liblz4 (pyarrow's lz4_raw codec): every block this encoder writes is decoded by liblz4 to the
input, every block liblz4 writes is decoded by this decoder; malformed blocks are refused.  The oracle's serde (its own plain encoder /
decoder) is crossed with both.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

pa = pytest.importorskip("pyarrow")


@pytest.fixture(scope="module")
def lib():
    from presto_b200._lib import load_library
    return load_library()


def compress(lib, data: bytes) -> bytes:
    cap = lib.lz4_compress_bound(len(data))
    out = C.create_string_buffer(max(cap, 1))
    n = lib.lz4_compress_host(data, len(data), out, cap)
    assert 0 <= n <= cap
    return out.raw[:n]


def decompress(lib, data: bytes, size: int):
    out = C.create_string_buffer(max(size, 1))
    n = lib.lz4_decompress_host(data, len(data), out, size)
    return None if n < 0 else out.raw[:n]


def samples():
    rng = np.random.default_rng(4)
    yield b""
    yield b"a"
    yield b"abcdefghijkl"                                           # shorter than MFLIMIT: literals only
    yield b"a" * 13
    yield b"a" * 100_000                                            # one long overlapping match per segment
    yield bytes(rng.integers(0, 256, 50_000).astype(np.uint8))      # incompressible
    yield bytes(rng.integers(0, 4, 70_001).astype(np.uint8))        # tiny alphabet: many short matches
    yield (b"lineitem|" + bytes(rng.integers(48, 58, 8).astype(np.uint8))) * 9000          # record-like, crosses many 8 KB segments
    yield np.arange(40_000, dtype=np.int64).tobytes()               # a BIGINT sequence column
    yield bytes(rng.integers(0, 256, 8192 * 3).astype(np.uint8)) + b"x" * 8192 * 2 + bytes(rng.integers(0, 256, 100).astype(np.uint8))   # segments without matches between compressible ones
    yield b"".join(bytes([i % 251]) * (i % 40) for i in range(3000))                        # literal / match length extension bytes (>= 15, >= 270)
    for n in (8191, 8192, 8193, 16384, 8192 + 11, 8192 + 12, 8192 + 13):                    # segment and end-of-block boundaries
        yield (b"0123456789abcdef" * (n // 16 + 1))[:n]


def test_blocks_of_this_encoder_are_decoded_by_liblz4_and_back(lib):
    codec = pa.Codec("lz4_raw")
    for data in samples():
        z = compress(lib, data)
        if data:
            assert codec.decompress(z, decompressed_size=len(data), asbytes=True) == data
        assert decompress(lib, z, len(data)) == data
        if len(data) > 1000 and len(set(data)) < 64:
            assert len(z) < len(data) * 0.8, (len(z), len(data))


def test_blocks_of_liblz4_are_decoded(lib):
    codec = pa.Codec("lz4_raw")
    for data in samples():
        if not data:
            continue
        z = codec.compress(data, asbytes=True)
        assert decompress(lib, z, len(data)) == data


def test_malformed_blocks_are_refused(lib):
    data = b"presto-b200 " * 500
    z = compress(lib, data)
    assert decompress(lib, z, len(data) - 1) is None                # decodes to more than the capacity
    assert decompress(lib, z[:-3], len(data)) != data               # truncated literals
    bad = bytearray(z)
    # an offset that reaches before the start of the output
    first_offset_at = 1 + (z[0] >> 4)
    bad[first_offset_at] = 0xff; bad[first_offset_at + 1] = 0xff
    assert decompress(lib, bytes(bad), len(data)) is None
    assert decompress(lib, b"\xf0", 100) is None                    # length extension runs off the end
    assert decompress(lib, b"\x10a\x00\x00", 100) is None           # offset 0


@pytest.mark.parametrize("checksum", [False, True])
def test_oracle_serde_lz4_pages(checksum):
    """PagesSerde.wrapSlice :118-143 in the oracle: compressible pages carry the COMPRESSED marker and a smaller slice, incompressible ones
    stay raw; both deserialize to the page; the compressed slice is a block liblz4 decodes to the raw page."""
    import oracle
    from presto_b200 import BIGINT, DOUBLE, Block, Page, PagesSerde
    from presto_b200 import _abi as abi
    be = oracle.backend()
    rng = np.random.default_rng(1)
    n = 20_000
    compressible = Page([Block.from_numpy(BIGINT, np.arange(n, dtype=np.int64) // 7), Block.from_strings([b"BUILDING", b"MACHINERY", None, b""] * (n // 4))])
    noise = Page([Block.from_numpy(DOUBLE, rng.random(n)), Block.from_numpy(BIGINT, rng.integers(-2**62, 2**62, n))])
    plain = PagesSerde(checksum_enabled=checksum, backend=be)
    lz = PagesSerde(checksum_enabled=checksum, compression=abi.COMPRESSION_LZ4, backend=be)
    for page, types, shrinks in ((compressible, [BIGINT, 6], True), (noise, [DOUBLE, BIGINT], False)):
        raw, wire = bytes(plain.serialize(page)), bytes(lz.serialize(page))
        markers = wire[4]
        assert bool(markers & abi.PAGE_MARKER_COMPRESSED) == shrinks
        assert wire[5:9] == raw[5:9]                                 # uncompressedSizeInBytes
        if shrinks:
            assert len(wire) < 0.8 * len(raw)
            assert pa.Codec("lz4_raw").decompress(wire[21:], decompressed_size=len(raw) - 21, asbytes=True) == raw[21:]
        else:
            assert wire[21:] == raw[21:]
        assert lz.deserialize(np.frombuffer(wire, dtype=np.uint8), types).rows() == page.rows()
        assert plain.deserialize(np.frombuffer(wire, dtype=np.uint8), types).rows() == page.rows()      # any serde READS compressed pages
    plain.close(); lz.close()


def test_corrupted_blocks_never_write_out_of_bounds(lib):
    """Fuzz: blocks with random byte flips / truncations decode to something or are refused, but never past the capacity (the output
    buffer is followed by a guard pattern) — the bounds checks of decodeBlock are the kernel's."""
    codec = pa.Codec("lz4_raw")
    rng = np.random.default_rng(2024)
    guard = b"\xa5" * 64
    for trial in range(300):
        n = int(rng.integers(1, 5000))
        alphabet = int(rng.integers(1, 40))
        data = bytes(rng.integers(0, alphabet, n).astype(np.uint8))
        z = bytearray(codec.compress(data, asbytes=True))
        for _ in range(int(rng.integers(1, 6))):
            z[int(rng.integers(0, len(z)))] = int(rng.integers(0, 256))
        if rng.random() < 0.3:
            z = z[:int(rng.integers(1, len(z) + 1))]
        cap = int(rng.integers(0, n + 10))
        out = C.create_string_buffer(cap + len(guard))
        C.memmove(C.addressof(out) + cap, guard, len(guard))
        got = lib.lz4_decompress_host(bytes(z), len(z), out, cap)
        assert -1 <= got <= cap
        assert out.raw[cap:cap + len(guard)] == guard, trial


def test_random_structured_data_round_trips_through_both_codecs(lib):
    codec = pa.Codec("lz4_raw")
    rng = np.random.default_rng(99)
    for trial in range(60):
        pieces = []
        for _ in range(int(rng.integers(1, 30))):
            kind = rng.integers(0, 4)
            m = int(rng.integers(1, 6000))
            if kind == 0: pieces.append(bytes(rng.integers(0, 256, m).astype(np.uint8)))
            elif kind == 1: pieces.append(bytes([int(rng.integers(0, 256))]) * m)
            elif kind == 2 and pieces: pieces.append(pieces[int(rng.integers(0, len(pieces)))][:m])
            else: pieces.append((b"%d|" % int(rng.integers(0, 1000))) * (m // 4 + 1))
        data = b"".join(pieces)
        z = compress(lib, data)
        assert codec.decompress(z, decompressed_size=len(data), asbytes=True) == data
        assert decompress(lib, codec.compress(data, asbytes=True), len(data)) == data
