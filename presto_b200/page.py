"""Page / Block — the host-side mirror of presto-common's columnar containers.

Reference: presto-common/src/main/java/com/facebook/presto/common/Page.java:45-498 and block/{LongArrayBlock.java:37-52,
IntArrayBlock.java:43-47, ByteArrayBlock.java:50-54, VariableWidthBlock.java:53-58, DictionaryBlock.java:59-64,
RunLengthEncodedBlock.java:48-49}.  A Block here is a thin owner of numpy arrays (host) or raw device pointers that knows
how to describe itself as a ``pg_block`` for the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _abi as abi

_NP_DTYPE = {abi.BIGINT: np.int64, abi.DOUBLE: np.float64, abi.INTEGER: np.int32, abi.DATE: np.int32, abi.BOOLEAN: np.int8,
             abi.LONG_DECIMAL: np.uint64}     # LONG_DECIMAL: two uint64 per position (low, high | sign << 63)
TYPE_WIDTH = {abi.BIGINT: 8, abi.DOUBLE: 8, abi.INTEGER: 4, abi.DATE: 4, abi.BOOLEAN: 1, abi.VARCHAR: 0, abi.LONG_DECIMAL: 16}


def _ptr(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


@dataclass
class Block:
    """One column.  ``values`` is the flat storage (or the dictionary's storage for DICTIONARY/RLE encodings)."""

    type: int
    position_count: int
    values: Optional[np.ndarray] = None          # fixed width: typed array; VARCHAR: uint8 bytes
    is_null: Optional[np.ndarray] = None         # uint8 per position (of the dictionary when encoded)
    offsets: Optional[np.ndarray] = None         # VARCHAR: int32[n+1]
    ids: Optional[np.ndarray] = None             # DICTIONARY: int32[position_count]
    encoding: int = abi.ENC_FLAT
    dictionary_size: int = 0
    # device-resident blocks carry raw addresses instead of numpy arrays
    device_ptrs: Optional[dict] = None
    _keepalive: list = field(default_factory=list, repr=False)

    # ---- constructors ------------------------------------------------------------------------------------------
    @staticmethod
    def from_values(type_: int, values: Sequence, nulls: Optional[Sequence[bool]] = None) -> "Block":
        """BlockAssertions.createLongsBlock & friends: python values with None for NULL."""
        if type_ == abi.VARCHAR:
            return Block.from_strings(values)
        vals = list(values)
        n = len(vals)
        is_null = np.zeros(n, dtype=np.uint8)
        arr = np.zeros(n, dtype=_NP_DTYPE[type_])
        for i, v in enumerate(vals):
            if v is None:
                is_null[i] = 1
            else:
                arr[i] = v
        if nulls is not None:
            is_null = np.asarray(nulls, dtype=np.uint8)
        return Block(type_, n, arr, is_null if is_null.any() else None)

    @staticmethod
    def from_numpy(type_: int, values: np.ndarray, is_null: Optional[np.ndarray] = None) -> "Block":
        values = np.ascontiguousarray(values, dtype=_NP_DTYPE[type_])
        if is_null is not None:
            is_null = np.ascontiguousarray(is_null, dtype=np.uint8)
        return Block(type_, int(values.shape[0]), values, is_null)

    @staticmethod
    def from_strings(strings: Sequence[Optional[bytes]]) -> "Block":
        n = len(strings)
        offsets = np.zeros(n + 1, dtype=np.int32)
        is_null = np.zeros(n, dtype=np.uint8)
        chunks = []
        total = 0
        for i, s in enumerate(strings):
            if s is None:
                is_null[i] = 1
            else:
                if isinstance(s, str):
                    s = s.encode()
                chunks.append(s)
                total += len(s)
            offsets[i + 1] = total
        data = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy() if total else np.zeros(1, dtype=np.uint8)
        return Block(abi.VARCHAR, n, data, is_null if is_null.any() else None, offsets)

    @staticmethod
    def from_decimals(values: Sequence[Optional[int]]) -> "Block":
        """Int128ArrayBlock of unscaled decimal(19..38) values (python ints, None for NULL): UnscaledDecimal128Arithmetic's sign-magnitude
        layout, low 64 bits then high 63 bits | sign << 63 (UnscaledDecimal128Arithmetic.java:172-180)."""
        n = len(values)
        arr = np.zeros(2 * n, dtype=np.uint64)
        is_null = np.zeros(n, dtype=np.uint8)
        for i, v in enumerate(values):
            if v is None:
                is_null[i] = 1
                continue
            m = abs(int(v))
            assert m < (1 << 127), "magnitude needs 127 bits"
            arr[2 * i] = m & 0xFFFFFFFFFFFFFFFF
            arr[2 * i + 1] = (m >> 64) | ((1 << 63) if v < 0 else 0)
        return Block(abi.LONG_DECIMAL, n, arr, is_null if is_null.any() else None)

    @staticmethod
    def dictionary(dictionary: "Block", ids: np.ndarray) -> "Block":
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        return Block(dictionary.type, int(ids.shape[0]), dictionary.values, dictionary.is_null, dictionary.offsets, ids,
                     abi.ENC_DICTIONARY, dictionary.position_count)

    @staticmethod
    def rle(value: "Block", position_count: int) -> "Block":
        assert value.position_count == 1
        return Block(value.type, position_count, value.values, value.is_null, value.offsets, None, abi.ENC_RLE, 1)

    @staticmethod
    def device(type_: int, position_count: int, values_ptr: int, is_null_ptr: int = 0, offsets_ptr: int = 0, keepalive=None) -> "Block":
        b = Block(type_, position_count)
        b.device_ptrs = {"values": values_ptr, "is_null": is_null_ptr or None, "offsets": offsets_ptr or None}
        if keepalive is not None:
            b._keepalive.append(keepalive)
        return b

    # ---- accessors ---------------------------------------------------------------------------------------------
    @property
    def on_device(self) -> bool:
        return self.device_ptrs is not None

    def flat_position(self, position: int) -> int:
        if self.encoding == abi.ENC_FLAT:
            return position
        if self.encoding == abi.ENC_RLE:
            return 0
        return int(self.ids[position])

    def to_pylist(self) -> list:
        """Materialise as python objects (None for NULL) — the MaterializedResult view the reference's tests compare."""
        assert not self.on_device
        out = []
        for p in range(self.position_count):
            q = self.flat_position(p)
            if self.is_null is not None and self.is_null[q]:
                out.append(None)
            elif self.type == abi.VARCHAR:
                out.append(bytes(self.values[self.offsets[q]:self.offsets[q + 1]]))
            elif self.type == abi.LONG_DECIMAL:
                low, high = int(self.values[2 * q]), int(self.values[2 * q + 1])
                m = ((high & 0x7FFFFFFFFFFFFFFF) << 64) | low
                out.append(-m if high >> 63 else m)
            elif self.type == abi.DOUBLE:
                out.append(float(self.values[q]))
            elif self.type == abi.BOOLEAN:
                out.append(bool(self.values[q]))
            else:
                out.append(int(self.values[q]))
        return out

    def fill(self, blk: abi.pg_block) -> None:
        blk.type = self.type
        blk.encoding = self.encoding
        blk.position_count = self.position_count
        blk.dictionary_size = self.dictionary_size
        if self.on_device:
            blk.values = self.device_ptrs["values"]
            blk.is_null = self.device_ptrs["is_null"]
            blk.offsets = self.device_ptrs["offsets"]
            blk.ids = None
        else:
            blk.values = _ptr(self.values)
            blk.is_null = _ptr(self.is_null)
            blk.offsets = _ptr(self.offsets)
            blk.ids = _ptr(self.ids)


class Page:
    """A batch of rows = a list of equally long Blocks (Page.java:45-498)."""

    def __init__(self, blocks: Sequence[Block], position_count: Optional[int] = None):
        self.blocks: List[Block] = list(blocks)
        if position_count is None:
            position_count = self.blocks[0].position_count if self.blocks else 0
        self.position_count = position_count
        for b in self.blocks:
            assert b.position_count == position_count, "blocks of a page must have the same position count"

    @property
    def channel_count(self) -> int:
        return len(self.blocks)

    @property
    def on_device(self) -> bool:
        return bool(self.blocks) and self.blocks[0].on_device

    def get_block(self, channel: int) -> Block:
        return self.blocks[channel]

    def as_c(self):
        """Returns (pg_page, keepalive) describing this page for the C ABI."""
        arr = (abi.pg_block * max(1, len(self.blocks)))()
        for i, b in enumerate(self.blocks):
            b.fill(arr[i])
        page = abi.pg_page()
        page.position_count = self.position_count
        page.channel_count = len(self.blocks)
        page.memory = abi.MEM_DEVICE if self.on_device else abi.MEM_HOST
        page.blocks = C.cast(arr, C.POINTER(abi.pg_block))
        return page, arr

    def rows(self) -> List[tuple]:
        cols = [b.to_pylist() for b in self.blocks]
        return [tuple(c[i] for c in cols) for i in range(self.position_count)]

    @staticmethod
    def from_c(page: abi.pg_page, backend=None, copy: bool = True) -> Optional["Page"]:
        """Copy a library-owned output page into numpy-owned Blocks (device pages are brought back with memcpy_d2h)."""
        if page.channel_count == 0:
            return None
        n = int(page.position_count)
        blocks = []
        for ch in range(page.channel_count):
            b = page.blocks[ch]
            t = int(b.type)
            if page.memory == abi.MEM_DEVICE and not copy:
                blocks.append(Block.device(t, n, b.values or 0, b.is_null or 0, b.offsets or 0))
                continue

            def fetch(ptr, nbytes, dtype):
                out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
                if nbytes == 0:
                    return out
                if page.memory == abi.MEM_DEVICE:
                    backend.check(backend.memcpy_d2h(out.ctypes.data, ptr, nbytes))
                else:
                    C.memmove(out.ctypes.data, ptr, nbytes)
                return out

            # storage positions: the block's own rows (FLAT), the dictionary's entries, or the single RLE value
            enc = int(b.encoding)
            m = n if enc == abi.ENC_FLAT else (int(b.dictionary_size) if enc == abi.ENC_DICTIONARY else 1)
            is_null = fetch(b.is_null, m, np.uint8) if b.is_null else None
            if t == abi.VARCHAR:
                offsets = fetch(b.offsets, 4 * (m + 1), np.int32)
                total = int(offsets[-1]) if m else 0
                data = fetch(b.values, total, np.uint8) if total else np.zeros(1, dtype=np.uint8)
                storage = Block(t, m, data, is_null, offsets)
            else:
                vals = fetch(b.values, m * TYPE_WIDTH[t], _NP_DTYPE[t])
                storage = Block(t, m, vals, is_null)
            if enc == abi.ENC_DICTIONARY:
                blocks.append(Block.dictionary(storage, fetch(b.ids, 4 * int(b.position_count), np.int32)))
            elif enc == abi.ENC_RLE:
                blocks.append(Block.rle(storage, int(b.position_count)))
            else:
                blocks.append(storage)
        return Page(blocks, n)


def concat_pages(pages: Sequence[Page]) -> List[tuple]:
    rows: List[tuple] = []
    for p in pages:
        rows.extend(p.rows())
    return rows
