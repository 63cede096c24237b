"""Test fixtures mirroring the reference's own: BlockAssertions, SequencePageBuilder, RowPagesBuilder, TypeUtils.getHashBlock
(presto-main-base/src/test/java/com/facebook/presto/{block/BlockAssertions.java, SequencePageBuilder.java, RowPagesBuilder.java})."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from presto_b200 import BIGINT, BOOLEAN, DATE, DOUBLE, INTEGER, VARCHAR, Block, Page
from presto_b200.hashing import get_hash_block


def create_longs_block(*values) -> Block:
    if len(values) == 1 and isinstance(values[0], (list, tuple)):
        values = values[0]
    return Block.from_values(BIGINT, list(values))


def create_long_sequence_block(start: int, end: int) -> Block:
    return Block.from_numpy(BIGINT, np.arange(start, end, dtype=np.int64))


def create_double_sequence_block(start: int, end: int) -> Block:
    return Block.from_numpy(DOUBLE, np.arange(start, end, dtype=np.float64))


def create_doubles_block(*values) -> Block:
    return Block.from_values(DOUBLE, list(values))


def create_string_sequence_block(start: int, end: int) -> Block:
    return Block.from_strings([str(i).encode() for i in range(start, end)])


def create_strings_block(*values) -> Block:
    return Block.from_strings(list(values))


def create_sequence_page(types: Sequence[int], length: int, *initial_values: int) -> Page:
    """SequencePageBuilder.createSequencePage."""
    if not initial_values:
        initial_values = tuple([0] * len(types))
    blocks = []
    for t, init in zip(types, initial_values):
        if t == BIGINT:
            blocks.append(create_long_sequence_block(init, init + length))
        elif t == DOUBLE:
            blocks.append(create_double_sequence_block(init, init + length))
        elif t == VARCHAR:
            blocks.append(create_string_sequence_block(init, init + length))
        elif t == BOOLEAN:
            blocks.append(Block.from_numpy(BOOLEAN, (np.arange(init, init + length) % 2 == 0).astype(np.int8)))
        elif t in (DATE, INTEGER):
            blocks.append(Block.from_numpy(t, np.arange(init, init + length, dtype=np.int32)))
        else:
            raise ValueError(t)
    return Page(blocks)


class RowPagesBuilder:
    """RowPagesBuilder: pages of rows with an optional trailing $hashvalue channel over hash_channels."""

    def __init__(self, types: Sequence[int], hash_enabled: bool = False, hash_channels: Optional[Sequence[int]] = None, backend=None):
        self.types = list(types)
        self.hash_enabled = hash_enabled
        self.hash_channels = list(hash_channels) if hash_channels is not None else list(range(len(types)))
        self.pages: List[Page] = []
        self.rows: List[tuple] = []
        self.backend = backend

    def row(self, *values) -> "RowPagesBuilder":
        self.rows.append(tuple(values))
        return self

    def page_break(self) -> "RowPagesBuilder":
        if self.rows:
            cols = list(zip(*self.rows))
            self.pages.append(Page([Block.from_values(t, list(c)) for t, c in zip(self.types, cols)]))
            self.rows = []
        return self

    def add_sequence_page(self, length: int, *initial_values: int) -> "RowPagesBuilder":
        self.page_break()
        self.pages.append(create_sequence_page(self.types, length, *initial_values))
        return self

    def add_block_page(self, *blocks: Block) -> "RowPagesBuilder":
        self.page_break()
        self.pages.append(Page(list(blocks)))
        return self

    def get_types(self) -> List[int]:
        return self.types + [BIGINT] if self.hash_enabled else list(self.types)

    def get_hash_channel(self) -> int:
        return len(self.types) if self.hash_enabled else -1

    def build(self) -> List[Page]:
        self.page_break()
        if not self.hash_enabled:
            return list(self.pages)
        out = []
        for p in self.pages:
            hb = get_hash_block([flatten(p.blocks[c]) for c in self.hash_channels], self.backend)
            out.append(Page(p.blocks + [hb]))
        return out


def flatten(b: Block) -> Block:
    if b.encoding == 0:
        return b
    return Block.from_values(b.type, b.to_pylist())


def rows_of(pages: Sequence[Page]) -> List[tuple]:
    out: List[tuple] = []
    for p in pages:
        if p.on_device:
            from presto_b200.exchange import page_to_host
            p = page_to_host(p)
        out.extend(p.rows())
    return out


def rows_close(actual: Sequence[tuple], expected: Sequence[tuple], rel: float = 1e-6) -> None:
    """Bit-exact for ints / bytes / None; doubles within `rel` relative (north_star: 1e-6 for double aggregates)."""
    assert len(actual) == len(expected), f"row count {len(actual)} != {len(expected)}"
    for i, (a, e) in enumerate(zip(actual, expected)):
        assert len(a) == len(e), f"row {i}: width {len(a)} != {len(e)}"
        for j, (x, y) in enumerate(zip(a, e)):
            if isinstance(y, float) and isinstance(x, float):
                if math.isnan(y):
                    assert math.isnan(x), f"row {i} col {j}: {x} != NaN"
                else:
                    assert abs(x - y) <= rel * max(abs(x), abs(y)) or x == y, f"row {i} col {j}: {x} != {y}"
            else:
                assert x == y and type(x) is type(y), f"row {i} col {j}: {x!r} != {y!r}"


def page_to_device(backend, page: Page, keep: list) -> Page:
    """A host page's FLAT blocks copied into device memory (values, null map, VARCHAR offsets); `keep` collects the allocations."""
    import ctypes as C

    def up(arr):
        p = C.c_void_p()
        backend.check(backend.device_alloc(max(arr.nbytes, 16), C.byref(p)))
        if arr.nbytes:
            backend.check(backend.memcpy_h2d(p, arr.ctypes.data, arr.nbytes))
        keep.append(p)
        return p.value

    blocks = []
    for b in page.blocks:
        blocks.append(Block.device(b.type, b.position_count, up(b.values), up(b.is_null) if b.is_null is not None else 0,
                                   up(b.offsets) if b.offsets is not None else 0))
    return Page(blocks, page.position_count)
