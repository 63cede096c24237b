"""The reference's own behavioural tests for the hot path, re-encoded backend-agnostically.

Each ``case_*`` takes a Backend (the CPU oracle in test_oracle_reference.py, the CUDA library in test_gpu_reference.py)
and asserts exactly what the cited Java test asserts.  MBT = presto-main-base/src/test/java/com/facebook/presto/.
Adaptations are stated per case (the GPU join path takes BIGINT/INTEGER/DATE keys, so VARCHAR join keys become BIGINT).
"""
from __future__ import annotations

import numpy as np
import pytest

from presto_b200 import (
    AGG_AVG_BIGINT, AGG_AVG_DOUBLE, AGG_COUNT_COLUMN, AGG_COUNT_STAR, AGG_SUM_BIGINT, AGG_SUM_DOUBLE, BIGINT, BOOLEAN, DATE, DOUBLE,
    INTEGER, JOIN_INNER, JOIN_PROBE_OUTER, VARCHAR, Aggregate, Block, FilterAndProjectOperator, GroupByHash,
    HashAggregationOperator, HashBuilderOperator, JoinBridge, LookupJoinOperator, Page, PartitionedOutputOperator, Program, to_pages,
)
from presto_b200.hashing import get_hash_block, get_partition, hash_long, local_partition

from util import (
    RowPagesBuilder, create_double_sequence_block, create_doubles_block, create_long_sequence_block, create_longs_block,
    create_string_sequence_block, create_strings_block, rows_close, rows_of,
)

MAX_GROUP_ID = 500   # MBT/operator/TestGroupByHash.java:58


# ---------------------------------------------------------------------------------------------------------------------
# TestGroupByHash
# ---------------------------------------------------------------------------------------------------------------------
def case_group_by_hash_add_page(backend, max_group_id=MAX_GROUP_ID, tries_per_value=10):
    """TestGroupByHash.testAddPage :68-94 — sequential values get groupId == value; re-adding changes nothing."""
    h = GroupByHash([BIGINT], [0], hash_channel=1, expected_size=100, backend=backend)
    for tries in range(2):
        for value in range(max_group_id):
            block = create_longs_block(value)
            page = Page([block, get_hash_block([block], backend)])
            for _ in range(tries_per_value):
                h.add_page(page)
                assert h.get_group_count() == (value + 1 if tries == 0 else max_group_id)
                ids = h.get_group_ids(page)
                assert h.get_group_count() == (value + 1 if tries == 0 else max_group_id)
                assert len(ids) == 1 and ids[0] == value
    h.close()


def case_group_by_hash_null_group(backend):
    """TestGroupByHash.testNullGroup :97-117 — null group, then 132 748 values forcing rehashes; contains(0) is false."""
    h = GroupByHash([BIGINT], [0], hash_channel=1, expected_size=100, backend=backend)
    block = create_longs_block(None)
    h.add_page(Page([block, get_hash_block([block], backend)]))
    block = create_long_sequence_block(1, 132748)
    h.add_page(Page([block, get_hash_block([block], backend)]))
    assert h.get_group_count() == 132748
    block = create_longs_block(0)
    assert not h.contains(Page([block, get_hash_block([block], backend)]))[0]
    block = create_longs_block(None, 5, 132747, 132748)
    assert list(h.contains(Page([block, get_hash_block([block], backend)]))) == [True, True, True, False]
    h.close()


def case_group_by_hash_get_group_ids(backend, max_group_id=MAX_GROUP_ID, tries_per_value=3):
    """TestGroupByHash.testGetGroupIds :119-139."""
    h = GroupByHash([BIGINT], [0], hash_channel=1, expected_size=100, backend=backend)
    for tries in range(2):
        for value in range(max_group_id):
            block = create_longs_block(value)
            page = Page([block, get_hash_block([block], backend)])
            for _ in range(tries_per_value):
                ids = h.get_group_ids(page)
                assert h.get_group_count() == (value + 1 if tries == 0 else max_group_id)
                assert len(ids) == 1 and ids[0] == value
    h.close()


def case_group_by_hash_append_to(backend, max_len=7):
    """TestGroupByHash.testAppendTo :149-176 — VARCHAR keys '0'..'99' + $hashvalue come back in group-id order.  max_len = 0 is the
    reference's unbounded VARCHAR (on the GPU: keys interned to value ids), 7 a varchar(7) that fits a packed key field."""
    values = create_string_sequence_block(0, 100)
    hash_block = get_hash_block([values], backend)
    h = GroupByHash([VARCHAR], [0], hash_channel=1, expected_size=100, max_lengths=[max_len], backend=backend)
    ids = h.get_group_ids(Page([values, hash_block]))
    assert list(ids) == list(range(100))
    assert h.get_group_count() == 100
    keys = h.get_keys()
    assert keys.position_count == 100 and keys.channel_count == 2
    assert keys.blocks[0].to_pylist() == values.to_pylist()
    assert keys.blocks[1].to_pylist() == hash_block.to_pylist()
    h.close()


def case_group_by_hash_append_to_multiple_tuples_per_group(backend):
    """TestGroupByHash.testAppendToMultipleTuplesPerGroup :178-200."""
    values = create_longs_block([i % 50 for i in range(100)])
    h = GroupByHash([BIGINT], [0], hash_channel=1, expected_size=100, backend=backend)
    ids = h.get_group_ids(Page([values, get_hash_block([values], backend)]))
    assert list(ids) == [i % 50 for i in range(100)]
    assert h.get_group_count() == 50
    keys = h.get_keys()
    assert keys.blocks[0].to_pylist() == list(range(50))
    assert keys.blocks[1].to_pylist() == [int(x) for x in hash_long(np.arange(50, dtype=np.int64))]
    h.close()


def case_group_by_hash_contains(backend):
    """TestGroupByHash.testContains :202-217 (DOUBLE key)."""
    values = create_double_sequence_block(0, 10)
    h = GroupByHash([DOUBLE], [0], hash_channel=1, expected_size=100, backend=backend)
    h.get_group_ids(Page([values, get_hash_block([values], backend)]))
    t = create_doubles_block(3.0)
    assert h.contains(Page([t, get_hash_block([t], backend)]))[0]
    t = create_doubles_block(11.0)
    assert not h.contains(Page([t, get_hash_block([t], backend)]))[0]
    h.close()


def case_group_by_hash_contains_multiple_columns(backend, max_len=7):
    """TestGroupByHash.testContainsMultipleColumns :219-234 (DOUBLE, VARCHAR)."""
    values, strings = create_double_sequence_block(0, 10), create_string_sequence_block(0, 10)
    h = GroupByHash([DOUBLE, VARCHAR], [0, 1], hash_channel=2, expected_size=100, max_lengths=[0, max_len], backend=backend)
    h.get_group_ids(Page([values, strings, get_hash_block([values, strings], backend)]))
    tv, ts = create_doubles_block(3.0), create_strings_block(b"3")
    assert h.contains(Page([tv, ts, get_hash_block([tv, ts], backend)]))[0]
    tv, ts = create_doubles_block(3.0), create_strings_block(b"4")
    assert not h.contains(Page([tv, ts, get_hash_block([tv, ts], backend)]))[0]
    h.close()


def case_group_by_hash_force_rehash(backend, max_len=7):
    """TestGroupByHash.testForceRehash :236-250 — expectedSize 4, 100 VARCHAR keys."""
    values = create_string_sequence_block(0, 100)
    page = Page([values, get_hash_block([values], backend)])
    h = GroupByHash([VARCHAR], [0], hash_channel=1, expected_size=4, max_lengths=[max_len], backend=backend)
    ids = h.get_group_ids(page)
    assert list(ids) == list(range(100))
    assert h.contains(page).all()
    h.close()


# ---------------------------------------------------------------------------------------------------------------------
# TestHashAggregationOperator
# ---------------------------------------------------------------------------------------------------------------------
def case_hash_aggregation(backend, hash_enabled: bool, number_of_rows: int = 40_000, max_len: int = 7):
    """TestHashAggregationOperator.testHashAggregation :162-217 — 3 sequence pages of 40 000 rows, VARCHAR group key;
    count(*), sum(bigint), avg(bigint), count(varchar), count(boolean).  Adaptation: max(varchar) is dropped (not on the GPU
    path); the result is compared in group-id order (stricter than the reference's assertPagesEqualIgnoreOrder).  max_len = 0: the
    reference's unbounded VARCHAR key type."""
    b = RowPagesBuilder([VARCHAR, VARCHAR, VARCHAR, BIGINT, BOOLEAN], hash_enabled, [1], backend)
    pages = (b.add_sequence_page(number_of_rows, 100, 0, 100_000, 0, 500)
              .add_sequence_page(number_of_rows, 100, 0, 200_000, 0, 500)
              .add_sequence_page(number_of_rows, 100, 0, 300_000, 0, 500).build())
    op = HashAggregationOperator(
        [VARCHAR], [1],
        [Aggregate(AGG_COUNT_STAR), Aggregate(AGG_SUM_BIGINT, 3), Aggregate(AGG_AVG_BIGINT, 3), Aggregate(AGG_COUNT_COLUMN, 0), Aggregate(AGG_COUNT_COLUMN, 4)],
        hash_channel=b.get_hash_channel(), expected_groups=100_000, group_max_lengths=[max_len], max_output_rows=16384, backend=backend)
    out = to_pages(op, pages)
    assert len(out) > 1, "Expected more than one output page"
    rows = rows_of(out)
    expected = []
    for i in range(number_of_rows):
        key = str(i).encode()
        row = (key,)
        if hash_enabled:
            row += (backend.hash_xxhash64(key, len(key)),)
        expected.append(row + (3, 3 * i, float(i), 3, 3))
    rows_close(rows, expected)
    op.close()


def case_hash_aggregation_null_inputs(backend):
    """DoubleSumAggregation / AverageAggregations / CountColumn null handling (AccumulatorCompiler.java:462-566): NULL inputs are
    skipped; sum over no non-null rows is NULL, avg is NULL, count is 0."""
    keys = Block.from_values(BIGINT, [1, 2, 1, None, 2, None, 3])
    vals = Block.from_values(DOUBLE, [1.5, None, 2.5, 4.0, None, 6.0, None])
    longs = Block.from_values(BIGINT, [10, 20, None, 40, None, 60, None])
    op = HashAggregationOperator(
        [BIGINT], [0],
        [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_AVG_DOUBLE, 1), Aggregate(AGG_COUNT_COLUMN, 1), Aggregate(AGG_SUM_BIGINT, 2),
         Aggregate(AGG_AVG_BIGINT, 2), Aggregate(AGG_COUNT_STAR)],
        expected_groups=16, backend=backend)
    rows = rows_of(to_pages(op, [Page([keys, vals, longs])]))
    rows_close(rows, [
        (1, 4.0, 2.0, 2, 10, 10.0, 2),
        (2, None, None, 0, 20, 20.0, 2),
        (None, 10.0, 5.0, 2, 100, 50.0, 2),
        (3, None, None, 0, None, None, 1),
    ])
    op.close()


def case_hash_aggregation_mask(backend):
    """Mask channel handling (TestHashAggregationOperator.testMask): rows whose BOOLEAN mask is false/NULL are not accumulated,
    but their group still exists."""
    keys = Block.from_values(BIGINT, [1, 1, 2, 2, 3])
    vals = Block.from_values(BIGINT, [5, 6, 7, 8, 9])
    mask = Block.from_values(BOOLEAN, [True, False, False, None, True])
    op = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_COUNT_STAR, -1, 2), Aggregate(AGG_SUM_BIGINT, 1, 2)], expected_groups=8, backend=backend)
    rows = rows_of(to_pages(op, [Page([keys, vals, mask])]))
    rows_close(rows, [(1, 1, 5), (2, 0, None), (3, 1, 9)])
    op.close()


def case_hash_aggregation_multi_key(backend):
    """MultiChannelGroupByHash with (BIGINT, DATE, INTEGER) keys — the Q3 shape — null == null grouping
    (MultiChannelGroupByHash.java:460-466) and first-appearance output order."""
    k0 = Block.from_values(BIGINT, [7, 7, 8, None, 7, None, 8])
    k1 = Block.from_values(DATE, [100, 100, 100, 5, 101, 5, 100])
    k2 = Block.from_values(INTEGER, [0, 0, None, 1, 0, 1, None])
    v = Block.from_values(DOUBLE, [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0])
    op = HashAggregationOperator([BIGINT, DATE, INTEGER], [0, 1, 2], [Aggregate(AGG_SUM_DOUBLE, 3), Aggregate(AGG_COUNT_STAR)], expected_groups=8, backend=backend)
    rows = rows_of(to_pages(op, [Page([k0, k1, k2, v])]))
    rows_close(rows, [(7, 100, 0, 3.0, 2), (8, 100, None, 10.0, 2), (None, 5, 1, 10.0, 2), (7, 101, 0, 5.0, 1)])
    op.close()


def case_hash_aggregation_empty(backend):
    """No input → no output page, operator finishes (HashAggregationOperator.java:451-501 with a null builder)."""
    op = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_COUNT_STAR)], backend=backend)
    assert to_pages(op, []) == []
    op.close()


def case_hash_aggregation_add_after_finish(backend):
    """checkState(!finishing) in addInput (HashAggregationOperator.java:415)."""
    from presto_b200 import PrestoException
    op = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_COUNT_STAR)], backend=backend)
    op.add_input(Page([create_longs_block(1, 2)]))
    op.finish()
    assert not op.needs_input()
    try:
        op.add_input(Page([create_longs_block(3)]))
    except PrestoException as e:
        assert e.code == 2
    else:
        raise AssertionError("addInput after finish must fail")
    op.close()


def case_bigint_sum_overflow(backend):
    """LongSumAggregation → BigintOperators.add → Math.addExact (BigintOperators.java:68-78): NUMERIC_VALUE_OUT_OF_RANGE."""
    from presto_b200 import PrestoException
    keys = create_longs_block(1, 1)
    vals = create_longs_block(2 ** 63 - 1, 1)
    op = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_BIGINT, 1)], backend=backend)
    try:
        to_pages(op, [Page([keys, vals])])
    except PrestoException as e:
        assert e.code == 6
    else:
        raise AssertionError("bigint sum overflow must raise")
    op.close()


# ---------------------------------------------------------------------------------------------------------------------
# TestHashJoinOperator (VARCHAR join keys of the reference become BIGINT: 'a','b','c' → 1,2,3; "20".."29" → 20..29)
# ---------------------------------------------------------------------------------------------------------------------
def _join(backend, build_pages, build_types, probe_pages, probe_types, join_type, build_hash, probe_hash, filter_program=None):
    bridge = JoinBridge(backend=backend)
    nb, npr = len(build_types) - (1 if build_hash else 0), len(probe_types) - (1 if probe_hash else 0)
    builder = HashBuilderOperator(bridge, build_types, [0], list(range(nb)), precomputed_hash_channel=nb if build_hash else -1, backend=backend,
                                  filter_program=filter_program)
    probe = LookupJoinOperator(bridge, probe_types, [0], list(range(npr)), join_type=join_type,
                               precomputed_hash_channel=npr if probe_hash else -1, backend=backend)
    assert not probe.needs_input(), "probe must wait for the lookup source"
    assert to_pages(builder, build_pages) == []
    out = to_pages(probe, probe_pages)
    rows = rows_of(out)
    probe.close(); builder.close(); bridge.close()
    return rows


def _key(key_type, v, sequence=False):
    """The reference's join tests key on VARCHAR ("20".."29", 'a' 'b' 'c'); the BIGINT variants of the cases keep the numbers."""
    if key_type != VARCHAR or v is None:
        return v
    return str(v).encode() if sequence else {1: b"a", 2: b"b", 3: b"c"}[v]


def case_inner_join(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """TestHashJoinOperator.testInnerJoin :171-207 — ordered expectation (key_type VARCHAR: the reference's own column types)."""
    bp = RowPagesBuilder([key_type, BIGINT, BIGINT], build_hash, [0], backend).add_sequence_page(10, 20, 30, 40)
    pp = RowPagesBuilder([key_type, BIGINT, BIGINT], probe_hash, [0], backend).add_sequence_page(1000, 0, 1000, 2000)
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_INNER, build_hash, probe_hash)
    k = lambda v: _key(key_type, v, True)
    assert rows == [(k(20 + i), 1020 + i, 2020 + i, k(20 + i), 30 + i, 40 + i) for i in range(10)]


def case_inner_join_with_null_probe(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """testInnerJoinWithNullProbe :712-750."""
    bp = RowPagesBuilder([key_type], build_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, 3))
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, None)).row(_key(key_type, None)).row(_key(key_type, 1)).row(_key(key_type, 2))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_INNER, build_hash, probe_hash)
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (1, 1), (2, 2)]]


def case_inner_join_with_null_build(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """testInnerJoinWithNullBuild :752-790 — build a,null,null,a,b ; probe a,b,c → (a,a),(a,a),(b,b)."""
    bp = RowPagesBuilder([key_type], build_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, None)).row(_key(key_type, None)).row(_key(key_type, 1)).row(_key(key_type, 2))
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, 3))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_INNER, build_hash, probe_hash)
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (1, 1), (2, 2)]]


def case_inner_join_with_null_on_both_sides(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """testInnerJoinWithNullOnBothSides :792-830."""
    bp = RowPagesBuilder([key_type], build_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, None)).row(_key(key_type, None)).row(_key(key_type, 1)).row(_key(key_type, 2))
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, None)).row(_key(key_type, 3))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_INNER, build_hash, probe_hash)
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (1, 1), (2, 2)]]


def case_probe_outer_join(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """testProbeOuterJoin :833-876."""
    bp = RowPagesBuilder([key_type, BIGINT, BIGINT], build_hash, [0], backend).add_sequence_page(10, 20, 30, 40)
    pp = RowPagesBuilder([key_type, BIGINT, BIGINT], probe_hash, [0], backend).add_sequence_page(15, 20, 1020, 2020)
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_PROBE_OUTER, build_hash, probe_hash)
    k = lambda v: _key(key_type, v, True)
    expected = [(k(20 + i), 1020 + i, 2020 + i, k(20 + i), 30 + i, 40 + i) for i in range(10)]
    expected += [(k(30 + i), 1030 + i, 2030 + i, None, None, None) for i in range(5)]
    assert rows == expected


def case_outer_join_with_null_on_both_sides(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """testOuterJoinWithNullOnBothSides :1101-1141 — probe a,b,null,c over build a,null,null,a,b."""
    bp = RowPagesBuilder([key_type], build_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, None)).row(_key(key_type, None)).row(_key(key_type, 1)).row(_key(key_type, 2))
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, None)).row(_key(key_type, 3))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_PROBE_OUTER, build_hash, probe_hash)
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (1, 1), (2, 2), (None, None), (3, None)]]


def case_outer_join_with_null_probe(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """testOuterJoinWithNullProbe :927-968 — build a,b,c ; probe a,null,null,a,b (PROBE_OUTER): NULL probe keys come out with a NULL build side."""
    bp = RowPagesBuilder([key_type], build_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, 3))
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(None).row(None).row(_key(key_type, 1)).row(_key(key_type, 2))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_PROBE_OUTER, build_hash, probe_hash)
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (None, None), (None, None), (1, 1), (2, 2)]]


def case_outer_join_with_null_build(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """testOuterJoinWithNullBuild :1014-1054 — build a,null,null,a,b ; probe a,b,c (PROBE_OUTER) → (a,a),(a,a),(b,b),(c,null)."""
    bp = RowPagesBuilder([key_type], build_hash, [0], backend).row(_key(key_type, 1)).row(None).row(None).row(_key(key_type, 1)).row(_key(key_type, 2))
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, 3))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_PROBE_OUTER, build_hash, probe_hash)
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (1, 1), (2, 2), (3, None)]]


def case_inner_join_with_non_empty_lookup_source_and_empty_probe(backend, key_type: int = BIGINT):
    """testInnerJoinWithNonEmptyLookupSourceAndEmptyProbe :1414-1452 — build a,b,null,c ; no probe page at all: no output, the operator finishes."""
    bp = RowPagesBuilder([key_type], False, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(None).row(_key(key_type, 3))
    assert _join(backend, bp.build(), bp.get_types(), [], [key_type], JOIN_INNER, False, False) == []


def _join_filter(build_types, probe_types, make):
    """A JoinFilterFunction: a Program over the build channels followed by the probe channels (createJoinSourcesLayout); `make(build,
    probe)` gets the two lists of input expressions and returns the BOOLEAN filter."""
    from presto_b200 import Program
    prog = Program(list(build_types) + list(probe_types))
    ins = [prog.input(i) for i in range(len(build_types) + len(probe_types))]
    prog.set_filter(make(ins[:len(build_types)], ins[len(build_types):]))
    return prog


def case_probe_outer_join_with_filter_function(backend, probe_hash: bool, build_hash: bool):
    """testProbeOuterJoinWithFilterFunction :879-925 — filter `probe.channel1 >= 1025`: matches below it lose their build side."""
    bp = RowPagesBuilder([BIGINT, BIGINT, BIGINT], build_hash, [0], backend).add_sequence_page(10, 20, 30, 40)
    pp = RowPagesBuilder([BIGINT, BIGINT, BIGINT], probe_hash, [0], backend).add_sequence_page(15, 20, 1020, 2020)
    f = _join_filter(bp.get_types(), pp.get_types(), lambda build, probe: probe[1] >= 1025)
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_PROBE_OUTER, build_hash, probe_hash, filter_program=f)
    expected = [(20 + i, 1020 + i, 2020 + i, None, None, None) for i in range(5)]
    expected += [(25 + i, 1025 + i, 2025 + i, 25 + i, 35 + i, 45 + i) for i in range(5)]
    expected += [(30 + i, 1030 + i, 2030 + i, None, None, None) for i in range(5)]
    assert rows == expected
    # the same filter on an INNER join keeps the five eligible matches only
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_INNER, build_hash, probe_hash, filter_program=f)
    assert rows == [(25 + i, 1025 + i, 2025 + i, 25 + i, 35 + i, 45 + i) for i in range(5)]


def case_outer_join_with_nulls_and_filter_function(backend, probe_hash: bool, build_hash: bool, key_type: int = BIGINT):
    """testOuterJoinWithNullProbeAndFilterFunction :970-1011, ...NullBuild... :1056-1098, ...NullOnBothSides... :1143-1187: the filter
    `probe.key = 'a'` on top of the NULL-key rules (key_type VARCHAR: the reference's own types — a VARCHAR input and a Slice constant inside
    the JoinFilterFunction; BIGINT: a -> 1, b -> 2, c -> 3)."""
    def f(bt, pt):
        return _join_filter(bt, pt, lambda build, probe: probe[0] == _key(key_type, 1))
    bp = RowPagesBuilder([key_type], build_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, 3))
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, None)).row(_key(key_type, None)).row(_key(key_type, 1)).row(_key(key_type, 2))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_PROBE_OUTER, build_hash, probe_hash, filter_program=f(bp.get_types(), pp.get_types()))
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (None, None), (None, None), (1, 1), (2, None)]]
    bp = RowPagesBuilder([key_type], build_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, None)).row(_key(key_type, None)).row(_key(key_type, 1)).row(_key(key_type, 2))
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, 3))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_PROBE_OUTER, build_hash, probe_hash, filter_program=f(bp.get_types(), pp.get_types()))
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (1, 1), (2, None), (3, None)]]
    pp = RowPagesBuilder([key_type], probe_hash, [0], backend).row(_key(key_type, 1)).row(_key(key_type, 2)).row(_key(key_type, None)).row(_key(key_type, 3))
    rows = _join(backend, bp.build(), bp.get_types(), pp.build(), pp.get_types(), JOIN_PROBE_OUTER, build_hash, probe_hash, filter_program=f(bp.get_types(), pp.get_types()))
    assert rows == [tuple(_key(key_type, v) for v in r) for r in [(1, 1), (1, 1), (2, None), (None, None), (3, None)]]


def case_join_filter_over_both_sides(backend):
    """A filter reading BOTH rows (the reference's own tests only look at the probe row): build.payload + 5 < probe.payload * 2.0 over
    duplicate chains, NULL payloads (NULL filter result = not eligible) — expectation from the definition of the join."""
    build_keys = [5, 7, 5, 9, 5, 7]
    build_pay = [10, 11, None, 13, 14, 15]
    probe_keys = [7, 5, 1, 9, 5]
    probe_pay = [9.0, 8.0, 1.0, None, 10.0]
    bp = Page([create_longs_block(*build_keys), create_longs_block(*build_pay)])
    pp = Page([create_longs_block(*probe_keys), Block.from_values(DOUBLE, probe_pay)])
    f = _join_filter([BIGINT, BIGINT], [BIGINT, DOUBLE], lambda build, probe: (build[1] + 5) < probe[1] * 2.0)
    def expect(outer):
        rows = []
        for pk, pv in zip(probe_keys, probe_pay):
            n = 0
            for pos in reversed(range(len(build_keys))):      # PositionLinks order: descending build position
                if build_keys[pos] == pk and build_pay[pos] is not None and pv is not None and build_pay[pos] + 5 < pv * 2.0:
                    rows.append((pk, pv, build_pay[pos])); n += 1
            if n == 0 and outer:
                rows.append((pk, pv, None))
        return rows
    for jt in (JOIN_INNER, JOIN_PROBE_OUTER):
        bridge = JoinBridge(backend=backend)
        builder = HashBuilderOperator(bridge, [BIGINT, BIGINT], [0], [1], backend=backend, filter_program=f)
        to_pages(builder, [bp])
        probe = LookupJoinOperator(bridge, [BIGINT, DOUBLE], [0], [0, 1], join_type=jt, backend=backend)
        assert rows_of(to_pages(probe, [pp])) == expect(jt == JOIN_PROBE_OUTER)
        probe.close(); builder.close(); bridge.close()


def case_join_varchar_outputs(backend):
    """VARCHAR channels travelling through the join (the reference's TestHashJoinOperator pages are mostly VARCHAR; here the KEYS stay BIGINT):
    LookupJoinPageBuilder.appendRow copies the probe row's slices (:90-119) and LookupSource.appendTo the build position's.  Duplicate
    chains (descending build position), NULL strings, empty strings, a probe page split in two, INNER / PROBE_OUTER and the outer operator."""
    from presto_b200 import JOIN_FULL_OUTER, LookupOuterOperator
    build_keys = [5, 7, 5, 9, None, 11]
    build_names = [b"five-a", b"", None, b"nine", b"null-key", b"eleven, the longest of them all"]
    probe_keys = [7, 5, 1, None, 9, 5]
    probe_names = [b"p7", None, b"p1", b"pnull", b"", b"p5 again"]
    bp = [Page([create_longs_block(*build_keys[:4]), create_strings_block(*build_names[:4])]), Page([create_longs_block(*build_keys[4:]), create_strings_block(*build_names[4:])])]
    pp = [Page([create_strings_block(*probe_names[:3]), create_longs_block(*probe_keys[:3])]), Page([create_strings_block(*probe_names[3:]), create_longs_block(*probe_keys[3:])])]

    def expect(outer):
        rows = []
        for pk, pn in zip(probe_keys, probe_names):
            hits = [pos for pos in reversed(range(len(build_keys))) if pk is not None and build_keys[pos] == pk]
            rows += [(pn, pk, build_names[pos], build_keys[pos]) for pos in hits]
            if not hits and outer:
                rows.append((pn, pk, None, None))
        return rows
    for jt in (JOIN_INNER, JOIN_PROBE_OUTER):
        bridge = JoinBridge(backend=backend)
        builder = HashBuilderOperator(bridge, [BIGINT, VARCHAR], [0], [1, 0], backend=backend)
        to_pages(builder, bp)
        probe = LookupJoinOperator(bridge, [VARCHAR, BIGINT], [1], [0, 1], join_type=jt, backend=backend)
        assert rows_of(to_pages(probe, pp)) == expect(jt == JOIN_PROBE_OUTER)
        probe.close(); builder.close(); bridge.close()
    # FULL OUTER: the unmatched build rows (positions 4 and 5) follow with NULL probe channels, VARCHAR ones included
    bridge = JoinBridge(backend=backend, outer=True)
    builder = HashBuilderOperator(bridge, [BIGINT, VARCHAR], [0], [1, 0], backend=backend)
    to_pages(builder, bp)
    probe = LookupJoinOperator(bridge, [VARCHAR, BIGINT], [1], [0, 1], join_type=JOIN_FULL_OUTER, backend=backend)
    outer = LookupOuterOperator(bridge, [VARCHAR, BIGINT], backend=backend, max_output_rows=1)
    assert rows_of(to_pages(probe, pp)) == expect(True)
    bridge.no_more_probe_operators()
    pages = []
    while not outer.is_finished():
        p = outer.get_output()
        if p is not None:
            pages.append(p)
    assert rows_of(pages) == [(None, None, b"null-key", None), (None, None, b"eleven, the longest of them all", 11)]
    for op in (probe, outer, builder):
        op.close()
    bridge.close()


def case_probe_outer_join_empty_lookup_source(backend):
    """testProbeOuterJoinWithEmptyLookupSource :1323+ — empty build: every probe row comes out with NULL build columns."""
    bridge = JoinBridge(backend=backend)
    builder = HashBuilderOperator(bridge, [BIGINT], [0], [0], backend=backend)
    assert to_pages(builder, []) == []
    probe = LookupJoinOperator(bridge, [BIGINT], [0], [0], join_type=JOIN_PROBE_OUTER, backend=backend)
    rows = rows_of(to_pages(probe, [Page([create_longs_block(1, 2, None, 3)])]))
    assert rows == [(1, None), (2, None), (None, None), (3, None)]
    inner = LookupJoinOperator(bridge, [BIGINT], [0], [0], join_type=JOIN_INNER, backend=backend)
    assert to_pages(inner, [Page([create_longs_block(1, 2, None, 3)])]) == []


def case_lookup_outer_join_empty_lookup_source(backend):
    """testLookupOuterJoinWithEmptyLookupSource :1285-1320 (RIGHT join, empty build): the probe produces nothing, and so does the outer
    operator; testFullOuterJoinWithEmptyLookupSource :1363-1402: every probe row comes out with NULL build columns.  (BIGINT keys for the
    reference's VARCHAR, like the other join cases.)"""
    from presto_b200 import JOIN_FULL_OUTER, JOIN_LOOKUP_OUTER, LookupOuterOperator
    for join_type, expected in ((JOIN_LOOKUP_OUTER, []), (JOIN_FULL_OUTER, [(1, None), (2, None), (None, None), (3, None)])):
        bridge = JoinBridge(backend=backend, outer=True)
        builder = HashBuilderOperator(bridge, [BIGINT], [0], [0], backend=backend)
        assert to_pages(builder, []) == []
        probe = LookupJoinOperator(bridge, [BIGINT], [0], [0], join_type=join_type, backend=backend)
        outer = LookupOuterOperator(bridge, [BIGINT], backend=backend)
        assert rows_of(to_pages(probe, [Page([create_longs_block(1, 2, None, 3)])])) == expected
        bridge.no_more_probe_operators()
        assert not outer.is_blocked() and outer.get_output() is None
        for op in (probe, outer, builder):
            op.close()
        bridge.close()


def case_lookup_outer_operator(backend):
    """RIGHT / FULL OUTER joins: LookupOuterOperator.java:183-214 + OuterLookupSource.java:100-151.  Two probe operators share the
    bridge; the outer operator is blocked (and returns no page) until the probe factory is closed AND both probes finished; then the
    build rows nobody matched come out in ascending build position with NULL probe channels.  Build keys include a NULL (never matches:
    stays unvisited) and duplicates (one visit marks only the positions actually emitted)."""
    from presto_b200 import JOIN_FULL_OUTER, JOIN_LOOKUP_OUTER, LookupOuterOperator
    build_keys = [5, 7, None, 9, 5, 11, 13]
    bp = Page([create_longs_block(*build_keys), create_long_sequence_block(100, 107)])
    for join_type in (JOIN_LOOKUP_OUTER, JOIN_FULL_OUTER):
        bridge = JoinBridge(backend=backend, outer=True)
        builder = HashBuilderOperator(bridge, [BIGINT, BIGINT], [0], [0, 1], backend=backend)
        probes = [LookupJoinOperator(bridge, [BIGINT, BIGINT], [0], [0, 1], join_type=join_type, backend=backend) for _ in range(2)]
        outer = LookupOuterOperator(bridge, [BIGINT, BIGINT], backend=backend, max_output_rows=2)
        assert outer.is_blocked() and not outer.needs_input() and outer.get_output() is None
        to_pages(builder, [bp])
        assert outer.is_blocked()                       # built, but the probes are still running
        r0 = rows_of(to_pages(probes[0], [Page([create_longs_block(7, 1, None), create_longs_block(70, 10, 0)])]))
        if join_type == JOIN_LOOKUP_OUTER:
            assert r0 == [(7, 70, 7, 101)]
        else:
            assert r0 == [(7, 70, 7, 101), (1, 10, None, None), (None, 0, None, None)]
        bridge.no_more_probe_operators()
        assert outer.is_blocked() and outer.get_output() is None      # probes[1] has not finished
        r1 = rows_of(to_pages(probes[1], [Page([create_longs_block(5, 13), create_longs_block(50, 130)])]))
        assert r1 == [(5, 50, 5, 104), (5, 50, 5, 100), (13, 130, 13, 106)]
        assert not outer.is_blocked()
        pages = []
        while not outer.is_finished():
            p = outer.get_output()
            if p is not None:
                pages.append(p)
        assert [p.position_count for p in pages] == [2, 1]            # pages of max_output_rows
        assert rows_of(pages) == [(None, None, None, 102), (None, None, 9, 103), (None, None, 11, 105)]
        for op in probes + [outer, builder]:
            op.close()
        bridge.close()
    # a LOOKUP_OUTER probe over a bridge without tracking is refused
    from presto_b200 import PrestoException
    bridge = JoinBridge(backend=backend)
    with pytest.raises(PrestoException):
        LookupJoinOperator(bridge, [BIGINT], [0], [0], join_type=JOIN_LOOKUP_OUTER, backend=backend)
    with pytest.raises(PrestoException):
        LookupOuterOperator(bridge, [BIGINT], backend=backend)
    bridge.close()


def case_join_duplicate_chain_order(backend):
    """PagesHash.java:108-123 + ArrayPositionLinks.java:45-50: matches of one probe row come out in DESCENDING build position."""
    build_keys = [5, 7, 5, 9, 5, 7]
    bp = Page([create_longs_block(build_keys), create_long_sequence_block(100, 106)])
    pp = Page([create_longs_block(7, 5, 1, 9)])
    bridge = JoinBridge(backend=backend)
    builder = HashBuilderOperator(bridge, [BIGINT, BIGINT], [0], [1], backend=backend)
    to_pages(builder, [bp])
    probe = LookupJoinOperator(bridge, [BIGINT], [0], [0], backend=backend)
    rows = rows_of(to_pages(probe, [pp]))
    assert rows == [(7, 105), (7, 101), (5, 104), (5, 102), (5, 100), (9, 103)]


# ---------------------------------------------------------------------------------------------------------------------
# PartitionedOutput (TestOptimizedPartitionedOutputOperator: per-partition rows equal the rows the partition function assigns)
# ---------------------------------------------------------------------------------------------------------------------
def case_partitioned_output(backend, partition_count: int = 4, n: int = 1000, nulls: bool = False, precomputed: bool = False, local: bool = False):
    """TestOptimizedPartitionedOutputOperator.testPartitionedSinglePagePrimitiveTypes / …WithNulls: every row lands in the
    partition HashGenerator.getPartition assigns, rows keep input order inside a partition."""
    rng = np.random.default_rng(11)
    c0 = rng.integers(-10**12, 10**12, n)
    c1 = rng.integers(0, 10**6, n).astype(np.int32)
    c2 = rng.random(n)
    null0 = (rng.random(n) < 0.1).astype(np.uint8) if nulls else None
    null2 = (rng.random(n) < 0.2).astype(np.uint8) if nulls else None
    blocks = [Block.from_numpy(BIGINT, c0, null0), Block.from_numpy(INTEGER, c1), Block.from_numpy(DOUBLE, c2, null2)]
    types = [BIGINT, INTEGER, DOUBLE]
    raw = get_hash_block(blocks[:2], backend)
    hash_channel = -1
    if precomputed:
        blocks.append(raw); types.append(BIGINT); hash_channel = 3
    # local = LocalPartitionGenerator (LocalExchange / PartitionedLookupSource, LocalPartitionGenerator.java:43-63) instead of
    # HashGenerator.getPartition; the expectation comes from the numpy restatement (cross-checked against python xxhash in test_abi)
    op = PartitionedOutputOperator(types, partition_count, [0, 1], precomputed_hash_channel=hash_channel, backend=backend,
                                   partition_function=1 if local else 0)
    page = Page(blocks)
    op.add_input(page)
    out = op.get_output()
    offsets = op.get_offsets()
    parts = local_partition(raw.values, partition_count) if local else get_partition(raw.values, partition_count)
    in_rows = page.rows()
    out_rows = out.rows()
    assert offsets[0] == 0 and offsets[-1] == n and out.position_count == n
    for p in range(partition_count):
        assert out_rows[offsets[p]:offsets[p + 1]] == [in_rows[i] for i in range(n) if parts[i] == p]
    op.finish()
    assert op.is_finished()
    op.close()


def case_min_max_aggregates(backend):
    """min / max over BIGINT and DOUBLE (AbstractMinMaxAggregationFunction.java:194-287): NULL inputs are skipped, a group whose inputs are
    all NULL yields NULL, NaN is the largest DOUBLE (TypeUtils.doubleCompare), SINGLE equals PARTIAL -> FINAL; expectations computed here
    from the definition, independent of both implementations."""
    from presto_b200 import AGG_MAX_BIGINT, AGG_MAX_DOUBLE, AGG_MIN_BIGINT, AGG_MIN_DOUBLE, STEP_FINAL, STEP_PARTIAL
    keys = [1, 2, 1, 3, 2, 1, 3, 4, 4]
    longs = [5, None, -7, None, 9, 2**62, None, -(2**63), 2**63 - 1]
    dbls = [1.5, float("nan"), -2.25, None, 7.0, None, None, float("-inf"), 3.0]
    page = Page([Block.from_values(BIGINT, keys), Block.from_values(BIGINT, longs), Block.from_values(DOUBLE, dbls)])
    aggs = [Aggregate(AGG_MIN_BIGINT, 1), Aggregate(AGG_MAX_BIGINT, 1), Aggregate(AGG_MIN_DOUBLE, 2), Aggregate(AGG_MAX_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)]
    op = HashAggregationOperator([BIGINT], [0], aggs, backend=backend)
    rows = rows_of(to_pages(op, [page]))
    op.close()
    nan = float("nan")
    expected = [(1, -7, 2**62, -2.25, 1.5, 3), (2, 9, 9, 7.0, nan, 2), (3, None, None, None, None, 2), (4, -(2**63), 2**63 - 1, float("-inf"), 3.0, 2)]
    rows_close(rows, expected, rel=0.0)
    # PARTIAL per page, FINAL over the intermediate pages (channels: key, then one per aggregate)
    partials = []
    for half in (page_slice(page, 0, 4), page_slice(page, 4, 9)):
        p = HashAggregationOperator([BIGINT], [0], aggs, step=STEP_PARTIAL, backend=backend)
        partials += to_pages(p, [half])
        p.close()
    f = HashAggregationOperator([BIGINT], [0], [Aggregate(a.function, 1 + i) for i, a in enumerate(aggs)], step=STEP_FINAL, backend=backend)
    rows_close(rows_of(to_pages(f, partials)), expected, rel=0.0)
    f.close()


def page_slice(page: Page, begin: int, end: int) -> Page:
    from presto_b200 import LONG_DECIMAL
    return Page([Block.from_decimals(b.to_pylist()[begin:end]) if b.type == LONG_DECIMAL else Block.from_values(b.type, b.to_pylist()[begin:end]) for b in page.blocks])


def _decimal_state(overflow: int, value: int) -> bytes:
    """LongDecimalWithOverflowStateSerializer.serialize :37-49: {overflow, low, high | sign << 63} as three little-endian longs."""
    import struct
    m = abs(value)
    return struct.pack("<qQQ", overflow, m & (2**64 - 1), (m >> 64) | ((1 << 63) if value < 0 else 0))


def _state_value(state: bytes) -> int:
    import struct
    overflow, low, high = struct.unpack("<qQQ", state)
    m = ((high & (2**63 - 1)) << 64) | low
    return overflow * 2**127 + (-m if high >> 63 else m)


def case_decimal_sum(backend):
    """sum(decimal) (DecimalSumAggregation.java:145-193): short decimals are unscaled longs, long decimals Int128 values; NULL inputs are
    skipped, a group without input is NULL; the result is decimal(38): exact python integers here.  SINGLE equals PARTIAL -> FINAL through
    the VARBINARY states {overflow, low, high} (LongDecimalWithOverflowStateSerializer.java:37-66)."""
    from presto_b200 import AGG_SUM_LONG_DECIMAL, AGG_SUM_SHORT_DECIMAL, STEP_FINAL, STEP_PARTIAL
    keys = [1, 2, 1, 3, 2, 1, 3, 4, 4, 4]
    shorts = [10**18 - 1, -5, 10**18 - 1, None, 7, -(10**18 - 1), None, 2**62, 2**62, 2**62]
    longs = [10**37, -(10**30), 8 * 10**37, None, 10**30, -3, None, -(10**38 - 1), 10**38 - 1, None]
    page = Page([Block.from_values(BIGINT, keys), Block.from_values(BIGINT, shorts), Block.from_decimals(longs)])
    aggs = [Aggregate(AGG_SUM_SHORT_DECIMAL, 1), Aggregate(AGG_SUM_LONG_DECIMAL, 2), Aggregate(AGG_COUNT_STAR)]
    expected = [(1, 10**18 - 1, 9 * 10**37 - 3, 3), (2, 2, 0, 2), (3, None, None, 2), (4, 3 * 2**62, 0, 3)]
    op = HashAggregationOperator([BIGINT], [0], aggs, backend=backend)
    assert rows_of(to_pages(op, [page])) == expected
    op.close()
    partials = []
    for half in (page_slice(page, 0, 4), page_slice(page, 4, 10)):
        p = HashAggregationOperator([BIGINT], [0], aggs, step=STEP_PARTIAL, backend=backend)
        out = to_pages(p, [half])
        for pg in out:      # the intermediate states are what the reference serialises: 24 bytes, the number they denote is the partial sum
            for row in pg.rows():
                assert all(s is None or len(s) == 24 for s in row[1:3])
        partials += out
        p.close()
    first = {r[0]: r for r in partials[0].rows()}
    assert _state_value(first[1][1]) == 2 * (10**18 - 1) and _state_value(first[1][2]) == 9 * 10**37 and first[3][1] is None
    f = HashAggregationOperator([BIGINT], [0], [Aggregate(a.function, 1 + i) for i, a in enumerate(aggs)], step=STEP_FINAL, backend=backend)
    assert rows_of(to_pages(f, partials)) == expected
    f.close()


def case_decimal_sum_overflow(backend):
    """TestDecimalSumAggregation.java:46-127: 2^126 + 2^126 leaves the 127-bit magnitude (overflow = 1, value = 0) and the output throws
    (testOverflowOnOutput); two such states combine to overflow 1 + value 2^126 (testCombineOverflow); a sum of magnitude >= 10^38 throws
    too (outputLongDecimal -> throwIfOverflows)."""
    from presto_b200 import AGG_SUM_LONG_DECIMAL, PrestoException, STEP_FINAL, STEP_PARTIAL
    two = 2
    def single(values):
        op = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_LONG_DECIMAL, 1)], backend=backend)
        try:
            return rows_of(to_pages(op, [Page([Block.from_values(BIGINT, [7] * len(values)), Block.from_decimals(values)])]))
        finally:
            op.close()
    for values in ([two**126, two**126], [-(two**126), -(two**126)], [9 * 10**37, 10**37], [-(9 * 10**37), -(10**37)]):
        with pytest.raises(PrestoException) as e:
            single(values)
        assert "Decimal overflow" in str(e.value)
    assert single([9 * 10**37, 10**37 - 1]) == [(7, 10**38 - 1)]
    # testUnderflowAfterOverflow: the overflow count goes back to 0 and the result is exact
    assert single([two**126, two**126, two**125, -(two**126), -(two**126), -(two**126)]) == [(7, -(two**125))]
    # testCombineOverflow through the operators: states (0, 2^125 + 2^126) twice -> FINAL holds overflow 1, value 2^126 -> throws on output
    p = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_LONG_DECIMAL, 1)], step=STEP_PARTIAL, backend=backend)
    state_pages = to_pages(p, [Page([Block.from_values(BIGINT, [7, 7]), Block.from_decimals([two**125, two**126])])])
    p.close()
    assert _state_value(state_pages[0].rows()[0][1]) == 3 * two**125
    f = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_LONG_DECIMAL, 1)], step=STEP_FINAL, backend=backend)
    with pytest.raises(PrestoException):
        to_pages(f, state_pages + state_pages)
    f.close()
    # a hand-written state with an overflow count that a negative state cancels (FINAL only sees states)
    f = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_LONG_DECIMAL, 1)], step=STEP_FINAL, backend=backend)
    states = Page([Block.from_values(BIGINT, [7, 7, 7]), Block.from_strings([_decimal_state(1, 5), _decimal_state(-1, -2), None])])
    assert rows_of(to_pages(f, [states])) == [(7, 3)]
    f.close()


def case_lookup_join_is_blocked_until_the_build_finishes(backend):
    """Operator.isBlocked() (Operator.java:32) as LookupJoinOperator uses it (:198-210): the probe is blocked — and does not need input —
    until HashBuilderOperator.finish() has published the LookupSource; every other operator is never blocked."""
    bridge = JoinBridge(backend=backend)
    build = HashBuilderOperator(bridge, [BIGINT, BIGINT], [0], [1], backend=backend)
    probe = LookupJoinOperator(bridge, [BIGINT], [0], [0], backend=backend)
    assert probe.is_blocked() and not probe.needs_input() and not build.is_blocked()
    build.add_input(Page([create_long_sequence_block(0, 10), create_long_sequence_block(100, 110)]))
    assert probe.is_blocked()
    build.finish()
    assert not probe.is_blocked() and probe.needs_input()
    probe.add_input(Page([create_longs_block(3, 11, 7)]))
    assert probe.get_output().rows() == [(3, 103), (7, 107)]
    for op in (build, probe):
        op.close()
    bridge.close()


def case_partitioned_output_varchar_channels(backend, partition_count: int = 5, replicate: bool = False):
    """TestOptimizedPartitionedOutputOperator.testPartitionedSinglePageVariableWidthTypes (VariableWidthBlockEncodingBuffer.java:1-293): VARCHAR
    channels travel with the row — every row lands in its partition with its slices intact, input order kept inside a partition; NULL
    slices, empty strings, multi-byte values; partition key = a BIGINT channel (or the precomputed $hashvalue)."""
    rng = np.random.default_rng(17)
    n = 4000
    words = [b"", b"A", b"BUILDING", b"AUTOMOBILE", "\u00e9t\u00e9".encode(), b"y" * 33]
    s1 = [None if rng.random() < 0.15 else words[rng.integers(len(words))] + (b"%d" % i) for i in range(n)]
    s2 = [words[rng.integers(len(words))] for _ in range(n)]
    k = rng.integers(-10**9, 10**9, n)
    kn = (rng.random(n) < 0.05).astype(np.uint8) if replicate else None
    v = rng.random(n)
    blocks = [Block.from_strings(s1), Block.from_numpy(BIGINT, k, kn), Block.from_strings(s2), Block.from_numpy(DOUBLE, v)]
    types = [VARCHAR, BIGINT, VARCHAR, DOUBLE]
    raw = get_hash_block([blocks[1]], backend)
    op = PartitionedOutputOperator(types, partition_count, [1], backend=backend, replicates_any_row=replicate, null_channel=1 if replicate else -1)
    page = Page(blocks)
    op.add_input(page)
    out = op.get_output()
    offsets = op.get_offsets()
    parts = get_partition(raw.values, partition_count)
    in_rows, out_rows = page.rows(), out.rows()
    for p in range(partition_count):
        if replicate:
            expected = [in_rows[i] for i in range(n) if i == 0 or in_rows[i][1] is None or parts[i] == p]
        else:
            expected = [in_rows[i] for i in range(n) if parts[i] == p]
        assert out_rows[offsets[p]:offsets[p + 1]] == expected
    op.close()


def case_partitioned_output_varchar_partition_channels(backend, partition_count: int = 7):
    """InterpretedHashGenerator.hashPosition :70-79 over VARCHAR partition channels (TypeUtils.hashPosition → VarcharType.hash = XxHash64 of the
    slice, AbstractVariableWidthBlock.java:102-106; NULL → 0), combined with a BIGINT channel by CombineHashFunction: every row lands in
    the partition the host-side restatement of the row hash names, in input order.  Slices of 0..70 bytes cross XXH64's 32-byte stripes."""
    rng = np.random.default_rng(23)
    n = 5000
    s1 = [None if rng.random() < 0.1 else bytes(rng.integers(32, 127, int(rng.integers(0, 71))).astype(np.uint8)) for _ in range(n)]
    k = rng.integers(-10**6, 10**6, n)
    blocks = [Block.from_strings(s1), Block.from_numpy(BIGINT, k), Block.from_numpy(DOUBLE, rng.random(n))]
    types = [VARCHAR, BIGINT, DOUBLE]
    page = Page(blocks)
    in_rows = page.rows()
    for channels in ([0], [0, 1]):
        raw = get_hash_block([blocks[c] for c in channels], backend)
        op = PartitionedOutputOperator(types, partition_count, channels, backend=backend)
        op.add_input(page)
        out_rows, offsets = op.get_output().rows(), op.get_offsets()
        parts = get_partition(raw.values, partition_count)
        for p in range(partition_count):
            assert out_rows[offsets[p]:offsets[p + 1]] == [in_rows[i] for i in range(n) if parts[i] == p]
        op.close()


def case_partitioned_output_bucket_table_and_round_robin(backend):
    """BucketPartitionFunction.java:25-54 — the bucket function's result goes through bucketToPartition (several buckets per partition, as
    NodePartitioningManager lays buckets over nodes) — over HashBucketFunction (SystemPartitioningHandle.java:311-329) and over
    RoundRobinBucketFunction (:282-305: a counter that runs on across pages and is not advanced by the replicated row)."""
    from presto_b200 import _abi as abi
    rng = np.random.default_rng(41)
    n, buckets, partitions = 3000, 64, 5
    table = rng.integers(0, partitions, buckets).astype(np.int32)
    table[:partitions] = np.arange(partitions)                 # every partition is hit
    pages = []
    for _ in range(3):
        pages.append(Page([Block.from_numpy(BIGINT, rng.integers(-10**9, 10**9, n)), Block.from_numpy(DOUBLE, rng.random(n))]))
    # hash buckets
    op = PartitionedOutputOperator([BIGINT, DOUBLE], partitions, [0], bucket_to_partition=table, backend=backend)
    for page in pages:
        op.add_input(page)
        out_rows, offsets = op.get_output().rows(), op.get_offsets()
        parts = table[get_partition(get_hash_block([page.blocks[0]], backend).values, buckets)]
        in_rows = page.rows()
        for p in range(partitions):
            assert out_rows[offsets[p]:offsets[p + 1]] == [in_rows[i] for i in range(n) if parts[i] == p]
    op.close()
    # round robin, with and without the replicated first row
    for replicate in (False, True):
        op = PartitionedOutputOperator([BIGINT, DOUBLE], partitions, [], bucket_to_partition=table, partition_function=abi.PARTITION_ROUND_ROBIN,
                                       replicates_any_row=replicate, backend=backend)
        counter = 0
        for k, page in enumerate(pages):
            op.add_input(page)
            out_rows, offsets = op.get_output().rows(), op.get_offsets()
            in_rows = page.rows()
            parts = np.full(n, -1)
            for i in range(n):
                if replicate and k == 0 and i == 0:
                    continue                                   # goes everywhere, does not call the function
                parts[i] = table[counter % buckets]; counter += 1
            for p in range(partitions):
                assert out_rows[offsets[p]:offsets[p + 1]] == [in_rows[i] for i in range(n) if parts[i] == p or parts[i] == -1]
        op.close()


def case_partitioned_output_double_zero_keys(backend):
    """DoubleType.hash → TypeUtils.doubleHashCode (PC/type/TypeUtils.java:232-238): -0.0 and +0.0 are equal keys, so they must reach the
    same partition (hash 0 → partition 0), as must every NaN payload."""
    import struct
    nan2 = struct.unpack("<d", struct.pack("<Q", 0x7FF0000000000123))[0]
    keys = Block.from_numpy(DOUBLE, np.array([0.0, -0.0, 1.5, float("nan"), nan2, -0.0, 2.5], dtype=np.float64))
    vals = create_long_sequence_block(0, 7)
    op = PartitionedOutputOperator([DOUBLE, BIGINT], 8, [0], backend=backend)
    op.add_input(Page([keys, vals]))
    out = op.get_output()
    offsets = op.get_offsets()
    seq = [r[1] for r in out.rows()]
    part_of = {}
    for p in range(8):
        for v in seq[offsets[p]:offsets[p + 1]]:
            part_of[v] = p
    assert part_of[0] == part_of[1] == part_of[5] == 0
    assert part_of[3] == part_of[4]
    op.close()


def case_partitioned_output_replicates(backend):
    """OptimizedPartitionedOutputOperator.java:504-530: replicatesAnyRow sends row 0 of the first page to every partition;
    rows whose nullChannel is NULL go to every partition."""
    keys = Block.from_values(BIGINT, [10, None, 11, 12, None, 13])
    vals = create_long_sequence_block(0, 6)
    op = PartitionedOutputOperator([BIGINT, BIGINT], 3, [0], replicates_any_row=True, null_channel=0, backend=backend)
    page = Page([keys, vals])
    op.add_input(page)
    out = op.get_output()
    offsets = op.get_offsets()
    raw = get_hash_block([keys], backend)
    parts = get_partition(raw.values, 3)
    in_rows = page.rows()
    out_rows = out.rows()
    for p in range(3):
        expected = [in_rows[0]]
        for i in range(1, 6):
            if in_rows[i][0] is None or parts[i] == p:
                expected.append(in_rows[i])
        assert out_rows[offsets[p]:offsets[p + 1]] == expected
    # second page: row 0 is no longer replicated
    op.add_input(Page([create_longs_block(10), create_longs_block(99)]))
    out2 = op.get_output()
    assert out2.position_count == 1
    op.close()


# ---------------------------------------------------------------------------------------------------------------------
# PageProcessor (TestPageProcessor / HandTpchQuery1)
# ---------------------------------------------------------------------------------------------------------------------
def case_page_processor_selection(backend):
    """TestPageProcessor.testSelectAllFilter / testSelectNoneFilter / testPartialFilter(AsList): SelectedPositions is a range
    when none or all positions are selected and a list otherwise (PageFilter.java:27-50)."""
    page = Page([create_long_sequence_block(0, 100)])
    for lo, expect_rows, expect_list in [(0, 100, False), (100, 0, False), (25, 75, True)]:
        prog = Program([BIGINT])
        x = prog.input(0)
        prog.set_filter(x >= lo)
        prog.project(x)
        op = FilterAndProjectOperator([BIGINT], prog, backend=backend)
        rows = rows_of(to_pages(op, [page]))
        assert rows == [(i,) for i in range(lo, 100)] and len(rows) == expect_rows
        op2 = FilterAndProjectOperator([BIGINT], prog, backend=backend)
        op2.add_input(page)
        is_list, size = op2.get_last_selection()
        assert (is_list, size) == (expect_list, expect_rows)
        op.close(); op2.close()


def q1_program():
    """PB/HandTpchQuery1.java:237-330 filterAndProjectRowOriented: shipdate <= 1998-09-02 (day 10471);
    out = returnflag, linestatus, quantity, extendedprice, ep*(1-d), ep*(1-d)*(1+t), discount.  returnflag/linestatus travel
    as dictionary-decoded INTEGER codes on the GPU path."""
    prog = Program([INTEGER, INTEGER, DOUBLE, DOUBLE, DOUBLE, DOUBLE, DATE])
    rf, ls, qty, ep, disc, tax, ship = [prog.input(i) for i in range(7)]
    prog.set_filter(ship <= 10471)
    disc_price = ep * (1.0 - disc)
    prog.project(rf, ls, qty, ep, disc_price, disc_price * (1.0 + tax), disc)
    return prog


def case_page_processor_q1(backend, n: int = 5000):
    rng = np.random.default_rng(5)
    def nullable(a, p=0.05):
        return (rng.random(n) < p).astype(np.uint8)
    rf = rng.integers(0, 3, n).astype(np.int32); ls = rng.integers(0, 2, n).astype(np.int32)
    qty = rng.integers(1, 51, n).astype(np.float64)
    ep = np.round(rng.random(n) * 1e5, 2); disc = rng.integers(0, 11, n) / 100.0; tax = rng.integers(0, 9, n) / 100.0
    ship = rng.integers(8035, 10600, n).astype(np.int32)
    nulls = [None, None, nullable(qty), nullable(ep), nullable(disc), nullable(tax), nullable(ship)]
    blocks = [Block.from_numpy(t, a, m) for t, a, m in zip([INTEGER, INTEGER, DOUBLE, DOUBLE, DOUBLE, DOUBLE, DATE], [rf, ls, qty, ep, disc, tax, ship], nulls)]
    op = FilterAndProjectOperator([INTEGER, INTEGER, DOUBLE, DOUBLE, DOUBLE, DOUBLE, DATE], q1_program(), backend=backend)
    rows = rows_of(to_pages(op, [Page(blocks)]))
    expected = []
    for i in range(n):   # HandTpchQuery1.filterAndProjectRowOriented, row by row
        if nulls[6][i] or ship[i] > 10471:
            continue
        q = None if nulls[2][i] else float(qty[i]); e = None if nulls[3][i] else float(ep[i])
        d = None if nulls[4][i] else float(disc[i]); t = None if nulls[5][i] else float(tax[i])
        dp = None if e is None or d is None else e * (1 - d)
        ch = None if dp is None or t is None else e * (1 - d) * (1 + t)
        expected.append((int(rf[i]), int(ls[i]), q, e, dp, ch, d))
    assert rows == expected   # IEEE double arithmetic is exact-reproducible: bit-for-bit
    op.close()


def case_page_processor_varchar_pass_through(backend):
    """InputPageProjection (MB/operator/project/InputPageProjection.java): VARCHAR input references pass the selected positions' slices
    through unchanged — with and without a filter, with NULL slices, empty strings and multi-byte values, between computed channels."""
    rng = np.random.default_rng(8)
    n = 3000
    words = [b"", b"A", b"BUILDING", b"AUTOMOBILE", "\u00e9t\u00e9".encode(), b"x" * 40]
    strs = [None if rng.random() < 0.1 else words[rng.integers(len(words))] for _ in range(n)]
    s_block = Block.from_strings(strs)
    v = rng.integers(-50, 50, n)
    page = Page([s_block, Block.from_numpy(BIGINT, v), Block.from_strings([b"k%d" % (i % 7) for i in range(n)])])
    for with_filter in (True, False):
        prog = Program([VARCHAR, BIGINT, VARCHAR])
        a, b, c = prog.input(0), prog.input(1), prog.input(2)
        if with_filter:
            prog.set_filter(b > 0)
        prog.project(c, b * 2, a, b)
        op = FilterAndProjectOperator([VARCHAR, BIGINT, VARCHAR], prog, backend=backend)
        rows = rows_of(to_pages(op, [page, page]))
        expected = [(b"k%d" % (i % 7), int(v[i]) * 2, strs[i], int(v[i])) for i in range(n) if (not with_filter or v[i] > 0)] * 2
        assert rows == expected
        op.close()


def case_page_processor_three_valued_logic(backend):
    """AND / OR / NOT / IS NULL / BETWEEN over NULLs (SQL three-valued logic)."""
    a = Block.from_values(BOOLEAN, [True, True, True, False, False, False, None, None, None])
    b = Block.from_values(BOOLEAN, [True, False, None, True, False, None, True, False, None])
    x = Block.from_values(BIGINT, [1, 5, 9, None, 3, 7, 10, 0, 5])
    prog = Program([BOOLEAN, BOOLEAN, BIGINT])
    ea, eb, ex = prog.input(0), prog.input(1), prog.input(2)
    prog.project(ea & eb, ea | eb, ~ea, ea.is_null(), ex.between(3, 7), ex + 1, ex * 2 - 3, -ex)
    op = FilterAndProjectOperator([BOOLEAN, BOOLEAN, BIGINT], prog, backend=backend)
    rows = rows_of(to_pages(op, [Page([a, b, x])]))
    T, F, N = True, False, None
    assert [r[0] for r in rows] == [T, F, N, F, F, F, N, F, N]
    assert [r[1] for r in rows] == [T, T, T, T, F, N, T, N, N]
    assert [r[2] for r in rows] == [F, F, F, T, T, T, N, N, N]
    assert [r[3] for r in rows] == [F, F, F, F, F, F, T, T, T]
    assert [r[4] for r in rows] == [F, T, F, N, T, T, F, F, T]
    assert [r[5] for r in rows] == [2, 6, 10, N, 4, 8, 11, 1, 6]
    assert [r[6] for r in rows] == [-1, 7, 15, N, 3, 11, 17, -3, 7]
    assert [r[7] for r in rows] == [-1, -5, -9, N, -3, -7, -10, 0, -5]
    op.close()


def case_page_processor_varchar_comparisons(backend, n: int = 6000):
    """VarcharOperators.java:54-105 (equal / notEqual / lessThan … between = Slice.equals / Slice.compareTo: unsigned bytes, then length) and
    :291 (xxHash64) inside a PageProcessor: TPC-H Q3's `c_mktsegment = 'BUILDING'` as the filter; comparisons against literals and
    between two VARCHAR channels as BOOLEAN projections with SQL NULL semantics; the selected slices pass through.  Expectations are
    computed here in python (bytes compare exactly like Slice.compareTo)."""
    import xxhash
    rng = np.random.default_rng(33)
    segments = [b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"HOUSEHOLD", b"MACHINERY", b"", b"BUILD", b"BUILDING ", b"\xc3\xa9t\xc3\xa9", b"building"]
    seg = [None if rng.random() < 0.1 else segments[rng.integers(len(segments))] for _ in range(n)]
    other = [None if rng.random() < 0.1 else segments[rng.integers(len(segments))] for _ in range(n)]
    key = rng.integers(0, 1 << 40, n)
    page = Page([Block.from_strings(seg), Block.from_numpy(BIGINT, key), Block.from_strings(other)])
    types = [VARCHAR, BIGINT, VARCHAR]

    def tv(f, *xs):     # SQL three-valued: NULL operand → NULL
        return None if any(x is None for x in xs) else f(*xs)

    # 1. the Q3 filter; key and slice pass through
    prog = Program(types)
    s = prog.input(0)
    prog.set_filter(s == b"BUILDING")
    prog.project(prog.input(1), s)
    op = FilterAndProjectOperator(types, prog, backend=backend)
    rows = rows_of(to_pages(op, [page]))
    assert rows == [(int(key[i]), b"BUILDING") for i in range(n) if seg[i] == b"BUILDING"] and len(rows) > 100
    op.close()

    # 2. every comparison as a projection (no filter), literal and column operands, BETWEEN, IS NULL, the type hash
    prog = Program(types)
    s, o = prog.input(0), prog.input(2)
    prog.project(s != b"BUILDING", s < b"FURNITURE", s <= b"BUILDING", s > b"BUILD", s >= "\u00e9", s == o, s < o,
                 s.between(b"AUTOMOBILE", b"HOUSEHOLD"), s.between(o, b"MACHINERY"), s.is_null(), s.hash(), prog.const(b"") == s)
    op = FilterAndProjectOperator(types, prog, backend=backend)
    rows = rows_of(to_pages(op, [page]))
    assert len(rows) == n
    signed = lambda h: h - (1 << 64) if h >= 1 << 63 else h
    for i in range(n):
        a, b = seg[i], other[i]
        want = (tv(lambda x: x != b"BUILDING", a), tv(lambda x: x < b"FURNITURE", a), tv(lambda x: x <= b"BUILDING", a), tv(lambda x: x > b"BUILD", a),
                tv(lambda x: x >= "\u00e9".encode(), a), tv(lambda x, y: x == y, a, b), tv(lambda x, y: x < y, a, b),
                tv(lambda x: b"AUTOMOBILE" <= x <= b"HOUSEHOLD", a),
                # value >= min AND value <= max with three-valued AND: FALSE wins over NULL
                (False if (a is not None and a > b"MACHINERY") or (a is not None and b is not None and a < b) else (None if a is None or b is None else True)),
                a is None, 0 if a is None else signed(xxhash.xxh64_intdigest(a)), tv(lambda x: x == b"", a))
        assert rows[i] == want, (i, a, b, rows[i], want)
    op.close()

    # 3. a selective VARCHAR range filter AND a numeric predicate; three outputs including the other VARCHAR channel
    prog = Program(types)
    s = prog.input(0)
    prog.set_filter((s >= b"BUILDING") & (s < b"HOUSEHOLD") & (prog.input(1) < (1 << 39)))
    prog.project(prog.input(2), s)
    op = FilterAndProjectOperator(types, prog, backend=backend)
    rows = rows_of(to_pages(op, [page]))
    assert rows == [(other[i], seg[i]) for i in range(n) if seg[i] is not None and b"BUILDING" <= seg[i] < b"HOUSEHOLD" and key[i] < (1 << 39)]
    op.close()


def case_varchar_operators(backend):
    """TestVarcharOperators.java:60-127 (testEqual … testBetween): the reference's own expectations for `=`, `<>`, `<`, `<=`, `>`, `>=` and
    BETWEEN over 'foo' / 'bar', evaluated through a PageProcessor over VARCHAR channels (and once against Slice constants)."""
    a = Block.from_strings([b"foo", b"foo", b"bar", b"bar"])
    b = Block.from_strings([b"foo", b"bar", b"foo", b"bar"])
    T, F = True, False
    expected = {                                             # rows: ('foo','foo'), ('foo','bar'), ('bar','foo'), ('bar','bar')
        "eq": [T, F, F, T], "ne": [F, T, T, F], "lt": [F, F, T, F], "le": [T, F, T, T], "gt": [F, T, F, F], "ge": [T, T, F, T],
    }
    prog = Program([VARCHAR, VARCHAR])
    x, y = prog.input(0), prog.input(1)
    prog.project(x == y, x != y, x < y, x <= y, x > y, x >= y)
    op = FilterAndProjectOperator([VARCHAR, VARCHAR], prog, backend=backend)
    rows = rows_of(to_pages(op, [Page([a, b])]))
    op.close()
    for j, name in enumerate(["eq", "ne", "lt", "le", "gt", "ge"]):
        assert [r[j] for r in rows] == expected[name], name
    # BETWEEN :113-127: value BETWEEN min AND max for the eight combinations, value and bounds as channels
    value = Block.from_strings([b"foo"] * 4 + [b"bar"] * 4)
    lo = Block.from_strings([b"foo", b"foo", b"bar", b"bar"] * 2)
    hi = Block.from_strings([b"foo", b"bar", b"foo", b"bar"] * 2)
    prog = Program([VARCHAR, VARCHAR, VARCHAR])
    v, l, h = [prog.input(i) for i in range(3)]
    prog.project(v.between(l, h), v.between(b"bar", b"foo"))
    op = FilterAndProjectOperator([VARCHAR, VARCHAR, VARCHAR], prog, backend=backend)
    rows = rows_of(to_pages(op, [Page([value, lo, hi])]))
    op.close()
    assert [r[0] for r in rows] == [T, F, T, F, F, F, T, T]
    assert [r[1] for r in rows] == [T] * 8                                   # 'foo' and 'bar' are both BETWEEN 'bar' AND 'foo'


def case_page_processor_hash_projection(backend):
    """The $hashvalue projection HashGenerationOptimizer.java:975-987 injects: combine_hash(combine_hash(0, hash(a)), hash(b))."""
    a = Block.from_values(BIGINT, [1, 2, None, 42])
    d = Block.from_values(DATE, [10, None, 30, 40])
    prog = Program([BIGINT, DATE])
    h = prog.const(0).combine_hash(prog.input(0).hash()).combine_hash(prog.input(1).hash())
    prog.project(h)
    op = FilterAndProjectOperator([BIGINT, DATE], prog, backend=backend)
    rows = rows_of(to_pages(op, [Page([a, d])]))
    assert [r[0] for r in rows] == get_hash_block([a, d], backend).to_pylist()
    op.close()
