"""CPU-side checks of the fused pipeline generator: fragments are lowered to CUDA C++ (pg_fused_aggregation_explain) and compiled by
NVRTC for sm_100a against the device headers embedded in libprestogpu.so (pg_jit_compile_check) — no GPU needed, exactly the code path
the operator takes at run time minus the module load.  Results are checked on the GPU by tests/test_gpu_fused.py."""
import ctypes as C

import pytest

from presto_b200 import (AGG_AVG_BIGINT, AGG_AVG_DOUBLE, AGG_COUNT_STAR, AGG_SUM_BIGINT, AGG_SUM_DOUBLE, BIGINT, BOOLEAN, DATE, DOUBLE, INTEGER, VARCHAR,
                         Aggregate, FusedAggregationOperator, FusedJoin, PrestoException, Program)
from presto_b200._lib import load_library
from presto_b200.pipelines import Q1_AGGREGATES, Q3_DATE, q1_program_typed


def compile_check(src: str) -> int:
    b = load_library()
    log = C.create_string_buffer(1 << 16)
    n = C.c_int64()
    status = b.jit_compile_check(src.encode(), log, 1 << 16, C.byref(n))
    assert status == 0, log.value.decode()
    return n.value


def join(types, key, outs):
    return FusedJoin(None, types, key, outs)


@pytest.mark.parametrize("key_type", [INTEGER, VARCHAR])
def test_q1_fragment_compiles(key_type):
    types = [key_type, key_type, DOUBLE, DOUBLE, DOUBLE, DOUBLE, DATE]
    src = FusedAggregationOperator.explain(types, q1_program_typed(key_type), [], [], [key_type, key_type], [0, 1], Q1_AGGREGATES,
                                           group_max_lengths=[1, 1] if key_type == VARCHAR else None)
    # every column is loaded once, ahead of the filter, as 128-bit / 64-bit pairs (VARCHAR slices are read where they are used)
    assert src.count("ldStream128") == 4 and ", R = 2," in src and "TIERS = true" in src
    assert "__dmul_rn" in src and "__dsub_rn" in src and "fma" not in src      # exact IEEE: never contracted
    assert compile_check(src) > 10_000


def test_q3_lineitem_fragment_compiles():
    types = [BIGINT, DOUBLE, DOUBLE, DATE]
    pl = Program(types)
    lk, ep, disc, ship = [pl.input(i) for i in range(4)]
    pl.set_filter(ship > Q3_DATE)
    pl.project(lk, ep * (1.0 - disc))
    src = FusedAggregationOperator.explain(types, pl, [join([BIGINT, DOUBLE], [0], [0, 1])], [[DATE, INTEGER]], [BIGINT, DATE, INTEGER], [0, 2, 3],
                                           [Aggregate(AGG_SUM_DOUBLE, 1)])
    # only the filter column and the probe key are loaded for every row; price and discount are read after the probe, for survivors only
    assert "static constexpr int NEAGER = 2" in src and ", R = 4," in src and "W = 3" in src and "TIERS = false" in src
    assert src.count("ldStream64((const uint64_t*) P.values[1] + row") == 4
    assert compile_check(src) > 10_000


def test_two_joins_nullable_inputs_and_key_kinds_compile():
    types = [BIGINT, BIGINT, DOUBLE, BOOLEAN, INTEGER]
    prog = Program(types)
    a, b, v, f, i = [prog.input(k) for k in range(5)]
    prog.set_filter(((v > 1.0) & f) | i.is_null())
    prog.project(a, b, v * 2.0 + i, f, (a + b) * 3)
    joins = [join([BIGINT, BIGINT, DOUBLE, BOOLEAN, BIGINT], [0], [1, 2, 3, 4]), join([BIGINT, DOUBLE, BOOLEAN, BIGINT, INTEGER, DOUBLE], [0], [1, 2, 3, 4, 5])]
    src = FusedAggregationOperator.explain(types, prog, joins, [[INTEGER, DOUBLE], []], [BOOLEAN, INTEGER, BIGINT], [1, 3, 2],
                                           [Aggregate(AGG_SUM_DOUBLE, 0), Aggregate(AGG_AVG_DOUBLE, 4), Aggregate(AGG_AVG_BIGINT, 2), Aggregate(AGG_COUNT_STAR)],
                                           nullable_inputs=[3])
    assert "errOv" in src and "P.nulls[3]" in src
    assert compile_check(src) > 10_000


def test_single_bigint_key_with_nulls_compiles():
    types = [BIGINT, DOUBLE]
    prog = Program(types)
    k, v = prog.input(0), prog.input(1)
    prog.project(k, v)
    src = FusedAggregationOperator.explain(types, prog, [], [], [BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1)], nullable_inputs=[0])
    assert "special = 1" in src and "W = 1" in src
    assert compile_check(src) > 10_000


def test_unfusable_fragments_are_declined_with_a_reason():
    types = [BIGINT, DOUBLE]
    prog = Program(types)
    k, v = prog.input(0), prog.input(1)
    prog.project(k, v)
    for aggs, nullable, why in [([Aggregate(AGG_SUM_BIGINT, 0)], [], "aggregate outside the fused set"), ([Aggregate(AGG_SUM_DOUBLE, 1)], [1], "nullable aggregate input")]:
        with pytest.raises(PrestoException) as e:
            FusedAggregationOperator.explain(types, prog, [], [], [BIGINT], [0], aggs, nullable_inputs=nullable)
        assert e.value.code == 3 and why in str(e.value)


def test_page_sink_fragment_compiles():
    """Q3's orders Driver: filter on orderdate -> probe the customer key set (no payload: pure membership) -> (orderkey, orderdate, shippriority)
    page; and a fragment whose join contributes payload columns with a nullable probe-side projection."""
    from presto_b200 import FusedPageOperator
    types = [BIGINT, BIGINT, DATE, INTEGER]
    po = Program(types)
    ok, ock, od, sp = [po.input(i) for i in range(4)]
    po.set_filter(od < Q3_DATE)
    po.project(ok, ock, od, sp)
    src = FusedPageOperator.explain(types, po, [join([BIGINT, BIGINT, DATE, INTEGER], [1], [0, 2, 3])], [[]])
    assert "fused_select" in src and "fused_emit" in src and ", R = 4;" in src and "P.outValues[2]" in src and "P.outValues[3]" not in src
    assert compile_check(src) > 10_000
    types = [BIGINT, DOUBLE, INTEGER]
    p2 = Program(types)
    k, v, i = [p2.input(c) for c in range(3)]
    p2.set_filter(v > 0.5)
    p2.project(k, v * 2.0, i + 1)
    src = FusedPageOperator.explain(types, p2, [join([BIGINT, DOUBLE, BIGINT], [0], [0, 1, 2])], [[DOUBLE, INTEGER]], nullable_inputs=[2])
    assert "P.outNulls[2]" in src and "P.outValues[4]" in src and "errOv" in src
    assert compile_check(src) > 10_000


def test_fragments_outside_the_generated_set_say_why():
    """Declined fragments run operator by operator on the GPU; the explain entry points name the reason (what `get_info()["reason"]` reports at
    run time): an unbounded VARCHAR group key (interned per chunk), a VARCHAR comparison in the page processor."""
    types = [VARCHAR, VARCHAR, DOUBLE]
    prog = Program(types)
    s, c, v = [prog.input(i) for i in range(3)]
    prog.set_filter(v > 1.0)
    prog.project(c, v)
    with pytest.raises(PrestoException, match="interned"):
        FusedAggregationOperator.explain(types, prog, [], [], [VARCHAR], [0], [Aggregate(AGG_SUM_DOUBLE, 1)])          # no declared length: value ids first
    assert "fused_kernel" in FusedAggregationOperator.explain(types, prog, [], [], [VARCHAR], [0], [Aggregate(AGG_SUM_DOUBLE, 1)], group_max_lengths=[7])
    prog = Program(types)
    s, c, v = [prog.input(i) for i in range(3)]
    prog.set_filter(s == b"BUILDING")
    prog.project(c, v)
    with pytest.raises(PrestoException, match="VARCHAR comparison"):
        FusedAggregationOperator.explain(types, prog, [], [], [VARCHAR], [0], [Aggregate(AGG_SUM_DOUBLE, 1)], group_max_lengths=[7])
