"""OperatorFactory lifecycle over the C ABI (pg_operator_factory_*, MB/operator/OperatorFactory.java:19-51).  The host-side rules need no
GPU: a closed factory refuses ("Factory is already closed", HashAggregationOperator.java:222), duplicate() gives an OPEN factory over
the same plan (:255-275).  -m gpu: operators handed out by a factory (from its own deep copy of the configuration) give the oracle's rows."""
import gc

import numpy as np
import pytest

from presto_b200 import (
    AGG_COUNT_STAR, AGG_SUM_DOUBLE, BIGINT, DOUBLE, STEP_PARTIAL, Aggregate, Block, FilterAndProjectOperator, HashAggregationOperator,
    HashBuilderOperator, JoinBridge, LookupJoinOperator, OperatorFactory, Page, PrestoException, Program, to_pages,
)
from presto_b200 import _abi as abi
from util import rows_close, rows_of


@pytest.fixture(scope="module")
def product():
    from presto_b200._lib import load_library
    return load_library()


def test_closed_factory_refuses_and_duplicate_is_open(product):
    f = OperatorFactory(HashAggregationOperator, [BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1)], backend=product)
    d_before = f.duplicate()
    f.no_more_operators()
    with pytest.raises(PrestoException) as e:
        f.create_operator()
    assert e.value.code == abi.PG_ERR_INVALID_STATE and "Factory is already closed" in str(e.value)
    for d in (d_before, f.duplicate()):           # duplicates are open whatever the state of the original
        try:
            op = d.create_operator()
            op.close()
        except PrestoException as e2:              # no GPU in this container: the operator's constructor fails, not the factory
            assert e2.code != abi.PG_ERR_INVALID_STATE
        d.close()
    assert f.created_count == 0
    f.close()


def test_factory_argument_errors(product):
    import ctypes as C
    out = C.c_void_p()
    cfg = abi.pg_hash_aggregation_config()
    assert product.operator_factory_create(99, C.byref(cfg), None, C.byref(out)) == abi.PG_ERR_INVALID_ARGUMENT
    assert product.operator_factory_create(1, None, None, C.byref(out)) == abi.PG_ERR_INVALID_ARGUMENT
    jc = abi.pg_lookup_join_config()
    assert product.operator_factory_create(4, C.byref(jc), None, C.byref(out)) == abi.PG_ERR_INVALID_ARGUMENT      # a join factory needs its bridge


@pytest.mark.gpu
def test_two_drivers_from_one_aggregation_factory(gpu_backend, oracle_backend):
    rng = np.random.default_rng(3)
    pages = [Page([Block.from_numpy(BIGINT, rng.integers(0, 40, 5000)), Block.from_numpy(DOUBLE, rng.random(5000))]) for _ in range(4)]
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)]
    f = OperatorFactory(HashAggregationOperator, [BIGINT], [0], aggs, step=STEP_PARTIAL, expected_groups=64, backend=gpu_backend)
    gc.collect()                                   # the python-side arrays of the configuration are gone: the factory owns a copy
    drivers = [f.create_operator(), f.create_operator(max_output_rows=7)]
    f.no_more_operators()
    assert f.created_count == 2
    for op, half in zip(drivers, (pages[:2], pages[2:])):
        got = rows_of(to_pages(op, half))
        want = rows_of(to_pages(HashAggregationOperator([BIGINT], [0], aggs, step=STEP_PARTIAL, expected_groups=64, backend=oracle_backend), half))
        rows_close(got, want)
        op.close()
    f.close()


@pytest.mark.gpu
def test_join_factories_share_the_bridge(gpu_backend, oracle_backend):
    rng = np.random.default_rng(4)
    build = Page([Block.from_numpy(BIGINT, rng.permutation(3000)[:2000].astype(np.int64)), Block.from_numpy(DOUBLE, rng.random(2000))])
    probes = [Page([Block.from_numpy(BIGINT, rng.integers(0, 3000, 4000)), Block.from_numpy(BIGINT, rng.integers(0, 9, 4000))]) for _ in range(2)]
    def run(backend, through_factories):
        bridge = JoinBridge(0, backend=backend)
        if through_factories:
            bf = OperatorFactory(HashBuilderOperator, bridge, [BIGINT, DOUBLE], [0], [1], backend=backend)
            pf = OperatorFactory(LookupJoinOperator, bridge, [BIGINT, BIGINT], [0], [0, 1], backend=backend)
            builder, probe_ops = bf.create_operator(), [pf.create_operator(), pf.duplicate().create_operator()]
        else:
            builder = HashBuilderOperator(bridge, [BIGINT, DOUBLE], [0], [1], backend=backend)
            probe_ops = [LookupJoinOperator(bridge, [BIGINT, BIGINT], [0], [0, 1], backend=backend) for _ in range(2)]
        builder.add_input(build); builder.finish()
        rows = [rows_of(to_pages(op, [p])) for op, p in zip(probe_ops, probes)]
        for op in probe_ops + [builder]:
            op.close()
        bridge.close()
        return rows
    got, want = run(gpu_backend, True), run(oracle_backend, False)
    assert got == want and len(got[0]) > 1000


@pytest.mark.gpu
def test_filter_project_factory(gpu_backend, oracle_backend):
    rng = np.random.default_rng(5)
    page = Page([Block.from_numpy(BIGINT, rng.integers(-50, 50, 3000)), Block.from_numpy(DOUBLE, rng.random(3000))])
    prog = Program([BIGINT, DOUBLE])
    a, b = prog.input(0), prog.input(1)
    prog.set_filter(a > 0)
    prog.project(a + 1, b * 2.0)
    f = OperatorFactory(FilterAndProjectOperator, [BIGINT, DOUBLE], prog, backend=gpu_backend)
    op = f.create_operator()
    want = rows_of(to_pages(FilterAndProjectOperator([BIGINT, DOUBLE], prog, backend=oracle_backend), [page]))
    assert rows_of(to_pages(op, [page])) == want
    op.close(); f.close()
