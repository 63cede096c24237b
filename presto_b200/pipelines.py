"""TPC-H Q1 / Q3 operator pipelines over the hot-path operators, the way LocalExecutionPlanner would chain them
(presto-benchto-benchmarks/src/main/resources/sql/presto/tpch/q01.sql, q03.sql; Q1 hand plan: presto-benchmark/.../HandTpchQuery1.java:50-137).

The functions only use the operator mirror (`Operator.add_input / get_output / finish`), so the same plan runs on the CUDA backend
with device-resident pages handed from operator to operator, and on any other Backend the caller passes (the tests pass the oracle).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

from . import _abi as abi
from .expressions import Program
from .operators import (Aggregate, FilterAndProjectOperator, HashAggregationOperator, HashBuilderOperator, JoinBridge, LookupJoinOperator)
from .page import Page

BUILDING = 1                 # dictionary code of c_mktsegment = 'BUILDING'
Q3_DATE = 9204               # DATE '1995-03-15'
Q1_DATE = 10471              # DATE '1998-09-02'


def _drain(op, pages: Sequence[Page], copy: bool) -> List[Page]:
    out = []
    for p in pages:
        op.add_input(p)
        o = op.get_output(copy=copy)
        if o is not None and o.position_count > 0:
            out.append(o)
    op.finish()
    while not op.is_finished():
        o = op.get_output(copy=copy)
        if o is not None and o.position_count > 0:
            out.append(o)
    return out


def q1_program() -> Program:
    """HandTpchQuery1.filterAndProjectRowOriented :237-330 (returnflag / linestatus as dictionary codes)."""
    prog = Program([abi.INTEGER, abi.INTEGER, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DATE])
    rf, ls, qty, ep, disc, tax, ship = [prog.input(i) for i in range(7)]
    prog.set_filter(ship <= Q1_DATE)
    dp = ep * (1.0 - disc)
    prog.project(rf, ls, qty, ep, dp, dp * (1.0 + tax), disc)
    return prog


def run_q1(lineitem_pages: Sequence[Page], backend=None, device_resident: bool = False, **opts) -> List[Page]:
    """lineitem(returnflag, linestatus, quantity, extendedprice, discount, tax, shipdate) → filter/project → GROUP BY returnflag,
    linestatus: sum(qty), sum(ep), sum(disc_price), sum(charge), avg(qty), avg(ep), avg(disc), count(*) (HandTpchQuery1.java:99-127)."""
    mem = abi.MEM_DEVICE if device_resident else abi.MEM_HOST
    fp = FilterAndProjectOperator([abi.INTEGER, abi.INTEGER, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DATE], q1_program(),
                                  backend=backend, output_memory=mem, **opts)
    agg = HashAggregationOperator([abi.INTEGER, abi.INTEGER], [0, 1],
                                  [Aggregate(abi.AGG_SUM_DOUBLE, 2), Aggregate(abi.AGG_SUM_DOUBLE, 3), Aggregate(abi.AGG_SUM_DOUBLE, 4), Aggregate(abi.AGG_SUM_DOUBLE, 5),
                                   Aggregate(abi.AGG_AVG_DOUBLE, 2), Aggregate(abi.AGG_AVG_DOUBLE, 3), Aggregate(abi.AGG_AVG_DOUBLE, 6), Aggregate(abi.AGG_COUNT_STAR)],
                                  expected_groups=10_000, backend=backend, max_output_rows=1 << 30, **opts)
    for page in lineitem_pages:
        fp.add_input(page)
        out = fp.get_output(copy=not device_resident)
        if out is not None and out.position_count:
            agg.add_input(out)
    fp.finish()
    agg.finish()
    result = []
    while not agg.is_finished():
        o = agg.get_output()
        if o is not None:
            result.append(o)
    fp.close(); agg.close()
    return result


def run_q3(customer: Sequence[Page], orders: Sequence[Page], lineitem: Sequence[Page], backend=None, device_resident: bool = False,
           stats: Optional[Dict] = None, **opts) -> List[Page]:
    """Q3 without the final ORDER BY / LIMIT (done on the host by the caller):
         customer(c_custkey BIGINT, c_mktsegment INTEGER code)   → filter mktsegment = BUILDING           → build #1 (custkey)
         orders(o_orderkey, o_custkey, o_orderdate DATE, o_shippriority INTEGER) → filter o_orderdate < 1995-03-15
               → probe #1 on custkey → build #2 (orderkey; payload orderdate, shippriority)
         lineitem(l_orderkey, l_extendedprice, l_discount, l_shipdate DATE) → filter l_shipdate > 1995-03-15, project revenue = ep*(1-d)
               → probe #2 on orderkey → GROUP BY (orderkey, orderdate, shippriority): sum(revenue)
    """
    mem = abi.MEM_DEVICE if device_resident else abi.MEM_HOST
    copy = not device_resident
    o = dict(backend=backend, **opts)
    # customer → build #1
    pc = Program([abi.BIGINT, abi.INTEGER])
    ck, seg = pc.input(0), pc.input(1)
    pc.set_filter(seg == BUILDING)
    pc.project(ck)
    fp_c = FilterAndProjectOperator([abi.BIGINT, abi.INTEGER], pc, output_memory=mem, **o)
    bridge1 = JoinBridge(opts.get("device", 0), backend=backend)
    build1 = HashBuilderOperator(bridge1, [abi.BIGINT], [0], [], **o)
    for p in _drain(fp_c, customer, copy):
        build1.add_input(p)
    build1.finish()
    # orders → probe #1 → build #2
    po = Program([abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER])
    ok, ock, od, sp = [po.input(i) for i in range(4)]
    po.set_filter(od < Q3_DATE)
    po.project(ok, ock, od, sp)
    fp_o = FilterAndProjectOperator([abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER], po, output_memory=mem, **o)
    probe1 = LookupJoinOperator(bridge1, [abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER], [1], [0, 2, 3], output_memory=mem, **o)
    bridge2 = JoinBridge(opts.get("device", 0), backend=backend)
    build2 = HashBuilderOperator(bridge2, [abi.BIGINT, abi.DATE, abi.INTEGER], [0], [1, 2], **o)
    for page in orders:
        fp_o.add_input(page)
        f = fp_o.get_output(copy=copy)
        if f is None or not f.position_count:
            continue
        probe1.add_input(f)
        j = probe1.get_output(copy=copy)
        if j is not None and j.position_count:
            build2.add_input(j)
    build2.finish()
    # lineitem → probe #2 → aggregation
    pl = Program([abi.BIGINT, abi.DOUBLE, abi.DOUBLE, abi.DATE])
    lk, ep, disc, ship = [pl.input(i) for i in range(4)]
    pl.set_filter(ship > Q3_DATE)
    pl.project(lk, ep * (1.0 - disc))
    fp_l = FilterAndProjectOperator([abi.BIGINT, abi.DOUBLE, abi.DOUBLE, abi.DATE], pl, output_memory=mem, **o)
    probe2 = LookupJoinOperator(bridge2, [abi.BIGINT, abi.DOUBLE], [0], [0, 1], output_memory=mem, **o)
    agg = HashAggregationOperator([abi.BIGINT, abi.DATE, abi.INTEGER], [0, 2, 3], [Aggregate(abi.AGG_SUM_DOUBLE, 1)], expected_groups=100_000,
                                  max_output_rows=1 << 30, **o)
    joined = 0
    for page in lineitem:
        fp_l.add_input(page)
        f = fp_l.get_output(copy=copy)
        if f is None or not f.position_count:
            continue
        probe2.add_input(f)
        j = probe2.get_output(copy=copy)
        if j is not None and j.position_count:
            joined += j.position_count
            agg.add_input(j)
    agg.finish()
    result = []
    while not agg.is_finished():
        r = agg.get_output()
        if r is not None:
            result.append(r)
    if stats is not None:
        stats["joined_rows"] = joined
        stats["kernel_launches"] = sum(op.get_stats()["kernel_launches"] for op in (fp_c, build1, fp_o, probe1, build2, fp_l, probe2, agg))
    for op in (fp_c, build1, fp_o, probe1, build2, fp_l, probe2, agg):
        op.close()
    bridge1.close(); bridge2.close()
    return result
