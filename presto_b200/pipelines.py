"""TPC-H Q1 / Q3 operator pipelines over the hot-path operators, the way LocalExecutionPlanner would chain them
(presto-benchto-benchmarks/src/main/resources/sql/presto/tpch/q01.sql, q03.sql; Q1 hand plan: presto-benchmark/.../HandTpchQuery1.java:50-137).

The functions only use the operator mirror (`Operator.add_input / get_output / finish`), so the same plan runs on the CUDA backend
with device-resident pages handed from operator to operator, and on any other Backend the caller passes (the tests pass the oracle).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

from . import _abi as abi
from .expressions import Program
from .operators import (Aggregate, FilterAndProjectOperator, FusedAggregationOperator, FusedJoin, FusedPageOperator, HashAggregationOperator,
                        HashBuilderOperator, JoinBridge, LookupJoinOperator)
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


def _detach(page: Page) -> Page:
    """A device page that outlives the operator that produced it: one device-to-device copy per block into torch-owned memory."""
    from .exchange import block_to_tensor, nulls_to_tensor, tensor_to_block
    blocks = []
    for b in page.blocks:
        n = nulls_to_tensor(b)
        blocks.append(tensor_to_block(b.type, block_to_tensor(b).clone(), None if n is None else n.clone()))
    return Page(blocks, page.position_count)


def q1_program() -> Program:
    """HandTpchQuery1.filterAndProjectRowOriented :237-330 (returnflag / linestatus as dictionary codes)."""
    prog = Program([abi.INTEGER, abi.INTEGER, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DATE])
    rf, ls, qty, ep, disc, tax, ship = [prog.input(i) for i in range(7)]
    prog.set_filter(ship <= Q1_DATE)
    dp = ep * (1.0 - disc)
    prog.project(rf, ls, qty, ep, dp, dp * (1.0 + tax), disc)
    return prog


Q1_AGGREGATES = [Aggregate(abi.AGG_SUM_DOUBLE, 2), Aggregate(abi.AGG_SUM_DOUBLE, 3), Aggregate(abi.AGG_SUM_DOUBLE, 4), Aggregate(abi.AGG_SUM_DOUBLE, 5),
                 Aggregate(abi.AGG_AVG_DOUBLE, 2), Aggregate(abi.AGG_AVG_DOUBLE, 3), Aggregate(abi.AGG_AVG_DOUBLE, 6), Aggregate(abi.AGG_COUNT_STAR)]


def q1_program_typed(key_type: int) -> Program:
    """q1_program with returnflag / linestatus of ``key_type``: INTEGER dictionary codes, or VARCHAR(1) — the reference's own types
    (HandTpchQuery1.java:99-108) — passed through the PageProcessor to become the group-by keys."""
    prog = Program([key_type, key_type, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DATE])
    rf, ls, qty, ep, disc, tax, ship = [prog.input(i) for i in range(7)]
    prog.set_filter(ship <= Q1_DATE)
    dp = ep * (1.0 - disc)
    prog.project(rf, ls, qty, ep, dp, dp * (1.0 + tax), disc)
    return prog


def run_q1(lineitem_pages: Sequence[Page], backend=None, device_resident: bool = False, fused: bool = False, key_type: int = abi.INTEGER,
           info: Optional[Dict] = None, **opts) -> List[Page]:
    """lineitem(returnflag, linestatus, quantity, extendedprice, discount, tax, shipdate) → filter/project → GROUP BY returnflag,
    linestatus: sum(qty), sum(ep), sum(disc_price), sum(charge), avg(qty), avg(ep), avg(disc), count(*) (HandTpchQuery1.java:99-127).
    ``fused``: the whole Driver pipeline as one FusedAggregationOperator (one generated kernel on the CUDA backend)."""
    mem = abi.MEM_DEVICE if device_resident else abi.MEM_HOST
    if fused:
        types = [key_type, key_type, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DATE]
        op = FusedAggregationOperator(types, q1_program_typed(key_type), [], [key_type, key_type], [0, 1], Q1_AGGREGATES, expected_groups=10_000,
                                      group_max_lengths=[1, 1] if key_type == abi.VARCHAR else None, backend=backend, max_output_rows=1 << 30, **opts)
        for page in lineitem_pages:
            op.add_input(page)
        op.finish()
        result = []
        while not op.is_finished():
            o = op.get_output()
            if o is not None:
                result.append(o)
        if info is not None:
            info.update(op.get_info()); info.update(op.get_stats())
        op.close()
        return result
    fp = FilterAndProjectOperator([key_type, key_type, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DOUBLE, abi.DATE], q1_program_typed(key_type),
                                  backend=backend, output_memory=mem, **opts)
    agg = HashAggregationOperator([key_type, key_type], [0, 1], Q1_AGGREGATES, expected_groups=10_000, backend=backend, max_output_rows=1 << 30,
                                  group_max_lengths=[1, 1] if key_type == abi.VARCHAR else None, **opts)
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
           stats: Optional[Dict] = None, fused: bool = False, **opts) -> List[Page]:
    """Q3 without the final ORDER BY / LIMIT (done on the host by the caller):
         customer(c_custkey BIGINT, c_mktsegment VARCHAR — the reference's type — or its INTEGER code) → filter c_mktsegment = 'BUILDING'
                                                                                                          → build #1 (custkey)
         orders(o_orderkey, o_custkey, o_orderdate DATE, o_shippriority INTEGER) → filter o_orderdate < 1995-03-15
               → probe #1 on custkey → build #2 (orderkey; payload orderdate, shippriority)
         lineitem(l_orderkey, l_extendedprice, l_discount, l_shipdate DATE) → filter l_shipdate > 1995-03-15, project revenue = ep*(1-d)
               → probe #2 on orderkey → GROUP BY (orderkey, orderdate, shippriority): sum(revenue)
    """
    mem = abi.MEM_DEVICE if device_resident else abi.MEM_HOST
    copy = not device_resident
    o = dict(backend=backend, **opts)
    # customer → build #1
    seg_type = customer[0].blocks[1].type if customer else abi.INTEGER
    pc = Program([abi.BIGINT, seg_type])
    ck, seg = pc.input(0), pc.input(1)
    pc.set_filter(seg == (b"BUILDING" if seg_type == abi.VARCHAR else BUILDING))       # VarcharOperators.equal over a Slice constant
    pc.project(ck)
    fp_c = FilterAndProjectOperator([abi.BIGINT, seg_type], pc, output_memory=mem, **o)
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
    bridge2 = JoinBridge(opts.get("device", 0), backend=backend)
    build2 = HashBuilderOperator(bridge2, [abi.BIGINT, abi.DATE, abi.INTEGER], [0], [1, 2], **o)
    if fused:
        # the orders Driver (filter → probe the BUILDING customers) as ONE operator feeding the build of the second join
        fp_o = FusedPageOperator([abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER], po, [FusedJoin(bridge1, [abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER], [1], [0, 2, 3])],
                                 output_memory=mem, **o)
        probe1 = fp_o
        for page in orders:
            fp_o.add_input(page)
            j = fp_o.get_output(copy=copy)
            if j is not None and j.position_count:
                build2.add_input(j)
        if stats is not None:
            stats.update({"orders_" + k: v for k, v in fp_o.get_info().items()})
    else:
        fp_o = FilterAndProjectOperator([abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER], po, output_memory=mem, **o)
        probe1 = LookupJoinOperator(bridge1, [abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER], [1], [0, 2, 3], output_memory=mem, **o)
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
    if fused:
        # the lineitem Driver (filter/project → probe → aggregate) as ONE operator: no filtered page, no joined page
        fa = FusedAggregationOperator([abi.BIGINT, abi.DOUBLE, abi.DOUBLE, abi.DATE], pl, [FusedJoin(bridge2, [abi.BIGINT, abi.DOUBLE], [0], [0, 1])],
                                      [abi.BIGINT, abi.DATE, abi.INTEGER], [0, 2, 3], [Aggregate(abi.AGG_SUM_DOUBLE, 1)],
                                      expected_groups=max(100_000, sum(p.position_count for p in lineitem) // 500),
                                      max_output_rows=1 << 30, output_memory=mem, **o)
        for page in lineitem:
            fa.add_input(page)
        fa.finish()
        result = []
        while not fa.is_finished():
            r = fa.get_output(copy=copy)
            if r is not None:
                result.append(_detach(r) if device_resident else r)
        if stats is not None:
            stats["joined_rows"] = -1
            stats["kernel_launches"] = sum(op.get_stats()["kernel_launches"] for op in (fp_c, build1, fp_o, build2, fa))
            stats.update({"fused_" + k: v for k, v in fa.get_info().items()})
            stats["kernel_paths"] = fa.get_stats()["kernel_paths"]
        for op in (fp_c, build1, fp_o, build2, fa):
            op.close()
        bridge1.close(); bridge2.close()
        return result
    fp_l = FilterAndProjectOperator([abi.BIGINT, abi.DOUBLE, abi.DOUBLE, abi.DATE], pl, output_memory=mem, **o)
    probe2 = LookupJoinOperator(bridge2, [abi.BIGINT, abi.DOUBLE], [0], [0, 1], output_memory=mem, **o)
    # the planner's group-count estimate (Q3: about one group per 500 lineitem rows — 1.1 M at SF100) sizes the table so that it
    # does not have to grow mid-page
    agg = HashAggregationOperator([abi.BIGINT, abi.DATE, abi.INTEGER], [0, 2, 3], [Aggregate(abi.AGG_SUM_DOUBLE, 1)],
                                  expected_groups=max(100_000, sum(p.position_count for p in lineitem) // 500),
                                  max_output_rows=1 << 30, output_memory=mem, **o)
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
        r = agg.get_output(copy=copy)
        if r is not None:
            result.append(_detach(r) if device_resident else r)     # the operator's buffers die with close() below
    if stats is not None:
        stats["joined_rows"] = joined
        stats["kernel_launches"] = sum(op.get_stats()["kernel_launches"] for op in (fp_c, build1, fp_o, probe1, build2, fp_l, probe2, agg))
    for op in (fp_c, build1, fp_o, probe1, build2, fp_l, probe2, agg):
        op.close()
    bridge1.close(); bridge2.close()
    return result


def run_q3_distributed(customer: Sequence[Page], orders: Sequence[Page], lineitem: Sequence[Page], backend=None, stats: Optional[Dict] = None,
                       group=None, fused: bool = True, exchanges=None, **opts) -> List[Page]:
    """Q3 across the GPUs of one box (BASELINE.json configs[3]); every rank holds a shard of each table, device resident.
         customer  : filtered, then REPLICATED (every rank receives all BUILDING customers: PartitionedOutput with one partition per rank
                     is not needed for a 1/5 x 15 M x 8 B build side — an all-gather of the filtered keys) → build #1 on every rank
         orders    : [filter → probe #1] (one FusedPageOperator) → PartitionedOutput on o_orderkey → exchange → build #2 (rank owns its partition)
         lineitem  : [filter/project] (FusedPageOperator) → PartitionedOutput on l_orderkey (same partition function) → exchange →
                     [probe #2 → aggregation] (one FusedAggregationOperator)
       The group-by keys start with orderkey, so the aggregation is co-partitioned with the join: no second exchange; every rank returns
       its own groups (the caller takes the global top-10).  Exactly one exchange per repartition boundary, nowhere else.
       ``fused`` (default): the repartition scatter writes straight into the destination GPU's receive columns (exchange.FusedExchange);
       otherwise a local PartitionedOutput slab + one all_to_all_single per column (exchange.all_to_all_pages).  ``exchanges`` =
       (orders FusedExchange, lineitem FusedExchange) reuses long-lived receive columns (a worker's output buffers outlive a query
       stage); by default they are created and released here."""
    import os
    import time
    import torch
    import torch.distributed as dist
    from .exchange import FusedExchange, all_to_all_pages, block_to_tensor, tensor_to_block
    from .operators import PartitionedOutputOperator
    world = dist.get_world_size(group)
    trace = [] if os.environ.get("PG_Q3_TRACE") else None      # wall-clock per phase (rank-local), printed by the caller

    def mark(name):
        if trace is not None:
            torch.cuda.synchronize()
            trace.append((name, time.perf_counter()))
    mark("start")
    mem = abi.MEM_DEVICE
    o = dict(backend=backend, **opts)
    dev = torch.device("cuda", torch.cuda.current_device())
    _TORCH = {abi.BIGINT: torch.int64, abi.DATE: torch.int32, abi.INTEGER: torch.int32, abi.DOUBLE: torch.float64}
    # ---- customer: filter, replicate the surviving keys, build #1 ----
    seg_type = customer[0].blocks[1].type if customer else abi.INTEGER
    pc = Program([abi.BIGINT, seg_type])
    ck, seg = pc.input(0), pc.input(1)
    pc.set_filter(seg == (b"BUILDING" if seg_type == abi.VARCHAR else BUILDING))       # VarcharOperators.equal over a Slice constant
    pc.project(ck)
    fp_c = FilterAndProjectOperator([abi.BIGINT, seg_type], pc, output_memory=mem, **o)
    local_keys = [block_to_tensor(p.blocks[0], dev).clone() for p in _drain(fp_c, customer, False)]
    mine = torch.cat(local_keys) if local_keys else torch.empty(0, dtype=torch.int64, device=dev)
    # one fixed-shape all-gather: [count, keys...] padded to the shard size (the filter keeps ~1/5)
    cap = max(sum(p.position_count for p in customer), 1)
    send = torch.zeros(cap + 1, dtype=torch.int64, device=dev)
    send[0] = mine.numel()
    send[1:1 + mine.numel()] = mine
    caps = torch.tensor([cap + 1], dtype=torch.int64, device=dev)
    dist.all_reduce(caps, op=dist.ReduceOp.MAX, group=group)
    width = int(caps.item())
    if width > send.numel():
        send = torch.cat([send, torch.zeros(width - send.numel(), dtype=torch.int64, device=dev)])
    gathered = torch.empty(world * width, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, send, group=group)
    gathered = gathered.view(world, width)
    counts = gathered[:, 0].cpu().tolist()
    all_keys = torch.cat([gathered[r, 1:1 + int(counts[r])] for r in range(world)])
    torch.cuda.current_stream().synchronize()      # torch produced the keys on its stream; the operators run on their own
    bridge1 = JoinBridge(opts.get("device", 0), backend=backend)
    build1 = HashBuilderOperator(bridge1, [abi.BIGINT], [0], [], **o)
    if all_keys.numel():
        build1.add_input(Page([tensor_to_block(abi.BIGINT, all_keys)]))
    build1.finish()
    mark("customer filter + replicate + build1")
    # ---- orders: [filter → probe #1] → repartition on orderkey → build #2 ----
    po = Program([abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER])
    ok, ock, od, sp = [po.input(i) for i in range(4)]
    po.set_filter(od < Q3_DATE)
    po.project(ok, ock, od, sp)
    fo = FusedPageOperator([abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER], po, [FusedJoin(bridge1, [abi.BIGINT, abi.BIGINT, abi.DATE, abi.INTEGER], [1], [0, 2, 3])],
                           output_memory=mem, **o)
    part_o = PartitionedOutputOperator([abi.BIGINT, abi.DATE, abi.INTEGER], world, [0], output_memory=mem, **o)
    bridge2 = JoinBridge(opts.get("device", 0), backend=backend)
    build2 = HashBuilderOperator(bridge2, [abi.BIGINT, abi.DATE, abi.INTEGER], [0], [1, 2], **o)
    exchanged = 0
    order_types = [abi.BIGINT, abi.DATE, abi.INTEGER]
    line_types = [abi.BIGINT, abi.DOUBLE]

    def empty(types):
        return Page([tensor_to_block(t, torch.empty(0, dtype=_TORCH[t], device=dev)) for t in types], 0)

    def shuffle(part, fx, page, types):
        """One repartition boundary for one page (collective: every rank calls it the same number of times)."""
        if page is None or not page.position_count:
            page = empty(types)
        if fx is not None:
            return fx.exchange(part, page)
        if not page.position_count:
            return all_to_all_pages(page, [0] * (world + 1), types, group=group)
        part.add_input(page)
        slab = part.get_output(copy=False)
        return all_to_all_pages(slab, part.get_offsets(), types, group=group)

    cap_o = max([p.position_count for p in orders] + [1])
    cap_l = max([p.position_count for p in lineitem] + [1])
    fx_o = fx_l = None
    if fused and exchanges is not None:
        fx_o, fx_l = exchanges
    elif fused:
        fx_o = FusedExchange(order_types, cap_o // 4 + 1024, group=group, backend=backend)
        fx_l = FusedExchange(line_types, cap_l + cap_l // 8 + 1024, group=group, backend=backend)
    for page in orders:            # every rank iterates the same number of pages: the exchange is collective
        fo.add_input(page)
        j = fo.get_output(copy=False)
        recv = shuffle(part_o, fx_o, j, order_types)
        exchanged += recv.position_count
        if recv.position_count:
            build2.add_input(recv)
    build2.finish()
    mark("orders [filter + probe1] + exchange + build2")
    # ---- lineitem: [filter/project] → repartition on orderkey → [probe #2 → aggregation] ----
    pl = Program([abi.BIGINT, abi.DOUBLE, abi.DOUBLE, abi.DATE])
    lk, ep, disc, ship = [pl.input(i) for i in range(4)]
    pl.set_filter(ship > Q3_DATE)
    pl.project(lk, ep * (1.0 - disc))
    fl = FusedPageOperator([abi.BIGINT, abi.DOUBLE, abi.DOUBLE, abi.DATE], pl, [], output_memory=mem, **o)
    part_l = PartitionedOutputOperator([abi.BIGINT, abi.DOUBLE], world, [0], output_memory=mem, **o)
    # after the exchange: probe #2 and the aggregation as ONE operator over the received (orderkey, revenue) rows
    pr = Program([abi.BIGINT, abi.DOUBLE])
    pr.project(pr.input(0), pr.input(1))
    fa = FusedAggregationOperator([abi.BIGINT, abi.DOUBLE], pr, [FusedJoin(bridge2, [abi.BIGINT, abi.DOUBLE], [0], [0, 1])],
                                  [abi.BIGINT, abi.DATE, abi.INTEGER], [0, 2, 3], [Aggregate(abi.AGG_SUM_DOUBLE, 1)],
                                  expected_groups=max(100_000, sum(p.position_count for p in lineitem) // 400),
                                  max_output_rows=1 << 30, output_memory=mem, **o)
    for page in lineitem:
        fl.add_input(page)
        f = fl.get_output(copy=False)
        mark("lineitem filter/project")
        recv = shuffle(part_l, fx_l, f, line_types)
        mark("lineitem exchange " + (str(getattr(fx_l, "trace_ms", "")) if fx_l is not None else "(nccl)"))
        exchanged += recv.position_count
        if not recv.position_count:
            continue
        fa.add_input(recv)
        mark("lineitem [probe2 + aggregation]")
    fa.finish()
    result = []
    while not fa.is_finished():
        r = fa.get_output(copy=False)
        if r is not None:
            result.append(_detach(r))      # device resident; the operator's buffers die with close() below
    mark("aggregation output")
    ops = (fp_c, build1, fo, part_o, build2, fl, part_l, fa)
    if stats is not None:
        stats["joined_rows"] = -1
        stats["exchanged_rows"] = exchanged
        stats["kernel_launches"] = sum(op.get_stats()["kernel_launches"] for op in ops)
        stats.update({"fused_" + k: v for k, v in fa.get_info().items()})
    for op in ops:
        op.close()
    for fx in (fx_o, fx_l):
        if fx is not None and exchanges is None:
            fx.close()
    bridge1.close(); bridge2.close()
    mark("close")
    if stats is not None and trace is not None:
        stats["trace_ms"] = [(b[0], round((b[1] - a[1]) * 1e3, 3)) for a, b in zip(trace[:-1], trace[1:])]
    return result
