"""-m gpu: fused pipeline fragments (FusedAggregationOperator = FilterAndProject → LookupJoin* → HashAggregation in one generated
kernel) against the SAME fragment run operator by operator on the oracle.  Bit-exact keys / counts / group order, double sums
within 1e-6 relative (north_star); the operator reports which path ran, so every case also pins that the kernel (or, for the
shapes that cannot be fused, the unfused chain) is what produced the result."""
import os

import numpy as np
import pytest

from presto_b200 import (
    AGG_AVG_BIGINT, AGG_AVG_DOUBLE, AGG_COUNT_STAR, AGG_SUM_BIGINT, AGG_SUM_DOUBLE, BIGINT, BOOLEAN, DATE, DOUBLE, INTEGER, STEP_FINAL, STEP_PARTIAL, VARCHAR,
    Aggregate, Block, FusedAggregationOperator, FusedJoin, HashAggregationOperator, HashBuilderOperator, JoinBridge, Page, PrestoException, Program, to_pages,
)
from presto_b200 import _abi as abi
from presto_b200.pipelines import Q1_AGGREGATES, q1_program_typed, run_q1, run_q3
from presto_b200.synthetic import q3_tables_np
from util import rows_close, rows_of

pytestmark = pytest.mark.gpu

FUSED = abi.PATH_FUSED_TIERS | abi.PATH_FUSED_DIRECT


def run_fused(backend, pages, *args, info=None, **kw):
    op = FusedAggregationOperator(*args, backend=backend, max_output_rows=1 << 30, **kw)
    rows = rows_of(to_pages(op, pages))
    if info is not None:
        info.update(op.get_info()); info.update(op.get_stats())
    op.close()
    return rows


def q1_arrays(rng, n, key_type, flags=3):
    rf, ls = rng.integers(0, flags, n), rng.integers(0, 2, n)
    if key_type == VARCHAR:
        keys = [Block.from_strings([b"ANR"[i:i + 1] for i in rf]), Block.from_strings([b"FO"[i:i + 1] for i in ls])]
    else:
        keys = [Block.from_numpy(INTEGER, rf.astype(np.int32)), Block.from_numpy(INTEGER, ls.astype(np.int32))]
    vals = [rng.integers(1, 51, n).astype(np.float64), np.round(rng.random(n) * 1e5, 2), rng.integers(0, 11, n) / 100.0, rng.integers(0, 9, n) / 100.0]
    ship = rng.integers(8035, 10600, n).astype(np.int32)
    return keys + [Block.from_numpy(DOUBLE, v) for v in vals] + [Block.from_numpy(DATE, ship)]


@pytest.mark.parametrize("key_type", [INTEGER, VARCHAR])
def test_q1_fragment(gpu_backend, oracle_backend, key_type):
    """TPC-H Q1 (HandTpchQuery1.java:99-127, 237-330) with INTEGER-coded and with the reference's VARCHAR(1) returnflag / linestatus."""
    rng = np.random.default_rng(11)
    pages = [Page(q1_arrays(rng, n, key_type)) for n in (70_001, 50_000, 3)]
    want = rows_of(run_q1(pages, backend=oracle_backend, fused=True, key_type=key_type))
    info = {}
    got = rows_of(run_q1(pages, backend=gpu_backend, fused=True, key_type=key_type, info=info))
    assert len(want) == 6
    rows_close(got, want)
    assert info["fused_pages"] == 3 and info["unfused_pages"] == 0 and info["kernel_paths"] & abi.PATH_FUSED_TIERS, info
    # the unfused plan (separate operators, intermediate pages; VARCHAR keys pass through the PageProcessor) gives the same rows
    rows_close(rows_of(run_q1(pages, backend=gpu_backend, fused=False, key_type=key_type)), want)


@pytest.mark.parametrize("groups,step", [(5, 0), (300, 0), (60_000, 0), (2_000_000, 0), (1000, STEP_PARTIAL)])
def test_filter_project_bigint_key(gpu_backend, oracle_backend, groups, step):
    """One BIGINT key computed by the page processor (k % groups would need a modulus: keys come in ready, the VALUES are computed),
    every cardinality regime: tiers, bypass of a full shared table, direct mode, table growth with deferred rows."""
    rng = np.random.default_rng(groups)
    n = 300_000 if groups < 1_000_000 else 1_500_000
    pages = []
    for _ in range(3):
        k = rng.integers(0, groups, n).astype(np.int64) * 7919 - 3
        a, b = rng.random(n) * 100, rng.integers(0, 50, n).astype(np.float64)
        flag = rng.integers(0, 100, n).astype(np.int32)
        pages.append(Page([Block.from_numpy(BIGINT, k), Block.from_numpy(DOUBLE, a), Block.from_numpy(DOUBLE, b), Block.from_numpy(INTEGER, flag)]))
    types = [BIGINT, DOUBLE, DOUBLE, INTEGER]
    prog = Program(types)
    k, a, b, f = [prog.input(i) for i in range(4)]
    prog.set_filter((f >= 10) & (a < 90.0))
    prog.project(k, a * b, a + 1.5, b)
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_AVG_DOUBLE, 3), Aggregate(AGG_COUNT_STAR)]
    args = (types, prog, [], [BIGINT], [0], aggs)
    want = run_fused(oracle_backend, pages, *args, expected_groups=64, step=step)
    info = {}
    got = run_fused(gpu_backend, pages, *args, expected_groups=64, step=step, info=info)
    rows_close(got, want)
    assert info["fused_pages"] == 3 and info["unfused_pages"] == 0 and info["kernel_paths"] & FUSED, info
    if groups >= 60_000:
        assert info["kernel_paths"] & abi.PATH_FUSED_DIRECT


def test_null_keys_null_filter_inputs_and_special_keys(gpu_backend, oracle_backend):
    """NULL group keys (the Bigint path's null group), NULL filter inputs (three-valued: NULL does not select), the key equal to the
    table's empty-slot sentinel."""
    rng = np.random.default_rng(3)
    n = 120_000
    k = rng.integers(0, 40, n).astype(np.int64)
    k[rng.integers(0, n, 500)] = -(1 << 63)
    kn = (rng.random(n) < 0.05).astype(np.uint8)
    f = rng.integers(0, 10, n).astype(np.int64); fn = (rng.random(n) < 0.2).astype(np.uint8)
    v = rng.random(n)
    page = Page([Block.from_numpy(BIGINT, k, kn), Block.from_numpy(BIGINT, f, fn), Block.from_numpy(DOUBLE, v)])
    types = [BIGINT, BIGINT, DOUBLE]
    prog = Program(types)
    ek, ef, ev = [prog.input(i) for i in range(3)]
    prog.set_filter((ef > 2) | ef.is_null())
    prog.project(ek, ev * 2.0)
    args = (types, prog, [], [BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)])
    want = run_fused(oracle_backend, [page, page], *args)
    info = {}
    got = run_fused(gpu_backend, [page, page], *args, info=info)
    rows_close(got, want)
    assert info["fused_pages"] == 2, info
    # multi-column key with NULLs and a BOOLEAN member
    kb = rng.integers(0, 2, n).astype(np.int8); kbn = (rng.random(n) < 0.1).astype(np.uint8)
    kd = rng.integers(8000, 8005, n).astype(np.int32)
    page2 = Page([Block.from_numpy(BOOLEAN, kb, kbn), Block.from_numpy(DATE, kd), Block.from_numpy(DOUBLE, v)])
    types2 = [BOOLEAN, DATE, DOUBLE]
    prog2 = Program(types2)
    a, b, c = [prog2.input(i) for i in range(3)]
    prog2.project(a, b, c)
    args2 = (types2, prog2, [], [BOOLEAN, DATE], [0, 1], [Aggregate(AGG_SUM_DOUBLE, 2)])
    rows_close(run_fused(gpu_backend, [page2], *args2, info=info), run_fused(oracle_backend, [page2], *args2))
    assert info["fused_pages"] == 1, info


def build_side(backend, keys, payload_cols, types):
    bridge = JoinBridge(backend=backend)
    b = HashBuilderOperator(bridge, types, [0], list(range(1, len(types))), backend=backend)
    to_pages(b, [Page([Block.from_numpy(types[0], keys)] + [Block.from_numpy(t, c) for t, c in zip(types[1:], payload_cols)])])
    return bridge, b


@pytest.mark.parametrize("n_orders", [3000, 60_000])
def test_q3_lineitem_fragment(gpu_backend, oracle_backend, n_orders):
    """lineitem filter/project → probe orders' dense index → GROUP BY (orderkey, orderdate, shippriority): sum(revenue) — the whole plan
    through pipelines.run_q3 (fused lineitem Driver) vs the oracle's operator chain."""
    customer, orders, lineitem = q3_tables_np(n_orders)
    def pages_of(types, arrays, n_pages):
        n = len(arrays[0]); cuts = np.linspace(0, n, n_pages + 1).astype(int)
        return [Page([Block.from_numpy(t, a[cuts[i]:cuts[i + 1]]) for t, a in zip(types, arrays)]) for i in range(n_pages)]
    cp = pages_of([BIGINT, INTEGER], customer, 1)
    op = pages_of([BIGINT, BIGINT, DATE, INTEGER], orders, 2)
    lp = pages_of([BIGINT, DOUBLE, DOUBLE, DATE], lineitem, 3)
    want = rows_of(run_q3(cp, op, lp, backend=oracle_backend))
    stats = {}
    got = rows_of(run_q3(cp, op, lp, backend=gpu_backend, fused=True, stats=stats))
    assert len(want) > 5
    rows_close(got, want)
    assert stats["fused_fused_pages"] == 3 and stats["fused_unfused_pages"] == 0 and stats["kernel_paths"] & FUSED, stats


def test_two_joins_and_build_nulls(gpu_backend, oracle_backend):
    """fact → join dim1 (payload with NULLs, used as group key) → join dim2 (no payload: pure membership) → aggregate."""
    rng = np.random.default_rng(9)
    n = 200_000
    d1_keys = rng.permutation(5000).astype(np.int64)[:4000] * 3 + 100
    d1_cat = rng.integers(0, 7, 4000).astype(np.int32); d1_null = (rng.random(4000) < 0.1).astype(np.uint8)
    d1_w = rng.random(4000)
    d2_keys = rng.permutation(3000).astype(np.int64)[:2000] + 50
    fk1 = rng.integers(0, 16000, n).astype(np.int64); fk2 = rng.integers(0, 3500, n).astype(np.int64)
    val = rng.random(n) * 10
    page = Page([Block.from_numpy(BIGINT, fk1), Block.from_numpy(BIGINT, fk2), Block.from_numpy(DOUBLE, val)])

    def run(backend, info=None):
        b1 = JoinBridge(backend=backend); b2 = JoinBridge(backend=backend)
        h1 = HashBuilderOperator(b1, [BIGINT, INTEGER, DOUBLE], [0], [1, 2], backend=backend)
        to_pages(h1, [Page([Block.from_numpy(BIGINT, d1_keys), Block.from_numpy(INTEGER, d1_cat, d1_null), Block.from_numpy(DOUBLE, d1_w)])])
        h2 = HashBuilderOperator(b2, [BIGINT], [0], [], backend=backend)
        to_pages(h2, [Page([Block.from_numpy(BIGINT, d2_keys)])])
        types = [BIGINT, BIGINT, DOUBLE]
        prog = Program(types)
        a, b, v = [prog.input(i) for i in range(3)]
        prog.set_filter(v > 1.0)
        prog.project(a, b, v)
        joins = [FusedJoin(b1, [BIGINT, BIGINT, DOUBLE], [0], [1, 2]),          # → (fk2, val, cat, w)
                 FusedJoin(b2, [BIGINT, DOUBLE, INTEGER, DOUBLE], [0], [1, 2, 3])]   # → (val, cat, w)
        rows = run_fused(backend, [page, page], types, prog, joins, [INTEGER], [1], [Aggregate(AGG_SUM_DOUBLE, 0), Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)], info=info)
        for x in (h1, h2): x.close()
        b1.close(); b2.close()
        return rows
    info = {}
    got, want = run(gpu_backend, info), run(oracle_backend)
    assert len(want) == 8      # 7 categories + the NULL group
    rows_close(got, want)
    assert info["fused_pages"] == 2 and info["unfused_pages"] == 0, info


def test_unfusable_fragments_run_the_operator_chain(gpu_backend, oracle_backend):
    """Shapes outside the fused set still give the chain's result, operator by operator on the GPU: a build side with duplicate keys
    (hash-table LookupSource), a nullable aggregate input, a bigint sum, PG_FUSED=0."""
    rng = np.random.default_rng(21)
    n = 50_000
    k = rng.integers(0, 300, n).astype(np.int64); v = rng.random(n); vn = (rng.random(n) < 0.1).astype(np.uint8)
    w = rng.integers(-1000, 1000, n).astype(np.int64)
    types = [BIGINT, DOUBLE, BIGINT]
    prog = Program(types)
    a, b, c = [prog.input(i) for i in range(3)]
    prog.set_filter(a >= 10)
    prog.project(a, b * 2.0, c)
    page_nulls = Page([Block.from_numpy(BIGINT, k), Block.from_numpy(DOUBLE, v, vn), Block.from_numpy(BIGINT, w)])
    page_plain = Page([Block.from_numpy(BIGINT, k), Block.from_numpy(DOUBLE, v), Block.from_numpy(BIGINT, w)])
    for page, aggs, expect_fused in [(page_nulls, [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)], 0),
                                     (page_plain, [Aggregate(AGG_SUM_BIGINT, 2), Aggregate(AGG_SUM_DOUBLE, 1)], 0),
                                     (page_plain, [Aggregate(AGG_AVG_BIGINT, 2), Aggregate(AGG_SUM_DOUBLE, 1)], 1)]:
        args = (types, prog, [], [BIGINT], [0], aggs)
        info = {}
        got = run_fused(gpu_backend, [page, page], *args, info=info)
        rows_close(got, run_fused(oracle_backend, [page, page], *args))
        assert (info["fused_pages"] == 2) == bool(expect_fused) and (info["unfused_pages"] == 2) == (not expect_fused), info
        if not expect_fused:
            assert info["kernel_paths"] & abi.PATH_UNFUSED_CHAIN and info["reason"], info
    # a page with a null map after pages without: the fragment is compiled per null signature; this one is not fusable (nullable input)
    args = (types, prog, [], [BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)])
    info = {}
    got = run_fused(gpu_backend, [page_plain, page_nulls, page_plain], *args, info=info)
    rows_close(got, run_fused(oracle_backend, [page_plain, page_nulls, page_plain], *args))
    assert info["fused_pages"] == 2 and info["unfused_pages"] == 1, info
    # duplicate build keys → the LookupSource is a hash table → the chain
    def dup(backend, info=None):
        br = JoinBridge(backend=backend)
        h = HashBuilderOperator(br, [BIGINT, DOUBLE], [0], [1], backend=backend)
        to_pages(h, [Page([Block.from_numpy(BIGINT, np.array([5, 7, 5, 9, 7, 7], dtype=np.int64) + 10), Block.from_numpy(DOUBLE, np.arange(6) * 1.0)])])
        rows = run_fused(backend, [page_plain], types, prog, [FusedJoin(br, [BIGINT, DOUBLE, BIGINT], [0], [0, 1])], [BIGINT], [0],
                         [Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)], info=info)
        h.close(); br.close()
        return rows
    info = {}
    rows_close(dup(gpu_backend, info), dup(oracle_backend))
    assert info["unfused_pages"] == 1 and "hash table" in info["reason"], info
    os.environ["PG_FUSED"] = "0"
    try:
        info = {}
        rows_close(run_fused(gpu_backend, [page_plain], *args, info=info), run_fused(oracle_backend, [page_plain], *args))
        assert info["unfused_pages"] == 1 and info["fused_pages"] == 0
    finally:
        del os.environ["PG_FUSED"]


def test_errors_inside_a_fused_fragment(gpu_backend):
    """BIGINT overflow / division by zero in the page processor half of the kernel raise like FilterAndProjectOperator does; rows the
    filter drops do not."""
    types = [BIGINT, DOUBLE]
    for build, code in [(lambda x: x * x * x * x * x, 6), (lambda x: x / (x - x), 1)]:
        prog = Program(types)
        x, v = prog.input(0), prog.input(1)
        prog.project(x, build(x).cast_double() + v)
        op = FusedAggregationOperator(types, prog, [], [BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1)], backend=gpu_backend)
        with pytest.raises(PrestoException) as e:
            op.add_input(Page([Block.from_values(BIGINT, [3, 100000, 7]), Block.from_values(DOUBLE, [1.0, 2.0, 3.0])]))
            op.finish(); op.get_output()
        assert e.value.code == code
        op.close()
    prog = Program(types)
    x, v = prog.input(0), prog.input(1)
    prog.set_filter(x < 10)
    prog.project(x, (x * x * x * x * x).cast_double() + v)
    rows = run_fused(gpu_backend, [Page([Block.from_values(BIGINT, [3, 100000, 2]), Block.from_values(DOUBLE, [1.0, 2.0, 3.0])])], types, prog, [], [BIGINT], [0],
                     [Aggregate(AGG_SUM_DOUBLE, 1)])
    assert rows == [(3, 244.0), (2, 35.0)]


def test_device_resident_pages(gpu_backend, oracle_backend):
    """Device pages are consumed in place (no staging copy); misaligned device columns fall back to the chain."""
    import ctypes as C
    rng = np.random.default_rng(4)
    n = 100_001
    host = Page(q1_arrays(rng, n + 1, INTEGER))
    keep = []
    def dev(page, skew):
        blocks = []
        for b in page.blocks:
            p = C.c_void_p()
            gpu_backend.check(gpu_backend.device_alloc(b.values.nbytes + 64, C.byref(p)))
            w = b.values.itemsize
            gpu_backend.check(gpu_backend.memcpy_h2d(p.value + skew * w, b.values.ctypes.data + skew * w, b.values.nbytes - skew * w))
            keep.append(p)
            blocks.append(Block.device(b.type, n, p.value + skew * w))
        return Page(blocks, n)
    for skew, expect_fused in [(0, 1), (1, 0)]:
        hp = Page([Block.from_numpy(b.type, b.values[skew:skew + n]) for b in host.blocks])
        want = rows_of(run_q1([hp], backend=oracle_backend, fused=True))
        info = {}
        got = rows_of(run_q1([dev(host, skew)], backend=gpu_backend, fused=True, device_resident=False, info=info))
        rows_close(got, want)
        assert info["fused_pages"] == expect_fused, info
    for p in keep:
        gpu_backend.device_free(p)


def run_page_op(backend, pages, types, prog, joins, info=None):
    from presto_b200 import FusedPageOperator
    op = FusedPageOperator(types, prog, joins, backend=backend)
    rows = rows_of(to_pages(op, pages))
    if info is not None:
        info.update(op.get_info())
    op.close()
    return rows


def test_page_sink_filter_join_pages(gpu_backend, oracle_backend):
    """FusedPageOperator: filter → probe → output page, rows in input order, bit-exact against the chained reference operators: pure
    membership join (Q3's orders ⋈ BUILDING customers), payload join with NULLs on both sides, no join at all, a page of all misses."""
    rng = np.random.default_rng(33)
    n = 150_000

    def run(backend, info=None):
        out = []
        b1 = JoinBridge(backend=backend); b2 = JoinBridge(backend=backend)
        h1 = HashBuilderOperator(b1, [BIGINT], [0], [], backend=backend)
        to_pages(h1, [Page([Block.from_numpy(BIGINT, (rng2.permutation(40_000)[:9000] + 7).astype(np.int64))])])
        pay_nulls = (rng2.random(5000) < 0.2).astype(np.uint8)
        h2 = HashBuilderOperator(b2, [BIGINT, DOUBLE, INTEGER], [0], [1, 2], backend=backend)
        to_pages(h2, [Page([Block.from_numpy(BIGINT, (rng2.permutation(20_000)[:5000] * 2).astype(np.int64)), Block.from_numpy(DOUBLE, rng2.random(5000), pay_nulls),
                            Block.from_numpy(INTEGER, rng2.integers(0, 9, 5000).astype(np.int32))])])
        # (1) membership join, Q3 orders shape
        types = [BIGINT, BIGINT, DATE, INTEGER]
        prog = Program(types)
        ok, ck, od, sp = [prog.input(i) for i in range(4)]
        prog.set_filter(od < 9204)
        prog.project(ok, ck, od, sp)
        pages = [Page([Block.from_numpy(BIGINT, rng2.integers(0, 10**9, m)), Block.from_numpy(BIGINT, rng2.integers(0, 45_000, m)),
                       Block.from_numpy(DATE, rng2.integers(8035, 10441, m).astype(np.int32)), Block.from_numpy(INTEGER, np.zeros(m, dtype=np.int32))]) for m in (n, 777, 1)]
        i1 = {}
        out.append(run_page_op(backend, pages, types, prog, [FusedJoin(b1, types, [1], [0, 2, 3])], i1))
        # (2) payload join, nullable probe key and a computed probe-side column
        types2 = [BIGINT, DOUBLE, INTEGER]
        prog2 = Program(types2)
        k, v, c = [prog2.input(i) for i in range(3)]
        prog2.set_filter((v > 0.25) | c.is_null())
        prog2.project(k, v * 2.0, c + 1)
        kn = (rng2.random(n) < 0.05).astype(np.uint8); cn = (rng2.random(n) < 0.1).astype(np.uint8)
        pages2 = [Page([Block.from_numpy(BIGINT, rng2.integers(0, 45_000, n), kn), Block.from_numpy(DOUBLE, rng2.random(n)), Block.from_numpy(INTEGER, rng2.integers(0, 100, n).astype(np.int32), cn)])]
        i2 = {}
        out.append(run_page_op(backend, pages2, types2, prog2, [FusedJoin(b2, [BIGINT, DOUBLE, BIGINT], [0], [0, 1, 2])], i2))
        # (3) no join: plain filter / project; (4) every probe misses
        i3 = {}
        out.append(run_page_op(backend, pages2, types2, prog2, [], i3))
        miss = [Page([Block.from_numpy(BIGINT, np.full(1000, 10**12, dtype=np.int64)), Block.from_numpy(DOUBLE, np.ones(1000)), Block.from_numpy(INTEGER, np.ones(1000, dtype=np.int32))])]
        out.append(run_page_op(backend, miss, types2, prog2, [FusedJoin(b2, [BIGINT, DOUBLE, BIGINT], [0], [0, 1, 2])]))
        for x in (h1, h2): x.close()
        b1.close(); b2.close()
        if info is not None:
            info.update({"i1": i1, "i2": i2, "i3": i3})
        return out
    rng2 = np.random.default_rng(34); want = run(oracle_backend)
    rng2 = np.random.default_rng(34); info = {}; got = run(gpu_backend, info)
    assert [len(x) for x in want] == [len(x) for x in got] and len(want[0]) > 1000 and len(want[1]) > 1000 and len(want[3]) == 0
    for g, w in zip(got, want):
        rows_close(g, w, rel=0.0)
    assert info["i1"]["fused_pages"] == 3 and info["i2"]["fused_pages"] == 1 and info["i3"]["fused_pages"] == 1, info


@pytest.mark.parametrize("key_type", [INTEGER, VARCHAR])
def test_tma_ring_path(gpu_backend, oracle_backend, key_type):
    """Pages large enough for the cp.async.bulk staging ring (fused_kernel_tma: eager columns land in shared-memory stages, a producer
    warp keeps them in flight): same rows as the oracle, and the operator reports that the ring kernel ran.  Page sizes that are not a
    multiple of the 512-row tile exercise the direct-load tail."""
    rng = np.random.default_rng(77)
    pages = [Page(q1_arrays(rng, n, key_type, flags=2)) for n in (330_001, 400_000, 512 * 700, 1023)]     # 4 groups, like TPC-H's Q1
    want = rows_of(run_q1(pages, backend=oracle_backend, fused=True, key_type=key_type))
    info = {}
    got = rows_of(run_q1(pages, backend=gpu_backend, fused=True, key_type=key_type, info=info))
    assert len(want) == 4
    rows_close(got, want)
    assert info["fused_pages"] == 4 and info["kernel_paths"] & abi.PATH_FUSED_TMA, info
    os.environ["PG_FUSED_TMA"] = "0"
    try:
        # a fresh program signature is not needed: the switch is read when the fragment is compiled per operator
        info2 = {}
        got2 = rows_of(run_q1(pages, backend=gpu_backend, fused=True, key_type=key_type, info=info2))
        rows_close(got2, want)
        assert not (info2["kernel_paths"] & abi.PATH_FUSED_TMA), info2
    finally:
        del os.environ["PG_FUSED_TMA"]


def test_varchar_predicate_and_long_varchar_key_run_the_operator_chain(gpu_backend, oracle_backend):
    """Q3's customer fragment shape — filter `mktsegment = 'BUILDING'` (VarcharOperators.equal over a Slice constant), GROUP BY a VARCHAR
    channel of unbounded length — is outside the generated kernels' set: the fused operator runs FilterAndProject (interpreter kernels, the
    literal adopted by the operator) → HashAggregation over interned keys, same rows as the oracle's chain."""
    rng = np.random.default_rng(5)
    n = 40_000
    segments = [b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"HOUSEHOLD", b"MACHINERY"]
    nations = [b"nation-%02d-with-a-long-name" % i for i in range(25)]
    seg = [segments[i] for i in rng.integers(0, 5, n)]
    nat = [None if rng.random() < 0.02 else nations[i] for i in rng.integers(0, 25, n)]
    bal = rng.random(n) * 1e4
    types = [VARCHAR, VARCHAR, DOUBLE]
    page = Page([Block.from_strings(seg), Block.from_strings(nat), Block.from_numpy(DOUBLE, bal)])
    prog = Program(types)
    s, c, b = [prog.input(i) for i in range(3)]
    prog.set_filter((s == b"BUILDING") & (b > 100.0))
    prog.project(c, b)
    args = (types, prog, [], [VARCHAR], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)])
    info = {}
    got = run_fused(gpu_backend, [page, page], *args, info=info)
    want = run_fused(oracle_backend, [page, page], *args)
    assert len(want) == 26
    rows_close(got, want)
    assert info["unfused_pages"] == 2 and info["kernel_paths"] & abi.PATH_UNFUSED_CHAIN and info["kernel_paths"] & abi.PATH_INTERNED_KEYS, info
