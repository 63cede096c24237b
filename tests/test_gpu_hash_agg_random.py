"""-m gpu: HashAggregationOperator on seeded random pages vs the oracle — every tier of the kernel (per-thread tier 0, the
CTA's shared table, the HBM table), the specialised fast kernel and the generic one, table growth and deferred rows.
Bit-exact keys / counts / group order; doubles within 1e-6 relative (north_star)."""
import numpy as np
import pytest

from presto_b200 import (
    AGG_AVG_BIGINT, AGG_AVG_DOUBLE, AGG_COUNT_COLUMN, AGG_COUNT_STAR, AGG_SUM_BIGINT, AGG_SUM_DOUBLE, BIGINT, DATE, DOUBLE, INTEGER,
    MEM_DEVICE, STEP_FINAL, STEP_PARTIAL, VARCHAR, Aggregate, Block, HashAggregationOperator, Page, to_pages,
)
from util import page_to_device, rows_close, rows_of

pytestmark = pytest.mark.gpu

LONG_MIN = -(2 ** 63)


def run(backend, pages, group_types, group_channels, aggs, **kw):
    op = HashAggregationOperator(group_types, group_channels, aggs, backend=backend, **kw)
    rows = rows_of(to_pages(op, pages))
    op.close()
    return rows


def make_pages(rng, n_pages, n, groups, n_vals=4, null_keys=False, null_vals=False, sentinel=False, misalign=False):
    pages = []
    for _ in range(n_pages):
        keys = rng.integers(0, groups, n + 1).astype(np.int64) * 7919 - 3
        if sentinel:
            keys[rng.integers(0, n, max(1, n // 50))] = LONG_MIN
        knull = (rng.random(n + 1) < 0.1).astype(np.uint8) if null_keys else None
        blocks = []
        o = 1 if misalign else 0   # a 1-element offset makes the column pointers 8- but not 16-byte aligned
        blocks.append(Block.from_numpy(BIGINT, keys[o:o + n], None if knull is None else knull[o:o + n]))
        for c in range(n_vals):
            v = rng.random(n + 1) * 1e5 - 2e4
            vn = (rng.random(n + 1) < 0.15).astype(np.uint8) if null_vals else None
            blocks.append(Block.from_numpy(DOUBLE, v[o:o + n], None if vn is None else vn[o:o + n]))
        pages.append(Page(blocks))
    return pages


@pytest.mark.parametrize("groups", [1, 4, 9, 30, 180, 700, 5000, 200_000])
@pytest.mark.parametrize("variant", ["fast", "fast_count", "misaligned", "null_keys", "null_vals", "sentinel"])
def test_bigint_key_double_sums(gpu_backend, oracle_backend, groups, variant):
    rng = np.random.default_rng(groups * 31 + len(variant))
    n = 60_000 if groups < 100_000 else 400_000
    pages = make_pages(rng, 3, n, groups, null_keys=variant == "null_keys", null_vals=variant == "null_vals",
                       sentinel=variant == "sentinel", misalign=variant == "misaligned")
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1 + c) for c in range(4)]
    if variant != "fast":
        aggs += [Aggregate(AGG_COUNT_STAR), Aggregate(AGG_AVG_DOUBLE, 2)]
    got = run(gpu_backend, pages, [BIGINT], [0], aggs, expected_groups=64, max_output_rows=1 << 30)
    want = run(oracle_backend, pages, [BIGINT], [0], aggs, expected_groups=64, max_output_rows=1 << 30)
    rows_close(got, want)


def test_table_growth_and_deferred_rows(gpu_backend, oracle_backend):
    """More distinct groups in one page than the table's claim limit: rows are deferred, the table grows, they are re-run."""
    rng = np.random.default_rng(77)
    n = 6_000_000
    keys = rng.permutation(n).astype(np.int64)          # every row its own group, scrambled
    vals = rng.random(n)
    page = Page([Block.from_numpy(BIGINT, keys), Block.from_numpy(DOUBLE, vals)])
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)]
    op = HashAggregationOperator([BIGINT], [0], aggs, backend=gpu_backend, expected_groups=16, max_output_rows=1 << 30)
    out = to_pages(op, [page, page])
    assert sum(p.position_count for p in out) == n
    k = np.concatenate([p.blocks[0].values for p in out]); s = np.concatenate([p.blocks[1].values for p in out]); c = np.concatenate([p.blocks[2].values for p in out])
    assert (k == keys).all(), "group ids must follow first-appearance order"
    assert (c == 2).all() and np.allclose(s, 2 * vals, rtol=1e-12)
    op.close()


@pytest.mark.parametrize("groups", [3, 300, 40_000])
def test_multi_channel_keys(gpu_backend, oracle_backend, groups):
    """(BIGINT, DATE, INTEGER) keys (TPC-H Q3 shape, W = 3 words) and two varchar(1) keys (Q1 shape, W = 1)."""
    rng = np.random.default_rng(groups)
    n = 80_000
    pages = []
    for _ in range(2):
        g = rng.integers(0, groups, n)
        k0 = Block.from_numpy(BIGINT, g * 1_000_003, (rng.random(n) < 0.02).astype(np.uint8))
        k1 = Block.from_numpy(DATE, (g % 2500 + 8000).astype(np.int32))
        k2 = Block.from_numpy(INTEGER, (g % 3).astype(np.int32), (rng.random(n) < 0.02).astype(np.uint8))
        v = Block.from_numpy(DOUBLE, rng.random(n) * 100)
        w = Block.from_numpy(BIGINT, rng.integers(-10**9, 10**9, n))
        pages.append(Page([k0, k1, k2, v, w]))
    aggs = [Aggregate(AGG_SUM_DOUBLE, 3), Aggregate(AGG_SUM_BIGINT, 4), Aggregate(AGG_AVG_BIGINT, 4), Aggregate(AGG_COUNT_COLUMN, 0), Aggregate(AGG_COUNT_STAR)]
    got = run(gpu_backend, pages, [BIGINT, DATE, INTEGER], [0, 1, 2], aggs, expected_groups=128, max_output_rows=1 << 30)
    want = run(oracle_backend, pages, [BIGINT, DATE, INTEGER], [0, 1, 2], aggs, expected_groups=128, max_output_rows=1 << 30)
    rows_close(got, want)


def test_q1_shape_varchar_keys(gpu_backend, oracle_backend):
    """TPC-H Q1: GROUP BY returnflag, linestatus (varchar(1) each) with the 7 aggregates of PB/HandTpchQuery1.java:113-127."""
    rng = np.random.default_rng(1)
    n = 200_000
    rf = [b"ANR"[i:i + 1] for i in rng.integers(0, 3, n)]
    ls = [b"FO"[i:i + 1] for i in rng.integers(0, 2, n)]
    qty = rng.integers(1, 51, n).astype(np.float64); ep = np.round(rng.random(n) * 1e5, 2); disc = rng.integers(0, 11, n) / 100; tax = rng.integers(0, 9, n) / 100
    page = Page([Block.from_strings(rf), Block.from_strings(ls), Block.from_numpy(DOUBLE, qty), Block.from_numpy(DOUBLE, ep),
                 Block.from_numpy(DOUBLE, ep * (1 - disc)), Block.from_numpy(DOUBLE, ep * (1 - disc) * (1 + tax)), Block.from_numpy(DOUBLE, disc)])
    aggs = [Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_SUM_DOUBLE, 3), Aggregate(AGG_SUM_DOUBLE, 4), Aggregate(AGG_AVG_DOUBLE, 2),
            Aggregate(AGG_AVG_DOUBLE, 5), Aggregate(AGG_AVG_DOUBLE, 6), Aggregate(AGG_COUNT_STAR)]
    kw = dict(expected_groups=10_000, group_max_lengths=[1, 1])
    got = run(gpu_backend, [page], [VARCHAR, VARCHAR], [0, 1], aggs, **kw)
    want = run(oracle_backend, [page], [VARCHAR, VARCHAR], [0, 1], aggs, **kw)
    assert len(got) == 6
    rows_close(got, want)


def test_partial_final_equals_single(gpu_backend, oracle_backend):
    """AggregationNode.Step PARTIAL → FINAL (what P GPUs + one merge do for Q1) equals SINGLE."""
    rng = np.random.default_rng(9)
    pages = make_pages(rng, 4, 50_000, 300, n_vals=2, null_vals=True)
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_AVG_DOUBLE, 2), Aggregate(AGG_COUNT_STAR), Aggregate(AGG_COUNT_COLUMN, 1)]
    single = run(oracle_backend, pages, [BIGINT], [0], aggs, expected_groups=512, max_output_rows=1 << 30)
    partials = []
    for half in (pages[:2], pages[2:]):
        op = HashAggregationOperator([BIGINT], [0], aggs, step=STEP_PARTIAL, backend=gpu_backend, expected_groups=512, max_output_rows=1 << 30)
        partials += to_pages(op, half)
        op.close()
    # FINAL input layout: key, then the intermediate columns of each aggregate in order (sum → 1, avg → 2, count → 1)
    final_aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_AVG_DOUBLE, 2), Aggregate(AGG_COUNT_STAR, 4), Aggregate(AGG_COUNT_COLUMN, 5)]
    got = run(gpu_backend, partials, [BIGINT], [0], final_aggs, step=STEP_FINAL, expected_groups=512, max_output_rows=1 << 30)
    rows_close(sorted(got, key=lambda r: (r[0] is None, r[0])), sorted(single, key=lambda r: (r[0] is None, r[0])))


def test_device_resident_pages(gpu_backend, oracle_backend):
    """Device pages in, device page out: chained GPU operators never round-trip through the host."""
    import ctypes as C
    rng = np.random.default_rng(5)
    n = 100_000
    keys = rng.integers(0, 50, n).astype(np.int64); vals = rng.random(n)
    ptrs = []
    for a in (keys, vals):
        p = C.c_void_p()
        gpu_backend.check(gpu_backend.device_alloc(a.nbytes, C.byref(p)))
        gpu_backend.check(gpu_backend.memcpy_h2d(p, a.ctypes.data, a.nbytes))
        ptrs.append(p)
    page = Page([Block.device(BIGINT, n, ptrs[0].value), Block.device(DOUBLE, n, ptrs[1].value)])
    op = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1)], backend=gpu_backend, output_memory=MEM_DEVICE)
    got = rows_of(to_pages(op, [page]))      # Page.from_c copies the device result back for comparison
    want = run(oracle_backend, [Page([Block.from_numpy(BIGINT, keys), Block.from_numpy(DOUBLE, vals)])], [BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1)])
    rows_close(got, want)
    op.close()
    for p in ptrs:
        gpu_backend.device_free(p)


@pytest.mark.parametrize("groups", [6, 25, 200, 5000], ids=["tier0", "warp_sets", "shared_table_per_sm", "generic_kernel"])
def test_two_32_bit_key_columns_on_the_fast_path(gpu_backend, oracle_backend, groups):
    """Q1's key shape (two INTEGER / DATE codes, no NULLs) with double sums + count(*): the specialised kernel packs the two columns
    into one 64-bit key and re-packs it into the table's two-word layout at the HBM table; beyond ~350 groups the generic kernel
    takes over.  Includes the pair (0, INT_MIN) whose packed value equals the kernel's empty-slot sentinel."""
    rng = np.random.default_rng(groups)
    n = 120_000
    pages = []
    for _ in range(2):
        g = rng.integers(0, groups, n)
        a = (g % 97 - 3).astype(np.int32)
        b = (g // 97 + 9000).astype(np.int32)
        a[:50] = 0; b[:50] = -2**31                       # col1 << 32 | col0 == 0x8000000000000000
        pages.append(Page([Block.from_numpy(INTEGER, a), Block.from_numpy(DATE, b), Block.from_numpy(DOUBLE, rng.random(n) * 10), Block.from_numpy(DOUBLE, rng.random(n))]))
    aggs = [Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_SUM_DOUBLE, 3), Aggregate(AGG_AVG_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)]
    got = run(gpu_backend, pages, [INTEGER, DATE], [0, 1], aggs, expected_groups=groups, max_output_rows=1 << 30)
    want = run(oracle_backend, pages, [INTEGER, DATE], [0, 1], aggs, expected_groups=groups, max_output_rows=1 << 30)
    assert len(want) >= min(groups, 6)
    rows_close(got, want)


def double_key_pages(rng):
    """DOUBLE group keys with both spellings of zero and NaN; the first row of the zero group is -0.0 on page 1 of column a and +0.0 of column b."""
    pages = []
    for p in range(3):
        n = 5000
        a = rng.integers(-3, 4, n).astype(np.float64)
        a[rng.random(n) < 0.3] = -0.0
        b = rng.integers(0, 3, n).astype(np.float64) * 0.5
        b[rng.random(n) < 0.2] = -0.0
        if p == 0:
            a[0], b[0] = -0.0, 0.0
            a[1], b[1] = 0.0, -0.0
        a[rng.random(n) < 0.05] = np.nan
        pages.append(Page([Block.from_numpy(DOUBLE, a), Block.from_numpy(DOUBLE, b), Block.from_numpy(DOUBLE, rng.random(n))]))
    return pages


def key_bits(rows, n_keys):
    import struct
    return [tuple(None if v is None else struct.unpack("<q", struct.pack("<d", v))[0] for v in r[:n_keys]) for r in rows]


def test_double_keys_keep_the_first_rows_zero_spelling(gpu_backend, oracle_backend):
    """DoubleType.equalTo groups -0.0 with +0.0, and the reference's key pages keep the FIRST row's bits
    (MultiChannelGroupByHash.addNewGroup :324-358): the output key column must match the oracle bit for bit, not just as a double."""
    rng = np.random.default_rng(2024)
    pages = double_key_pages(rng)
    aggs = [Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)]
    for types, channels in (([DOUBLE, DOUBLE], [0, 1]), ([DOUBLE], [0])):
        got = run(gpu_backend, pages, types, channels, aggs, expected_groups=64, max_output_rows=1 << 30)
        want = run(oracle_backend, pages, types, channels, aggs, expected_groups=64, max_output_rows=1 << 30)
        rows_close(got, want)
        assert key_bits(got, len(types)) == key_bits(want, len(types))
        assert any(b == (-(1 << 63),) * 1 or (-(1 << 63)) in b for b in key_bits(want, len(types))), "the case must contain a -0.0 first row"


@pytest.mark.parametrize("groups", [3, 400, 150_000])
def test_min_max_random(gpu_backend, oracle_backend, groups):
    """min / max of BIGINT and DOUBLE columns with NULLs next to sums, through tier 0, the shared table and the HBM table (table growth);
    bit-exact against the oracle (min / max pick an input value: no rounding)."""
    from presto_b200 import AGG_MAX_BIGINT, AGG_MAX_DOUBLE, AGG_MIN_BIGINT, AGG_MIN_DOUBLE
    rng = np.random.default_rng(groups + 5)
    pages = []
    for _ in range(3):
        n = 90_000
        k = rng.integers(0, groups, n).astype(np.int64)
        a = rng.integers(-10**15, 10**15, n); an = (rng.random(n) < 0.2).astype(np.uint8)
        d = rng.normal(0, 1e6, n); d[rng.random(n) < 0.01] = np.nan; d[rng.random(n) < 0.01] = np.inf
        d[d == 0] = 1.0             # (+0 / -0 ties keep the first seen zero in the reference; not reproduced)
        dn = (rng.random(n) < 0.3).astype(np.uint8)
        pages.append(Page([Block.from_numpy(BIGINT, k), Block.from_numpy(BIGINT, a, an), Block.from_numpy(DOUBLE, d, dn)]))
    aggs = [Aggregate(AGG_MIN_BIGINT, 1), Aggregate(AGG_MAX_BIGINT, 1), Aggregate(AGG_MIN_DOUBLE, 2), Aggregate(AGG_MAX_DOUBLE, 2), Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)]
    got = run(gpu_backend, pages, [BIGINT], [0], aggs, expected_groups=16, max_output_rows=1 << 30)
    want = run(oracle_backend, pages, [BIGINT], [0], aggs, expected_groups=16, max_output_rows=1 << 30)
    rows_close(got, want)
    assert [r[:5] for r in got] == [r[:5] for r in want] or all((x == y) or (x != x and y != y) for g, w in zip(got, want) for x, y in zip(g[:5], w[:5]))


@pytest.mark.parametrize("groups", [5, 400, 30_000])
def test_decimal_sums_random(gpu_backend, oracle_backend, groups):
    """sum(decimal): short (unscaled longs over the whole 18-digit range) and long (Int128, up to 10^36 so that 10^5 rows stay inside
    decimal(38)) with NULLs, one and three key words; SINGLE and PARTIAL -> FINAL through the serialized states.  Exact integers."""
    from presto_b200 import AGG_SUM_LONG_DECIMAL, AGG_SUM_SHORT_DECIMAL
    rng = np.random.default_rng(groups + 17)
    n = 50_000
    pages = []
    for _ in range(3):
        g = rng.integers(0, groups, n)
        shorts = rng.integers(-(10**18) + 1, 10**18, n)
        longs = [int(a) * 10**18 + int(b) for a, b in zip(rng.integers(-(10**17), 10**17, n), rng.integers(0, 10**18, n))]
        longs = [None if rng.random() < 0.1 else v for v in longs]
        pages.append(Page([Block.from_numpy(BIGINT, g * 1_000_003), Block.from_numpy(DATE, (g % 2500 + 8000).astype(np.int32)),
                           Block.from_numpy(BIGINT, shorts, (rng.random(n) < 0.1).astype(np.uint8)), Block.from_decimals(longs)]))
    aggs = [Aggregate(AGG_SUM_SHORT_DECIMAL, 2), Aggregate(AGG_SUM_LONG_DECIMAL, 3), Aggregate(AGG_COUNT_COLUMN, 3)]
    for types, channels in (([BIGINT], [0]), ([BIGINT, DATE], [0, 1])):
        got = run(gpu_backend, pages, types, channels, aggs, expected_groups=64, max_output_rows=1 << 30)
        want = run(oracle_backend, pages, types, channels, aggs, expected_groups=64, max_output_rows=1 << 30)
        assert got == want
        partials = []
        for half in (pages[:1], pages[1:]):
            op = HashAggregationOperator(types, channels, aggs, step=STEP_PARTIAL, backend=gpu_backend, expected_groups=64, max_output_rows=1 << 30)
            partials += to_pages(op, half)
            op.close()
        k = len(channels)
        final_aggs = [Aggregate(AGG_SUM_SHORT_DECIMAL, k), Aggregate(AGG_SUM_LONG_DECIMAL, k + 1), Aggregate(AGG_COUNT_COLUMN, k + 2)]
        final = run(gpu_backend, partials, types, list(range(k)), final_aggs, step=STEP_FINAL, expected_groups=64, max_output_rows=1 << 30)
        assert sorted(final, key=repr) == sorted(want, key=repr)
        # and the oracle's FINAL accepts the GPU's states (the wire format of the intermediate column is the reference's)
        mixed = run(oracle_backend, partials, types, list(range(k)), final_aggs, step=STEP_FINAL, expected_groups=64, max_output_rows=1 << 30)
        assert sorted(mixed, key=repr) == sorted(want, key=repr)


# ---- VARCHAR keys of any length (MultiChannelGroupByHash over VariableWidthBlock keys; GPU: interned to value ids first) -------------
def _words(rng, count, min_len=0, max_len=40):
    """`count` distinct byte strings, lengths spread over [min_len, max_len], sharing long prefixes (so a comparison has to read them)."""
    out = set()
    while len(out) < count:
        n = int(rng.integers(min_len, max_len + 1))
        out.add((b"presto-b200-key-prefix-" * 2)[:max(0, n - 3)] + bytes(rng.integers(97, 123, min(3, n)).astype(np.uint8)))
    return sorted(out, key=lambda _: rng.random())


@pytest.mark.parametrize("groups", [1, 7, 500, 60_000])
@pytest.mark.parametrize("variant", ["plain", "nulls", "hash_channel", "two_keys", "device_pages"])
def test_long_varchar_keys(gpu_backend, oracle_backend, groups, variant):
    from presto_b200 import _abi as abi
    rng = np.random.default_rng(groups * 13 + len(variant))
    words = _words(rng, groups, 0 if groups > 1 else 12, 45)
    n = 30_000 if groups < 10_000 else 150_000
    pages = []
    for _ in range(3):
        pick = rng.integers(0, groups, n)
        strings = [words[i] for i in pick]
        if variant == "nulls":
            strings = [None if rng.random() < 0.05 else s for s in strings]
        cols = [Block.from_strings(strings), Block.from_numpy(DOUBLE, rng.random(n) * 100), Block.from_numpy(BIGINT, rng.integers(-1000, 1000, n))]
        if variant == "two_keys":
            cols.append(Block.from_numpy(INTEGER, (pick % 3).astype(np.int32), (rng.random(n) < 0.03).astype(np.uint8)))
            cols.append(Block.from_strings([words[(i * 7) % groups][:5] for i in pick]))
        pages.append(Page(cols))
    types, chans = ([VARCHAR, INTEGER, VARCHAR], [0, 3, 4]) if variant == "two_keys" else ([VARCHAR], [0])
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_SUM_BIGINT, 2), Aggregate(AGG_COUNT_COLUMN, 0), Aggregate(AGG_COUNT_STAR)]
    kw = dict(expected_groups=64, max_output_rows=1 << 30)
    if variant == "two_keys":
        kw["group_max_lengths"] = [0, 0, 5]          # the third key fits a packed field, the first does not
    want = run(oracle_backend, pages, types, chans, aggs, **kw)
    if variant == "hash_channel":
        # $hashvalue of a VARCHAR key = XXH64 of the slice (AbstractVariableWidthBlock.hash): precomputed on input, re-emitted on output
        import xxhash
        hashed = []
        for p in pages:
            col = p.blocks[0]
            hv = np.array([np.uint64(xxhash.xxh64_intdigest(bytes(col.values[col.offsets[i]:col.offsets[i + 1]]))).astype(np.int64) for i in range(p.position_count)], dtype=np.int64)
            hashed.append(Page(list(p.blocks) + [Block.from_numpy(BIGINT, hv)]))
        got = run(gpu_backend, hashed, types, chans, aggs, hash_channel=3, **kw)
        want = run(oracle_backend, hashed, types, chans, aggs, hash_channel=3, **kw)
    elif variant == "device_pages":
        op = HashAggregationOperator(types, chans, aggs, backend=gpu_backend, **kw)
        keep = []
        dev = [page_to_device(gpu_backend, p, keep) for p in pages]
        got = rows_of(to_pages(op, dev))
        assert op.get_stats()["kernel_paths"] & abi.PATH_INTERNED_KEYS
        op.close()
        for ptr in keep:
            gpu_backend.device_free(ptr)
    else:
        got = run(gpu_backend, pages, types, chans, aggs, **kw)
    assert len(got) == len(want)
    rows_close(got, want)


def test_long_varchar_keys_partial_final(gpu_backend, oracle_backend):
    """PARTIAL → FINAL through pages with an interned key: the intermediate pages carry the slices, every operator interns its own."""
    rng = np.random.default_rng(21)
    words = _words(rng, 900, 9, 30)
    pages = []
    for _ in range(4):
        pick = rng.integers(0, 900, 40_000)
        pages.append(Page([Block.from_strings([words[i] for i in pick]), Block.from_numpy(DOUBLE, rng.random(40_000)), ]))
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_AVG_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)]
    single = run(oracle_backend, pages, [VARCHAR], [0], aggs, expected_groups=512, max_output_rows=1 << 30)
    partials = []
    for half in (pages[:2], pages[2:]):
        op = HashAggregationOperator([VARCHAR], [0], aggs, step=STEP_PARTIAL, backend=gpu_backend, expected_groups=512, max_output_rows=1 << 30)
        partials += to_pages(op, half)
        op.close()
    final_aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_AVG_DOUBLE, 2), Aggregate(AGG_COUNT_STAR, 4)]
    got = run(gpu_backend, partials, [VARCHAR], [0], final_aggs, step=STEP_FINAL, expected_groups=512, max_output_rows=1 << 30)
    rows_close(sorted(got), sorted(single))


def test_long_varchar_keys_many_distinct_values(gpu_backend, oracle_backend):
    """Every row a new value over several chunks of the interner (table / pool / entry arrays grow and are re-hashed)."""
    n = 300_000
    pages = []
    for p in range(2):
        strings = [b"order-comment-%09d-%s" % (i, b"x" * (i % 23)) for i in range(p * n // 2, p * n // 2 + n)]   # the second page repeats half of the first
        pages.append(Page([Block.from_strings(strings), Block.from_numpy(BIGINT, np.arange(n, dtype=np.int64))]))
    aggs = [Aggregate(AGG_SUM_BIGINT, 1), Aggregate(AGG_COUNT_STAR)]
    got = run(gpu_backend, pages, [VARCHAR], [0], aggs, expected_groups=16, max_output_rows=1 << 30)
    want = run(oracle_backend, pages, [VARCHAR], [0], aggs, expected_groups=16, max_output_rows=1 << 30)
    assert len(got) == n * 3 // 2
    assert got == want
