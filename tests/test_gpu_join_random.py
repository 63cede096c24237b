"""-m gpu: HashBuilderOperator + LookupJoinOperator on seeded random pages vs the oracle, bit-exact including output order.
Covers both LookupSource structures of the GPU path (bitmap + rank index for unique dense keys, hash table + position links
otherwise), NULL keys on either side, duplicate chains, INNER and PROBE_OUTER, multi-page build and probe."""
import numpy as np
import pytest

from presto_b200 import (
    BIGINT, DATE, DOUBLE, INTEGER, JOIN_INNER, JOIN_PROBE_OUTER, Block, HashBuilderOperator, JoinBridge, LookupJoinOperator, Page, to_pages,
)
from util import rows_of

pytestmark = pytest.mark.gpu


def join(backend, build_pages, build_types, probe_pages, probe_types, join_type, build_out, probe_out, key_type=BIGINT):
    bridge = JoinBridge(backend=backend)
    builder = HashBuilderOperator(bridge, build_types, [0], build_out, backend=backend)
    assert to_pages(builder, build_pages) == []
    probe = LookupJoinOperator(bridge, probe_types, [0], probe_out, join_type=join_type, backend=backend)
    rows = rows_of(to_pages(probe, probe_pages))
    probe.close(); builder.close(); bridge.close()
    return rows


def build_side(rng, kind, n, key_type):
    if kind == "dense_unique":          # TPC-H order keys: 8 of every 32 integers, shuffled
        i = rng.permutation(n) + 1
        keys = ((i >> 3) << 5) + (i & 7)
    elif kind == "sparse_unique":       # unique but spread over 2^40: hash table path
        keys = rng.choice(1 << 40, n, replace=False)
    elif kind == "duplicates":          # ~4 rows per key: position-link chains
        keys = rng.integers(0, max(n // 4, 1), n)
    else:
        raise ValueError(kind)
    if key_type != BIGINT:
        keys = keys % (1 << 31)
        if kind != "duplicates":
            keys = np.unique(keys); rng.shuffle(keys); n = len(keys)
    return keys.astype(np.int64), n


@pytest.mark.parametrize("kind", ["dense_unique", "sparse_unique", "duplicates"])
@pytest.mark.parametrize("join_type", [JOIN_INNER, JOIN_PROBE_OUTER])
@pytest.mark.parametrize("nulls", [False, True])
def test_join_matches_oracle(gpu_backend, oracle_backend, kind, join_type, nulls):
    rng = np.random.default_rng(sum(map(ord, kind)) * 4 + join_type * 2 + int(nulls))
    nb = 30_000
    keys, nb = build_side(rng, kind, nb, BIGINT)
    knull = (rng.random(nb) < 0.05).astype(np.uint8) if nulls else None
    pay_a = rng.integers(8000, 10500, nb).astype(np.int32)
    pay_b = rng.random(nb)
    pnull = (rng.random(nb) < 0.1).astype(np.uint8) if nulls else None
    half = nb // 2
    build_pages = [Page([Block.from_numpy(BIGINT, keys[s], None if knull is None else knull[s]), Block.from_numpy(DATE, pay_a[s]),
                         Block.from_numpy(DOUBLE, pay_b[s], None if pnull is None else pnull[s])]) for s in (slice(0, half), slice(half, nb))]
    probe_pages = []
    for _ in range(3):
        npr = 50_000
        hit = keys[rng.integers(0, nb, npr)]
        miss = rng.integers(-1000, 1 << 41, npr)
        pk = np.where(rng.random(npr) < 0.7, hit, miss)
        pkn = (rng.random(npr) < 0.05).astype(np.uint8) if nulls else None
        probe_pages.append(Page([Block.from_numpy(BIGINT, pk, pkn), Block.from_numpy(DOUBLE, rng.random(npr)), Block.from_numpy(INTEGER, rng.integers(0, 100, npr).astype(np.int32))]))
    args = (build_pages, [BIGINT, DATE, DOUBLE], probe_pages, [BIGINT, DOUBLE, INTEGER], join_type, [1, 2, 0], [0, 1, 2])
    got = join(gpu_backend, *args)
    want = join(oracle_backend, *args)
    assert len(got) == len(want) and len(got) > 0
    assert got == want   # byte/integer values, double payload passthrough and the row ORDER are all exact


def test_sorted_probe_against_dense_index(gpu_backend, oracle_backend):
    """orders ⋈ lineitem shape: unique sparse order keys (shuffled build), probe clustered by key with 1..7 rows per key."""
    rng = np.random.default_rng(12)
    nb = 200_000
    i = np.arange(1, nb + 1)
    okey = ((i >> 3) << 5) + (i & 7)
    perm = rng.permutation(nb)
    build = Page([Block.from_numpy(BIGINT, okey[perm]), Block.from_numpy(DATE, rng.integers(8035, 10441, nb).astype(np.int32)), Block.from_numpy(INTEGER, np.zeros(nb, dtype=np.int32))])
    lkey = np.repeat(okey, rng.integers(1, 8, nb))
    npr = len(lkey)
    probe = Page([Block.from_numpy(BIGINT, lkey), Block.from_numpy(DOUBLE, rng.random(npr) * 1e5), Block.from_numpy(DOUBLE, rng.integers(0, 11, npr) / 100)])
    args = ([build], [BIGINT, DATE, INTEGER], [probe], [BIGINT, DOUBLE, DOUBLE], JOIN_INNER, [1, 2], [0, 1, 2])
    got = join(gpu_backend, *args)
    want = join(oracle_backend, *args)
    assert len(got) == npr and got == want


@pytest.mark.parametrize("key_type", [INTEGER, DATE])
def test_int32_join_keys(gpu_backend, oracle_backend, key_type):
    rng = np.random.default_rng(3)
    keys, nb = build_side(rng, "dense_unique", 5000, key_type)
    build = Page([Block.from_numpy(key_type, keys.astype(np.int32)), Block.from_numpy(BIGINT, np.arange(nb))])
    pk = rng.integers(0, int(keys.max()) + 50, 20_000).astype(np.int32)
    probe = Page([Block.from_numpy(key_type, pk)])
    args = ([build], [key_type, BIGINT], [probe], [key_type], JOIN_PROBE_OUTER, [1], [0])
    assert join(gpu_backend, *args) == join(oracle_backend, *args)


@pytest.mark.parametrize("payload", ["date", "date_int", "double_nullable_date", "bigint_double_nullable_int"])
@pytest.mark.parametrize("with_duplicate", [False, True])
def test_range_partitioned_build_of_the_dense_index(gpu_backend, oracle_backend, monkeypatch, payload, with_duplicate):
    """Large dense builds first split the rows into key-range buckets (hash_join.cu DenseBucketParams); forced here on a small
    build.  Payload rows of 1, 2 and 4 words, NULL keys and NULL payload values, a range that is not a multiple of the bucket
    width, and a duplicate key detected by the partitioning pass (→ falls back to the hash table, same result)."""
    monkeypatch.setenv("PG_JOIN_BUCKET_MIN_ROWS", "1")
    rng = np.random.default_rng(len(payload) * 2 + int(with_duplicate))
    nb = 70_001
    keys, nb = build_side(rng, "dense_unique", nb, BIGINT)
    keys = keys + 12345
    if with_duplicate:
        keys[nb // 2] = keys[7]
    knull = (rng.random(nb) < 0.02).astype(np.uint8)
    cols = {"date": (DATE, rng.integers(8000, 10500, nb).astype(np.int32), None),
            "int": (INTEGER, rng.integers(-5, 5, nb).astype(np.int32), None),
            "double_nullable": (DOUBLE, rng.random(nb), (rng.random(nb) < 0.1).astype(np.uint8)),
            "bigint": (BIGINT, rng.integers(-2**62, 2**62, nb), None),
            "nullable_int": (INTEGER, rng.integers(0, 1000, nb).astype(np.int32), (rng.random(nb) < 0.3).astype(np.uint8))}
    chosen = {"date": ["date"], "date_int": ["date", "int"], "double_nullable_date": ["double_nullable", "date"],
              "bigint_double_nullable_int": ["bigint", "double_nullable", "nullable_int"]}[payload]
    types = [BIGINT] + [cols[c][0] for c in chosen]
    cuts = [0, nb // 3, nb]
    build_pages = [Page([Block.from_numpy(BIGINT, keys[a:b], knull[a:b])] + [Block.from_numpy(cols[c][0], cols[c][1][a:b], None if cols[c][2] is None else cols[c][2][a:b]) for c in chosen])
                   for a, b in zip(cuts[:-1], cuts[1:])]
    npr = 120_000
    pk = np.where(rng.random(npr) < 0.6, keys[rng.integers(0, nb, npr)], rng.integers(0, int(keys.max()) + 1000, npr))
    probe_pages = [Page([Block.from_numpy(BIGINT, pk), Block.from_numpy(DOUBLE, rng.random(npr))])]
    args = (build_pages, types, probe_pages, [BIGINT, DOUBLE], JOIN_INNER, list(range(1, len(types))), [0, 1])
    got = join(gpu_backend, *args)
    want = join(oracle_backend, *args)
    assert len(got) > 0 and got == want


@pytest.mark.parametrize("kind", ["dense_unique", "sparse_unique", "duplicates"])
@pytest.mark.parametrize("full", [False, True], ids=["lookup_outer", "full_outer"])
def test_build_outer_joins_match_oracle(gpu_backend, oracle_backend, kind, full):
    """RIGHT / FULL OUTER: two probe operators over one outer bridge, then the LookupOuterOperator's pages — bit-exact vs the oracle,
    output order included (unvisited build positions ascending).  `dense_unique` keys would pick the bitmap index for an inner join;
    an outer bridge keeps the position-ordered hash path."""
    from presto_b200 import JOIN_FULL_OUTER, JOIN_LOOKUP_OUTER, LookupOuterOperator
    rng = np.random.default_rng(sum(map(ord, kind)) + int(full))
    keys, nb = build_side(rng, kind, 40_000, BIGINT)
    knull = (rng.random(nb) < 0.05).astype(np.uint8)
    pay = rng.random(nb); pnull = (rng.random(nb) < 0.1).astype(np.uint8)
    cut = nb // 3
    build_pages = [Page([Block.from_numpy(BIGINT, keys[s], knull[s]), Block.from_numpy(DOUBLE, pay[s], pnull[s])]) for s in (slice(0, cut), slice(cut, nb))]
    probe_pages = []
    for _ in range(2):
        npr = 30_000
        pk = np.where(rng.random(npr) < 0.5, keys[rng.integers(0, nb // 2, npr)], rng.integers(-1000, 1 << 41, npr))      # half of the build side is never probed
        probe_pages.append(Page([Block.from_numpy(BIGINT, pk, (rng.random(npr) < 0.05).astype(np.uint8)), Block.from_numpy(INTEGER, rng.integers(0, 100, npr).astype(np.int32))]))

    def run(backend):
        bridge = JoinBridge(backend=backend, outer=True)
        builder = HashBuilderOperator(bridge, [BIGINT, DOUBLE], [0], [0, 1], backend=backend)
        assert to_pages(builder, build_pages) == []
        jt = JOIN_FULL_OUTER if full else JOIN_LOOKUP_OUTER
        probes = [LookupJoinOperator(bridge, [BIGINT, INTEGER], [0], [1, 0], join_type=jt, backend=backend) for _ in range(2)]
        outer = LookupOuterOperator(bridge, [INTEGER, BIGINT], backend=backend, max_output_rows=7000)
        bridge.no_more_probe_operators()
        rows = [rows_of(to_pages(op, [p])) for op, p in zip(probes, probe_pages)]
        assert not outer.is_blocked()
        pages = []
        while not outer.is_finished():
            p = outer.get_output()
            if p is not None:
                pages.append(p)
        rows.append(rows_of(pages))
        for op in probes + [outer, builder]:
            op.close()
        bridge.close()
        return rows
    got, want = run(gpu_backend), run(oracle_backend)
    assert got == want
    assert len(want[2]) > nb // 8


@pytest.mark.parametrize("payload", ["date", "date_int", "nullable_int"])
@pytest.mark.parametrize("with_duplicate", [False, True])
def test_sorted_build_of_the_dense_index(gpu_backend, oracle_backend, monkeypatch, payload, with_duplicate):
    """Sparse unique keys in arrival order with a one-word payload (TPC-H Q3's filtered orders) are built by sorting (key - min, payload
    word) pairs (hash_join.cu densePackPairsKernel); forced here on a small build.  NULL keys, NULL payload values, a duplicate key found
    as two equal neighbours of the sorted keys (→ falls back to the hash table, same result)."""
    monkeypatch.setenv("PG_JOIN_SORTED_MIN_ROWS", "1")
    rng = np.random.default_rng(len(payload) * 2 + int(with_duplicate))
    nb = 50_003
    keys = (rng.choice(1 << 21, nb, replace=False).astype(np.int64) * 37) - 999        # ~1 key per 1500 values: far from the range-partitioned build
    if with_duplicate:
        keys[nb // 2] = keys[7]
    knull = (rng.random(nb) < 0.02).astype(np.uint8)
    cols = {"date": [(DATE, rng.integers(8000, 10500, nb).astype(np.int32), None)],
            "date_int": [(DATE, rng.integers(8000, 10500, nb).astype(np.int32), None), (INTEGER, rng.integers(-5, 5, nb).astype(np.int32), None)],
            "nullable_int": [(INTEGER, rng.integers(0, 1000, nb).astype(np.int32), (rng.random(nb) < 0.3).astype(np.uint8))]}[payload]
    types = [BIGINT] + [c[0] for c in cols]
    cuts = [0, nb // 3, nb]
    build_pages = [Page([Block.from_numpy(BIGINT, keys[a:b], knull[a:b])] + [Block.from_numpy(t, v[a:b], None if m is None else m[a:b]) for t, v, m in cols]) for a, b in zip(cuts[:-1], cuts[1:])]
    npr = 100_000
    pk = np.where(rng.random(npr) < 0.6, keys[rng.integers(0, nb, npr)], rng.integers(int(keys.min()) - 50, int(keys.max()) + 50, npr))
    probe_pages = [Page([Block.from_numpy(BIGINT, pk), Block.from_numpy(DOUBLE, rng.random(npr))])]
    for jt in (JOIN_INNER, JOIN_PROBE_OUTER):
        args = (build_pages, types, probe_pages, [BIGINT, DOUBLE], jt, list(range(1, len(types))), [0, 1])
        got, want = join(gpu_backend, *args), join(oracle_backend, *args)
        assert len(got) > 0 and got == want


def _random_strings(rng, n, null_rate=0.1, max_len=12):
    from presto_b200 import Block
    lens = rng.integers(0, max_len + 1, n)
    return Block.from_strings([None if rng.random() < null_rate else bytes(rng.integers(97, 123, int(k)).astype(np.uint8)) for k in lens])


@pytest.mark.parametrize("kind", ["dense_unique", "sparse_unique", "duplicates"])
@pytest.mark.parametrize("join_type", [JOIN_INNER, JOIN_PROBE_OUTER])
@pytest.mark.parametrize("build_varchar", [False, True], ids=["probe_varchar_only", "varchar_on_both_sides"])
def test_varchar_channels_through_the_join(gpu_backend, oracle_backend, kind, join_type, build_varchar):
    """VARCHAR output channels: probe-side ones ride on either LookupSource (bitmap index for dense_unique, hash table otherwise); a VARCHAR
    build payload keeps the position-ordered hash path.  NULL / empty strings, NULL keys, duplicate chains, several pages."""
    from presto_b200 import VARCHAR
    rng = np.random.default_rng(sum(map(ord, kind)) + join_type * 2 + int(build_varchar))
    keys, nb = build_side(rng, kind, 20_000, BIGINT)
    cut = nb // 2
    build_pages = []
    for s in (slice(0, cut), slice(cut, nb)):
        m = len(keys[s])
        blocks = [Block.from_numpy(BIGINT, keys[s], (rng.random(m) < 0.03).astype(np.uint8)), Block.from_numpy(DATE, rng.integers(8000, 9000, m).astype(np.int32))]
        if build_varchar:
            blocks.append(_random_strings(rng, m))
        build_pages.append(Page(blocks))
    build_types = [BIGINT, DATE] + ([VARCHAR] if build_varchar else [])
    probe_pages = []
    for _ in range(2):
        npr = 15_000
        pk = np.where(rng.random(npr) < 0.6, keys[rng.integers(0, nb, npr)], rng.integers(-1000, 1 << 41, npr))
        probe_pages.append(Page([_random_strings(rng, npr), Block.from_numpy(BIGINT, pk, (rng.random(npr) < 0.05).astype(np.uint8)), _random_strings(rng, npr, null_rate=0.0, max_len=3)]))

    def run(backend):
        bridge = JoinBridge(backend=backend)
        builder = HashBuilderOperator(bridge, build_types, [0], list(range(1, len(build_types))), backend=backend)
        assert to_pages(builder, build_pages) == []
        probe = LookupJoinOperator(bridge, [VARCHAR, BIGINT, VARCHAR], [1], [2, 1, 0], join_type=join_type, backend=backend)
        rows = rows_of(to_pages(probe, probe_pages))
        probe.close(); builder.close(); bridge.close()
        return rows
    got, want = run(gpu_backend), run(oracle_backend)
    assert len(want) > 5000 and got == want


# ---- VARCHAR join keys (the reference's TestHashJoinOperator keys on VARCHAR; GPU: slices interned to value ids at build time) ---------
@pytest.mark.parametrize("shape", ["unique", "duplicates", "two_keys", "full_outer"])
@pytest.mark.parametrize("nulls", [False, True])
def test_varchar_join_keys_match_oracle(gpu_backend, oracle_backend, shape, nulls):
    from presto_b200 import JOIN_FULL_OUTER, VARCHAR, LookupOuterOperator
    rng = np.random.default_rng(len(shape) * 2 + int(nulls))
    nb = 20_000
    vocabulary = [b"customer#%09d-%s" % (i, b"z" * (i % 19)) for i in range(nb)]
    vocabulary[0] = b""                                                         # the empty slice is a value, not a NULL
    pick = rng.permutation(nb) if shape in ("unique", "full_outer") else rng.integers(0, nb // 4, nb)
    bkeys = [vocabulary[i] for i in pick]
    if nulls:
        bkeys = [None if rng.random() < 0.05 else k for k in bkeys]
    second = rng.integers(0, 3, nb).astype(np.int32)
    pay = rng.random(nb)
    half = nb // 2
    build_pages = [Page([Block.from_strings(bkeys[s]), Block.from_numpy(INTEGER, second[s]), Block.from_numpy(DOUBLE, pay[s])]) for s in (slice(0, half), slice(half, nb))]
    probe_pages = []
    for _ in range(2):
        npr = 30_000
        pk = [vocabulary[i] if rng.random() < 0.7 else b"no-such-customer-%d" % i for i in rng.integers(0, nb, npr)]
        if nulls:
            pk = [None if rng.random() < 0.05 else k for k in pk]
        probe_pages.append(Page([Block.from_strings(pk), Block.from_numpy(INTEGER, rng.integers(0, 3, npr).astype(np.int32)), Block.from_numpy(BIGINT, rng.integers(0, 1 << 40, npr))]))
    key_channels = [0, 1] if shape == "two_keys" else [0]

    def run(backend):
        bridge = JoinBridge(backend=backend, outer=shape == "full_outer")
        builder = HashBuilderOperator(bridge, [VARCHAR, INTEGER, DOUBLE], key_channels, [0, 2], backend=backend)
        assert to_pages(builder, build_pages) == []
        jt = JOIN_FULL_OUTER if shape == "full_outer" else (JOIN_PROBE_OUTER if nulls else JOIN_INNER)
        probe = LookupJoinOperator(bridge, [VARCHAR, INTEGER, BIGINT], key_channels, [2, 0], join_type=jt, backend=backend)
        rows = rows_of(to_pages(probe, probe_pages))
        if shape == "full_outer":
            bridge.no_more_probe_operators()
            probe.finish()
            outer = LookupOuterOperator(bridge, [BIGINT, VARCHAR], backend=backend)
            while not outer.is_finished():
                p = outer.get_output()
                if p is not None:
                    rows += rows_of([p])
            outer.close()
        probe.close(); builder.close(); bridge.close()
        return rows

    got, want = run(gpu_backend), run(oracle_backend)
    assert len(got) == len(want) and len(got) > 0
    assert got == want
