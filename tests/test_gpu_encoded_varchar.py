"""-m gpu: DictionaryBlock / RunLengthEncodedBlock over VARCHAR (how Presto hands low-cardinality strings such as TPC-H
l_returnflag / l_linestatus to an operator) are flattened at ingest on the device; results must equal the oracle's, which walks the
encodings position by position like Block.getSlice(position) does."""
import numpy as np
import pytest

from presto_b200 import AGG_COUNT_STAR, AGG_SUM_DOUBLE, DOUBLE, Aggregate, Block, HashAggregationOperator, Page, PagesSerde, to_pages
from util import rows_close, rows_of

pytestmark = pytest.mark.gpu


def q1_like_pages(rng, n_pages, n):
    pages = []
    for _ in range(n_pages):
        flags = Block.from_strings([b"A", b"N", b"R", None])
        status = Block.from_strings([b"F", b"O"])
        pages.append(Page([Block.dictionary(flags, rng.integers(0, 4, n).astype(np.int32)),
                           Block.dictionary(status, rng.integers(0, 2, n).astype(np.int32)),
                           Block.from_numpy(DOUBLE, rng.random(n) * 100)], n))
    return pages


def aggregate(backend, pages):
    op = HashAggregationOperator([6, 6], [0, 1], [Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)], group_max_lengths=[1, 1],
                                 expected_groups=16, backend=backend)
    return rows_of(to_pages(op, pages))


def test_group_by_on_dictionary_encoded_varchar_keys(gpu_backend, oracle_backend):
    pages = q1_like_pages(np.random.default_rng(2), 3, 50_000)
    got, want = aggregate(gpu_backend, pages), aggregate(oracle_backend, pages)
    assert len(want) == 8                      # {A, N, R, NULL} x {F, O}, in first-appearance order
    rows_close(got, want)


def test_rle_varchar_key(gpu_backend, oracle_backend):
    rng = np.random.default_rng(4)
    pages = [Page([Block.rle(Block.from_strings([w]), 1000), Block.dictionary(Block.from_strings([b"F", b"O"]), rng.integers(0, 2, 1000).astype(np.int32)),
                   Block.from_numpy(DOUBLE, rng.random(1000))], 1000) for w in (b"A", None, b"R", b"A")]
    rows_close(aggregate(gpu_backend, pages), aggregate(oracle_backend, pages))


def test_serialized_page_of_encoded_varchar(gpu_backend, oracle_backend):
    page = q1_like_pages(np.random.default_rng(6), 1, 10_007)[0]
    g, o = PagesSerde(backend=gpu_backend), PagesSerde(backend=oracle_backend)
    from serde_cases import same_wire_up_to_dictionary_ids
    wire = g.serialize(page)
    assert same_wire_up_to_dictionary_ids(wire, o.serialize(page))      # DICTIONARY channels carry a per-writer instance id
    assert g.deserialize(wire, [6, 6, DOUBLE]).rows() == page.rows()
    g.close(); o.close()


# ---- VARCHAR keys of unbounded length arriving as DictionaryBlock / RunLengthEncodedBlock: the dictionary is interned, the rows only mapped
# (MultiChannelGroupByHash.java:543-552, 628-714 works per dictionary entry as well) ---------------------------------------------------------
def _long_key_pages(rng, n):
    from presto_b200 import BIGINT
    words = [b"customer-segment-%03d-%s" % (i, b"q" * (i % 17)) for i in range(40)]
    d1 = Block.from_strings(words[:25] + [None] + [b"never-referenced"])
    d2 = Block.from_strings(words[20:] + [b""])                                    # overlaps d1's values: same groups across pages
    flat = [None if rng.random() < 0.05 else words[i] for i in rng.integers(0, 40, n)]
    mk = lambda key: Page([key, Block.from_numpy(DOUBLE, rng.random(n) * 10), Block.from_numpy(BIGINT, rng.integers(0, 5, n))], n)
    return [mk(Block.dictionary(d1, rng.integers(0, 26, n).astype(np.int32))),      # ids 0..25: the NULL entry is used, the last entry never
            mk(Block.from_strings(flat)),
            mk(Block.dictionary(d2, rng.integers(0, 21, n).astype(np.int32))),
            mk(Block.rle(Block.from_strings([words[3]]), n)),
            mk(Block.rle(Block.from_strings([None]), n))]


def test_group_by_on_dictionary_encoded_long_varchar_keys(gpu_backend, oracle_backend):
    from presto_b200 import _abi as abi, AGG_COUNT_COLUMN, BIGINT, VARCHAR
    pages = _long_key_pages(np.random.default_rng(8), 20_000)

    def run(backend, aggs, keys=([VARCHAR], [0])):
        op = HashAggregationOperator(keys[0], keys[1], aggs, expected_groups=16, backend=backend, max_output_rows=1 << 30)
        rows = rows_of(to_pages(op, pages))
        stats = op.get_stats()
        op.close()
        return rows, stats

    aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)]
    (got, stats), (want, _) = run(gpu_backend, aggs), run(oracle_backend, aggs)
    assert len(want) == 42                                    # 40 words + NULL + the empty string
    rows_close(got, want)
    assert stats["kernel_paths"] & abi.PATH_DICTIONARY_KEYS and stats["kernel_paths"] & abi.PATH_INTERNED_KEYS, stats
    # count(key column) reads the channel's per-row null map: the channel is flattened at ingest instead (no dictionary path), same rows
    aggs = [Aggregate(AGG_COUNT_COLUMN, 0), Aggregate(AGG_SUM_DOUBLE, 1)]
    (got, stats), (want, _) = run(gpu_backend, aggs), run(oracle_backend, aggs)
    rows_close(got, want)
    assert not (stats["kernel_paths"] & abi.PATH_DICTIONARY_KEYS), stats
    # two keys: the encoded VARCHAR and a BIGINT channel
    (got, _), (want, _) = run(gpu_backend, [Aggregate(AGG_COUNT_STAR)], ([VARCHAR, BIGINT], [0, 2])), run(oracle_backend, [Aggregate(AGG_COUNT_STAR)], ([VARCHAR, BIGINT], [0, 2]))
    assert len(want) > 150
    rows_close(got, want)


def test_group_by_hash_ids_of_dictionary_encoded_long_varchar_keys(gpu_backend, oracle_backend):
    from presto_b200 import VARCHAR, GroupByHash
    pages = _long_key_pages(np.random.default_rng(10), 5_000)

    def run(backend):
        h = GroupByHash([VARCHAR], [0], expected_size=8, backend=backend)
        ids = [list(h.get_group_ids(Page([p.blocks[0]], p.position_count))) for p in pages]
        contains = [list(h.contains(Page([p.blocks[0]], p.position_count))) for p in pages[:3]]
        absent = list(h.contains(Page([Block.dictionary(Block.from_strings([b"not-a-key", b"customer-segment-003-qqq", None]), np.array([0, 1, 2, 1, 0], dtype=np.int32))], 5)))
        keys = h.get_keys().rows()
        h.close()
        return ids, contains, absent, keys

    got, want = run(gpu_backend), run(oracle_backend)
    assert got == want
    assert want[2] == [False, True, True, True, False]
