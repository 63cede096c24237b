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
