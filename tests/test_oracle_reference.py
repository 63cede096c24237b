"""Pins the CPU oracle: the reference's behavioural tests (tests/reference_cases.py) + known-answer vectors for the hash
functions (SURVEY.md §8c "derived known-answer vectors", XXH64 cross-checked against the python xxhash package)."""
import ctypes as C
import struct

import numpy as np
import pytest

import reference_cases as rc
from presto_b200 import hashing


# ---- hash known-answer vectors ---------------------------------------------------------------
KAT = [
    # v, murmurHash3(v), AbstractLongType.hash(v), XXH64(8 LE bytes of v)
    (0, 0, 0, 3803688792395291579),
    (1, -5451962507482445012, 9155312661752487122, -6977822845260490347),
    (2, 4233148493373801447, -2783790655855066069, -1529185102714098256),
    (42, -9148929187392628276, 1279370338401826588, -5379971487550586029),
    (-1, 7256831767414464289, -6507640756101998425, -8804195676797548855),
    (123456789012345, -241788548253407999, -8770786616159739318, 5003376228732448344),
]


def test_hash_known_answers(oracle_backend):
    b = oracle_backend
    for v, murmur, type_hash, xx in KAT:
        assert b.hash_murmur3_fmix64(v) == murmur
        assert b.hash_bigint(v) == type_hash
        data = struct.pack("<q", v)
        assert b.hash_xxhash64(data, 8) == xx
    # HashGenerator.getPartition(P=8) on AbstractLongType.hash(v): v=0 → 0, v=1 → 4, v=42 → 2
    assert [b.hash_partition(b.hash_bigint(v), 8) for v in (0, 1, 42)] == [0, 4, 2]
    assert b.hash_xxhash64(b"A", 1) == 1371800463213966980
    assert b.hash_xxhash64(b"", 0) == -1205034819632174695


def test_xxhash64_matches_python_xxhash(oracle_backend):
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(3)
    for n in list(range(0, 70)) + [100, 255, 1000]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        expect = xxhash.xxh64_intdigest(data, seed=0)
        expect = expect - (1 << 64) if expect >= (1 << 63) else expect
        assert oracle_backend.hash_xxhash64(data, n) == expect


def test_numpy_hashing_matches_oracle(oracle_backend):
    rng = np.random.default_rng(4)
    v = rng.integers(-2**63, 2**63 - 1, 1000, dtype=np.int64)
    assert list(hashing.hash_long(v)) == [oracle_backend.hash_bigint(int(x)) for x in v]
    assert list(hashing.murmur_hash3(v)) == [oracle_backend.hash_murmur3_fmix64(int(x)) for x in v]
    for p in (1, 2, 3, 8, 100, 256):
        assert list(hashing.get_partition(v, p)) == [oracle_backend.hash_partition(int(x), p) for x in v]


def test_array_size(oracle_backend):
    """fastutil HashCommon.arraySize(expected, 0.75f): next power of two of ceil(expected/0.75), min 2, > 2^30 throws."""
    import oracle
    lib = oracle.backend().lib
    out = C.c_int32()
    for expected, size in [(1, 2), (3, 4), (4, 8), (6, 8), (7, 16), (100, 256), (10_000, 16384), (150_000_000, 1 << 28)]:
        assert lib.or_array_size(expected, C.byref(out)) == 0 and out.value == size
    assert lib.or_array_size(900_000_000, C.byref(out)) == 5


# ---- the reference's behavioural tests against the oracle -----------------------------------------------------------
def test_group_by_hash_add_page(oracle_backend): rc.case_group_by_hash_add_page(oracle_backend)
def test_group_by_hash_null_group(oracle_backend): rc.case_group_by_hash_null_group(oracle_backend)
def test_group_by_hash_get_group_ids(oracle_backend): rc.case_group_by_hash_get_group_ids(oracle_backend, tries_per_value=10)
def test_group_by_hash_append_to(oracle_backend): rc.case_group_by_hash_append_to(oracle_backend)
def test_group_by_hash_append_to_multiple(oracle_backend): rc.case_group_by_hash_append_to_multiple_tuples_per_group(oracle_backend)
def test_group_by_hash_contains(oracle_backend): rc.case_group_by_hash_contains(oracle_backend)
def test_group_by_hash_contains_multiple_columns(oracle_backend): rc.case_group_by_hash_contains_multiple_columns(oracle_backend)
def test_group_by_hash_force_rehash(oracle_backend): rc.case_group_by_hash_force_rehash(oracle_backend)


@pytest.mark.parametrize("hash_enabled", [False, True])
def test_hash_aggregation(oracle_backend, hash_enabled): rc.case_hash_aggregation(oracle_backend, hash_enabled)
def test_hash_aggregation_null_inputs(oracle_backend): rc.case_hash_aggregation_null_inputs(oracle_backend)
def test_hash_aggregation_mask(oracle_backend): rc.case_hash_aggregation_mask(oracle_backend)
def test_hash_aggregation_multi_key(oracle_backend): rc.case_hash_aggregation_multi_key(oracle_backend)
def test_hash_aggregation_empty(oracle_backend): rc.case_hash_aggregation_empty(oracle_backend)
def test_hash_aggregation_add_after_finish(oracle_backend): rc.case_hash_aggregation_add_after_finish(oracle_backend)
def test_bigint_sum_overflow(oracle_backend): rc.case_bigint_sum_overflow(oracle_backend)


HASH_FLAGS = [(False, False), (True, False), (False, True), (True, True)]


@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_inner_join(oracle_backend, probe_hash, build_hash): rc.case_inner_join(oracle_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_inner_join_with_null_probe(oracle_backend, probe_hash, build_hash): rc.case_inner_join_with_null_probe(oracle_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_inner_join_with_null_build(oracle_backend, probe_hash, build_hash): rc.case_inner_join_with_null_build(oracle_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_inner_join_with_null_on_both_sides(oracle_backend, probe_hash, build_hash): rc.case_inner_join_with_null_on_both_sides(oracle_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_probe_outer_join(oracle_backend, probe_hash, build_hash): rc.case_probe_outer_join(oracle_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_outer_join_with_null_on_both_sides(oracle_backend, probe_hash, build_hash): rc.case_outer_join_with_null_on_both_sides(oracle_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_probe_outer_join_with_filter_function(oracle_backend, probe_hash, build_hash): rc.case_probe_outer_join_with_filter_function(oracle_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_outer_join_with_nulls_and_filter_function(oracle_backend, probe_hash, build_hash): rc.case_outer_join_with_nulls_and_filter_function(oracle_backend, probe_hash, build_hash)
def test_join_filter_over_both_sides(oracle_backend): rc.case_join_filter_over_both_sides(oracle_backend)
def test_join_varchar_outputs(oracle_backend): rc.case_join_varchar_outputs(oracle_backend)
# TestHashJoinOperator with the reference's own key type: VARCHAR join channels (GPU: interned to value ids when the build finishes)
@pytest.mark.parametrize("probe_hash,build_hash", [(False, False), (True, True)])
@pytest.mark.parametrize("case", ["inner_join", "inner_join_with_null_probe", "inner_join_with_null_build", "inner_join_with_null_on_both_sides",
                                  "probe_outer_join", "outer_join_with_null_on_both_sides", "outer_join_with_nulls_and_filter_function"])
def test_join_varchar_keys(oracle_backend, case, probe_hash, build_hash): getattr(rc, "case_" + case)(oracle_backend, probe_hash, build_hash, key_type=rc.VARCHAR)
@pytest.mark.parametrize("key_type", [rc.BIGINT, rc.VARCHAR])
@pytest.mark.parametrize("probe_hash,build_hash", [(False, False), (True, True)])
def test_outer_join_with_null_probe_and_null_build(oracle_backend, probe_hash, build_hash, key_type):
    rc.case_outer_join_with_null_probe(oracle_backend, probe_hash, build_hash, key_type)
    rc.case_outer_join_with_null_build(oracle_backend, probe_hash, build_hash, key_type)
    rc.case_inner_join_with_non_empty_lookup_source_and_empty_probe(oracle_backend, key_type)
def test_probe_outer_join_empty_lookup_source(oracle_backend): rc.case_probe_outer_join_empty_lookup_source(oracle_backend)
def test_join_duplicate_chain_order(oracle_backend): rc.case_join_duplicate_chain_order(oracle_backend)
def test_lookup_outer_join_empty_lookup_source(oracle_backend): rc.case_lookup_outer_join_empty_lookup_source(oracle_backend)
def test_lookup_outer_operator(oracle_backend): rc.case_lookup_outer_operator(oracle_backend)


@pytest.mark.parametrize("partition_count,nulls,precomputed", [(1, False, False), (4, False, False), (7, True, False), (8, True, True), (256, False, True)])
def test_partitioned_output(oracle_backend, partition_count, nulls, precomputed):
    rc.case_partitioned_output(oracle_backend, partition_count, 1000, nulls, precomputed)
def test_partitioned_output_replicates(oracle_backend): rc.case_partitioned_output_replicates(oracle_backend)
@pytest.mark.parametrize("replicate", [False, True])
def test_partitioned_output_varchar_channels(oracle_backend, replicate): rc.case_partitioned_output_varchar_channels(oracle_backend, 5, replicate)
def test_min_max_aggregates(oracle_backend): rc.case_min_max_aggregates(oracle_backend)
def test_decimal_sum(oracle_backend): rc.case_decimal_sum(oracle_backend)
def test_decimal_sum_overflow(oracle_backend): rc.case_decimal_sum_overflow(oracle_backend)
def test_lookup_join_is_blocked_until_the_build_finishes(oracle_backend): rc.case_lookup_join_is_blocked_until_the_build_finishes(oracle_backend)
def test_partitioned_output_varchar_partition_channels(oracle_backend): rc.case_partitioned_output_varchar_partition_channels(oracle_backend)
def test_partitioned_output_bucket_table_and_round_robin(oracle_backend): rc.case_partitioned_output_bucket_table_and_round_robin(oracle_backend)
def test_partitioned_output_double_zero_keys(oracle_backend): rc.case_partitioned_output_double_zero_keys(oracle_backend)
@pytest.mark.parametrize("partition_count", [1, 2, 8, 64])
def test_partitioned_output_local_partition_generator(oracle_backend, partition_count):
    rc.case_partitioned_output(oracle_backend, partition_count, 1000, True, False, local=True)
    rc.case_partitioned_output(oracle_backend, partition_count, 1000, False, True, local=True)


def test_page_processor_selection(oracle_backend): rc.case_page_processor_selection(oracle_backend)
def test_page_processor_q1(oracle_backend): rc.case_page_processor_q1(oracle_backend)
def test_page_processor_three_valued_logic(oracle_backend): rc.case_page_processor_three_valued_logic(oracle_backend)
def test_page_processor_hash_projection(oracle_backend): rc.case_page_processor_hash_projection(oracle_backend)
def test_page_processor_varchar_comparisons(oracle_backend): rc.case_page_processor_varchar_comparisons(oracle_backend)
def test_varchar_operators(oracle_backend): rc.case_varchar_operators(oracle_backend)
def test_page_processor_varchar_pass_through(oracle_backend): rc.case_page_processor_varchar_pass_through(oracle_backend)


@pytest.mark.parametrize("seed", range(8))
def test_page_processor_vectorised_equals_row_at_a_time(oracle_backend, seed, monkeypatch):
    """The oracle evaluates expressions node by node over a page (evalNodes); OR_ROW_AT_A_TIME=1 switches to the row-at-a-time
    statement of the same semantics (evalRow).  Random DAGs with nulls: identical rows, or the same error."""
    from presto_b200 import BIGINT, BOOLEAN, DATE, DOUBLE, INTEGER, Block, FilterAndProjectOperator, Page, PrestoException, to_pages
    from test_gpu_page_processor import random_program
    from util import rows_of
    rng = np.random.default_rng(7000 + seed)
    n = 3000
    types = [BIGINT, INTEGER, DOUBLE, DOUBLE, DATE, BOOLEAN]
    def nulls(p): return (rng.random(n) < p).astype(np.uint8) if p else None
    page = Page([Block.from_numpy(BIGINT, rng.integers(-100, 100, n), nulls(0.1)), Block.from_numpy(INTEGER, rng.integers(-100, 100, n).astype(np.int32), nulls(0.0)),
                 Block.from_numpy(DOUBLE, np.round(rng.normal(0, 50, n), 3), nulls(0.2)), Block.from_numpy(DOUBLE, rng.integers(-3, 4, n).astype(np.float64), nulls(0.0)),
                 Block.from_numpy(DATE, rng.integers(8000, 8100, n).astype(np.int32), nulls(0.05)), Block.from_numpy(BOOLEAN, rng.integers(0, 2, n).astype(np.int8), nulls(0.1))])
    prog = random_program(rng, types)
    def run():
        try:
            return rows_of(to_pages(FilterAndProjectOperator(types, prog, backend=oracle_backend), [page]))
        except PrestoException as e:
            return ("error", e.code)
    vectorised = run()
    monkeypatch.setenv("OR_ROW_AT_A_TIME", "1")
    assert repr(run()) == repr(vectorised)       # repr: NaN results compare equal to themselves
