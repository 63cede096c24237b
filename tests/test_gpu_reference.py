"""-m gpu: the reference's behavioural tests (tests/reference_cases.py) against the CUDA product through the C ABI."""
import pytest

from conftest import unverified_on_gpu
import reference_cases as rc

pytestmark = pytest.mark.gpu


def test_group_by_hash_add_page(gpu_backend): rc.case_group_by_hash_add_page(gpu_backend, max_group_id=120, tries_per_value=2)
def test_group_by_hash_null_group(gpu_backend): rc.case_group_by_hash_null_group(gpu_backend)
def test_group_by_hash_get_group_ids(gpu_backend): rc.case_group_by_hash_get_group_ids(gpu_backend, max_group_id=120, tries_per_value=2)
def test_group_by_hash_append_to(gpu_backend): rc.case_group_by_hash_append_to(gpu_backend)
def test_group_by_hash_append_to_multiple(gpu_backend): rc.case_group_by_hash_append_to_multiple_tuples_per_group(gpu_backend)
def test_group_by_hash_contains(gpu_backend): rc.case_group_by_hash_contains(gpu_backend)
def test_group_by_hash_contains_multiple_columns(gpu_backend): rc.case_group_by_hash_contains_multiple_columns(gpu_backend)
def test_group_by_hash_force_rehash(gpu_backend): rc.case_group_by_hash_force_rehash(gpu_backend)
# the same cases with the reference's unbounded VARCHAR type: keys interned to value ids (varchar_interner.cu)
def test_group_by_hash_append_to_unbounded_varchar(gpu_backend): rc.case_group_by_hash_append_to(gpu_backend, max_len=0)
def test_group_by_hash_contains_multiple_columns_unbounded_varchar(gpu_backend): rc.case_group_by_hash_contains_multiple_columns(gpu_backend, max_len=0)
def test_group_by_hash_force_rehash_unbounded_varchar(gpu_backend): rc.case_group_by_hash_force_rehash(gpu_backend, max_len=0)


@pytest.mark.parametrize("hash_enabled", [False, True])
def test_hash_aggregation(gpu_backend, hash_enabled): rc.case_hash_aggregation(gpu_backend, hash_enabled)
@pytest.mark.parametrize("hash_enabled", [False, True])
def test_hash_aggregation_unbounded_varchar_key(gpu_backend, hash_enabled): rc.case_hash_aggregation(gpu_backend, hash_enabled, max_len=0)
def test_hash_aggregation_null_inputs(gpu_backend): rc.case_hash_aggregation_null_inputs(gpu_backend)
def test_hash_aggregation_mask(gpu_backend): rc.case_hash_aggregation_mask(gpu_backend)
def test_hash_aggregation_multi_key(gpu_backend): rc.case_hash_aggregation_multi_key(gpu_backend)
def test_hash_aggregation_empty(gpu_backend): rc.case_hash_aggregation_empty(gpu_backend)
def test_hash_aggregation_add_after_finish(gpu_backend): rc.case_hash_aggregation_add_after_finish(gpu_backend)
def test_bigint_sum_overflow(gpu_backend): rc.case_bigint_sum_overflow(gpu_backend)


HASH_FLAGS = [(False, False), (True, True)]


@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_inner_join(gpu_backend, probe_hash, build_hash): rc.case_inner_join(gpu_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_inner_join_with_null_probe(gpu_backend, probe_hash, build_hash): rc.case_inner_join_with_null_probe(gpu_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_inner_join_with_null_build(gpu_backend, probe_hash, build_hash): rc.case_inner_join_with_null_build(gpu_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_inner_join_with_null_on_both_sides(gpu_backend, probe_hash, build_hash): rc.case_inner_join_with_null_on_both_sides(gpu_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_probe_outer_join(gpu_backend, probe_hash, build_hash): rc.case_probe_outer_join(gpu_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_outer_join_with_null_on_both_sides(gpu_backend, probe_hash, build_hash): rc.case_outer_join_with_null_on_both_sides(gpu_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_probe_outer_join_with_filter_function(gpu_backend, probe_hash, build_hash): rc.case_probe_outer_join_with_filter_function(gpu_backend, probe_hash, build_hash)
@pytest.mark.parametrize("probe_hash,build_hash", HASH_FLAGS)
def test_outer_join_with_nulls_and_filter_function(gpu_backend, probe_hash, build_hash): rc.case_outer_join_with_nulls_and_filter_function(gpu_backend, probe_hash, build_hash)
def test_join_filter_over_both_sides(gpu_backend): rc.case_join_filter_over_both_sides(gpu_backend)
def test_join_varchar_outputs(gpu_backend): rc.case_join_varchar_outputs(gpu_backend)
# TestHashJoinOperator with the reference's own key type: VARCHAR join channels (GPU: interned to value ids when the build finishes)
@pytest.mark.parametrize("probe_hash,build_hash", [(False, False), (True, True)])
@pytest.mark.parametrize("case", ["inner_join", "inner_join_with_null_probe", "inner_join_with_null_build", "inner_join_with_null_on_both_sides",
                                  "probe_outer_join", "outer_join_with_null_on_both_sides"])
def test_join_varchar_keys(gpu_backend, case, probe_hash, build_hash): getattr(rc, "case_" + case)(gpu_backend, probe_hash, build_hash, key_type=rc.VARCHAR)
# the join filter `probe.key = 'a'` with the reference's VARCHAR key and a Slice constant inside the JoinFilterFunction
@unverified_on_gpu
@pytest.mark.parametrize("probe_hash,build_hash", [(False, False), (True, True)])
def test_join_varchar_filter_function(gpu_backend, probe_hash, build_hash): rc.case_outer_join_with_nulls_and_filter_function(gpu_backend, probe_hash, build_hash, key_type=rc.VARCHAR)
@unverified_on_gpu
@pytest.mark.parametrize("key_type", [rc.BIGINT, rc.VARCHAR])
@pytest.mark.parametrize("probe_hash,build_hash", [(False, False), (True, True)])
def test_outer_join_with_null_probe_and_null_build(gpu_backend, probe_hash, build_hash, key_type):
    rc.case_outer_join_with_null_probe(gpu_backend, probe_hash, build_hash, key_type)
    rc.case_outer_join_with_null_build(gpu_backend, probe_hash, build_hash, key_type)
    rc.case_inner_join_with_non_empty_lookup_source_and_empty_probe(gpu_backend, key_type)
def test_probe_outer_join_empty_lookup_source(gpu_backend): rc.case_probe_outer_join_empty_lookup_source(gpu_backend)
def test_join_duplicate_chain_order(gpu_backend): rc.case_join_duplicate_chain_order(gpu_backend)
def test_lookup_outer_join_empty_lookup_source(gpu_backend): rc.case_lookup_outer_join_empty_lookup_source(gpu_backend)
def test_lookup_outer_operator(gpu_backend): rc.case_lookup_outer_operator(gpu_backend)


@pytest.mark.parametrize("partition_count,nulls,precomputed", [(1, False, False), (4, False, False), (7, True, False), (8, True, True), (256, False, True)])
def test_partitioned_output(gpu_backend, partition_count, nulls, precomputed):
    rc.case_partitioned_output(gpu_backend, partition_count, 1000, nulls, precomputed)
def test_partitioned_output_replicates(gpu_backend): rc.case_partitioned_output_replicates(gpu_backend)
@pytest.mark.parametrize("replicate", [False, True])
def test_partitioned_output_varchar_channels(gpu_backend, replicate): rc.case_partitioned_output_varchar_channels(gpu_backend, 5, replicate)
def test_min_max_aggregates(gpu_backend): rc.case_min_max_aggregates(gpu_backend)
def test_decimal_sum(gpu_backend): rc.case_decimal_sum(gpu_backend)
def test_decimal_sum_overflow(gpu_backend): rc.case_decimal_sum_overflow(gpu_backend)
def test_lookup_join_is_blocked_until_the_build_finishes(gpu_backend): rc.case_lookup_join_is_blocked_until_the_build_finishes(gpu_backend)
def test_partitioned_output_varchar_partition_channels(gpu_backend): rc.case_partitioned_output_varchar_partition_channels(gpu_backend)
def test_partitioned_output_bucket_table_and_round_robin(gpu_backend): rc.case_partitioned_output_bucket_table_and_round_robin(gpu_backend)
def test_partitioned_output_double_zero_keys(gpu_backend): rc.case_partitioned_output_double_zero_keys(gpu_backend)
@pytest.mark.parametrize("partition_count", [1, 2, 8, 64])
def test_partitioned_output_local_partition_generator(gpu_backend, partition_count):
    rc.case_partitioned_output(gpu_backend, partition_count, 1000, True, False, local=True)
    rc.case_partitioned_output(gpu_backend, partition_count, 1000, False, True, local=True)


def test_page_processor_selection(gpu_backend): rc.case_page_processor_selection(gpu_backend)
def test_page_processor_q1(gpu_backend): rc.case_page_processor_q1(gpu_backend)
def test_page_processor_three_valued_logic(gpu_backend): rc.case_page_processor_three_valued_logic(gpu_backend)
def test_page_processor_hash_projection(gpu_backend): rc.case_page_processor_hash_projection(gpu_backend)
def test_page_processor_varchar_comparisons(gpu_backend): rc.case_page_processor_varchar_comparisons(gpu_backend)
@unverified_on_gpu
def test_varchar_operators(gpu_backend): rc.case_varchar_operators(gpu_backend)
