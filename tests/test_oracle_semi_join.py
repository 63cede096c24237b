"""CPU: the oracle's HashSemiJoinOperator against the expectations of the reference's TestHashSemiJoinOperator."""
import pytest

from semi_join_cases import REFERENCE_CASES, run_semi_join


@pytest.mark.parametrize("name", sorted(REFERENCE_CASES))
def test_reference_case(oracle_backend, name):
    set_values, probe, expected = REFERENCE_CASES[name]
    assert run_semi_join(oracle_backend, set_values, probe) == expected


def test_empty_set_answers_false_even_for_null(oracle_backend):
    assert run_semi_join(oracle_backend, [], [(1,), (None,)]) == [(1, False), (None, False)]      # HashSemiJoinOperator.java:161-164


@pytest.mark.parametrize("name", sorted(REFERENCE_CASES))
def test_reference_case_varchar_channel(oracle_backend, name):
    """The reference's cases over a VARCHAR join channel (TestHashSemiJoinOperator is typed BIGINT; HashSemiJoinOperator itself is generic)."""
    from semi_join_cases import varchar_key
    set_values, probe, expected = REFERENCE_CASES[name]
    assert run_semi_join(oracle_backend, set_values, probe, varchar=True) == [(varchar_key(r[0]),) + tuple(r[1:]) for r in expected]
