"""-m gpu: HashSemiJoinOperator on the CUDA backend vs the oracle: the reference's own cases, both LookupSource structures (dense
index for unique keys, hash table with duplicate chains), NULLs on either side, the empty set, multi-page sets."""
import numpy as np
import pytest

from conftest import unverified_on_gpu
from semi_join_cases import REFERENCE_CASES, run_semi_join

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(REFERENCE_CASES))
def test_reference_case(gpu_backend, name):
    set_values, probe, expected = REFERENCE_CASES[name]
    assert run_semi_join(gpu_backend, set_values, probe) == expected


def test_empty_set(gpu_backend):
    assert run_semi_join(gpu_backend, [], [(1,), (None,)]) == [(1, False), (None, False)]


@pytest.mark.parametrize("kind", ["unique_dense", "duplicates", "sparse"])
@pytest.mark.parametrize("set_nulls", [False, True])
def test_random(gpu_backend, oracle_backend, kind, set_nulls):
    rng = np.random.default_rng(len(kind) + int(set_nulls))
    n = 20_000
    if kind == "unique_dense":
        values = list(rng.permutation(n) * 3)
    elif kind == "duplicates":
        values = list(rng.integers(0, n // 4, n))
    else:
        values = list(rng.choice(1 << 40, n, replace=False))
    values = [int(v) for v in values]
    if set_nulls:
        for i in rng.integers(0, n, 50):
            values[i] = None
    pool = [v for v in values if v is not None]
    probe = []
    for i in range(30_000):
        r = rng.random()
        key = None if r < 0.05 else (pool[rng.integers(0, len(pool))] if r < 0.6 else int(rng.integers(-100, 1 << 41)))
        probe.append((key, i, None if i % 7 == 0 else i * 2))
    got = run_semi_join(gpu_backend, values, probe, set_pages=3)
    want = run_semi_join(oracle_backend, values, probe, set_pages=3)
    assert got == want and any(r[-1] is True for r in got) and any(r[-1] is False or r[-1] is None for r in got)


# ---- VARCHAR join channel (the set's slices are interned when the SetBuilder finishes; the probe looks its slices up) -----------------
@unverified_on_gpu
@pytest.mark.parametrize("name", sorted(REFERENCE_CASES))
def test_reference_case_varchar_channel(gpu_backend, name):
    from semi_join_cases import varchar_key
    set_values, probe, expected = REFERENCE_CASES[name]
    assert run_semi_join(gpu_backend, set_values, probe, varchar=True) == [(varchar_key(r[0]),) + tuple(r[1:]) for r in expected]


@unverified_on_gpu
@pytest.mark.parametrize("set_nulls", [False, True])
def test_random_varchar_channel(gpu_backend, oracle_backend, set_nulls):
    rng = np.random.default_rng(40 + int(set_nulls))
    n = 10_000
    values = [int(v) for v in rng.integers(0, n // 3, n)]
    if set_nulls:
        for i in rng.integers(0, n, 30):
            values[i] = None
    probe = [(None if rng.random() < 0.05 else int(rng.integers(0, n // 2)), i, None if i % 5 == 0 else -i) for i in range(15_000)]
    got = run_semi_join(gpu_backend, values, probe, set_pages=2, varchar=True)
    want = run_semi_join(oracle_backend, values, probe, set_pages=2, varchar=True)
    assert got == want and any(r[-1] is True for r in got) and any(r[-1] is not True for r in got)
    assert run_semi_join(gpu_backend, [], [(1,), (None,)], varchar=True) == [(b"key-1", False), (None, False)]
