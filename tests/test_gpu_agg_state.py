"""-m gpu: PARTIAL -> FINAL hand-off without pages (pg_hash_aggregation_export_state / merge_states): the state of source operators,
merged into a destination operator of the same configuration, must give exactly what ONE operator gives over the destination's pages
followed by the sources' pages in order (the oracle's SINGLE step is the checker): keys, counts and group order bit-exact, doubles 1e-6."""
import ctypes as C

import numpy as np
import pytest

from presto_b200 import (
    AGG_AVG_BIGINT, AGG_AVG_DOUBLE, AGG_COUNT_COLUMN, AGG_COUNT_STAR, AGG_MAX_BIGINT, AGG_MAX_DOUBLE, AGG_MIN_BIGINT, AGG_MIN_DOUBLE,
    AGG_SUM_BIGINT, AGG_SUM_DOUBLE, BIGINT, DATE, DOUBLE, INTEGER, STEP_PARTIAL, VARCHAR, Aggregate, Block, FusedAggregationOperator,
    HashAggregationOperator, Page, PrestoException, to_pages,
)
from test_gpu_hash_agg_random import double_key_pages, key_bits, make_pages
from util import rows_close, rows_of

pytestmark = pytest.mark.gpu


class DeviceBytes:
    def __init__(self, backend, nbytes):
        self.b, self.p, self.nbytes = backend, C.c_void_p(), nbytes
        backend.check(backend.device_alloc(nbytes, C.byref(self.p)))
        zeros = np.zeros(nbytes, dtype=np.uint8)
        backend.check(backend.memcpy_h2d(self.p, zeros.ctypes.data, nbytes))
    @property
    def ptr(self): return self.p.value
    def close(self): self.b.device_free(self.p)


def merged(backend, make_op, own_pages, source_pages, max_groups, *, first_slot_empty=False):
    """Destination aggregates `own_pages`; every entry of `source_pages` is aggregated by its own operator and handed over as a state."""
    dst = make_op(backend)
    for p in own_pages:
        dst.add_input(p)
    stride = (dst.state_bytes(max_groups) + 255) // 256 * 256
    n_slots = len(source_pages) + (1 if first_slot_empty else 0)
    buf = DeviceBytes(backend, stride * n_slots)
    sources = []
    for j, pages in enumerate(source_pages):
        src = make_op(backend, step=STEP_PARTIAL)
        for p in pages:
            src.add_input(p)
        src.export_state(buf.ptr + (j + (1 if first_slot_empty else 0)) * stride, stride, max_groups)
        sources.append(src)
    backend.check(backend.device_synchronize())      # the operators own their streams in this test
    dst.merge_states(buf.ptr, n_slots, stride, max_groups, order_base=1)
    dst.finish()
    rows = []
    while True:
        out = dst.get_output()
        if out is None:
            break
        rows += rows_of([out])
    for s in sources:
        s.close()
    dst.close()
    buf.close()
    return rows


def single(backend, make_op, pages):
    op = make_op(backend)
    rows = rows_of(to_pages(op, pages))
    op.close()
    return rows


@pytest.mark.parametrize("groups,variant", [(4, "plain"), (300, "null_keys"), (300, "null_vals"), (5000, "sentinel"), (200_000, "plain")])
def test_bigint_key_states(gpu_backend, oracle_backend, groups, variant):
    rng = np.random.default_rng(groups + len(variant))
    n = 50_000 if groups < 100_000 else 300_000
    pages = make_pages(rng, 6, n, groups, n_vals=3, null_keys=variant == "null_keys", null_vals=variant == "null_vals", sentinel=variant == "sentinel")
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_AVG_DOUBLE, 2), Aggregate(AGG_COUNT_STAR), Aggregate(AGG_COUNT_COLUMN, 3), Aggregate(AGG_MIN_DOUBLE, 1),
            Aggregate(AGG_MAX_DOUBLE, 3)]
    def make_op(backend, step=None, **kw):
        kw = dict(expected_groups=64, max_output_rows=1 << 30, backend=backend, **kw)
        if step is not None:
            kw["step"] = step
        return HashAggregationOperator([BIGINT], [0], aggs, **kw)
    got = merged(gpu_backend, make_op, pages[:2], [pages[2:3], pages[3:5], pages[5:]], max_groups=1 << 18, first_slot_empty=True)
    want = single(oracle_backend, make_op, pages)
    rows_close(got, want)


def test_lazy_counts_differ_between_source_and_destination(gpu_backend, oracle_backend):
    """Only one source sees NULL inputs: its non-null counts are active, everybody else's are still implied by the row count."""
    rng = np.random.default_rng(12)
    clean = make_pages(rng, 2, 20_000, 50, n_vals=2)
    nulls = make_pages(rng, 1, 20_000, 80, n_vals=2, null_vals=True)
    aggs = [Aggregate(AGG_AVG_DOUBLE, 1), Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_COUNT_COLUMN, 2)]
    def make_op(backend, step=None, **kw):
        kw = dict(expected_groups=64, max_output_rows=1 << 30, backend=backend, **kw)
        if step is not None:
            kw["step"] = step
        return HashAggregationOperator([BIGINT], [0], aggs, **kw)
    for own, sources in ((clean[:1], [nulls, clean[1:]]), (nulls, [clean[:1], clean[1:]])):
        got = merged(gpu_backend, make_op, own, sources, max_groups=1024)
        want = single(oracle_backend, make_op, own + [p for s in sources for p in s])
        rows_close(got, want)


def test_multi_word_keys_and_integer_aggregates(gpu_backend, oracle_backend):
    rng = np.random.default_rng(31)
    n, groups = 60_000, 3000
    pages = []
    for _ in range(4):
        g = rng.integers(0, groups, n)
        pages.append(Page([Block.from_numpy(BIGINT, g * 1_000_003, (rng.random(n) < 0.02).astype(np.uint8)), Block.from_numpy(DATE, (g % 2500 + 8000).astype(np.int32)),
                           Block.from_numpy(INTEGER, (g % 3).astype(np.int32), (rng.random(n) < 0.02).astype(np.uint8)), Block.from_numpy(DOUBLE, rng.random(n) * 100),
                           Block.from_numpy(BIGINT, rng.integers(-10**9, 10**9, n))]))
    aggs = [Aggregate(AGG_SUM_DOUBLE, 3), Aggregate(AGG_SUM_BIGINT, 4), Aggregate(AGG_AVG_BIGINT, 4), Aggregate(AGG_COUNT_COLUMN, 0), Aggregate(AGG_COUNT_STAR),
            Aggregate(AGG_MIN_BIGINT, 4), Aggregate(AGG_MAX_BIGINT, 4)]
    def make_op(backend, step=None, **kw):
        kw = dict(expected_groups=128, max_output_rows=1 << 30, backend=backend, **kw)
        if step is not None:
            kw["step"] = step
        return HashAggregationOperator([BIGINT, DATE, INTEGER], [0, 1, 2], aggs, **kw)
    got = merged(gpu_backend, make_op, pages[:1], [pages[1:2], pages[2:]], max_groups=8192)
    want = single(oracle_backend, make_op, pages)
    rows_close(got, want)


def test_double_keys_keep_the_first_sources_zero_spelling(gpu_backend, oracle_backend):
    rng = np.random.default_rng(2024)
    pages = double_key_pages(rng)
    aggs = [Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)]
    def make_op(backend, step=None, **kw):
        kw = dict(expected_groups=64, max_output_rows=1 << 30, backend=backend, **kw)
        if step is not None:
            kw["step"] = step
        return HashAggregationOperator([DOUBLE, DOUBLE], [0, 1], aggs, **kw)
    got = merged(gpu_backend, make_op, [], [pages[:1], pages[1:2], pages[2:]], max_groups=256)
    want = single(oracle_backend, make_op, pages)
    assert key_bits(got, 2) == key_bits(want, 2)
    rows_close(got, want)


def test_fused_q1_states(gpu_backend, oracle_backend):
    """The Q1 fragment (filter/project -> aggregation, VARCHAR(1) keys) on three 'GPUs': the destination's own shard + two states."""
    from presto_b200.pipelines import Q1_AGGREGATES, q1_program_typed
    rng = np.random.default_rng(5)
    def shard(n):
        rf = [b"ANR"[i:i + 1] for i in rng.integers(0, 3, n)]; ls = [b"FO"[i:i + 1] for i in rng.integers(0, 2, n)]
        return Page([Block.from_strings(rf), Block.from_strings(ls), Block.from_numpy(DOUBLE, rng.integers(1, 51, n).astype(np.float64)),
                     Block.from_numpy(DOUBLE, np.round(rng.random(n) * 1e5, 2)), Block.from_numpy(DOUBLE, rng.integers(0, 11, n) / 100),
                     Block.from_numpy(DOUBLE, rng.integers(0, 9, n) / 100), Block.from_numpy(DATE, rng.integers(8035, 10592, n).astype(np.int32))])
    pages = [shard(70_000), shard(50_000), shard(90_000)]
    types = [VARCHAR, VARCHAR, DOUBLE, DOUBLE, DOUBLE, DOUBLE, DATE]
    def make_op(backend, step=None, **kw):
        kw = dict(expected_groups=1024, group_max_lengths=[1, 1], max_output_rows=1 << 30, backend=backend, **kw)
        if step is not None:
            kw["step"] = step
        return FusedAggregationOperator(types, q1_program_typed(VARCHAR), [], [VARCHAR, VARCHAR], [0, 1], Q1_AGGREGATES, **kw)
    got = merged(gpu_backend, make_op, pages[:1], [pages[1:2], pages[2:]], max_groups=64)
    want = single(oracle_backend, make_op, pages)
    assert len(got) == 6
    rows_close(got, want)


def test_state_errors(gpu_backend):
    rng = np.random.default_rng(1)
    pages = make_pages(rng, 1, 10_000, 500, n_vals=1)
    aggs = [Aggregate(AGG_SUM_DOUBLE, 1)]
    src = HashAggregationOperator([BIGINT], [0], aggs, backend=gpu_backend)
    src.add_input(pages[0])
    stride = src.state_bytes(100)
    buf = DeviceBytes(gpu_backend, stride)
    with pytest.raises(PrestoException):                      # buffer smaller than state_bytes(max_groups)
        src.export_state(buf.ptr, stride - 8, 100)
    src.export_state(buf.ptr, stride, 100)                    # 500 groups into a 100-group state: marked, rejected at the merge
    gpu_backend.check(gpu_backend.device_synchronize())
    dst = HashAggregationOperator([BIGINT], [0], aggs, backend=gpu_backend)
    with pytest.raises(PrestoException) as e:
        dst.merge_states(buf.ptr, 1, stride, 100)
    assert "groups" in str(e.value)
    other = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)], backend=gpu_backend)
    big = DeviceBytes(gpu_backend, src.state_bytes(1000))
    src.export_state(big.ptr, big.nbytes, 1000)
    gpu_backend.check(gpu_backend.device_synchronize())
    with pytest.raises(PrestoException):                      # a differently planned operator
        other.merge_states(big.ptr, 1, other.state_bytes(1000), 1000)
    dst.merge_states(big.ptr, 1, big.nbytes, 1000)
    assert len(rows_of(to_pages(dst, []))) == 500
    for o in (src, dst, other):
        o.close()
    buf.close(); big.close()
