"""-m gpu: the TPC-H Q1 and Q3 operator pipelines (presto_b200/pipelines.py) on the CUDA backend with device-resident pages chained
between operators, against the same plan run on the oracle.  Keys, counts and row ORDER bit-exact; double sums within 1e-6."""
import ctypes as C

import numpy as np
import pytest

from presto_b200 import BIGINT, DATE, DOUBLE, INTEGER, Block, Page
from presto_b200.pipelines import run_q1, run_q3
from presto_b200.synthetic import q3_tables_np
from util import rows_close, rows_of

pytestmark = pytest.mark.gpu


def to_device(backend, page: Page, keep: list) -> Page:
    blocks = []
    for b in page.blocks:
        p = C.c_void_p()
        backend.check(backend.device_alloc(max(b.values.nbytes, 16), C.byref(p)))
        backend.check(backend.memcpy_h2d(p, b.values.ctypes.data, b.values.nbytes))
        keep.append(p)
        blocks.append(Block.device(b.type, b.position_count, p.value))
    return Page(blocks, page.position_count)


def pages_of(types, arrays, n_pages=3):
    n = len(arrays[0])
    cuts = np.linspace(0, n, n_pages + 1).astype(int)
    return [Page([Block.from_numpy(t, a[cuts[i]:cuts[i + 1]]) for t, a in zip(types, arrays)]) for i in range(n_pages)]


@pytest.mark.parametrize("device_resident", [False, True])
def test_q3_pipeline(gpu_backend, oracle_backend, device_resident):
    customer, orders, lineitem = q3_tables_np(40_000)
    cp = pages_of([BIGINT, INTEGER], customer, 1)
    op = pages_of([BIGINT, BIGINT, DATE, INTEGER], orders, 2)
    lp = pages_of([BIGINT, DOUBLE, DOUBLE, DATE], lineitem, 3)
    want = rows_of(run_q3(cp, op, lp, backend=oracle_backend))
    keep = []
    if device_resident:
        cp, op, lp = [[to_device(gpu_backend, p, keep) for p in ps] for ps in (cp, op, lp)]
    stats = {}
    got = rows_of(run_q3(cp, op, lp, backend=gpu_backend, device_resident=device_resident, stats=stats))
    assert len(want) > 50 and stats["joined_rows"] >= len(want)
    rows_close(got, want)
    for p in keep:
        gpu_backend.device_free(p)


@pytest.mark.parametrize("device_resident", [False, True])
def test_q1_pipeline(gpu_backend, oracle_backend, device_resident):
    rng = np.random.default_rng(2)
    n = 150_000
    arrays = [rng.integers(0, 3, n).astype(np.int32), rng.integers(0, 2, n).astype(np.int32), rng.integers(1, 51, n).astype(np.float64),
              np.round(rng.random(n) * 1e5, 2), rng.integers(0, 11, n) / 100.0, rng.integers(0, 9, n) / 100.0, rng.integers(8035, 10600, n).astype(np.int32)]
    types = [INTEGER, INTEGER, DOUBLE, DOUBLE, DOUBLE, DOUBLE, DATE]
    pages = pages_of(types, arrays, 4)
    want = rows_of(run_q1(pages, backend=oracle_backend))
    keep = []
    if device_resident:
        pages = [to_device(gpu_backend, p, keep) for p in pages]
    got = rows_of(run_q1(pages, backend=gpu_backend, device_resident=device_resident))
    assert len(want) == 6
    rows_close(got, want)
    for p in keep:
        gpu_backend.device_free(p)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("device_resident", [False, True])
def test_q3_pipeline_with_varchar_mktsegment(gpu_backend, oracle_backend, fused, device_resident):
    """Q3 with the reference's customer types: c_mktsegment is a VARCHAR channel and the filter is `c_mktsegment = 'BUILDING'` (a Slice
    constant through VarcharOperators.equal), not a dictionary code.  Same groups as the INTEGER-coded plan, same rows as the oracle."""
    from presto_b200 import VARCHAR
    from presto_b200.synthetic import mktsegment_varchar_np
    from util import page_to_device
    customer, orders, lineitem = q3_tables_np(30_000)
    data, offsets = mktsegment_varchar_np(customer[1])
    cp = [Page([Block.from_numpy(BIGINT, customer[0]), Block(VARCHAR, len(customer[0]), data, None, offsets)])]
    coded = pages_of([BIGINT, INTEGER], customer, 1)
    op = pages_of([BIGINT, BIGINT, DATE, INTEGER], orders, 2)
    lp = pages_of([BIGINT, DOUBLE, DOUBLE, DATE], lineitem, 3)
    want = rows_of(run_q3(cp, op, lp, backend=oracle_backend))
    assert want == rows_of(run_q3(coded, op, lp, backend=oracle_backend)) and len(want) > 50
    keep = []
    if device_resident:
        cp, op, lp = [[page_to_device(gpu_backend, p, keep) for p in ps] for ps in (cp, op, lp)]
    got = rows_of(run_q3(cp, op, lp, backend=gpu_backend, device_resident=device_resident, fused=fused))
    rows_close(got, want)
    for p in keep:
        gpu_backend.device_free(p)
