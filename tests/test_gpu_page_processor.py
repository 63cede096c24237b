"""-m gpu: PageProcessor on the CUDA backend in BOTH lowering modes — RowExpression → generated PTX (driver JIT) and the
interpreter kernels (PG_FP_INTERPRETER=1) — against the oracle: the reference cases plus seeded random expression DAGs with
NULLs.  IEEE double arithmetic is reproducible, so values are compared exactly (NaN == NaN)."""
import os

import numpy as np
import pytest

import reference_cases as rc
from presto_b200 import BIGINT, BOOLEAN, DATE, DOUBLE, INTEGER, Block, FilterAndProjectOperator, Page, Program, to_pages
from conftest import unverified_on_gpu
from util import rows_close, rows_of

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["ptx", "interpreter"])
def mode(request):
    old = os.environ.get("PG_FP_INTERPRETER")
    os.environ["PG_FP_INTERPRETER"] = "1" if request.param == "interpreter" else "0"
    yield request.param
    if old is None:
        os.environ.pop("PG_FP_INTERPRETER", None)
    else:
        os.environ["PG_FP_INTERPRETER"] = old


def test_reference_cases_both_modes(gpu_backend, mode):
    rc.case_page_processor_selection(gpu_backend)
    rc.case_page_processor_q1(gpu_backend, n=20_000)
    rc.case_page_processor_three_valued_logic(gpu_backend)
    rc.case_page_processor_hash_projection(gpu_backend)
    rc.case_page_processor_varchar_pass_through(gpu_backend)
    rc.case_page_processor_varchar_comparisons(gpu_backend, n=3000)


@pytest.mark.parametrize("literal", [b"", b"A", b"BUILDING", b"a-literal-of-thirty-one-bytes..", b"thirty-three bytes: beyond the PTX lowering's limit"[:33]])
def test_varchar_equals_literal_both_modes(gpu_backend, mode, literal):
    """`slice = literal` / `slice <> literal` (VarcharOperators.equal / notEqual) with NULL slices, slices that share a prefix with the literal
    or extend it, against python's own comparison.  With PG_FP_VARCHAR_JIT=1 the "ptx" mode lowers literals of up to 32 bytes to PTX byte
    compares against immediates (opt-in: not yet run on a GPU); otherwise both modes use the interpreter kernels."""
    from presto_b200 import VARCHAR
    from presto_b200 import _abi as abi
    rng = np.random.default_rng(len(literal))
    variants = [literal, literal + b"x", literal[:-1], literal[:-1] + b"~" if literal else b"~", b"", literal.lower() if literal else b"q", None]
    n = 9000
    col = [variants[i] for i in rng.integers(0, len(variants), n)]
    page = Page([Block.from_strings(col), Block.from_numpy(BIGINT, np.arange(n))])
    for negate in (False, True):
        prog = Program([VARCHAR, BIGINT])
        s = prog.input(0)
        prog.set_filter((s != literal) if negate else (s == literal))
        prog.project(prog.input(1), s.is_null())
        op = FilterAndProjectOperator([VARCHAR, BIGINT], prog, backend=gpu_backend)
        rows = rows_of(to_pages(op, [page]))
        paths = op.get_stats()["kernel_paths"]
        op.close()
        want = [(i, False) for i in range(n) if col[i] is not None and ((col[i] != literal) if negate else (col[i] == literal))]
        assert rows == want and len(rows) > 500
        # the PTX lowering of slice = literal is opt-in (PG_FP_VARCHAR_JIT=1: written after the round's last GPU run); by default both modes
        # evaluate VARCHAR predicates in the interpreter kernels
        if os.environ.get("PG_FP_VARCHAR_JIT") == "1":
            expect_jit = mode == "ptx" and len(literal) <= 32
            assert bool(paths & abi.PATH_FP_JIT) == expect_jit and bool(paths & abi.PATH_FP_INTERPRETER) == (not expect_jit), (paths, mode, literal)


@unverified_on_gpu
@pytest.mark.parametrize("literal", [b"", b"BUILDING", b"a-literal-of-thirty-one-bytes.."])
def test_varchar_equals_literal_ptx_lowering_on_the_device(gpu_backend, literal, monkeypatch):
    """The PTX lowering of `slice = literal` (opt-in: PG_FP_VARCHAR_JIT=1) through the driver JIT on the device; on the CPU the same PTX is
    executed by tests/ptx_interp.py against the oracle (tests/test_ptx_semantics.py)."""
    monkeypatch.setenv("PG_FP_VARCHAR_JIT", "1")
    monkeypatch.setenv("PG_FP_INTERPRETER", "0")
    test_varchar_equals_literal_both_modes(gpu_backend, "ptx", literal)


def random_program(rng, types):
    prog = Program(types)
    ins = [prog.input(i) for i in range(len(types))]
    nums = [e for e, t in zip(ins, types) if t in (BIGINT, INTEGER, DATE, DOUBLE)]
    bools = [e for e, t in zip(ins, types) if t == BOOLEAN]
    for _ in range(int(rng.integers(3, 9))):
        kind = rng.integers(0, 9)
        a, b = nums[rng.integers(len(nums))], nums[rng.integers(len(nums))]
        if kind <= 1:
            c = float(rng.integers(-5, 6)) if rng.random() < 0.5 else int(rng.integers(-5, 6))
            nums.append([a + b, a - b, a * b, a + c, c - a, a * c][rng.integers(6)])
        elif kind == 2:
            bools.append([a < b, a <= b, a > b, a >= b, a == b, a != b][rng.integers(6)])
        elif kind == 3 and len(bools) >= 2:
            p, q = bools[rng.integers(len(bools))], bools[rng.integers(len(bools))]
            bools.append([p & q, p | q, ~p][rng.integers(3)])
        elif kind == 4:
            bools.append(a.is_null())
        elif kind == 5:
            bools.append(a.between(int(rng.integers(-50, 0)), int(rng.integers(0, 50))))
        elif kind == 6:
            nums.append(a.cast_double() * 0.5)
        elif kind == 7:
            nums.append(a.hash().combine_hash(b.hash()))
        else:
            nums.append(-a)
    if bools and rng.random() < 0.8:
        prog.set_filter(bools[rng.integers(len(bools))])
    outs = [nums[i] for i in rng.choice(len(nums), size=min(4, len(nums)), replace=False)]
    outs += [bools[i] for i in rng.choice(len(bools), size=min(2, len(bools)), replace=False)] if bools else []
    prog.project(*outs)
    return prog


@pytest.mark.parametrize("seed", range(12))
def test_random_expressions(gpu_backend, oracle_backend, mode, seed):
    rng = np.random.default_rng(1000 + seed)
    n = 4000
    types = [BIGINT, INTEGER, DOUBLE, DOUBLE, DATE, BOOLEAN]
    def nulls(p): return (rng.random(n) < p).astype(np.uint8) if p else None
    blocks = [Block.from_numpy(BIGINT, rng.integers(-100, 100, n), nulls(0.1)), Block.from_numpy(INTEGER, rng.integers(-100, 100, n).astype(np.int32), nulls(0.0)),
              Block.from_numpy(DOUBLE, np.round(rng.normal(0, 50, n), 3), nulls(0.2)), Block.from_numpy(DOUBLE, rng.integers(-3, 4, n).astype(np.float64), nulls(0.0)),
              Block.from_numpy(DATE, rng.integers(8000, 8100, n).astype(np.int32), nulls(0.05)), Block.from_numpy(BOOLEAN, rng.integers(0, 2, n).astype(np.int8), nulls(0.1))]
    page = Page(blocks)
    prog = random_program(rng, types)
    from presto_b200 import PrestoException
    got_op = FilterAndProjectOperator(types, prog, backend=gpu_backend)
    try:
        want = rows_of(to_pages(FilterAndProjectOperator(types, prog, backend=oracle_backend), [page]))
    except PrestoException as e:      # a random DAG may genuinely overflow BIGINT: the GPU path must report the same error
        with pytest.raises(PrestoException) as g:
            to_pages(got_op, [page])
        assert g.value.code == e.code
        return
    got = rows_of(to_pages(got_op, [page]))
    rows_close(got, want, rel=0.0)
    got_op.close()


def test_overflow_and_division_by_zero_are_reported(gpu_backend, mode):
    from presto_b200 import PrestoException
    # x^5 overflows; x / 0; Long.MIN_VALUE / -1 (BigintOperators.divide :106-117: NUMERIC_VALUE_OUT_OF_RANGE)
    for build, code, vals in [(lambda x: x * x * x * x * x, 6, [3, 100000, 7]), (lambda x: x / (x - x), 1, [3, 100000, 7]),
                              (lambda x: x / -1, 6, [3, -(1 << 63), 7])]:
        prog = Program([BIGINT])
        prog.project(build(prog.input(0)))
        op = FilterAndProjectOperator([BIGINT], prog, backend=gpu_backend)
        with pytest.raises(PrestoException) as e:
            op.add_input(Page([Block.from_values(BIGINT, vals)]))
        assert e.value.code == code
        op.close()
    # NULL operands and filtered-out rows never raise
    prog = Program([BIGINT])
    x = prog.input(0)
    prog.set_filter(x < 10)
    prog.project(x * x * x * x * x)
    op = FilterAndProjectOperator([BIGINT], prog, backend=gpu_backend)
    rows = rows_of(to_pages(op, [Page([Block.from_values(BIGINT, [3, 100000, None, 2])])]))
    assert rows == [(243,), (32,)]
    op.close()
