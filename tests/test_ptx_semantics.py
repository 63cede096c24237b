"""CPU: the PTX that fp_codegen.cpp generates is EXECUTED by a small PTX interpreter (tests/ptx_interp.py, one warp in lock step) and its
results are compared with the oracle's PageProcessor on the same pages — the generated code's semantics, not only that ptxas assembles it
(tests/test_fp_codegen.py).  This is how the lowering of `slice = literal` (written after the round's last GPU run, opt-in on the device:
PG_FP_VARCHAR_JIT=1) is checked: NULLs, prefixes, longer and shorter slices, the empty literal."""
import numpy as np
import pytest

from presto_b200 import BIGINT, BOOLEAN, DATE, DOUBLE, INTEGER, VARCHAR, Block, FilterAndProjectOperator, Page, Program, to_pages
from ptx_interp import Arena, Kernel
from util import rows_of

WIDTH = {BIGINT: 8, DOUBLE: 8, INTEGER: 4, DATE: 4, BOOLEAN: 1}


def run_generated_ptx(types, program, page: Page):
    """fp_filter → exclusive scan of the unit counts (the operator does it with CUB) → fp_project, as FilterProjectOp::addInput drives them."""
    n = page.position_count
    nullable = tuple(ch for ch, b in enumerate(page.blocks) if b.is_null is not None)
    ptx = FilterAndProjectOperator.explain(types, program, nullable_inputs=nullable)
    arena = Arena()
    units = (n + 255) // 256
    inputs = {}
    for ch, b in enumerate(page.blocks):
        inputs[f"p_in{ch}_v"] = arena.put(b.values)
        if b.type == VARCHAR:
            inputs[f"p_in{ch}_o"] = arena.put(b.offsets)
        if b.is_null is not None:
            inputs[f"p_in{ch}_n"] = arena.put(b.is_null)
    sel = arena.alloc(((n + 31) // 32 + 1) * 4)
    cnt = arena.alloc((units + 1) * 4)
    err = arena.alloc(16)
    common = {"p_n": n, "p_units": units}
    total, off = n, 0
    opcodes = set()
    if program.filter >= 0:
        k = Kernel(ptx, "fp_filter")
        opcodes |= k.opcodes
        k.run(arena, dict({p: inputs[p] for p in k.params if p in inputs}, p_sel=sel, p_cnt=cnt, p_err=err, **common))
        counts = arena.read(cnt, units * 4, np.uint32).copy()
        offsets = np.zeros(units + 1, dtype=np.uint32)
        np.cumsum(counts, out=offsets[1:])
        total = int(offsets[-1])
        off = arena.put(offsets)
    k = Kernel(ptx, "fp_project")
    opcodes |= k.opcodes
    outs = {}
    proj_types = [program.nodes[p][1] for p in program.projections if program.nodes[p][1] != VARCHAR]
    for j, t in enumerate(proj_types):
        outs[f"p_out{j}_v"] = arena.alloc(max(total, 1) * WIDTH[t])
        if f"p_out{j}_n" in k.params:
            outs[f"p_out{j}_n"] = arena.alloc(max(total, 1))
    k.run(arena, dict({p: inputs[p] for p in k.params if p in inputs}, p_sel=sel, p_off=off, p_err=err, **outs, **common))
    columns = []
    for j, t in enumerate(proj_types):
        raw = arena.read(outs[f"p_out{j}_v"], total * WIDTH[t], {BIGINT: np.int64, DOUBLE: np.float64, INTEGER: np.int32, DATE: np.int32, BOOLEAN: np.uint8}[t]).copy()
        nulls = arena.read(outs[f"p_out{j}_n"], total, np.uint8).copy() if f"p_out{j}_n" in outs else np.zeros(total, dtype=np.uint8)
        conv = (lambda v: bool(v)) if t == BOOLEAN else (lambda v: float(v)) if t == DOUBLE else (lambda v: int(v))
        columns.append([None if z else conv(v) for v, z in zip(raw, nulls)])
    return list(zip(*columns)) if columns else [], arena.load(err, 4), opcodes


def oracle_rows(oracle_backend, types, program, page):
    op = FilterAndProjectOperator(types, program, backend=oracle_backend)
    rows = rows_of(to_pages(op, [page]))
    op.close()
    return rows


def same(a, b):
    if len(a) != len(b):
        return False
    for r, s in zip(a, b):
        for x, y in zip(r, s):
            if isinstance(x, float) and isinstance(y, float):
                if not (x == y or (x != x and y != y)):
                    return False
            elif x != y or type(x) is not type(y):
                return False
    return True


def test_q1_and_q3_programs_execute_like_the_oracle(oracle_backend):
    """The existing lowerings (IEEE double arithmetic, DATE comparisons, NULL inputs) through the same harness: the interpreter agrees with
    what has been measured on the device for a year of tests, which is what makes its verdict on new lowerings worth something."""
    rng = np.random.default_rng(3)
    n = 700                     # two full 256-row units and a partial one
    types = [BIGINT, DOUBLE, DOUBLE, DATE]
    page = Page([Block.from_numpy(BIGINT, rng.integers(0, 1 << 40, n)), Block.from_numpy(DOUBLE, np.round(rng.random(n) * 1e5, 2), (rng.random(n) < 0.1).astype(np.uint8)),
                 Block.from_numpy(DOUBLE, rng.integers(0, 11, n) / 100.0), Block.from_numpy(DATE, rng.integers(8035, 10600, n).astype(np.int32))])
    prog = Program(types)
    k, e, d, s = [prog.input(i) for i in range(4)]
    prog.set_filter((s > 9204) & (d >= 0.02))
    prog.project(k, e * (1.0 - d), (e * (1.0 - d)) * (1.0 + d), s, e.is_null())
    got, err, _ = run_generated_ptx(types, prog, page)
    want = oracle_rows(oracle_backend, types, prog, page)
    assert err == 0 and len(want) > 100 and same(got, want)
    # integer arithmetic with overflow / division flags and three-valued logic, no filter
    types = [BIGINT, BIGINT, INTEGER]
    a = rng.integers(-50, 50, n); b = rng.integers(1, 9, n)
    page = Page([Block.from_numpy(BIGINT, a, (rng.random(n) < 0.1).astype(np.uint8)), Block.from_numpy(BIGINT, b), Block.from_numpy(INTEGER, rng.integers(-5, 5, n).astype(np.int32))])
    prog = Program(types)
    x, y, z = [prog.input(i) for i in range(3)]
    prog.project(x * y + z, x / y, (x < z) | (y > 4), ~(x == z) & (z >= 0), x.between(-10, 10), (x + y).hash().combine_hash(z.hash()), -x)
    got, err, _ = run_generated_ptx(types, prog, page)
    assert err == 0 and same(got, oracle_rows(oracle_backend, types, prog, page))


@pytest.mark.parametrize("literal", [b"", b"A", b"BUILDING", b"a-literal-of-thirty-one-bytes.."])
@pytest.mark.parametrize("with_nulls", [False, True])
def test_slice_equals_literal_lowering_executes_like_the_oracle(oracle_backend, literal, with_nulls):
    rng = np.random.default_rng(len(literal) + 7 * int(with_nulls))
    variants = [literal, literal + b"x", literal[:-1], (literal[:-1] + b"~") if literal else b"~", b"", literal.lower() if literal else b"q"]
    n = 600
    col = [variants[i] for i in rng.integers(0, len(variants), n)]
    if with_nulls:
        col = [None if rng.random() < 0.15 else c for c in col]
    types = [VARCHAR, BIGINT, DOUBLE]
    page = Page([Block.from_strings(col), Block.from_numpy(BIGINT, np.arange(n)), Block.from_numpy(DOUBLE, rng.random(n))])
    for build in (lambda p, s, k, v: (p.set_filter(s == literal), p.project(k, v * 2.0)),
                  lambda p, s, k, v: (p.set_filter((s != literal) & (v > 0.25)), p.project(k, s == literal, s.is_null())),
                  lambda p, s, k, v: (p.project(s == literal, s != literal, (s == literal) | (v > 0.5), k),)):
        prog = Program(types)
        build(prog, prog.input(0), prog.input(1), prog.input(2))
        got, err, opcodes = run_generated_ptx(types, prog, page)
        want = oracle_rows(oracle_backend, types, prog, page)
        assert err == 0 and len(want) > 20 and same(got, want), (literal, with_nulls)
        assert "ld.global.nc.u8" in opcodes or not literal          # the byte compares really ran through the lowering


@pytest.mark.parametrize("seed", range(10))
def test_random_expression_dags_execute_like_the_oracle(oracle_backend, seed):
    """The seeded random RowExpression DAGs of tests/test_gpu_page_processor.py (arithmetic, comparisons, three-valued logic, BETWEEN, casts,
    hashes, NULL inputs) through generated PTX on the interpreter vs the oracle — including the error word when a DAG overflows BIGINT."""
    from presto_b200 import PrestoException
    from test_gpu_page_processor import random_program
    rng = np.random.default_rng(1000 + seed)
    n = 600
    types = [BIGINT, INTEGER, DOUBLE, DOUBLE, DATE, BOOLEAN]
    def nulls(p): return (rng.random(n) < p).astype(np.uint8) if p else None
    blocks = [Block.from_numpy(BIGINT, rng.integers(-100, 100, n), nulls(0.1)), Block.from_numpy(INTEGER, rng.integers(-100, 100, n).astype(np.int32), nulls(0.0)),
              Block.from_numpy(DOUBLE, np.round(rng.normal(0, 50, n), 3), nulls(0.2)), Block.from_numpy(DOUBLE, rng.integers(-3, 4, n).astype(np.float64), nulls(0.0)),
              Block.from_numpy(DATE, rng.integers(8000, 8100, n).astype(np.int32), nulls(0.05)), Block.from_numpy(BOOLEAN, rng.integers(0, 2, n).astype(np.int8), nulls(0.1))]
    page = Page(blocks)
    prog = random_program(rng, types)
    got, err, _ = run_generated_ptx(types, prog, page)
    try:
        want = oracle_rows(oracle_backend, types, prog, page)
    except PrestoException as e:
        assert err != 0, "the oracle raised, the generated code did not flag it"
        assert (e.code == 6) == bool(err & 1) or (e.code == 1) == bool(err & 2)
        return
    assert err == 0 and same(got, want)
