"""CPU: the PTX the PageProcessor code generator emits (pg_filter_project_explain, no GPU needed) assembles for sm_100a with
ptxas, for the TPC-H Q1 / Q3 programs, nullable inputs, integer arithmetic with overflow checks and three-valued logic.  Results
are checked against the oracle in tests/test_gpu_page_processor.py on the GPU; this test keeps the generator honest on the CPU box."""
import shutil
import subprocess

import pytest

from presto_b200 import BIGINT, BOOLEAN, DATE, DOUBLE, INTEGER, FilterAndProjectOperator, Program

PTXAS = shutil.which("ptxas") or "/usr/local/cuda/bin/ptxas"


def q1_program():
    types = [INTEGER, INTEGER, DOUBLE, DOUBLE, DOUBLE, DOUBLE, DATE]
    p = Program(types)
    c = [p.input(i) for i in range(7)]
    p.set_filter(c[6] <= 10471)
    dp = c[3] * (1.0 - c[4])
    p.project(c[0], c[1], c[2], c[3], dp, dp * (1.0 + c[5]), c[4])
    return types, p


def q3_lineitem_program():
    types = [BIGINT, DOUBLE, DOUBLE, DATE]
    p = Program(types)
    k, ep, disc, ship = [p.input(i) for i in range(4)]
    p.set_filter(ship > 9204)
    p.project(k, ep * (1.0 - disc))
    return types, p


def integer_logic_program():
    types = [BIGINT, BIGINT, DOUBLE, BOOLEAN]
    p = Program(types)
    a, b, d, f = [p.input(i) for i in range(4)]
    p.set_filter(((a + b) > 10) | (f & (d < 0.5)))
    p.project(a * b - 7, a / b, (a < b) & f, ~f, d + a, -a)
    return types, p


def projection_only_program():
    types = [DOUBLE, DOUBLE]
    p = Program(types)
    x, y = p.input(0), p.input(1)
    p.project(x * y, x / y, x - y)
    return types, p


def q3_customer_program():
    """c_mktsegment = 'BUILDING' (VARCHAR slice = literal, lowered to byte compares against immediates) AND a numeric predicate; the key passes."""
    from presto_b200 import VARCHAR
    types = [BIGINT, VARCHAR, DOUBLE]
    p = Program(types)
    k, seg, bal = p.input(0), p.input(1), p.input(2)
    p.set_filter((seg == b"BUILDING") & (bal > 0.0) & (seg != b""))
    p.project(k, bal * 2.0)
    return types, p


CASES = [(q3_customer_program, ()), (q3_customer_program, (1,)), (q1_program, ()), (q1_program, (2, 6)), (q3_lineitem_program, ()), (q3_lineitem_program, (0, 1, 2, 3)),
         (integer_logic_program, ()), (integer_logic_program, (0, 1, 2, 3)), (projection_only_program, (1,))]


@pytest.mark.parametrize("make,nullable", CASES, ids=[f"{m.__name__}-nulls{''.join(map(str, n)) or 'none'}" for m, n in CASES])
def test_generated_ptx_assembles_for_sm_100a(make, nullable, tmp_path):
    types, program = make()
    ptx = FilterAndProjectOperator.explain(types, program, nullable_inputs=nullable)
    assert ".target sm_100a" in ptx and ".entry fp_project" in ptx
    assert (".entry fp_filter" in ptx) == (program.filter >= 0)
    # eight rounds of 32 rows (one 256-row unit) per loop iteration, each with its own registers: the loads of all of them are issued together
    assert "%row7" in ptx and "%row8" not in ptx
    src = tmp_path / "k.ptx"
    src.write_text(ptx)
    r = subprocess.run([PTXAS, "-arch=sm_100a", "-v", "-o", str(tmp_path / "k.cubin"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "bytes spill stores" not in r.stderr or " 0 bytes spill stores" in r.stderr
