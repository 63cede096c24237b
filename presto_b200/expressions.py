"""RowExpression builder → flat register program for pg_filter_project_create.

The reference compiles RowExpressions to JVM bytecode (sql/gen/PageFunctionCompiler.java:195-861); the GPU path lowers the
same tree to a topologically ordered node list (one node = one virtual register) that the fused filter→project kernel
interprets with warp-uniform control flow.
"""
from __future__ import annotations

from typing import List

from . import _abi as abi


class Expr:
    def __init__(self, program: "Program", index: int, type_: int):
        self.program, self.index, self.type = program, index, type_

    def _bin(self, op: int, other, result_type=None) -> "Expr":
        other = self.program.lift(other, self.type)
        if result_type is None:
            result_type = abi.DOUBLE if abi.DOUBLE in (self.type, other.type) else abi.BIGINT
        return self.program.add(op, result_type, self.index, other.index)

    def __add__(self, o): return self._bin(abi.EX_ADD, o)
    def __sub__(self, o): return self._bin(abi.EX_SUB, o)
    def __mul__(self, o): return self._bin(abi.EX_MUL, o)
    def __truediv__(self, o): return self._bin(abi.EX_DIV, o)
    def __radd__(self, o): return self.program.lift(o, self.type)._bin(abi.EX_ADD, self)
    def __rsub__(self, o): return self.program.lift(o, self.type)._bin(abi.EX_SUB, self)
    def __rmul__(self, o): return self.program.lift(o, self.type)._bin(abi.EX_MUL, self)
    def __neg__(self): return self.program.add(abi.EX_NEG, self.type, self.index)
    def __eq__(self, o): return self._bin(abi.EX_EQ, o, abi.BOOLEAN)      # type: ignore[override]
    def __ne__(self, o): return self._bin(abi.EX_NE, o, abi.BOOLEAN)      # type: ignore[override]
    def __lt__(self, o): return self._bin(abi.EX_LT, o, abi.BOOLEAN)
    def __le__(self, o): return self._bin(abi.EX_LE, o, abi.BOOLEAN)
    def __gt__(self, o): return self._bin(abi.EX_GT, o, abi.BOOLEAN)
    def __ge__(self, o): return self._bin(abi.EX_GE, o, abi.BOOLEAN)
    def __and__(self, o): return self._bin(abi.EX_AND, o, abi.BOOLEAN)
    def __or__(self, o): return self._bin(abi.EX_OR, o, abi.BOOLEAN)
    def __invert__(self): return self.program.add(abi.EX_NOT, abi.BOOLEAN, self.index)
    __hash__ = None  # type: ignore[assignment]

    def is_null(self) -> "Expr": return self.program.add(abi.EX_IS_NULL, abi.BOOLEAN, self.index)
    def cast_double(self) -> "Expr": return self.program.add(abi.EX_CAST_DOUBLE, abi.DOUBLE, self.index)
    def between(self, lo, hi) -> "Expr":
        lo, hi = self.program.lift(lo, self.type), self.program.lift(hi, self.type)
        return self.program.add(abi.EX_BETWEEN, abi.BOOLEAN, self.index, lo.index, hi.index)
    def hash(self) -> "Expr": return self.program.add(abi.EX_HASH, abi.BIGINT, self.index)
    def combine_hash(self, other: "Expr") -> "Expr": return self.program.add(abi.EX_COMBINE_HASH, abi.BIGINT, self.index, other.index)


class Program:
    """A PageProcessor: optional filter + projections over shared sub-expressions."""

    def __init__(self, input_types: List[int]):
        self.input_types = list(input_types)
        self.nodes: List[tuple] = []
        self.filter = -1
        self.projections: List[int] = []
        self.literals = {}          # node index -> bytes of a VARCHAR literal (kept alive with the program)

    def add(self, op, type_, a=-1, b=-1, c=-1, imm_i=0, imm_d=0.0) -> Expr:
        self.nodes.append((op, type_, a, b, c, imm_i, imm_d))
        return Expr(self, len(self.nodes) - 1, type_)

    def input(self, channel: int) -> Expr:
        return self.add(abi.EX_INPUT, self.input_types[channel], channel)

    def const(self, value, type_=None) -> Expr:
        if value is None:
            return self.add(abi.EX_CONST_NULL, type_ or abi.BIGINT)
        if isinstance(value, (bytes, str)):
            # a Slice constant: the node carries the address and the length of its bytes (c_nodes), the operator copies them at create
            e = self.add(abi.EX_CONST_VARCHAR, abi.VARCHAR)
            self.literals[e.index] = value.encode() if isinstance(value, str) else bytes(value)
            return e
        if isinstance(value, bool):
            return self.add(abi.EX_CONST_BOOLEAN, abi.BOOLEAN, imm_i=int(value))
        if isinstance(value, float) or type_ == abi.DOUBLE:
            return self.add(abi.EX_CONST_DOUBLE, abi.DOUBLE, imm_d=float(value))
        return self.add(abi.EX_CONST_BIGINT, type_ or abi.BIGINT, imm_i=int(value))

    def lift(self, v, like_type) -> Expr:
        if isinstance(v, Expr):
            return v
        if isinstance(v, (bytes, str)):
            return self.const(v)
        if isinstance(v, float):
            return self.const(v, abi.DOUBLE)
        if like_type == abi.DOUBLE and not isinstance(v, bool):
            return self.const(float(v), abi.DOUBLE)
        return self.const(v, like_type if like_type in (abi.BIGINT, abi.INTEGER, abi.DATE) and not isinstance(v, bool) else None)

    def set_filter(self, e: Expr) -> None:
        self.filter = e.index

    def project(self, *exprs: Expr) -> None:
        self.projections.extend(e.index for e in exprs)

    def c_nodes(self):
        import ctypes as C
        arr = (abi.pg_expr_node * max(1, len(self.nodes)))()
        keep = []
        for i, (op, t, a, b, c, ii, dd) in enumerate(self.nodes):
            arr[i].op, arr[i].type, arr[i].a, arr[i].b, arr[i].c, arr[i].imm_i, arr[i].imm_d = op, t, a, b, c, ii, dd
            if op == abi.EX_CONST_VARCHAR:
                buf = C.create_string_buffer(self.literals[i], max(1, len(self.literals[i])))
                keep.append(buf)
                arr[i].imm_i, arr[i].reserved = C.addressof(buf), len(self.literals[i])
        arr._literal_buffers = keep      # the bytes live as long as the node array
        return arr
