// presto_b200 — ownership of VARCHAR literals inside an expression program (PG_EX_CONST_VARCHAR: imm_i = address of the bytes in the
// caller's memory, reserved = length).  Whoever keeps a program beyond the call that handed it over (fused operators that build their
// FilterAndProjectOperator later, operator factories) adopts the literals: the bytes are copied and imm_i repointed at the copy.
#pragma once
#include "../../include/presto_gpu.h"
#include <cstdint>
#include <memory>
#include <vector>

namespace pg {

struct ExprLiterals {
    std::vector<std::unique_ptr<uint8_t[]>> blobs;
    void adopt(pg_expr_node* nodes, size_t n)
    {
        for (size_t i = 0; i < n; i++) {
            if (nodes[i].op != PG_EX_CONST_VARCHAR || nodes[i].reserved <= 0 || nodes[i].imm_i == 0) continue;
            std::unique_ptr<uint8_t[]> copy(new uint8_t[(size_t) nodes[i].reserved]);
            const uint8_t* src = (const uint8_t*) (uintptr_t) nodes[i].imm_i;
            for (int32_t b = 0; b < nodes[i].reserved; b++) copy[b] = src[b];
            nodes[i].imm_i = (int64_t) (uintptr_t) copy.get();
            blobs.push_back(std::move(copy));
        }
    }
    void adopt(std::vector<pg_expr_node>& nodes) { adopt(nodes.data(), nodes.size()); }
};

// What a VARCHAR-typed operand may feed on the GPU path: a comparison (all operands VARCHAR), BETWEEN, IS NULL, the type hash
// (VarcharOperators.java:54-105,291).  VARCHAR-valued nodes are input references and literals.  Returns nullptr when the program is fine.
inline const char* varcharUseProblem(const pg_expr_node* nodes, int n)
{
    for (int i = 0; i < n; i++) {
        const pg_expr_node& nd = nodes[i];
        if (nd.type == PG_TYPE_VARCHAR && nd.op != PG_EX_INPUT && nd.op != PG_EX_CONST_VARCHAR) return "VARCHAR-valued expressions other than input references and literals are not on the GPU path";
        if (nd.op == PG_EX_CONST_VARCHAR) { if (nd.reserved < 0 || (nd.reserved > 0 && nd.imm_i == 0)) return "CONST_VARCHAR: imm_i = address of the bytes, reserved = length"; continue; }
        if (nd.op == PG_EX_INPUT || nd.op == PG_EX_CONST_BIGINT || nd.op == PG_EX_CONST_DOUBLE || nd.op == PG_EX_CONST_BOOLEAN || nd.op == PG_EX_CONST_NULL) continue;
        int nVarchar = 0, nOperands = 0;
        for (int o : {nd.a, nd.b, nd.c}) if (o >= 0 && o < i) { nOperands++; nVarchar += nodes[o].type == PG_TYPE_VARCHAR; }
        if (!nVarchar) continue;
        const bool comparison = (nd.op >= PG_EX_EQ && nd.op <= PG_EX_GE) || nd.op == PG_EX_BETWEEN;
        if (!((comparison && nVarchar == nOperands) || nd.op == PG_EX_IS_NULL || nd.op == PG_EX_HASH))
            return "VARCHAR operands: only comparisons among VARCHAR values, BETWEEN, IS NULL and the type hash are on the GPU path";
    }
    return nullptr;
}

}  // namespace pg
