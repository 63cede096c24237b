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

}  // namespace pg
