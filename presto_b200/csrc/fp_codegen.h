// presto_b200 — RowExpression → PTX code generator for the fused filter / project kernels.
//
// The reference turns a PageProcessor's RowExpressions into JVM bytecode (MB/sql/gen/PageFunctionCompiler.java:195-861,
// ExpressionCompiler.compilePageProcessor :144-189).  The B200-native equivalent emits PTX for sm_100a — two kernels per
// (program, set of nullable inputs): `fp_filter` (selection bitmask + per-unit counts) and `fp_project` (projections of the
// selected rows written at scan offsets) — and hands it to the driver's JIT (cuModuleLoadDataEx; libcuda is dlopen'ed, there is
// no NVRTC in this image).  Straight-line code per row: no interpreter dispatch, no local-memory register file.
//
// Semantics are those of the interpreter in filter_project.cu (and of the oracle): IEEE double + - * / with explicit .rn (never
// contracted to FMA), exact BIGINT arithmetic with overflow / division-by-zero flags, SQL three-valued AND / OR / NOT.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/presto_gpu.h"

namespace pg {

struct FpProgramDesc {
    std::vector<pg_expr_node> nodes;
    std::vector<int> inputTypes;          // per input channel
    std::vector<bool> inputHasNulls;      // per input channel, for THIS page
    int filterNode = -1;
    std::vector<int> projections;
    std::vector<std::string> literals;    // per node: the bytes of a PG_EX_CONST_VARCHAR (empty vector: the program has none)
};

struct FpKernelSignature {
    // Parameter order of both kernels: n (u64), units (u64), then for every used input channel (ascending): values ptr [, offsets ptr if
    // the channel is VARCHAR] [, nulls ptr if it has nulls]; fp_filter adds: selection, counts.  fp_project adds: for every projection: values ptr [, nulls ptr if nullable];
    // then selection, offsets, error.
    std::vector<int> usedInputs;
    std::vector<bool> projectionNullable;
    bool ok = false;
    std::string reason;                   // why codegen declined (fallback to the interpreter)
};

// Returns the PTX text of both kernels ("fp_filter" only when the program has a filter) or an empty string when unsupported.
std::string generateFilterProjectPtx(const FpProgramDesc& prog, FpKernelSignature& sig);

constexpr int kJitUnitRows = 256;   // rows per warp unit (must match kRowsPerFpUnit)

}  // namespace pg
