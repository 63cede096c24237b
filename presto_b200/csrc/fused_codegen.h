// presto_b200 — pipeline fragment → CUDA C++ (compiled at run time by NVRTC, jit_runtime.h).
//
// The reference runs FilterAndProjectOperator → LookupJoinOperator* → HashAggregationOperator as separate operators of one Driver
// handing ≤ 1 MB pages to each other inside the CPU cache (MB/operator/Driver.java:402-536, LocalExecutionPlanner.java:1633-1666,
// 2650-2686, 3544-3640; scan+filter+project are already one operator there: ScanFilterAndProjectOperator).  On the GPU the
// equivalent of "stays in cache" is "stays in registers": the PageProcessor's RowExpression program (the thing
// PageFunctionCompiler.java:195-861 turns into bytecode), the probes of the dense join indexes and the aggregation's key packing are
// generated as ONE row program that fused_agg.cuh's kernel runs per tile — no intermediate page is ever written to HBM.
#pragma once
#include <string>
#include <vector>
#include "fp_codegen.h"

namespace pg {

struct FusedBuildColumn {          // one build OUTPUT column inside the dense index's payload row (hash_join.cu PayloadLayout)
    int type = 0, width = 0;       // pg_type, bytes
    int word = 0, shift = 0;       // 64-bit word of the payload row, bit offset
    int nullWord = -1, nullShift = 0;
};

struct FusedJoinDesc {             // inner LookupJoin against a dense unique-key index (LookupJoinOperator.java:481-616)
    int keyChannel = 0;            // probe key: channel of the layout entering this join
    std::vector<int> probeOutputChannels;   // channels of the entering layout that survive
    std::vector<FusedBuildColumn> buildOutputs;
    int strideWords = 0;           // payload row size in 64-bit words (0: no build output, membership only)
};

struct FusedKeyField { int channel = 0, type = 0, word = 0, shift = 0, bits = 0, maxLen = 0; };

struct FusedProgramDesc {
    FpProgramDesc fp;              // PageProcessor over the raw input page; its projections are layout 1
    std::vector<FusedJoinDesc> joins;
    // aggregation over the final layout (key_layout.h plan)
    std::vector<FusedKeyField> keys;
    bool bigintPath = false;
    int W = 1, nullWord = 0, nullShift = 0;
    struct Acc { int channel; bool fromBigint; };   // sum(double) / avg inputs: final-layout channel (BIGINT inputs are converted: avg(bigint))
    std::vector<Acc> accs;
    bool countRows = false;        // one plain row count column after the sums
    bool tma = false;              // eager columns staged by a cp.async.bulk ring instead of register double buffering
};

struct FusedKernelInfo {
    bool ok = false;
    std::string reason;            // why the fragment cannot be fused (the operator then runs the unfused chain, still on the GPU)
    std::vector<int> usedInputs;   // raw input channels the kernel reads
    std::vector<int> eagerInputs;  // subset loaded for every row ahead of time (the rest is loaded on demand after the filter / probe)
    int rowsPerThread = 2;
    int nv = 0;
    bool cnt = false;
    bool queue = false;            // survivors of the probes are queued per warp and aggregated 32 at a time (fragments with joins)
    size_t queueBytes = 0;         // shared memory of the queues (per CTA)
    bool tma = false;              // the translation unit also holds `fused_kernel_tma` (eager columns through a cp.async.bulk ring)
    int tmaStageBytes = 0;
    bool tiers = false;            // the on-chip tiers are available (one-word key, or a key whose fields squeeze into 64 bits)
};

// Returns the CUDA C++ translation unit (#include "fused_agg.cuh" + struct Prog + an explicit kernel instantiation named
// `fused_kernel`); empty when the fragment is not fusable (info.reason says why).
std::string generateFusedAggSource(const FusedProgramDesc& d, FusedKernelInfo& info);

// FilterAndProject → LookupJoin* materialised as a page (the aggregation fields of `d` are ignored): two kernels, `fused_select` and
// `fused_emit` (fused_agg.cuh, page sink).  info.outputTypes / outputNullable describe the channels of the output page.
struct FusedPageInfo {
    bool ok = false;
    std::string reason;
    std::vector<int> usedInputs, eagerInputs;
    int rowsPerThread = 2;
    std::vector<int> outputTypes;
    std::vector<bool> outputNullable;
};
std::string generateFusedPageSource(const FusedProgramDesc& d, FusedPageInfo& info);

}  // namespace pg
