// presto_b200 — what the operator translation units expose to each other inside libprestogpu.so (nothing here crosses the C ABI).
#pragma once
#include "common.cuh"
#include "dense_index.cuh"
#include "fused_codegen.h"

namespace pg {

struct JoinBridge;

Operator* createFilterProject(const pg_filter_project_config* c);
Operator* createLookupJoin(const pg_lookup_join_config* c, JoinBridge* bridge);

// The LookupSource behind a bridge as the fused pipeline kernels see it (hash_join.cu).
struct LookupSourceView {
    bool ready = false;            // the build finished (LookupSourceFactory future complete, LookupJoinOperator.java:127)
    bool dense = false;            // unique integer keys behind the bitmap + rank index (the only kind the fused kernels probe)
    int keyType = 0;
    int nKeys = 0;
    DenseIndex D{};
    const uint64_t* rows = nullptr;
    int strideWords = 0;
    std::vector<FusedBuildColumn> buildOutputs;
};
void describeLookupSource(JoinBridge* bridge, LookupSourceView& out);

}  // namespace pg
