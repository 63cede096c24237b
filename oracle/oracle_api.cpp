/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 * C API of liboracle.so: the same shapes as include/presto_gpu.h with the prefix or_ so the python test harness can drive
 * the oracle and the CUDA product through one marshalling layer.
 */
#include "oracle_common.h"

namespace oracle {
struct JoinBridge;
Operator* createHashAggregation(const pg_hash_aggregation_config* c);
JoinBridge* createJoinBridge();
void destroyJoinBridge(JoinBridge* b);
int64_t joinBridgePositionCount(JoinBridge* b);
Operator* createHashBuilder(const pg_hash_builder_config* c, JoinBridge* b);
Operator* createLookupJoin(const pg_lookup_join_config* c, JoinBridge* b);
Operator* createLookupOuter(const pg_lookup_outer_config* c, JoinBridge* b);
void joinBridgeSetOuter(JoinBridge* b);
void joinBridgeNoMoreProbes(JoinBridge* b);
Operator* createSemiJoin(const pg_semi_join_config* c, JoinBridge* b);
Operator* createPartitionedOutput(const pg_partitioned_output_config* c);
Operator* createFilterProject(const pg_filter_project_config* c);
void partitionedOutputOffsets(Operator* op, int64_t* out);
void filterProjectSelection(Operator* op, int32_t* isList, int64_t* size);
}

namespace oracle {
// The operators of one Driver chained the way Driver.processInternal does (MB/operator/Driver.java:402-536): a page is pushed through
// FilterAndProjectOperator -> LookupJoinOperator* -> HashAggregationOperator, each operator's output page being the next one's input.
// This is the reference semantics of include/presto_gpu.h's pg_fused_aggregation_create (which the product runs as one kernel).
struct DriverChain : Operator {
    std::unique_ptr<Operator> fp;
    std::vector<std::unique_ptr<Operator>> joins;
    std::unique_ptr<Operator> agg;
    OwnedPage a, b;
    bool finishing = false, hasPending = false;
    OwnedPage pending;
    explicit DriverChain(const pg_fused_aggregation_config* c)
    {
        fp.reset(createFilterProject(&c->page_processor));
        for (int j = 0; j < c->n_joins; j++) joins.emplace_back(createLookupJoin(&c->joins[j], (JoinBridge*) c->bridges[j]));
        agg.reset(createHashAggregation(&c->aggregation));
    }
    explicit DriverChain(const pg_fused_page_config* c)      // no aggregation: the chain's output is the last operator's page
    {
        fp.reset(createFilterProject(&c->page_processor));
        for (int j = 0; j < c->n_joins; j++) joins.emplace_back(createLookupJoin(&c->joins[j], (JoinBridge*) c->bridges[j]));
    }
    bool needsInput() override { return agg ? agg->needsInput() : (!finishing && !hasPending); }
    void addInput(const pg_page* page) override
    {
        stats.input_positions += page->position_count;
        fp->addInput(page);
        OwnedPage* cur = &a; OwnedPage* next = &b;
        if (!fp->getOutput(*cur) || cur->positionCount == 0) return;
        for (auto& j : joins) {
            pg_page in{}; cur->fill(&in);
            j->addInput(&in);
            if (!j->getOutput(*next) || next->positionCount == 0) return;
            std::swap(cur, next);
        }
        if (!agg) { pending.columns.swap(cur->columns); pending.positionCount = cur->positionCount; hasPending = true; return; }
        pg_page in{}; cur->fill(&in);
        agg->addInput(&in);
    }
    bool getOutput(OwnedPage& out) override
    {
        if (agg) return agg->getOutput(out);
        if (!hasPending) return false;
        out.columns.swap(pending.columns); out.positionCount = pending.positionCount; hasPending = false;
        return true;
    }
    void finish() override { if (agg) agg->finish(); else finishing = true; }
    bool isFinished() override { return agg ? agg->isFinished() : (finishing && !hasPending); }
};
}

using namespace oracle;

// The operators build every output page in fresh std::vectors.  With glibc's defaults each of those (> 128 KB) is its own mmap / page
// fault / munmap cycle, and 128 driver threads doing that serialise on the process's mm lock (measured: 0.3 M rows/s per thread instead
// of 3.7 M).  Keep freed blocks in the arenas instead — the JVM's heap does the same for the reference.
#include <malloc.h>
__attribute__((constructor)) static void oracleMallocSetup()
{
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_ARENA_MAX, 256);
}

static thread_local std::string g_lastError;

extern "C" {

void or_set_error(const char* msg) { g_lastError = msg; }
const char* or_last_error(void) { return g_lastError.c_str(); }
const char* or_version(void) { return "presto oracle (CPU restatement of reference Java operators)"; }

#define OR_TRY try {
#define OR_CATCH } catch (const OracleError& e) { or_set_error(e.what()); return e.code; } catch (const std::exception& e) { or_set_error(e.what()); return PG_ERR_INVALID_STATE; } return PG_OK;

int64_t or_hash_bigint(int64_t v) { return hashLong(v); }
int64_t or_hash_combine(int64_t p, int64_t h) { return combineHash(p, h); }
int64_t or_hash_murmur3_fmix64(int64_t x) { return murmurHash3(x); }
int32_t or_hash_partition(int64_t rawHash, int32_t partitionCount) { return hashPartition(rawHash, partitionCount); }
int64_t or_hash_xxhash64(const void* data, size_t len) { return xxHash64((const uint8_t*) data, len); }
int64_t or_hash_double(double d) { return doubleHashCode(d); }
int32_t or_hash_local_partition(int64_t rawHash, int32_t partitionCount)
{
    if (partitionCount <= 0 || (partitionCount & (partitionCount - 1)) != 0) return -1;
    return localPartition(rawHash, partitionCount);
}
int32_t or_array_size(int64_t expected, int32_t* out) { OR_TRY *out = arraySize(expected, 0.75f); OR_CATCH }

int32_t or_hash_aggregation_create(const pg_hash_aggregation_config* c, void** out) { OR_TRY *out = createHashAggregation(c); OR_CATCH }
int32_t or_join_bridge_create(int32_t, void** out) { OR_TRY *out = createJoinBridge(); OR_CATCH }
int32_t or_join_bridge_destroy(void* b) { destroyJoinBridge((JoinBridge*) b); return PG_OK; }
int32_t or_join_bridge_get_position_count(void* b, int64_t* out) { OR_TRY *out = joinBridgePositionCount((JoinBridge*) b); OR_CATCH }
int32_t or_hash_builder_create(const pg_hash_builder_config* c, void* bridge, void** out) { OR_TRY *out = createHashBuilder(c, (JoinBridge*) bridge); OR_CATCH }
int32_t or_join_bridge_set_outer(void* b) { OR_TRY joinBridgeSetOuter((JoinBridge*) b); OR_CATCH }
int32_t or_join_bridge_no_more_probe_operators(void* b) { OR_TRY joinBridgeNoMoreProbes((JoinBridge*) b); OR_CATCH }
int32_t or_lookup_outer_create(const pg_lookup_outer_config* c, void* bridge, void** out) { OR_TRY *out = createLookupOuter(c, (JoinBridge*) bridge); OR_CATCH }
int32_t or_lookup_join_create(const pg_lookup_join_config* c, void* bridge, void** out) { OR_TRY *out = createLookupJoin(c, (JoinBridge*) bridge); OR_CATCH }
int32_t or_hash_semi_join_create(const pg_semi_join_config* c, void* set, void** out) { OR_TRY *out = createSemiJoin(c, (JoinBridge*) set); OR_CATCH }
int32_t or_partitioned_output_create(const pg_partitioned_output_config* c, void** out) { OR_TRY *out = createPartitionedOutput(c); OR_CATCH }
int32_t or_partitioned_output_get_offsets(void* op, int64_t* out) { OR_TRY partitionedOutputOffsets((Operator*) op, out); OR_CATCH }
int32_t or_filter_project_create(const pg_filter_project_config* c, void** out) { OR_TRY *out = createFilterProject(c); OR_CATCH }
int32_t or_filter_project_get_last_selection(void* op, int32_t* isList, int64_t* size) { OR_TRY filterProjectSelection((Operator*) op, isList, size); OR_CATCH }

int32_t or_fused_aggregation_create(const pg_fused_aggregation_config* c, void** out) { OR_TRY *out = new DriverChain(c); OR_CATCH }
int32_t or_fused_page_create(const pg_fused_page_config* c, void** out) { OR_TRY *out = new DriverChain(c); OR_CATCH }

int32_t or_operator_needs_input(void* op, int32_t* out) { OR_TRY *out = ((Operator*) op)->needsInput() ? 1 : 0; OR_CATCH }
int32_t or_operator_is_blocked(void* op, int32_t* out) { OR_TRY *out = ((Operator*) op)->isBlocked() ? 1 : 0; OR_CATCH }
int32_t or_operator_add_input(void* op, const pg_page* page)
{
    OR_TRY
    if (page->memory != PG_MEM_HOST) throw OracleError(PG_ERR_INVALID_ARGUMENT, "the oracle only takes host pages");
    ((Operator*) op)->addInput(page);
    OR_CATCH
}
int32_t or_operator_get_output(void* opv, pg_page* out)
{
    OR_TRY
    auto* op = (Operator*) opv;
    if (op->getOutput(op->lastOutput)) op->lastOutput.fill(out);
    else { out->position_count = 0; out->channel_count = 0; out->memory = PG_MEM_HOST; out->blocks = nullptr; }
    OR_CATCH
}
int32_t or_operator_finish(void* op) { OR_TRY ((Operator*) op)->finish(); OR_CATCH }
int32_t or_operator_is_finished(void* op, int32_t* out) { OR_TRY *out = ((Operator*) op)->isFinished() ? 1 : 0; OR_CATCH }
int32_t or_operator_get_stats(void* op, pg_operator_stats* out) { OR_TRY *out = ((Operator*) op)->stats; OR_CATCH }
int32_t or_operator_destroy(void* op) { delete (Operator*) op; return PG_OK; }

}  // extern "C"
