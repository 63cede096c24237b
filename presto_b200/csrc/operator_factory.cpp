// presto_b200 — OperatorFactory lifecycle (MB/operator/OperatorFactory.java:19-51): a factory holds one plan node's immutable
// configuration, hands out one operator per Driver (createOperator :21), refuses after noMoreOperators (:33; the reference's
// factories checkState(!closed, "Factory is already closed"): HashAggregationOperator.java:222,249-252), and duplicate()s itself for
// another pipeline instance (:50; HashAggregationOperator.java:255-275 builds a fresh OPEN factory from the same lists,
// LookupJoinOperatorFactory shares the JoinBridgeManager).  Host code only: the operators themselves are created by the
// pg_*_create entry points.
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/presto_gpu.h"

namespace pg { void setLastError(const std::string& msg); }

namespace {

struct FactoryError { int code; std::string message; };

template <typename T>
const T* keep(std::vector<std::vector<char>>& arena, const T* p, size_t n)
{
    if (!p || n == 0) return nullptr;
    arena.emplace_back((const char*) p, (const char*) p + n * sizeof(T));
    return (const T*) arena.back().data();
}

// Deep copies of the configuration structs: every array a config points to moves into `arena`.
struct Plan {
    int kind = 0;
    std::vector<std::vector<char>> arena;
    pg_hash_aggregation_config aggregation{};
    pg_filter_project_config filterProject{};
    pg_hash_builder_config builder{};
    pg_lookup_join_config join{};
    pg_semi_join_config semiJoin{};
    pg_partitioned_output_config output{};
    pg_fused_aggregation_config fusedAggregation{};
    pg_fused_page_config fusedPage{};
    pg_join_bridge* bridge = nullptr;            // shared, never copied (JoinBridgeManager)

    void copyAggregation(pg_hash_aggregation_config& d, const pg_hash_aggregation_config& s)
    {
        d = s;
        d.group_channels = keep(arena, s.group_channels, s.n_group_channels);
        d.group_types = keep(arena, s.group_types, s.n_group_channels);
        d.group_max_lengths = keep(arena, s.group_max_lengths, s.n_group_channels);
        d.aggregates = keep(arena, s.aggregates, s.n_aggregates);
    }
    void copyFilterProject(pg_filter_project_config& d, const pg_filter_project_config& s)
    {
        d = s;
        d.input_types = keep(arena, s.input_types, s.n_input_channels);
        d.nodes = keep(arena, s.nodes, s.n_nodes);
        d.projection_nodes = keep(arena, s.projection_nodes, s.n_projections);
        // VARCHAR literals live in the caller's memory: the copied nodes point at copies of their bytes
        for (int i = 0; i < s.n_nodes; i++) {
            pg_expr_node& n = const_cast<pg_expr_node&>(d.nodes[i]);
            if (n.op != PG_EX_CONST_VARCHAR || n.reserved <= 0 || n.imm_i == 0) continue;
            n.imm_i = (int64_t) (uintptr_t) keep(arena, (const char*) (uintptr_t) n.imm_i, (size_t) n.reserved);
        }
    }
    void copyJoin(pg_lookup_join_config& d, const pg_lookup_join_config& s)
    {
        d = s;
        d.types = keep(arena, s.types, s.n_channels);
        d.hash_channels = keep(arena, s.hash_channels, s.n_hash_channels);
        d.output_channels = keep(arena, s.output_channels, s.n_output_channels);
    }
    const pg_lookup_join_config* copyJoins(const pg_lookup_join_config* s, int n)
    {
        if (n <= 0 || !s) return nullptr;
        std::vector<pg_lookup_join_config> joins(n);
        for (int j = 0; j < n; j++) copyJoin(joins[j], s[j]);
        return keep(arena, joins.data(), n);
    }

    Plan(int k, const void* config, pg_join_bridge* b) : kind(k), bridge(b)
    {
        if (!config) throw FactoryError{PG_ERR_INVALID_ARGUMENT, "null configuration"};
        switch (kind) {
            case PG_FACTORY_HASH_AGGREGATION: copyAggregation(aggregation, *(const pg_hash_aggregation_config*) config); break;
            case PG_FACTORY_FILTER_PROJECT: copyFilterProject(filterProject, *(const pg_filter_project_config*) config); break;
            case PG_FACTORY_HASH_BUILDER: {
                const auto& s = *(const pg_hash_builder_config*) config;
                builder = s;
                builder.types = keep(arena, s.types, s.n_channels);
                builder.hash_channels = keep(arena, s.hash_channels, s.n_hash_channels);
                builder.output_channels = keep(arena, s.output_channels, s.n_output_channels);
                if (s.filter) {      // the join filter function belongs to the build side's plan node
                    pg_filter_project_config f{};
                    copyFilterProject(f, *s.filter);
                    builder.filter = keep(arena, &f, 1);
                }
                break;
            }
            case PG_FACTORY_LOOKUP_JOIN: copyJoin(join, *(const pg_lookup_join_config*) config); break;
            case PG_FACTORY_HASH_SEMI_JOIN: {
                const auto& s = *(const pg_semi_join_config*) config;
                semiJoin = s;
                semiJoin.types = keep(arena, s.types, s.n_channels);
                break;
            }
            case PG_FACTORY_PARTITIONED_OUTPUT: {
                const auto& s = *(const pg_partitioned_output_config*) config;
                output = s;
                output.types = keep(arena, s.types, s.n_channels);
                output.partition_channels = keep(arena, s.partition_channels, s.n_partition_channels);
                output.bucket_to_partition = s.bucket_to_partition ? keep(arena, s.bucket_to_partition, (size_t) std::max(s.bucket_count, 0)) : nullptr;
                break;
            }
            case PG_FACTORY_FUSED_AGGREGATION: {
                const auto& s = *(const pg_fused_aggregation_config*) config;
                fusedAggregation = s;
                copyFilterProject(fusedAggregation.page_processor, s.page_processor);
                fusedAggregation.joins = copyJoins(s.joins, s.n_joins);
                fusedAggregation.bridges = keep(arena, s.bridges, s.n_joins);
                copyAggregation(fusedAggregation.aggregation, s.aggregation);
                break;
            }
            case PG_FACTORY_FUSED_PAGE: {
                const auto& s = *(const pg_fused_page_config*) config;
                fusedPage = s;
                copyFilterProject(fusedPage.page_processor, s.page_processor);
                fusedPage.joins = copyJoins(s.joins, s.n_joins);
                fusedPage.bridges = keep(arena, s.bridges, s.n_joins);
                break;
            }
            default: throw FactoryError{PG_ERR_INVALID_ARGUMENT, "unknown pg_operator_kind " + std::to_string(kind)};
        }
        if ((kind == PG_FACTORY_HASH_BUILDER || kind == PG_FACTORY_LOOKUP_JOIN || kind == PG_FACTORY_HASH_SEMI_JOIN) && !bridge)
            throw FactoryError{PG_ERR_INVALID_ARGUMENT, "this operator kind needs a join bridge"};
    }

    // the per-Driver part of the configuration (DriverContext): device, stream, output placement
    int32_t create(const pg_operator_options* o, pg_operator** out) const
    {
        switch (kind) {
            case PG_FACTORY_HASH_AGGREGATION: { auto c = aggregation; if (o) c.options = *o; return pg_hash_aggregation_create(&c, out); }
            case PG_FACTORY_FILTER_PROJECT: { auto c = filterProject; if (o) c.options = *o; return pg_filter_project_create(&c, out); }
            case PG_FACTORY_HASH_BUILDER: { auto c = builder; if (o) c.options = *o; return pg_hash_builder_create(&c, bridge, out); }
            case PG_FACTORY_LOOKUP_JOIN: { auto c = join; if (o) c.options = *o; return pg_lookup_join_create(&c, bridge, out); }
            case PG_FACTORY_HASH_SEMI_JOIN: { auto c = semiJoin; if (o) c.options = *o; return pg_hash_semi_join_create(&c, bridge, out); }
            case PG_FACTORY_PARTITIONED_OUTPUT: { auto c = output; if (o) c.options = *o; return pg_partitioned_output_create(&c, out); }
            case PG_FACTORY_FUSED_AGGREGATION: { auto c = fusedAggregation; if (o) c.aggregation.options = *o; return pg_fused_aggregation_create(&c, out); }
            default: { auto c = fusedPage; if (o) c.options = *o; return pg_fused_page_create(&c, out); }
        }
    }
};

}  // namespace

struct pg_operator_factory {
    std::shared_ptr<const Plan> plan;     // duplicates share the immutable plan
    std::mutex mu;
    bool closed = false;
    int64_t created = 0;
};

#define FACTORY_TRY try {
#define FACTORY_CATCH                                                                                     \
    }                                                                                                     \
    catch (const FactoryError& e) { pg::setLastError(e.message); return e.code; }                         \
    catch (const std::bad_alloc&) { pg::setLastError("host out of memory"); return PG_ERR_OUT_OF_MEMORY; } \
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }          \
    return PG_OK;

extern "C" {

int32_t pg_operator_factory_create(int32_t kind, const void* config, pg_join_bridge* bridge, pg_operator_factory** out)
{
    FACTORY_TRY
    if (!out) throw FactoryError{PG_ERR_INVALID_ARGUMENT, "null argument"};
    auto* f = new pg_operator_factory();
    try { f->plan = std::make_shared<const Plan>(kind, config, bridge); }
    catch (...) { delete f; throw; }
    *out = f;
    FACTORY_CATCH
}

int32_t pg_operator_factory_create_operator(pg_operator_factory* f, const pg_operator_options* options, pg_operator** out)
{
    FACTORY_TRY
    if (!f || !out) throw FactoryError{PG_ERR_INVALID_ARGUMENT, "null argument"};
    {
        std::lock_guard<std::mutex> lock(f->mu);
        if (f->closed) throw FactoryError{PG_ERR_INVALID_STATE, "Factory is already closed"};     // checkState(!closed, ...) of every reference factory
    }
    const int32_t rc = f->plan->create(options, out);
    if (rc != PG_OK) return rc;        // pg_last_error already set by the operator's constructor
    std::lock_guard<std::mutex> lock(f->mu);
    f->created++;
    FACTORY_CATCH
}

int32_t pg_operator_factory_no_more_operators(pg_operator_factory* f)
{
    FACTORY_TRY
    if (!f) throw FactoryError{PG_ERR_INVALID_ARGUMENT, "null argument"};
    std::lock_guard<std::mutex> lock(f->mu);
    f->closed = true;
    FACTORY_CATCH
}

int32_t pg_operator_factory_duplicate(pg_operator_factory* f, pg_operator_factory** out)
{
    FACTORY_TRY
    if (!f || !out) throw FactoryError{PG_ERR_INVALID_ARGUMENT, "null argument"};
    auto* d = new pg_operator_factory();
    d->plan = f->plan;              // an open factory even when `f` is closed (HashAggregationOperator.java:255-275)
    *out = d;
    FACTORY_CATCH
}

int32_t pg_operator_factory_get_created_count(pg_operator_factory* f, int64_t* count)
{
    FACTORY_TRY
    if (!f || !count) throw FactoryError{PG_ERR_INVALID_ARGUMENT, "null argument"};
    std::lock_guard<std::mutex> lock(f->mu);
    *count = f->created;
    FACTORY_CATCH
}

int32_t pg_operator_factory_destroy(pg_operator_factory* f)
{
    delete f;
    return PG_OK;
}

}  // extern "C"
