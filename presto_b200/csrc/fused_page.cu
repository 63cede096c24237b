// presto_b200 — FusedPageOperator: FilterAndProjectOperator → LookupJoinOperator* of one Driver as ONE operator whose output is the
// page the last of them would emit (include/presto_gpu.h pg_fused_page_create; reference: MB/operator/Driver.java:402-536 chaining
// FilterAndProjectOperator.java:88-99 and LookupJoinOperator.java:481-616).  Two launches of one generated row program
// (fused_codegen.cpp generateFusedPageSource, fused_agg.cuh page sink): selection pass → exclusive scan of the per-unit survivor
// counts → emit pass.  The filtered page and every intermediate join page never exist; rows keep their input order, so the result is
// the chain's, row for row.  Fragments outside the fused set (hash-table LookupSource, VARCHAR channels, outer joins) run the chain
// operator by operator on the GPU.
#include "common.cuh"
#include "fused_agg.cuh"
#include "jit_runtime.h"
#include "operators_internal.h"
#include "expr_literals.h"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <map>

namespace pg {

class FusedPageOp : public Operator {
    struct JoinCfg { JoinBridge* bridge; int joinType; std::vector<int32_t> types, hashChannels, outputChannels; int precomputedHash; };
    std::vector<int32_t> inputTypes, projections;
    std::vector<pg_expr_node> nodes;
    ExprLiterals literals;
    int filterNode;
    std::vector<JoinCfg> joins;
    std::vector<int> rawChannels;
    std::vector<LookupSourceView> views;
    struct Compiled { const JitKernel* select = nullptr; const JitKernel* emit = nullptr; FusedPageInfo info; std::string reason; };
    std::map<uint64_t, Compiled> compiled;
    PageStager stager;
    bool finishing = false, hasPending = false;
    DeviceBuffer masks, counts, offsets, scanTemp, errorWord;
    std::vector<DeviceBuffer> outValues, outNulls;
    std::vector<PinnedBuffer> hostValues, hostNulls;
    std::vector<pg_block> outBlocks;
    pg_page pendingPage{};
    std::unique_ptr<Operator> fpOp;
    std::vector<std::unique_ptr<Operator>> joinOps;
    int64_t fusedPages = 0, unfusedPages = 0;
    std::string lastReason;

public:
    explicit FusedPageOp(const pg_fused_page_config* c) : Operator(c->options), stager(this)
    {
        const pg_filter_project_config& f = c->page_processor;
        PG_CHECK_ARG(f.n_input_channels >= 0 && f.n_input_channels <= kFusedMaxInputs, "at most 32 input channels");
        PG_CHECK_ARG(f.n_nodes >= 0 && f.n_nodes <= 64 && f.n_projections >= 0 && f.n_projections <= 32, "page processor program too large");
        PG_CHECK_ARG(c->n_joins >= 0 && c->n_joins <= kFusedMaxJoins, "at most 2 joins in a fused fragment");
        inputTypes.assign(f.input_types, f.input_types + f.n_input_channels);
        nodes.assign(f.nodes, f.nodes + f.n_nodes);
        literals.adopt(nodes);       // the unfused chain builds its FilterAndProjectOperator from `nodes` later
        filterNode = f.filter_node;
        projections.assign(f.projection_nodes, f.projection_nodes + f.n_projections);
        PG_CHECK_ARG(filterNode < f.n_nodes, "filter node out of range");
        for (int p : projections) PG_CHECK_ARG(p >= 0 && p < f.n_nodes, "projection node out of range");
        for (int i = 0; i < f.n_nodes; i++) {
            const pg_expr_node& n = nodes[i];
            if (n.op == PG_EX_INPUT) PG_CHECK_ARG(n.a >= 0 && n.a < f.n_input_channels, "input channel out of range");
            else for (int o : {n.a, n.b, n.c}) PG_CHECK_ARG(o < i, "expression nodes must be topologically ordered");
        }
        for (int j = 0; j < c->n_joins; j++) {
            const pg_lookup_join_config& jc = c->joins[j];
            PG_CHECK_ARG(c->bridges && c->bridges[j], "join bridge is null");
            JoinCfg J;
            J.bridge = (JoinBridge*) c->bridges[j]; J.joinType = jc.join_type; J.precomputedHash = jc.precomputed_hash_channel;
            J.types.assign(jc.types, jc.types + jc.n_channels);
            J.hashChannels.assign(jc.hash_channels, jc.hash_channels + jc.n_hash_channels);
            J.outputChannels.assign(jc.output_channels, jc.output_channels + jc.n_output_channels);
            joins.push_back(J);
        }
        for (int ch = 0; ch < f.n_input_channels; ch++) {
            bool used = false;
            for (const auto& n : nodes) if (n.op == PG_EX_INPUT && n.a == ch) used = true;
            if (used) rawChannels.push_back(ch);
        }
        views.resize(joins.size());
        for (DeviceBuffer* d : {&masks, &counts, &offsets, &scanTemp, &errorWord}) d->accounting = &stats.device_bytes;
        errorWord.reserve(16);
        PG_CUDA(cudaMemsetAsync(errorWord.ptr, 0, 16, stream));
    }

    bool isBlocked() override
    {
        for (size_t j = 0; j < joins.size(); j++) { if (!views[j].ready) describeLookupSource(joins[j].bridge, views[j]); if (!views[j].ready) return true; }
        return false;
    }
    bool needsInput() override
    {
        if (finishing || hasPending) return false;
        for (size_t j = 0; j < joins.size(); j++) { if (!views[j].ready) describeLookupSource(joins[j].bridge, views[j]); if (!views[j].ready) return false; }
        return true;
    }

    Compiled& kernelFor(const pg_page* page)
    {
        uint64_t sig = 0;
        std::vector<bool> hasNulls(inputTypes.size(), false);
        for (int ch : rawChannels) if (page->blocks[ch].is_null) { sig |= 1ULL << ch; hasNulls[ch] = true; }
        auto it = compiled.find(sig);
        if (it != compiled.end()) return it->second;
        Compiled& c = compiled[sig];
        const char* off = getenv("PG_FUSED");
        if (off && off[0] == '0') { c.reason = "PG_FUSED=0"; return c; }
        FusedProgramDesc D;
        D.fp.nodes = nodes; D.fp.inputTypes.assign(inputTypes.begin(), inputTypes.end()); D.fp.inputHasNulls = hasNulls;
        D.fp.filterNode = filterNode; D.fp.projections.assign(projections.begin(), projections.end());
        for (size_t j = 0; j < joins.size(); j++) {
            const JoinCfg& J = joins[j];
            const LookupSourceView& v = views[j];
            if (J.joinType != PG_JOIN_INNER) { c.reason = "outer join"; return c; }
            if (J.hashChannels.size() != 1 || v.nKeys != 1) { c.reason = "multi-column join key"; return c; }
            if (!v.dense) { c.reason = "lookup source is a hash table (duplicate or sparse keys), not the dense index"; return c; }
            FusedJoinDesc fj;
            fj.keyChannel = J.hashChannels[0];
            fj.probeOutputChannels.assign(J.outputChannels.begin(), J.outputChannels.end());
            fj.buildOutputs = v.buildOutputs; fj.strideWords = v.strideWords;
            D.joins.push_back(fj);
        }
        const std::string src = generateFusedPageSource(D, c.info);
        if (!c.info.ok) { c.reason = c.info.reason; return c; }
        std::string err;
        c.select = getOrCompileKernel(options.device, src, "fused_select", &err);
        if (c.select) c.emit = getOrCompileKernel(options.device, src, "fused_emit", &err);
        if (!c.select || !c.emit) { c.select = c.emit = nullptr; c.reason = err; }
        return c;
    }

    void launch(const JitKernel* k, FusedPageParams& L, int rowsPerThread)
    {
        DriverApi& api = DriverApi::get();
        int perSM = 0;
        api.occupancyMaxActiveBlocks(&perSM, k->function, 256, 0);
        perSM = std::max(1, std::min(perSM, 8));
        const int64_t tiles = divUp(L.n, 256 * rowsPerThread);
        const unsigned grid = (unsigned) std::min<int64_t>((int64_t) numSMs * perSM, std::max<int64_t>(tiles, 1));
        void* args[1] = {&L};
        const CUresult r = api.launchKernel(k->function, grid, 1, 1, 256, 1, 1, 0, (CUstream) stream, args, nullptr);
        if (r != CUDA_SUCCESS) throw Error(PG_ERR_CUDA, "cuLaunchKernel of a fused page kernel failed (" + std::to_string((int) r) + ")");
        stats.kernel_launches++;
    }

    void addInput(const pg_page* page) override
    {
        PG_CHECK_STATE(needsInput(), "fused page operator does not need input (lookup sources still building, finishing, or output pending)");
        PG_CHECK_ARG(page->channel_count == (int) inputTypes.size(), "page channel count differs from the page processor's input types");
        const int64_t n = page->position_count;
        stats.input_positions += n;
        if (n == 0) return;
        PG_CHECK_ARG(n < (int64_t(1) << 31), "page exceeds 2^31 positions");
        for (int ch : rawChannels) PG_CHECK_ARG(page->blocks[ch].type == inputTypes[ch], "input channel type differs from the operator's configuration");
        Compiled& c = kernelFor(page);
        bool fusable = c.select != nullptr;
        if (fusable && page->memory == PG_MEM_DEVICE)
            for (int ch : c.info.eagerInputs) if (((uintptr_t) page->blocks[ch].values % 16) != 0 || page->blocks[ch].encoding != PG_ENC_FLAT) { fusable = false; lastReason = "misaligned or encoded device column"; }
        if (!fusable) { if (!c.reason.empty()) lastReason = c.reason; addInputUnfused(page); return; }
        std::vector<DeviceColumn> cols;
        const int slot = stager.stage(page, rawChannels, cols);
        FusedPageParams L{};
        L.n = n;
        for (size_t i = 0; i < rawChannels.size(); i++) { L.values[rawChannels[i]] = cols[i].values; L.nulls[rawChannels[i]] = cols[i].nulls; L.offsets[rawChannels[i]] = cols[i].offsets; }
        for (size_t j = 0; j < joins.size(); j++) { L.joins[j].D = views[j].D; L.joins[j].rows = views[j].rows; }
        L.fpError = errorWord.as<unsigned int>();
        const int64_t units = divUp(n, 64);
        masks.reserve((size_t) units * 8 + 16); counts.reserve((size_t) (units + 1) * 4); offsets.reserve((size_t) (units + 1) * 4);
        PG_CUDA(cudaMemsetAsync(counts.as<uint32_t>() + units, 0, 4, stream));
        L.masks = masks.as<uint32_t>(); L.counts = counts.as<uint32_t>();
        launch(c.select, L, c.info.rowsPerThread);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.as<uint32_t>(), offsets.as<uint32_t>(), (int) (units + 1), stream);
        scanTemp.reserve(tb);
        cub::DeviceScan::ExclusiveSum(scanTemp.ptr, tb, counts.as<uint32_t>(), offsets.as<uint32_t>(), (int) (units + 1), stream);
        stats.kernel_launches += 2;
        struct { uint32_t total; uint32_t err; } hostWord{};
        PG_CUDA(cudaMemcpyAsync(&hostWord.total, offsets.as<uint32_t>() + units, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaMemcpyAsync(&hostWord.err, errorWord.ptr, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        stats.d2h_bytes += 8;
        checkError(hostWord.err);
        const int64_t total = hostWord.total;
        fusedPages++;
        if (total == 0) { stager.markCompute(slot); return; }
        const size_t nOut = c.info.outputTypes.size();
        if (outValues.size() < nOut) { outValues.resize(nOut); outNulls.resize(nOut); for (auto& d : outValues) d.accounting = &stats.device_bytes; for (auto& d : outNulls) d.accounting = &stats.device_bytes; }
        outBlocks.assign(nOut, pg_block{});
        L.unitOffsets = offsets.as<uint32_t>();
        for (size_t i = 0; i < nOut; i++) {
            outValues[i].reserve((size_t) total * typeWidth(c.info.outputTypes[i]));
            L.outValues[i] = outValues[i].ptr; L.outNulls[i] = nullptr;
            if (c.info.outputNullable[i]) { outNulls[i].reserve((size_t) total); L.outNulls[i] = outNulls[i].as<uint8_t>(); }
        }
        launch(c.emit, L, c.info.rowsPerThread);
        stager.markCompute(slot);
        const bool toHost = options.output_memory == PG_MEM_HOST;
        if (toHost && hostValues.size() < nOut) { hostValues.resize(nOut); hostNulls.resize(nOut); }
        for (size_t i = 0; i < nOut; i++) {
            pg_block& b = outBlocks[i];
            const size_t vbytes = (size_t) total * typeWidth(c.info.outputTypes[i]);
            b.type = c.info.outputTypes[i]; b.encoding = PG_ENC_FLAT; b.position_count = total;
            if (toHost) {
                hostValues[i].reserve(std::max<size_t>(vbytes, 16));
                PG_CUDA(cudaMemcpyAsync(hostValues[i].ptr, outValues[i].ptr, vbytes, cudaMemcpyDeviceToHost, stream));
                b.values = hostValues[i].ptr; b.is_null = nullptr;
                stats.d2h_bytes += (int64_t) vbytes;
                if (L.outNulls[i]) { hostNulls[i].reserve((size_t) total); PG_CUDA(cudaMemcpyAsync(hostNulls[i].ptr, outNulls[i].ptr, (size_t) total, cudaMemcpyDeviceToHost, stream)); b.is_null = hostNulls[i].as<uint8_t>(); stats.d2h_bytes += total; }
            }
            else { b.values = outValues[i].ptr; b.is_null = L.outNulls[i]; }
            stats.output_bytes += (int64_t) vbytes;
        }
        uint32_t err = 0;
        PG_CUDA(cudaMemcpyAsync(&err, errorWord.ptr, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        checkError(err);
        pendingPage.position_count = total; pendingPage.channel_count = (int32_t) nOut; pendingPage.memory = options.output_memory; pendingPage.blocks = outBlocks.data();
        hasPending = true;
        stats.output_positions += total;
    }

    void checkError(uint32_t e)
    {
        if (!e) return;
        PG_CUDA(cudaMemsetAsync(errorWord.ptr, 0, 4, stream));
        if (e & 1u) throw Error(PG_ERR_NUMERIC_OVERFLOW, "bigint arithmetic overflow");
        if (e & 2u) throw Error(PG_ERR_INVALID_ARGUMENT, "Division by zero");
    }

    void addInputUnfused(const pg_page* page)
    {
        pg_operator_options o = options;
        o.cuda_stream = (void*) stream;
        if (!fpOp) {
            pg_operator_options inner = o; inner.output_memory = joins.empty() ? options.output_memory : PG_MEM_DEVICE;
            pg_filter_project_config f{};
            f.n_input_channels = (int) inputTypes.size(); f.input_types = inputTypes.data(); f.n_nodes = (int) nodes.size(); f.nodes = nodes.data();
            f.filter_node = filterNode; f.n_projections = (int) projections.size(); f.projection_nodes = projections.data(); f.options = inner;
            fpOp.reset(createFilterProject(&f));
            for (size_t j = 0; j < joins.size(); j++) {
                auto& J = joins[j];
                pg_lookup_join_config jc{};
                jc.join_type = J.joinType; jc.n_channels = (int) J.types.size(); jc.types = J.types.data();
                jc.n_hash_channels = (int) J.hashChannels.size(); jc.hash_channels = J.hashChannels.data();
                jc.n_output_channels = (int) J.outputChannels.size(); jc.output_channels = J.outputChannels.data();
                jc.precomputed_hash_channel = J.precomputedHash; jc.options = o;
                jc.options.output_memory = j + 1 == joins.size() ? options.output_memory : PG_MEM_DEVICE;
                joinOps.emplace_back(createLookupJoin(&jc, J.bridge));
            }
        }
        unfusedPages++;
        fpOp->addInput(page);
        pg_page cur{};
        bool has = fpOp->getOutput(&cur);
        for (auto& j : joinOps) {
            if (!has || cur.position_count == 0) { has = false; break; }
            j->addInput(&cur);
            has = j->getOutput(&cur);
        }
        if (has && cur.position_count > 0) { pendingPage = cur; hasPending = true; stats.output_positions += cur.position_count; }
    }

    bool getOutput(pg_page* out) override
    {
        if (!hasPending) return false;
        *out = pendingPage;
        hasPending = false;
        return true;
    }
    void finish() override { finishing = true; }
    bool isFinished() override { return finishing && !hasPending; }
    void info(int64_t* fused, int64_t* unfused, char* reason, size_t cap) const
    {
        if (fused) *fused = fusedPages;
        if (unfused) *unfused = unfusedPages;
        if (reason && cap) snprintf(reason, cap, "%s", lastReason.c_str());
    }
};

}  // namespace pg

extern "C" int32_t pg_fused_page_create(const pg_fused_page_config* config, pg_operator** out)
{
    try {
        if (!config || !out) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "null argument");
        *out = (pg_operator*) new pg::FusedPageOp(config);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
extern "C" int32_t pg_fused_page_get_info(pg_operator* op, int64_t* fused, int64_t* unfused, char* reason, size_t capacity)
{
    try {
        auto* p = dynamic_cast<pg::FusedPageOp*>((pg::Operator*) op);
        if (!p) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a fused page operator");
        p->info(fused, unfused, reason, capacity);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    return PG_OK;
}
// The CUDA C++ of both passes for a fragment, without a device (join build sides described by their output channel types, payload rows
// laid out as HashBuilderOperator would: widest columns first, no NULL bytes) — same contract as pg_fused_aggregation_explain.
extern "C" int32_t pg_fused_page_explain(const pg_fused_page_config* c, uint64_t nullableInputs, const int32_t* const* buildTypes, const int32_t* nBuildTypes,
                                         char* buffer, size_t capacity, size_t* needed)
{
    try {
        using namespace pg;
        PG_CHECK_ARG(c && needed, "null argument");
        const pg_filter_project_config& f = c->page_processor;
        FusedProgramDesc D;
        D.fp.nodes.assign(f.nodes, f.nodes + f.n_nodes);
        D.fp.inputTypes.assign(f.input_types, f.input_types + f.n_input_channels);
        D.fp.filterNode = f.filter_node;
        D.fp.projections.assign(f.projection_nodes, f.projection_nodes + f.n_projections);
        D.fp.inputHasNulls.assign(D.fp.inputTypes.size(), false);
        for (size_t i = 0; i < D.fp.inputTypes.size(); i++) D.fp.inputHasNulls[i] = (nullableInputs >> i) & 1ULL;
        for (int j = 0; j < c->n_joins; j++) {
            const pg_lookup_join_config& jc = c->joins[j];
            if (jc.join_type != PG_JOIN_INNER || jc.n_hash_channels != 1) throw Error(PG_ERR_UNSUPPORTED, "fragment is not fusable: join shape");
            FusedJoinDesc fj;
            fj.keyChannel = jc.hash_channels[0];
            fj.probeOutputChannels.assign(jc.output_channels, jc.output_channels + jc.n_output_channels);
            const int nb = nBuildTypes ? nBuildTypes[j] : 0;
            fj.buildOutputs.resize(nb);
            int off = 0;
            for (int width : {8, 4, 1})
                for (int i = 0; i < nb; i++)
                    if (typeWidth(buildTypes[j][i]) == width) { FusedBuildColumn& b = fj.buildOutputs[i]; b.type = buildTypes[j][i]; b.width = width; b.word = off >> 3; b.shift = 8 * (off & 7); off += width; }
            fj.strideWords = nb ? std::max(1, (off + 7) / 8) : 0;
            if (fj.strideWords == 3) fj.strideWords = 4;
            D.joins.push_back(fj);
        }
        FusedPageInfo info;
        const std::string src = generateFusedPageSource(D, info);
        if (!info.ok) throw Error(PG_ERR_UNSUPPORTED, "fragment is not fusable: " + info.reason);
        *needed = src.size() + 1;
        if (buffer && capacity >= src.size() + 1) memcpy(buffer, src.c_str(), src.size() + 1);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
