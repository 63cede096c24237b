// presto_b200 — FilterAndProjectOperator / PageProcessor on sm_100a.
//
// Replaces (MB = presto-main-base/src/main/java/com/facebook/presto/):
//   MB/operator/FilterAndProjectOperator.java:88-99, MB/operator/project/PageProcessor.java:112-148,180-343
//   MB/operator/project/PageFilter.java:27-50 (boolean[] → SelectedPositions), PageProjection.java:24-31
//   MB/sql/gen/PageFunctionCompiler.java:195-861 — the reference JIT-compiles the RowExpression tree to JVM bytecode;
//   here the tree arrives as a flat node program (pg_expr_node) and ONE kernel interprets it with warp-uniform control flow.
//
//   pass A  every thread evaluates the filter for its row; a warp-unit's selected count goes to counts[unit]
//   scan    exclusive sum → first output row of every unit (order preserving)
//   pass B  every selected row re-evaluates the program and writes its projections at offset + rank-in-unit
// Arithmetic follows MB/type/DoubleOperators.java:62-81 (plain IEEE + - * /) and BigintOperators.java:68-130 (exact, overflow →
// NUMERIC_VALUE_OUT_OF_RANGE); AND / OR / NOT are SQL three-valued.
#include "common.cuh"
#include "fp_interp.cuh"
#include "fp_codegen.h"
#include <cub/device/device_scan.cuh>
#include "jit_runtime.h"
#include "operators_internal.h"
#include <algorithm>
#include <map>
#include <mutex>

namespace pg {

constexpr int kMaxProjections = 32;
constexpr int kRowsPerFpUnit = 256;   // rows owned by one warp per step (8 rounds of 32)

struct FpOutput { void* values; uint8_t* nulls; int32_t node; int32_t width; };

struct FpParams {
    FpNode nodes[kMaxNodes];
    int32_t nNodes;
    FpInput inputs[kMaxInputs];
    int32_t filterNode;
    FpOutput outs[kMaxProjections];
    int32_t nOuts;
    int64_t n, units;
    uint32_t* counts;          // pass A: [units + 1]
    const uint32_t* offsets;   // pass B
    uint32_t* selection;       // one bit per row: pass A writes, pass B reads
    uint64_t needFilter, needProject;   // nodes each pass has to evaluate (sub-DAG of the filter / of the projections)
    unsigned int* error;
};

__device__ __forceinline__ Val evalNode(const FpParams& P, const FpNode& nd, int64_t row, const Val& x, const Val& y, const Val& z)
{
    return evalNodeAt(P.nodes, P.inputs, P.error, nd, row, x, y, z);
}

// Evaluates the nodes of `needMask` for R rows of one thread at once (one switch dispatch per node for R rows); register files
// live in local memory (L1 resident), the program is warp-uniform.  Rows that are out of range read nothing (active[r] false).
template <int R>
__device__ __forceinline__ void evalRows(const FpParams& P, uint64_t needMask, const int64_t (&row)[R], const bool (&active)[R],
                                         uint64_t (&regs)[kMaxNodes][R], uint64_t (&nullBits)[R])
{
#pragma unroll
    for (int r = 0; r < R; r++) nullBits[r] = 0;
    for (int k = 0; k < P.nNodes; k++) {
        if (!((needMask >> k) & 1ULL)) continue;
        const FpNode& nd = P.nodes[k];
        const bool isInput = nd.op == PG_EX_INPUT;
#pragma unroll
        for (int r = 0; r < R; r++) {
            Val x{0, false}, y{0, false}, z{0, false};
            if (!isInput) {
                if (nd.a >= 0) x = Val{regs[nd.a][r], (bool) ((nullBits[r] >> nd.a) & 1ULL)};
                if (nd.b >= 0) y = Val{regs[nd.b][r], (bool) ((nullBits[r] >> nd.b) & 1ULL)};
                if (nd.c >= 0) z = Val{regs[nd.c][r], (bool) ((nullBits[r] >> nd.c) & 1ULL)};
            }
            Val v{0, false};
            if (active[r]) v = evalNode(P, nd, row[r], x, y, z);
            regs[k][r] = v.bits;
            if (v.isNull) nullBits[r] |= 1ULL << k;
        }
    }
}

constexpr int kFpRows = 2;   // rows per thread per dispatch

// pass A: selection bitmask (one 32-bit word per 32 rows) + selected count per unit
__global__ void __launch_bounds__(256) fpFilterKernel(const __grid_constant__ FpParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    uint64_t regs[kMaxNodes][kFpRows];
    for (int64_t unit = warpId; unit < P.units; unit += nWarps) {
        const int64_t begin = unit * kRowsPerFpUnit, end = min(begin + (int64_t) kRowsPerFpUnit, P.n);
        uint32_t selected = 0;
        for (int64_t base = begin; base < end; base += 32 * kFpRows) {
            int64_t row[kFpRows]; bool active[kFpRows]; uint64_t nb[kFpRows];
#pragma unroll
            for (int r = 0; r < kFpRows; r++) { row[r] = base + r * 32 + lane; active[r] = row[r] < end; }
            evalRows<kFpRows>(P, P.needFilter, row, active, regs, nb);
#pragma unroll
            for (int r = 0; r < kFpRows; r++) {
                const bool sel = active[r] && !((nb[r] >> P.filterNode) & 1ULL) && regs[P.filterNode][r] != 0;   // only TRUE selects
                const unsigned m = __ballot_sync(0xffffffffu, sel);
                if (lane == 0 && base + r * 32 < end) P.selection[(base + r * 32) >> 5] = m;
                selected += __popc(m);
            }
        }
        if (lane == 0) P.counts[unit] = selected;
    }
}

// pass B: projections of the selected rows, written at offsets[unit] + rank inside the unit
__global__ void __launch_bounds__(256) fpProjectKernel(const __grid_constant__ FpParams P)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    uint64_t regs[kMaxNodes][kFpRows];
    for (int64_t unit = warpId; unit < P.units; unit += nWarps) {
        const int64_t begin = unit * kRowsPerFpUnit, end = min(begin + (int64_t) kRowsPerFpUnit, P.n);
        uint32_t cursor = P.filterNode >= 0 ? P.offsets[unit] : (uint32_t) begin;
        for (int64_t base = begin; base < end; base += 32 * kFpRows) {
            int64_t row[kFpRows]; bool active[kFpRows]; uint64_t nb[kFpRows]; unsigned masks[kFpRows];
#pragma unroll
            for (int r = 0; r < kFpRows; r++) {
                row[r] = base + r * 32 + lane;
                const bool in = row[r] < end;
                masks[r] = 0;
                if (base + r * 32 < end) masks[r] = P.filterNode >= 0 ? P.selection[(base + r * 32) >> 5] : __ballot_sync(0xffffffffu, in);
                active[r] = in && ((masks[r] >> lane) & 1u);     // unselected rows load nothing
            }
            evalRows<kFpRows>(P, P.needProject, row, active, regs, nb);
#pragma unroll
            for (int r = 0; r < kFpRows; r++) {
                if (active[r]) {
                    const uint32_t dest = cursor + __popc(masks[r] & lt);
                    for (int j = 0; j < P.nOuts; j++) {
                        const FpOutput& o = P.outs[j];
                        const bool isNull = (nb[r] >> o.node) & 1ULL;
                        const uint64_t v = isNull ? 0 : regs[o.node][r];
                        if (o.width == 8) ((uint64_t*) o.values)[dest] = v;
                        else if (o.width == 4) ((uint32_t*) o.values)[dest] = (uint32_t) v;
                        else ((uint8_t*) o.values)[dest] = (uint8_t) v;
                        if (o.nulls) o.nulls[dest] = isNull ? 1 : 0;
                    }
                }
                cursor += __popc(masks[r]);
            }
        }
    }
}

// ---- VARCHAR input references (InputPageProjection, MB/operator/project/InputPageProjection.java): the selected positions' slices pass
// through.  The fixed-width machinery above never sees them: pass 1 writes the selected rows' lengths at their output positions, a scan
// turns them into the output offsets, pass 2 copies the bytes.  `selection` == nullptr: every row is selected (dest = row).
__global__ void varcharSelectLensKernel(const uint32_t* selection, const uint32_t* unitOffsets, const int32_t* inOffsets, const uint8_t* inNulls, int64_t n,
                                        int32_t* outLens, uint8_t* outNulls)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const int64_t units = (n + kRowsPerFpUnit - 1) / kRowsPerFpUnit;
    for (int64_t unit = warpId; unit < units; unit += nWarps) {
        const int64_t begin = unit * kRowsPerFpUnit, end = min(begin + (int64_t) kRowsPerFpUnit, n);
        uint32_t cursor = selection ? unitOffsets[unit] : (uint32_t) begin;
        for (int64_t base = begin; base < end; base += 32) {
            const int64_t row = base + lane;
            const unsigned m = selection ? selection[base >> 5] : __ballot_sync(0xffffffffu, row < end);
            if (row < end && ((m >> lane) & 1u)) {
                const uint32_t dest = cursor + __popc(m & lt);
                const bool isNull = inNulls && inNulls[row];
                outLens[dest] = isNull ? 0 : inOffsets[row + 1] - inOffsets[row];
                if (outNulls) outNulls[dest] = isNull ? 1 : 0;
            }
            cursor += __popc(m);
        }
    }
}
__global__ void varcharSelectCopyKernel(const uint32_t* selection, const uint32_t* unitOffsets, const uint8_t* inBytes, const int32_t* inOffsets, const uint8_t* inNulls,
                                        int64_t n, const int32_t* outOffsets, uint8_t* outBytes)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const int64_t units = (n + kRowsPerFpUnit - 1) / kRowsPerFpUnit;
    for (int64_t unit = warpId; unit < units; unit += nWarps) {
        const int64_t begin = unit * kRowsPerFpUnit, end = min(begin + (int64_t) kRowsPerFpUnit, n);
        uint32_t cursor = selection ? unitOffsets[unit] : (uint32_t) begin;
        for (int64_t base = begin; base < end; base += 32) {
            const int64_t row = base + lane;
            const unsigned m = selection ? selection[base >> 5] : __ballot_sync(0xffffffffu, row < end);
            if (row < end && ((m >> lane) & 1u) && !(inNulls && inNulls[row])) {
                const uint32_t dest = cursor + __popc(m & lt);
                const int32_t from = inOffsets[row], len = inOffsets[row + 1] - from, to = outOffsets[dest];
                for (int32_t i = 0; i < len; i++) outBytes[to + i] = inBytes[from + i];
            }
            cursor += __popc(m);
        }
    }
}

// ---- driver-API JIT of generated PTX (DriverApi: jit_runtime.h; libcuda is dlopen'ed so the library still loads without a driver) ----
struct JitProgram {
    CUmodule module = nullptr;
    CUfunction filter = nullptr, project = nullptr;
    FpKernelSignature sig;
    bool ok = false;
    std::string log;
};

class FilterProjectOp : public Operator {
    std::vector<int> inputTypes, staged;
    std::vector<pg_expr_node> nodes;
    std::vector<bool> used;
    int filterNode;
    std::vector<int> projections;              // the fixed-width projections (what the generated / interpreter kernels evaluate)
    struct OutSlot { bool varchar; int index; };   // output channel → fixed-width projection index, or VARCHAR pass-through index
    std::vector<OutSlot> outSlots;
    std::vector<int> varcharChannels;          // input channel of each VARCHAR pass-through output
    std::vector<DeviceBuffer> vcLens, vcOffsets, vcBytes, vcNulls;
    std::vector<PinnedBuffer> vcHostBytes, vcHostOffsets, vcHostNulls;
    DeviceBuffer vcScanTemp;
    std::vector<pg_block> publicBlocks;
    PageStager stager;
    bool finishing = false, hasPending = false;
    DeviceBuffer counts, offsets, scanTemp, errorWord, selection;
    uint64_t needFilter = 0, needProject = 0;
    std::vector<DeviceBuffer> outValues, outNulls;
    std::vector<PinnedBuffer> hostValues, hostNulls;
    std::vector<pg_block> outBlocks;
    int64_t pendingRows = 0;
    int lastIsList = 0; int64_t lastSize = 0;
    std::map<uint64_t, JitProgram> jitCache;   // keyed by which inputs carry a null map on this page
    int64_t jitLaunches = 0, interpreterLaunches = 0;
    DeviceBuffer literalPool;                  // bytes of the CONST_VARCHAR nodes
    std::vector<std::string> literalHost;      // the same per node, for the PTX generator (slice = literal compares against immediates)
    std::vector<int64_t> literalOffset;        // per node: offset of its literal in the pool (-1: not a literal)
    bool varcharExpressions = false;           // a comparison / hash reads slices: interpreter kernels only

public:
    explicit FilterProjectOp(const pg_filter_project_config* c) : Operator(c->options), stager(this)
    {
        PG_CHECK_ARG(c->n_input_channels >= 0 && c->n_input_channels <= kMaxInputs, "at most 32 input channels");
        PG_CHECK_ARG(c->n_nodes >= 0 && c->n_nodes <= kMaxNodes, "at most 64 expression nodes");
        PG_CHECK_ARG(c->n_projections >= 0 && c->n_projections <= kMaxProjections, "at most 32 projections");
        inputTypes.assign(c->input_types, c->input_types + c->n_input_channels);
        nodes.assign(c->nodes, c->nodes + c->n_nodes);
        filterNode = c->filter_node;
        PG_CHECK_ARG(filterNode < c->n_nodes, "filter node out of range");
        if (filterNode >= 0) PG_CHECK_ARG(nodes[filterNode].type == PG_TYPE_BOOLEAN, "filter must be BOOLEAN");
        // VARCHAR: input references only (InputPageProjection: the selected positions' slices pass through); they are split off here and
        // never reach the expression kernels
        for (int j = 0; j < c->n_projections; j++) {
            const int p = c->projection_nodes[j];
            PG_CHECK_ARG(p >= 0 && p < c->n_nodes, "projection node out of range");
            if (nodes[p].type == PG_TYPE_VARCHAR) {
                if (nodes[p].op != PG_EX_INPUT) throw Error(PG_ERR_UNSUPPORTED, "VARCHAR projections other than input references are not on the GPU path");
                outSlots.push_back({true, (int) varcharChannels.size()});
                varcharChannels.push_back(nodes[p].a);
            }
            else { outSlots.push_back({false, (int) projections.size()}); projections.push_back(p); }
        }
        for (int i = 0; i < c->n_nodes; i++) {
            const pg_expr_node& n = nodes[i];
            if (n.type == PG_TYPE_VARCHAR && n.op != PG_EX_INPUT && n.op != PG_EX_CONST_VARCHAR)
                throw Error(PG_ERR_UNSUPPORTED, "VARCHAR-valued expressions other than input references and literals are not on the GPU path");
            if (n.op == PG_EX_CONST_VARCHAR) {
                PG_CHECK_ARG(n.type == PG_TYPE_VARCHAR && n.reserved >= 0 && (n.reserved == 0 || n.imm_i != 0), "CONST_VARCHAR: imm_i = address of the bytes, reserved = length");
                continue;
            }
            if (n.op == PG_EX_INPUT) {
                PG_CHECK_ARG(n.a >= 0 && n.a < c->n_input_channels, "input channel out of range");
                PG_CHECK_ARG(inputTypes[n.a] == n.type, "input node type differs from the channel's type");
                if (std::find(staged.begin(), staged.end(), n.a) == staged.end()) staged.push_back(n.a);
            }
            else {
                for (int o : {n.a, n.b, n.c}) PG_CHECK_ARG(o < i, "expression nodes must be topologically ordered");
                if (n.op == PG_EX_CAST_BIGINT) throw Error(PG_ERR_UNSUPPORTED, "CAST to BIGINT is not on the GPU path yet");
                // VARCHAR operands: comparisons between two VARCHAR values, BETWEEN over three, IS_NULL, HASH (VarcharOperators.java:54-105,291)
                int nVarchar = 0, nOperands = 0;
                for (int o : {n.a, n.b, n.c}) if (o >= 0) { nOperands++; nVarchar += nodes[o].type == PG_TYPE_VARCHAR; }
                if (nVarchar) {
                    const bool comparison = (n.op >= PG_EX_EQ && n.op <= PG_EX_GE) || n.op == PG_EX_BETWEEN;
                    if (!((comparison && nVarchar == nOperands) || n.op == PG_EX_IS_NULL || n.op == PG_EX_HASH))
                        throw Error(PG_ERR_UNSUPPORTED, "VARCHAR operands: only comparisons among VARCHAR values, BETWEEN, IS NULL and the type hash are on the GPU path");
                    if (n.op != PG_EX_IS_NULL) varcharExpressions = true;
                }
            }
        }
        {   // the literals go to the device once
            literalOffset.assign(c->n_nodes, -1);
            std::vector<uint8_t> bytes;
            for (int i = 0; i < c->n_nodes; i++) {
                if (nodes[i].op != PG_EX_CONST_VARCHAR) continue;
                literalOffset[i] = (int64_t) bytes.size();
                const uint8_t* src = (const uint8_t*) (uintptr_t) nodes[i].imm_i;
                bytes.insert(bytes.end(), src, src + nodes[i].reserved);
                if (literalHost.empty()) literalHost.resize((size_t) c->n_nodes);
                literalHost[(size_t) i].assign((const char*) src, (size_t) nodes[i].reserved);
                nodes[i].imm_i = 0; nodes[i].a = nodes[i].b = nodes[i].c = -1;        // the caller's memory is not referenced after create
            }
            if (!bytes.empty()) {
                literalPool.accounting = &stats.device_bytes;
                literalPool.reserve(bytes.size());
                PG_CUDA(cudaMemcpyAsync(literalPool.ptr, bytes.data(), bytes.size(), cudaMemcpyHostToDevice, stream));
                PG_CUDA(cudaStreamSynchronize(stream));
            }
        }
        for (DeviceBuffer* d : {&counts, &offsets, &scanTemp, &errorWord, &selection}) d->accounting = &stats.device_bytes;
        // sub-DAGs: which nodes the filter needs, which the projections need
        auto closure = [&](std::vector<int> roots) {
            uint64_t mask = 0;
            while (!roots.empty()) {
                int k = roots.back(); roots.pop_back();
                if (k < 0 || ((mask >> k) & 1ULL)) continue;
                mask |= 1ULL << k;
                if (nodes[k].op != PG_EX_INPUT) for (int o : {nodes[k].a, nodes[k].b, nodes[k].c}) roots.push_back(o);
            }
            return mask;
        };
        if (filterNode >= 0) needFilter = closure({filterNode});
        needProject = closure(projections);
        errorWord.reserve(16);
        PG_CUDA(cudaMemsetAsync(errorWord.ptr, 0, 16, stream));
    }

    bool needsInput() override { return !finishing && !hasPending; }

    // PTX for this program specialised on which inputs carry a null map; compiled once per variant, PG_FP_INTERPRETER=1 disables it
    JitProgram& getJit(const std::vector<DeviceColumn>& cols)
    {
        uint64_t key = 0;
        for (size_t i = 0; i < staged.size(); i++) if (cols[i].nulls) key |= 1ULL << staged[i];
        auto it = jitCache.find(key);
        if (it != jitCache.end()) return it->second;
        JitProgram& jp = jitCache[key];
        const char* off = getenv("PG_FP_INTERPRETER");
        DriverApi& api = DriverApi::get();
        if ((off && off[0] == '1') || !api.ok) { jp.log = "JIT disabled or libcuda not loadable"; return jp; }
        // The PTX lowering of `slice = / <> literal` (fp_codegen.cpp) assembles and is reviewed, but was written after the round's last GPU
        // run: it is opt-in (PG_FP_VARCHAR_JIT=1) until tests/test_gpu_page_processor.py::test_varchar_equals_literal_both_modes has run with
        // it on a B200.  Default: every program with a VARCHAR literal or comparison runs in the interpreter kernels (measured: Q3's
        // customer predicate over 20 M rows in 0.70 ms).
        const char* vj = getenv("PG_FP_VARCHAR_JIT");
        if ((varcharExpressions || !literalHost.empty()) && !(vj && vj[0] == '1')) { jp.log = "VARCHAR comparison / hash: interpreter kernels"; return jp; }
        FpProgramDesc d;
        d.nodes = nodes; d.inputTypes = inputTypes; d.filterNode = filterNode; d.projections = projections;
        d.literals = literalHost;      // slice = / <> literal is lowered to PTX; any other VARCHAR use makes the generator decline → interpreter kernels
        d.inputHasNulls.assign(inputTypes.size(), false);
        for (size_t i = 0; i < staged.size(); i++) d.inputHasNulls[staged[i]] = cols[i].nulls != nullptr;
        const std::string ptx = generateFilterProjectPtx(d, jp.sig);
        if (!jp.sig.ok) { jp.log = jp.sig.reason; return jp; }
        jp.module = getOrLoadPtxModule(options.device, ptx, &jp.log);      // shared by every operator of the same plan shape
        if (!jp.module) return jp;
        bool fine = api.moduleGetFunction(&jp.project, jp.module, "fp_project") == CUDA_SUCCESS;
        if (filterNode >= 0) fine = fine && api.moduleGetFunction(&jp.filter, jp.module, "fp_filter") == CUDA_SUCCESS;
        jp.ok = fine;
        if (!fine) jp.log = "generated module lacks an entry point";
        return jp;
    }

    void launchJit(CUfunction f, int grid, std::vector<uint64_t>& values)
    {
        std::vector<void*> args(values.size());
        for (size_t i = 0; i < values.size(); i++) args[i] = &values[i];
        CUresult r = DriverApi::get().launchKernel(f, (unsigned) grid, 1, 1, 256, 1, 1, 0, (CUstream) stream, args.data(), nullptr);
        if (r != CUDA_SUCCESS) throw Error(PG_ERR_CUDA, "cuLaunchKernel of a generated filter/project kernel failed (" + std::to_string((int) r) + ")");
        stats.kernel_launches++;
        stats.kernel_paths |= PG_PATH_FP_JIT;
        jitLaunches++;
    }
    void inputArgs(const JitProgram& jp, const FpParams& P, std::vector<uint64_t>& v) const
    {
        v.push_back((uint64_t) P.n); v.push_back((uint64_t) P.units);
        for (int ch : jp.sig.usedInputs) {
            v.push_back((uint64_t) (uintptr_t) P.inputs[ch].values);
            if (P.inputs[ch].type == PG_TYPE_VARCHAR) v.push_back((uint64_t) (uintptr_t) P.inputs[ch].offsets);
            if (P.inputs[ch].nulls) v.push_back((uint64_t) (uintptr_t) P.inputs[ch].nulls);
        }
    }

    void addInput(const pg_page* page) override
    {
        PG_CHECK_STATE(needsInput(), "filter/project does not need input (finishing or output pending)");
        const int64_t n = page->position_count;
        stats.input_positions += n;
        lastIsList = 0; lastSize = 0;
        if (n == 0) return;
        PG_CHECK_ARG(n < (int64_t(1) << 31), "page exceeds 2^31 positions");
        std::vector<DeviceColumn> cols;
        int slot = stager.stage(page, staged, cols);
        JitProgram& jit = getJit(cols);
        FpParams P{};
        P.nNodes = (int) nodes.size();
        for (size_t i = 0; i < nodes.size(); i++) {
            FpNode& d = P.nodes[i]; const pg_expr_node& s = nodes[i];
            d.op = s.op; d.type = s.type; d.a = s.a; d.b = s.b; d.c = s.c; d.isDouble = s.type == PG_TYPE_DOUBLE;
            if (s.op == PG_EX_CONST_DOUBLE) memcpy(&d.immI, &s.imm_d, 8); else d.immI = s.imm_i;
            if (s.op == PG_EX_CONST_VARCHAR) { d.immI = (int64_t) (uintptr_t) ((const uint8_t*) literalPool.ptr + literalOffset[i]); d.len = s.reserved; d.a = d.b = d.c = -1; }
        }
        std::vector<bool> nullable(nodes.size(), false);
        for (size_t i = 0; i < staged.size(); i++) {
            PG_CHECK_ARG(cols[i].type == inputTypes[staged[i]], "input channel type differs from the operator's configuration");
            P.inputs[staged[i]].values = cols[i].values; P.inputs[staged[i]].nulls = cols[i].nulls; P.inputs[staged[i]].type = cols[i].type; P.inputs[staged[i]].offsets = cols[i].offsets;
        }
        for (size_t i = 0; i < nodes.size(); i++) {
            const pg_expr_node& s = nodes[i];
            switch (s.op) {
                case PG_EX_INPUT: nullable[i] = P.inputs[s.a].nulls != nullptr; break;
                case PG_EX_CONST_NULL: nullable[i] = true; break;
                case PG_EX_CONST_BIGINT: case PG_EX_CONST_DOUBLE: case PG_EX_CONST_BOOLEAN: case PG_EX_CONST_VARCHAR: case PG_EX_IS_NULL: case PG_EX_HASH: case PG_EX_COMBINE_HASH: nullable[i] = false; break;
                default: nullable[i] = (s.a >= 0 && nullable[s.a]) || (s.b >= 0 && nullable[s.b]) || (s.c >= 0 && nullable[s.c]); break;
            }
        }
        P.filterNode = filterNode; P.n = n; P.units = divUp(n, kRowsPerFpUnit); P.error = errorWord.as<unsigned int>();
        P.needFilter = needFilter; P.needProject = needProject;
        if (filterNode >= 0) { selection.reserve((size_t) divUp(n, 32) * 4 + 16); P.selection = selection.as<uint32_t>(); }
        const int grid = (int) std::min<int64_t>(divUp(P.units, 8), (int64_t) numSMs * 8);
        int64_t total = n;
        if (filterNode >= 0) {
            counts.reserve((size_t) (P.units + 1) * 4); offsets.reserve((size_t) (P.units + 1) * 4);
            PG_CUDA(cudaMemsetAsync(counts.as<uint32_t>() + P.units, 0, 4, stream));
            P.counts = counts.as<uint32_t>();
            if (jit.ok) {
                std::vector<uint64_t> v;
                inputArgs(jit, P, v);
                v.push_back((uint64_t) (uintptr_t) P.selection); v.push_back((uint64_t) (uintptr_t) P.counts); v.push_back((uint64_t) (uintptr_t) P.error);
                launchJit(jit.filter, grid, v);
            }
            else { PG_LAUNCH(this, fpFilterKernel, grid, 256, 0, P); interpreterLaunches++; stats.kernel_paths |= PG_PATH_FP_INTERPRETER; }
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.as<uint32_t>(), offsets.as<uint32_t>(), (int) (P.units + 1), stream);
            scanTemp.reserve(tb);
            cub::DeviceScan::ExclusiveSum(scanTemp.ptr, tb, counts.as<uint32_t>(), offsets.as<uint32_t>(), (int) (P.units + 1), stream);
            stats.kernel_launches += 2;
            uint32_t t = 0;
            PG_CUDA(cudaMemcpyAsync(&t, offsets.as<uint32_t>() + P.units, 4, cudaMemcpyDeviceToHost, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
            stats.d2h_bytes += 4;
            total = t;
            P.offsets = offsets.as<uint32_t>();
        }
        // PageFilter.positionsArrayToSelectedPositions :27-50
        lastSize = total; lastIsList = (total == 0 || total == n) ? 0 : 1;
        checkError();
        if (total == 0) { stager.markCompute(slot); return; }
        const size_t nOut = projections.size();
        if (outValues.size() < nOut) { outValues.resize(nOut); outNulls.resize(nOut); for (auto& d : outValues) d.accounting = &stats.device_bytes; for (auto& d : outNulls) d.accounting = &stats.device_bytes; }
        outBlocks.assign(nOut, pg_block{});
        P.nOuts = (int) nOut;
        for (size_t j = 0; j < nOut; j++) {
            FpOutput& o = P.outs[j];
            o.node = projections[j]; o.width = typeWidth(nodes[o.node].type);
            outValues[j].reserve((size_t) total * o.width);
            o.values = outValues[j].ptr; o.nulls = nullptr;
            if (jit.ok ? (bool) jit.sig.projectionNullable[j] : (bool) nullable[o.node]) { outNulls[j].reserve(total); o.nulls = outNulls[j].as<uint8_t>(); }
        }
        if (nOut && jit.ok) {
            std::vector<uint64_t> v;
            inputArgs(jit, P, v);
            for (size_t j = 0; j < nOut; j++) { v.push_back((uint64_t) (uintptr_t) P.outs[j].values); if (jit.sig.projectionNullable[j]) v.push_back((uint64_t) (uintptr_t) P.outs[j].nulls); }
            v.push_back((uint64_t) (uintptr_t) P.selection); v.push_back((uint64_t) (uintptr_t) P.offsets); v.push_back((uint64_t) (uintptr_t) P.error);
            launchJit(jit.project, grid, v);
        }
        else if (nOut) { PG_LAUNCH(this, fpProjectKernel, grid, 256, 0, P); interpreterLaunches++; stats.kernel_paths |= PG_PATH_FP_INTERPRETER; }
        // VARCHAR pass-through channels: lengths of the selected rows → output offsets (scan) → bytes
        const size_t nVc = varcharChannels.size();
        std::vector<int32_t> vcTotalBytes(nVc, 0);
        std::vector<bool> vcHasNulls(nVc, false);
        if (nVc) {
            if (vcLens.size() < nVc) {
                vcLens.resize(nVc); vcOffsets.resize(nVc); vcBytes.resize(nVc); vcNulls.resize(nVc);
                for (auto* vec : {&vcLens, &vcOffsets, &vcBytes, &vcNulls}) for (auto& d : *vec) d.accounting = &stats.device_bytes;
            }
            const uint32_t* sel = filterNode >= 0 ? selection.as<uint32_t>() : nullptr;
            for (size_t v = 0; v < nVc; v++) {
                const DeviceColumn& vc = cols[std::find(staged.begin(), staged.end(), varcharChannels[v]) - staged.begin()];
                PG_CHECK_ARG(vc.offsets != nullptr, "VARCHAR block without offsets");
                vcLens[v].reserve((size_t) (total + 1) * 4); vcOffsets[v].reserve((size_t) (total + 1) * 4);
                PG_CUDA(cudaMemsetAsync(vcLens[v].as<int32_t>() + total, 0, 4, stream));
                uint8_t* on = nullptr;
                if (vc.nulls) { vcNulls[v].reserve((size_t) total); on = vcNulls[v].as<uint8_t>(); vcHasNulls[v] = true; }
                PG_LAUNCH(this, varcharSelectLensKernel, grid, 256, 0, sel, P.offsets, vc.offsets, vc.nulls, n, vcLens[v].as<int32_t>(), on);
                size_t tb = 0;
                cub::DeviceScan::ExclusiveSum(nullptr, tb, vcLens[v].as<int32_t>(), vcOffsets[v].as<int32_t>(), (int) (total + 1), stream);
                vcScanTemp.reserve(tb);
                cub::DeviceScan::ExclusiveSum(vcScanTemp.ptr, tb, vcLens[v].as<int32_t>(), vcOffsets[v].as<int32_t>(), (int) (total + 1), stream);
                stats.kernel_launches += 2;
                PG_CUDA(cudaMemcpyAsync(&vcTotalBytes[v], vcOffsets[v].as<int32_t>() + total, 4, cudaMemcpyDeviceToHost, stream));
            }
            PG_CUDA(cudaStreamSynchronize(stream));
            for (size_t v = 0; v < nVc; v++) {
                const DeviceColumn& vc = cols[std::find(staged.begin(), staged.end(), varcharChannels[v]) - staged.begin()];
                vcBytes[v].reserve((size_t) vcTotalBytes[v] + 16);
                PG_LAUNCH(this, varcharSelectCopyKernel, grid, 256, 0, sel, P.offsets, (const uint8_t*) vc.values, vc.offsets, vc.nulls, n, vcOffsets[v].as<int32_t>(), vcBytes[v].as<uint8_t>());
            }
        }
        stager.markCompute(slot);
        const bool toHost = options.output_memory == PG_MEM_HOST;
        if (toHost && hostValues.size() < nOut) { hostValues.resize(nOut); hostNulls.resize(nOut); }
        for (size_t j = 0; j < nOut; j++) {
            pg_block& b = outBlocks[j];
            b.type = nodes[projections[j]].type; b.encoding = PG_ENC_FLAT; b.position_count = total;
            const size_t vbytes = (size_t) total * P.outs[j].width;
            if (toHost) {
                hostValues[j].reserve(std::max<size_t>(vbytes, 16));
                PG_CUDA(cudaMemcpyAsync(hostValues[j].ptr, outValues[j].ptr, vbytes, cudaMemcpyDeviceToHost, stream));
                b.values = hostValues[j].ptr; b.is_null = nullptr;
                stats.d2h_bytes += (int64_t) vbytes;
                if (P.outs[j].nulls) { hostNulls[j].reserve(total); PG_CUDA(cudaMemcpyAsync(hostNulls[j].ptr, outNulls[j].ptr, total, cudaMemcpyDeviceToHost, stream)); b.is_null = hostNulls[j].as<uint8_t>(); stats.d2h_bytes += total; }
            }
            else { b.values = outValues[j].ptr; b.is_null = P.outs[j].nulls; }
            stats.output_bytes += (int64_t) vbytes;
        }
        // the public page: fixed-width and VARCHAR channels in the order the projections were given
        if (toHost && vcHostBytes.size() < nVc) { vcHostBytes.resize(nVc); vcHostOffsets.resize(nVc); vcHostNulls.resize(nVc); }
        publicBlocks.assign(outSlots.size(), pg_block{});
        for (size_t c = 0; c < outSlots.size(); c++) {
            if (!outSlots[c].varchar) { publicBlocks[c] = outBlocks[outSlots[c].index]; continue; }
            const int v = outSlots[c].index;
            pg_block& b = publicBlocks[c];
            b.type = PG_TYPE_VARCHAR; b.encoding = PG_ENC_FLAT; b.position_count = total;
            if (toHost) {
                vcHostBytes[v].reserve((size_t) vcTotalBytes[v] + 16); vcHostOffsets[v].reserve((size_t) (total + 1) * 4);
                if (vcTotalBytes[v]) PG_CUDA(cudaMemcpyAsync(vcHostBytes[v].ptr, vcBytes[v].ptr, (size_t) vcTotalBytes[v], cudaMemcpyDeviceToHost, stream));
                PG_CUDA(cudaMemcpyAsync(vcHostOffsets[v].ptr, vcOffsets[v].ptr, (size_t) (total + 1) * 4, cudaMemcpyDeviceToHost, stream));
                b.values = vcHostBytes[v].ptr; b.offsets = vcHostOffsets[v].as<int32_t>(); b.is_null = nullptr;
                if (vcHasNulls[v]) { vcHostNulls[v].reserve((size_t) total); PG_CUDA(cudaMemcpyAsync(vcHostNulls[v].ptr, vcNulls[v].ptr, (size_t) total, cudaMemcpyDeviceToHost, stream)); b.is_null = vcHostNulls[v].as<uint8_t>(); }
                stats.d2h_bytes += (int64_t) vcTotalBytes[v] + (total + 1) * 4;
            }
            else { b.values = vcBytes[v].ptr; b.offsets = vcOffsets[v].as<int32_t>(); b.is_null = vcHasNulls[v] ? vcNulls[v].as<uint8_t>() : nullptr; }
            stats.output_bytes += (int64_t) vcTotalBytes[v] + (total + 1) * 4;
        }
        PG_CUDA(cudaStreamSynchronize(stream));
        checkError();
        pendingRows = total;
        hasPending = true;
        stats.output_positions += total;
    }

    void checkError()
    {
        unsigned int e = 0;
        PG_CUDA(cudaMemcpyAsync(&e, errorWord.ptr, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        if (e) PG_CUDA(cudaMemsetAsync(errorWord.ptr, 0, 4, stream));
        if (e & FP_ERR_OVERFLOW) throw Error(PG_ERR_NUMERIC_OVERFLOW, "bigint arithmetic overflow");
        if (e & FP_ERR_DIV_ZERO) throw Error(PG_ERR_INVALID_ARGUMENT, "Division by zero");
    }

    ~FilterProjectOp() override
    {
        // (the modules belong to the process-wide cache)
    }
    // 1 = the last page ran through generated PTX, 0 = through the interpreter kernels
    int usedJit() const { return jitLaunches > 0 && interpreterLaunches == 0 ? 1 : 0; }

    bool getOutput(pg_page* out) override
    {
        if (!hasPending) return false;
        out->position_count = pendingRows; out->channel_count = (int32_t) publicBlocks.size(); out->memory = options.output_memory; out->blocks = publicBlocks.data();
        hasPending = false;
        return true;
    }
    void finish() override { finishing = true; }
    bool isFinished() override { return finishing && !hasPending; }
    void lastSelection(int32_t* isList, int64_t* size) const { *isList = lastIsList; *size = lastSize; }
};

Operator* createFilterProject(const pg_filter_project_config* c) { return new FilterProjectOp(c); }

}  // namespace pg

extern "C" {
int32_t pg_filter_project_create(const pg_filter_project_config* config, pg_operator** out)
{
    try { *out = (pg_operator*) new pg::FilterProjectOp(config); }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
int32_t pg_filter_project_explain(const pg_filter_project_config* c, uint64_t nullableInputs, char* buffer, size_t capacity, size_t* needed)
{
    try {
        PG_CHECK_ARG(c && needed, "null argument");
        PG_CHECK_ARG(c->n_input_channels >= 0 && c->n_input_channels <= pg::kMaxInputs && c->n_nodes >= 0 && c->n_nodes <= pg::kMaxNodes, "program too large");
        pg::FpProgramDesc d;
        d.nodes.assign(c->nodes, c->nodes + c->n_nodes);
        d.inputTypes.assign(c->input_types, c->input_types + c->n_input_channels);
        d.filterNode = c->filter_node;
        for (int j = 0; j < c->n_projections; j++)      // VARCHAR input references pass through outside the generated kernels
            if (c->nodes[c->projection_nodes[j]].type != PG_TYPE_VARCHAR) d.projections.push_back(c->projection_nodes[j]);
        for (int i = 0; i < c->n_nodes; i++)
            if (c->nodes[i].op == PG_EX_CONST_VARCHAR) {
                if (d.literals.empty()) d.literals.resize((size_t) c->n_nodes);
                d.literals[(size_t) i].assign((const char*) (uintptr_t) c->nodes[i].imm_i, (size_t) std::max(c->nodes[i].reserved, 0));
            }
        d.inputHasNulls.assign(d.inputTypes.size(), false);
        for (size_t i = 0; i < d.inputTypes.size(); i++) d.inputHasNulls[i] = (nullableInputs >> i) & 1ULL;
        pg::FpKernelSignature sig;
        const std::string ptx = pg::generateFilterProjectPtx(d, sig);
        if (!sig.ok) throw pg::Error(PG_ERR_UNSUPPORTED, "no PTX for this program (interpreter path): " + sig.reason);
        *needed = ptx.size() + 1;
        if (buffer && capacity >= ptx.size() + 1) memcpy(buffer, ptx.c_str(), ptx.size() + 1);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
int32_t pg_filter_project_get_last_selection(pg_operator* op, int32_t* isList, int64_t* size)
{
    try {
        auto* p = dynamic_cast<pg::FilterProjectOp*>((pg::Operator*) op);
        if (!p) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a filter/project operator");
        p->lastSelection(isList, size);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    return PG_OK;
}
}
