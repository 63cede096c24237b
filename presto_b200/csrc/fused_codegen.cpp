// presto_b200 — pipeline fragment → CUDA C++ (see fused_codegen.h).
//
// The generated translation unit defines `struct Prog` with
//   Tile                 the eagerly loaded columns of one tile, R rows per thread (registers; double buffered by the kernel)
//   load(P, base, t)     128-bit / 64-bit streaming loads of those columns
//   process(P, t, base, sink)   the row program, emitted stage by stage with the R rows of a thread unrolled INSIDE each stage:
//         A  validity, deferred-row bitmap, the filter                        (PageProcessor.java:118-148 / PageFilter.java:27-50)
//         B  per join: probe key, bitmap sector + directory word loads of all R rows, then ranks, then payload rows
//                                                                             (JoinProbe.java:74-119, PagesHash.java:169-181)
//         C  group key packing + aggregate inputs → sink.add(...)             (MultiChannelGroupByHash.putIfAbsent / AccumulatorCompiler loops)
//    so that the independent loads of all R rows of a stage are in flight together, and columns that only surviving rows need are
//    loaded after the filter / the probe (a 1 %-selective join never touches 99 % of its payload sectors).
// Arithmetic is the PageProcessor's: IEEE double through __dadd_rn / __dmul_rn … (never contracted), exact BIGINT with overflow /
// division-by-zero flags, SQL three-valued logic — the same semantics as fp_codegen.cpp's PTX and the interpreter kernels.
#include "fused_codegen.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace pg {
namespace {

struct Val {
    char kind = 'i';        // 'i' long long, 'd' double, 'b' bool, 's' VARCHAR slice (pass-through only)
    std::string v;          // C++ expression / variable holding the value
    std::string z;          // bool variable "is NULL"; empty = never NULL
    int channel = -1;       // 's': raw input channel
};

struct ChanRef { int kind = 0; int node = -1; int join = -1; int col = -1; };   // kind 0: expression node of the PageProcessor, 1: build column

std::string hex64(uint64_t v) { char b[40]; snprintf(b, sizeof b, "0x%016llxULL", (unsigned long long) v); return b; }
int widthOf(int t) { return (t == PG_TYPE_BIGINT || t == PG_TYPE_DOUBLE) ? 8 : (t == PG_TYPE_BOOLEAN ? 1 : (t == PG_TYPE_VARCHAR ? 0 : 4)); }

struct Gen {
    const FusedProgramDesc& D;
    const FpProgramDesc& P;
    int R = 2;
    std::vector<bool> eager;                 // per raw input channel
    std::ostringstream out;                  // current stage's code
    std::vector<std::vector<Val>> vals;      // [row slot][node]
    std::vector<std::vector<bool>> done;
    std::string fail;
    bool usesErr = false;

    explicit Gen(const FusedProgramDesc& d) : D(d), P(d.fp) {}
    void emit(const std::string& s) { out << "        " << s << "\n"; }
    static std::string sfx(int k, int u) { return std::to_string(k) + "_" + std::to_string(u); }
    std::string orz(const std::string& a, const std::string& b, const std::string& name)
    {
        if (a.empty() && b.empty()) return "";
        if (a.empty()) return b;
        if (b.empty()) return a;
        emit("const bool " + name + " = " + a + " || " + b + ";");
        return name;
    }
    std::string asDouble(const Val& x)
    {
        if (x.kind == 'd') return x.v;
        if (x.kind != 'i') { fail = "boolean operand in arithmetic"; return x.v; }
        return "__ll2double_rn(" + x.v + ")";
    }
    // raise an error flag for rows that are evaluated (guard) and whose operands are not NULL
    void raise(const char* flag, const std::string& cond, const std::string& z, const std::string& guard)
    {
        usesErr = true;
        emit(std::string(flag) + " |= " + guard + " && (" + cond + ")" + (z.empty() ? "" : " && !" + z) + ";");
    }

    Val compare(int op, const Val& a, const Val& b, const std::string& id)
    {
        Val r; r.kind = 'b'; r.v = "v" + id;
        r.z = orz(a.z, b.z, "z" + id);
        if (a.kind == 's' || b.kind == 's') { fail = "VARCHAR comparison"; return r; }
        if (a.kind == 'b' || b.kind == 'b') {
            if (a.kind != 'b' || b.kind != 'b' || (op != PG_EX_EQ && op != PG_EX_NE)) { fail = "boolean comparison"; return r; }
            emit("const bool " + r.v + " = " + a.v + (op == PG_EX_EQ ? " == " : " != ") + b.v + ";");
            return r;
        }
        const char* cc = op == PG_EX_EQ ? "==" : op == PG_EX_NE ? "!=" : op == PG_EX_LT ? "<" : op == PG_EX_LE ? "<=" : op == PG_EX_GT ? ">" : ">=";
        if (a.kind == 'd' || b.kind == 'd') emit("const bool " + r.v + " = " + asDouble(a) + " " + cc + " " + asDouble(b) + ";");
        else emit("const bool " + r.v + " = " + a.v + " " + cc + " " + b.v + ";");
        return r;
    }
    Val logic(bool isOr, const Val& a, const Val& b, const std::string& id)
    {
        Val r; r.kind = 'b'; r.v = "v" + id;
        if (a.kind != 'b' || b.kind != 'b') { fail = "AND/OR over non-boolean"; return r; }
        if (a.z.empty() && b.z.empty()) { emit("const bool " + r.v + " = " + a.v + (isOr ? " || " : " && ") + b.v + ";"); return r; }
        // decided: an operand that is not NULL and alone fixes the result (OR: true, AND: false)
        auto decided = [&](const Val& x) { return std::string("(") + (x.z.empty() ? "" : "!" + x.z + " && ") + (isOr ? "" : "!") + x.v + ")"; };
        emit("const bool y" + id + " = " + decided(a) + " || " + decided(b) + ";");
        r.z = "z" + id;
        emit("const bool " + r.z + " = (" + (a.z.empty() ? "false" : a.z) + " || " + (b.z.empty() ? "false" : b.z) + ") && !y" + id + ";");
        emit("const bool " + r.v + " = y" + id + " ? " + (isOr ? "true" : "false") + " : (" + a.v + (isOr ? " || " : " && ") + b.v + ");");
        return r;
    }

    // Evaluates node k for row slot u under `guard` (a bool variable that is true when the row is still alive).
    const Val& node(int k, int u, const std::string& guard)
    {
        if (done[u][k]) return vals[u][k];
        const pg_expr_node& nd = P.nodes[k];
        const std::string id = sfx(k, u);
        const std::string row = "row" + std::to_string(u);
        Val r;
        auto A = [&]() -> Val { return node(nd.a, u, guard); };
        auto B = [&]() -> Val { return node(nd.b, u, guard); };
        switch (nd.op) {
            case PG_EX_INPUT: {
                const int ch = nd.a, t = P.inputTypes[ch], w = widthOf(t);
                const std::string cs = std::to_string(ch), us = std::to_string(u);
                if (t == PG_TYPE_VARCHAR) { r.kind = 's'; r.channel = ch; r.v = ""; }
                else {
                    std::string raw;
                    if (eager[ch]) raw = "t.c" + cs + "[" + us + "]";
                    else {
                        raw = "w" + id;
                        const char* ty = w == 8 ? "unsigned long long" : (w == 4 ? "unsigned int" : "unsigned char");
                        const char* ld = w == 8 ? "ldStream64((const uint64_t*) P.values[" : (w == 4 ? "ldStream32((const uint32_t*) P.values[" : "((const uint8_t*) P.values[");
                        if (w == 1) emit(std::string("const ") + ty + " " + raw + " = " + guard + " ? " + ld + cs + "])[" + row + "] : (unsigned char) 0;");
                        else emit(std::string("const ") + ty + " " + raw + " = " + guard + " ? " + ld + cs + "] + " + row + ") : 0;");
                    }
                    r.v = "v" + id;
                    if (t == PG_TYPE_DOUBLE) { r.kind = 'd'; emit("const double " + r.v + " = __longlong_as_double((long long) " + raw + ");"); }
                    else if (t == PG_TYPE_BIGINT) { r.kind = 'i'; emit("const long long " + r.v + " = (long long) " + raw + ";"); }
                    else if (t == PG_TYPE_BOOLEAN) { r.kind = 'b'; emit("const bool " + r.v + " = " + raw + " != 0;"); }
                    else { r.kind = 'i'; emit("const long long " + r.v + " = (long long) (int) " + raw + ";"); }
                }
                if (P.inputHasNulls[ch]) { r.z = "z" + id; emit("const bool " + r.z + " = " + guard + " && P.nulls[" + cs + "][" + row + "] != 0;"); }
                break;
            }
            case PG_EX_CONST_BIGINT: r.kind = 'i'; r.v = "((long long) " + hex64((uint64_t) nd.imm_i) + ")"; break;
            case PG_EX_CONST_DOUBLE: { uint64_t b; memcpy(&b, &nd.imm_d, 8); r.kind = 'd'; r.v = "__longlong_as_double((long long) " + hex64(b) + ")"; break; }
            case PG_EX_CONST_BOOLEAN: r.kind = 'b'; r.v = nd.imm_i ? "true" : "false"; break;
            case PG_EX_CONST_NULL:
                r.z = "true";
                if (nd.type == PG_TYPE_BOOLEAN) { r.kind = 'b'; r.v = "false"; }
                else if (nd.type == PG_TYPE_DOUBLE) { r.kind = 'd'; r.v = "0.0"; }
                else { r.kind = 'i'; r.v = "0LL"; }
                break;
            case PG_EX_ADD: case PG_EX_SUB: case PG_EX_MUL: case PG_EX_DIV: {
                const Val a = A(), b = B();
                r.z = orz(a.z, b.z, "z" + id); r.v = "v" + id;
                if (nd.type == PG_TYPE_DOUBLE) {
                    r.kind = 'd';
                    const char* fn = nd.op == PG_EX_ADD ? "__dadd_rn" : nd.op == PG_EX_SUB ? "__dsub_rn" : nd.op == PG_EX_MUL ? "__dmul_rn" : "__ddiv_rn";
                    emit("const double " + r.v + " = " + fn + "(" + asDouble(a) + ", " + asDouble(b) + ");");
                }
                else {
                    if (a.kind != 'i' || b.kind != 'i') { fail = "integer arithmetic over non-integer operands"; break; }
                    r.kind = 'i';
                    if (nd.op == PG_EX_ADD) {
                        emit("const long long " + r.v + " = (long long) ((unsigned long long) " + a.v + " + (unsigned long long) " + b.v + ");");
                        raise("errOv", "((" + a.v + " ^ " + r.v + ") & (" + b.v + " ^ " + r.v + ")) < 0", r.z, guard);
                    }
                    else if (nd.op == PG_EX_SUB) {
                        emit("const long long " + r.v + " = (long long) ((unsigned long long) " + a.v + " - (unsigned long long) " + b.v + ");");
                        raise("errOv", "((" + a.v + " ^ " + b.v + ") & (" + a.v + " ^ " + r.v + ")) < 0", r.z, guard);
                    }
                    else if (nd.op == PG_EX_MUL) {
                        emit("const long long " + r.v + " = (long long) ((unsigned long long) " + a.v + " * (unsigned long long) " + b.v + ");");
                        raise("errOv", "__mul64hi(" + a.v + ", " + b.v + ") != (" + r.v + " >> 63)", r.z, guard);
                    }
                    else {
                        raise("errDz", b.v + " == 0", r.z, guard);
                        raise("errOv", a.v + " == LLONG_MIN && " + b.v + " == -1", r.z, guard);     // BigintOperators.divide :106-117
                        emit("const long long " + r.v + " = (" + b.v + " == 0 || (" + a.v + " == LLONG_MIN && " + b.v + " == -1)) ? 0 : " + a.v + " / " + b.v + ";");
                    }
                }
                break;
            }
            case PG_EX_NEG: {
                const Val a = A();
                r.z = a.z; r.kind = a.kind; r.v = "v" + id;
                if (a.kind == 'd') emit("const double " + r.v + " = __longlong_as_double(__double_as_longlong(" + a.v + ") ^ (long long) 0x8000000000000000ULL);");
                else if (a.kind == 'i') {
                    raise("errOv", a.v + " == LLONG_MIN", r.z, guard);
                    emit("const long long " + r.v + " = (long long) (0ULL - (unsigned long long) " + a.v + ");");
                }
                else fail = "negation of a non-number";
                break;
            }
            case PG_EX_EQ: case PG_EX_NE: case PG_EX_LT: case PG_EX_LE: case PG_EX_GT: case PG_EX_GE: { const Val a = A(), b = B(); r = compare(nd.op, a, b, id); break; }
            case PG_EX_BETWEEN: {
                const Val x = A(), lo = B(), hi = node(nd.c, u, guard);
                const Val c1 = compare(PG_EX_GE, x, lo, id + "a"), c2 = compare(PG_EX_LE, x, hi, id + "b");
                r = logic(false, c1, c2, id);
                break;
            }
            case PG_EX_AND: { const Val a = A(), b = B(); r = logic(false, a, b, id); break; }
            case PG_EX_OR: { const Val a = A(), b = B(); r = logic(true, a, b, id); break; }
            case PG_EX_NOT: {
                const Val a = A();
                if (a.kind != 'b') { fail = "NOT over non-boolean"; break; }
                r.kind = 'b'; r.z = a.z; r.v = "v" + id;
                emit("const bool " + r.v + " = !" + a.v + ";");
                break;
            }
            case PG_EX_IS_NULL: { const Val a = A(); r.kind = 'b'; r.v = a.z.empty() ? "false" : a.z; break; }
            case PG_EX_CAST_DOUBLE: {
                const Val a = A();
                r.kind = 'd'; r.z = a.z; r.v = "v" + id;
                emit("const double " + r.v + " = " + asDouble(a) + ";");
                break;
            }
            case PG_EX_HASH: {
                const Val a = A();
                r.kind = 'i'; r.v = "v" + id;
                std::string h;
                if (a.kind == 'b') h = "(" + a.v + " ? 1231LL : 1237LL)";
                else if (a.kind == 'd') h = "hashLong(doubleHashBits(__double_as_longlong(" + a.v + ")))";
                else if (a.kind == 'i') h = "hashLong(" + a.v + ")";
                else { fail = "hash of a VARCHAR expression"; break; }
                emit("const long long " + r.v + " = " + (a.z.empty() ? h : a.z + " ? 0LL : " + h) + ";");      // TypeUtils.hashPosition: NULL → 0
                break;
            }
            case PG_EX_COMBINE_HASH: {
                const Val a = A(), b = B();
                if (a.kind != 'i' || b.kind != 'i') { fail = "combine_hash over non-integers"; break; }
                r.kind = 'i'; r.v = "v" + id;
                emit("const long long " + r.v + " = (long long) (31ULL * (unsigned long long) " + a.v + " + (unsigned long long) " + b.v + ");");
                break;
            }
            default: fail = "expression op " + std::to_string(nd.op) + " cannot be fused"; break;
        }
        vals[u][k] = r; done[u][k] = true;
        return vals[u][k];
    }
};


// ---- what both sinks (aggregation, page) share: layouts, eager / lazy inputs, the tile loads, stages A (filter) and B (probes) ----
struct Frag {
    const FusedProgramDesc& D;
    const FpProgramDesc& P;
    int nNodes, nIn;
    std::vector<std::vector<ChanRef>> layouts;
    std::vector<bool> used, eager;
    std::vector<bool> eagerOffsets;          // VARCHAR channels whose offsets ride in the tile (the slices themselves are read by the row program)
    std::vector<int> usedInputs, eagerInputs;
    int R = 2;
    static constexpr int THREADS = 256;
    Gen g;
    std::string fail;

    explicit Frag(const FusedProgramDesc& d) : D(d), P(d.fp), nNodes((int) d.fp.nodes.size()), nIn((int) d.fp.inputTypes.size()), g(d) {}
    const std::vector<ChanRef>& fin() const { return layouts.back(); }

    bool buildLayouts()
    {
        layouts.assign(1, {});
        for (int p : P.projections) { ChanRef c; c.kind = 0; c.node = p; layouts[0].push_back(c); }
        for (size_t j = 0; j < D.joins.size(); j++) {
            const FusedJoinDesc& J = D.joins[j];
            const std::vector<ChanRef>& in = layouts.back();
            std::vector<ChanRef> next;
            if (J.keyChannel < 0 || J.keyChannel >= (int) in.size()) { fail = "join key channel out of range"; return false; }
            for (int c : J.probeOutputChannels) { if (c < 0 || c >= (int) in.size()) { fail = "probe output channel out of range"; return false; } next.push_back(in[c]); }
            for (size_t c = 0; c < J.buildOutputs.size(); c++) { ChanRef r; r.kind = 1; r.join = (int) j; r.col = (int) c; next.push_back(r); }
            layouts.push_back(next);
        }
        return true;
    }
    std::vector<bool> closure(std::vector<int> roots) const
    {
        std::vector<bool> seen(nNodes, false);
        while (!roots.empty()) {
            const int k = roots.back(); roots.pop_back();
            if (k < 0 || k >= nNodes || seen[k]) continue;
            seen[k] = true;
            if (P.nodes[k].op != PG_EX_INPUT) for (int o : {P.nodes[k].a, P.nodes[k].b, P.nodes[k].c}) roots.push_back(o);
        }
        return seen;
    }
    std::vector<bool> inputsOf(const std::vector<bool>& seen) const
    {
        std::vector<bool> in(nIn, false);
        for (int k = 0; k < nNodes; k++) if (seen[k] && P.nodes[k].op == PG_EX_INPUT) in[P.nodes[k].a] = true;
        return in;
    }
    // `finalChannels`: channels of the last layout the sink consumes.  Inputs of the filter and of the first probe key are loaded
    // eagerly (for every row, ahead of time); everything else after the filter / the probes, for surviving rows only — unless there is
    // no join and `allEagerWithoutJoins` (aggregation sink: almost every row reaches it).
    void planInputs(const std::vector<int>& finalChannels, bool allEagerWithoutJoins)
    {
        std::vector<int> allRoots;
        if (P.filterNode >= 0) allRoots.push_back(P.filterNode);
        auto rootOf = [&](const ChanRef& c) { if (c.kind == 0) allRoots.push_back(c.node); };
        for (size_t j = 0; j < D.joins.size(); j++) rootOf(layouts[j][D.joins[j].keyChannel]);
        for (int c : finalChannels) rootOf(fin()[c]);
        used = inputsOf(closure(allRoots));
        std::vector<int> eagerRoots;
        if (P.filterNode >= 0) eagerRoots.push_back(P.filterNode);
        if (!D.joins.empty()) { const ChanRef& c = layouts[0][D.joins[0].keyChannel]; if (c.kind == 0) eagerRoots.push_back(c.node); }
        else if (allEagerWithoutJoins) eagerRoots = allRoots;
        eager = inputsOf(closure(eagerRoots));
        if (P.filterNode < 0 && D.joins.empty() && allEagerWithoutJoins) eager = used;
        int eagerBytes = 0;
        for (int c = 0; c < nIn; c++) {
            if (P.inputTypes[c] == PG_TYPE_VARCHAR) eager[c] = false;       // slices are read where they are used
            if (eager[c]) eagerBytes += widthOf(P.inputTypes[c]);
        }
        // VARCHAR inputs of a fragment without joins (every row reaches the sink): their offsets are loaded with the tile, so that only
        // ONE dependent load (the bytes) is left when the row is processed
        eagerOffsets.assign(nIn, false);
        if (D.joins.empty() && allEagerWithoutJoins)
            for (int c = 0; c < nIn; c++) if (used[c] && P.inputTypes[c] == PG_TYPE_VARCHAR) eagerOffsets[c] = true;
        usedInputs.clear(); eagerInputs.clear();
        for (int c = 0; c < nIn; c++) if (used[c]) usedInputs.push_back(c);
        for (int c = 0; c < nIn; c++) if (eager[c]) eagerInputs.push_back(c);
        R = eagerBytes <= 16 ? 4 : 2;
        if (const char* e = getenv("PG_FUSED_ROWS")) { const int r = atoi(e); if (r == 2 || r == 4 || r == 8) R = r; }     // experiment knob: rows per thread per tile
        resetGen();
    }
    void resetGen()
    {
        g.R = R; g.eager = eager;
        g.vals.assign(R, std::vector<Val>(nNodes)); g.done.assign(R, std::vector<bool>(nNodes, false));
        g.out.str(""); g.out.clear();
    }
    void flush(std::ostringstream& body) { body << g.out.str(); g.out.str(""); g.out.clear(); }

    void emitTileAndLoad(std::ostringstream& src, const char* paramsType)
    {
        src << "    struct Tile {\n";
        for (int c : eagerInputs) {
            const int w = widthOf(P.inputTypes[c]);
            src << "        " << (w == 8 ? "unsigned long long" : w == 4 ? "unsigned int" : "unsigned char") << " c" << c << "[R];\n";
        }
        for (int c = 0; c < nIn; c++) if (eagerOffsets[c]) src << "        int o" << c << "[R / 2][3];\n";
        if (eagerInputs.empty()) src << "        int unused;\n";
        src << "    };\n";
        src << "    static __device__ __forceinline__ void load(const " << paramsType << "& P, long long base, Tile& t)\n    {\n";
        for (int p = 0; p < R / 2; p++) {
            src << "        {\n            const long long r0 = base + " << p * 2 * THREADS << " + 2 * (long long) threadIdx.x;\n"
                << "            if (r0 + 1 < P.n) {\n";
            for (int c : eagerInputs) {
                const int w = widthOf(P.inputTypes[c]);
                if (w == 8) src << "                { const ulonglong2 x = ldStream128((const uint64_t*) P.values[" << c << "] + r0); t.c" << c << "[" << 2 * p << "] = x.x; t.c" << c << "[" << 2 * p + 1 << "] = x.y; }\n";
                else if (w == 4) src << "                { const uint64_t x = ldStream64((const uint32_t*) P.values[" << c << "] + r0); t.c" << c << "[" << 2 * p << "] = (unsigned int) x; t.c" << c << "[" << 2 * p + 1 << "] = (unsigned int) (x >> 32); }\n";
                else src << "                { const unsigned short x = *((const unsigned short*) ((const uint8_t*) P.values[" << c << "] + r0)); t.c" << c << "[" << 2 * p << "] = (unsigned char) x; t.c" << c << "[" << 2 * p + 1 << "] = (unsigned char) (x >> 8); }\n";
            }
            for (int c = 0; c < nIn; c++) if (eagerOffsets[c])
                src << "                { const int* op = P.offsets[" << c << "] + r0; t.o" << c << "[" << p << "][0] = __ldg(op); t.o" << c << "[" << p << "][1] = __ldg(op + 1); t.o" << c << "[" << p << "][2] = __ldg(op + 2); }\n";
            src << "            }\n            else {\n";
            for (int c : eagerInputs) {
                const int w = widthOf(P.inputTypes[c]);
                const char* ty = w == 8 ? "uint64_t" : w == 4 ? "uint32_t" : "uint8_t";
                src << "                t.c" << c << "[" << 2 * p << "] = r0 < P.n ? ((const " << ty << "*) P.values[" << c << "])[r0] : 0; t.c" << c << "[" << 2 * p + 1 << "] = 0;\n";
            }
            for (int c = 0; c < nIn; c++) if (eagerOffsets[c])
                src << "                { const bool in = r0 < P.n; t.o" << c << "[" << p << "][0] = in ? P.offsets[" << c << "][r0] : 0; t.o" << c << "[" << p << "][1] = in ? P.offsets[" << c << "][r0 + 1] : 0; t.o" << c << "[" << p << "][2] = t.o" << c << "[" << p << "][1]; }\n";
            src << "            }\n        }\n";
        }
        src << "    }\n";
    }

    // The cp.async.bulk ring's view of a tile (fused_agg.cuh fusedAggBodyTma): every eager column's TILE rows back to back in one
    // shared-memory stage.  Returns the stage size in bytes.
    int emitTma(std::ostringstream& src, const char* paramsType)
    {
        const int TILE = THREADS * R;
        struct Part { int ch; int width; bool offsets; int at; int bytes; };
        std::vector<Part> parts;
        int at = 0;
        for (int c : eagerInputs) { const int w = widthOf(P.inputTypes[c]); parts.push_back({c, w, false, at, TILE * w}); at += (TILE * w + 15) & ~15; }
        for (int c = 0; c < nIn; c++) if (eagerOffsets[c]) { parts.push_back({c, 4, true, at, TILE * 4 + 16}); at += TILE * 4 + 16; }
        const int stageBytes = (at + 127) & ~127;
        src << "    static constexpr int STAGE_BYTES = " << stageBytes << ";\n";
        src << "    static __device__ __forceinline__ void issueStage(const " << paramsType << "& P, long long base, unsigned char* stage, uint64_t* bar)\n    {\n";
        int tx = 0;
        for (const Part& p : parts) {
            if (p.offsets) src << "        bulkCopyToShared(stage + " << p.at << ", (const char*) (P.offsets[" << p.ch << "] + base), " << p.bytes << ", bar);\n";
            else src << "        bulkCopyToShared(stage + " << p.at << ", (const char*) P.values[" << p.ch << "] + base * " << p.width << ", " << p.bytes << ", bar);\n";
            tx += p.bytes;
        }
        src << "    }\n    static constexpr int STAGE_TX_BYTES = " << tx << ";\n";
        src << "    static __device__ __forceinline__ void loadSmem(const unsigned char* stage, Tile& t)\n    {\n";
        for (int p = 0; p < R / 2; p++) {
            src << "        {\n            const int r = " << p * 2 * THREADS << " + 2 * (int) threadIdx.x;\n";
            for (const Part& q : parts) {
                if (q.offsets) src << "            { const int* op = (const int*) (stage + " << q.at << ") + r; const int2 a = *(const int2*) op; t.o" << q.ch << "[" << p << "][0] = a.x; t.o" << q.ch << "[" << p << "][1] = a.y; t.o" << q.ch << "[" << p << "][2] = op[2]; }\n";
                else if (q.width == 8) src << "            { const ulonglong2 x = *(const ulonglong2*) (stage + " << q.at << " + r * 8); t.c" << q.ch << "[" << 2 * p << "] = x.x; t.c" << q.ch << "[" << 2 * p + 1 << "] = x.y; }\n";
                else if (q.width == 4) src << "            { const uint2 x = *(const uint2*) (stage + " << q.at << " + r * 4); t.c" << q.ch << "[" << 2 * p << "] = x.x; t.c" << q.ch << "[" << 2 * p + 1 << "] = x.y; }\n";
                else src << "            { const unsigned short x = *(const unsigned short*) (stage + " << q.at << " + r); t.c" << q.ch << "[" << 2 * p << "] = (unsigned char) x; t.c" << q.ch << "[" << 2 * p + 1 << "] = (unsigned char) (x >> 8); }\n";
            }
            src << "        }\n";
        }
        src << "    }\n";
        return stageBytes;
    }

    // stage A: row index and validity of every row slot (+ the deferred-row bitmap of the aggregation sink)
    void stageValidity(std::ostringstream& body, bool withBitmapIn)
    {
        for (int u = 0; u < R; u++) {
            const std::string us = std::to_string(u);
            const int p = u / 2, q = u % 2;
            g.emit("const long long row" + us + " = base + " + std::to_string(p * 2 * THREADS + q) + " + 2 * (long long) threadIdx.x;");
            g.emit("bool live" + us + " = row" + us + " < P.n;");
            if (withBitmapIn) g.emit("if (P.bitmapIn && live" + us + ") live" + us + " = (P.bitmapIn[row" + us + " >> 5] >> (row" + us + " & 31)) & 1u;");
        }
        flush(body);
    }
    bool stageFilter(std::ostringstream& body)
    {
        if (P.filterNode < 0) return true;
        for (int u = 0; u < R; u++) {
            const std::string us = std::to_string(u), guard = "live" + us;
            // the filter's inputs are evaluated under the validity guard; a copy of the guard keeps `live` assignable
            g.emit("const bool in" + us + " = live" + us + ";");
            const Val f = g.node(P.filterNode, u, "in" + us);
            if (!g.fail.empty()) { fail = g.fail; return false; }
            if (f.kind != 'b') { fail = "filter is not boolean"; return false; }
            g.emit(guard + " = in" + us + " && " + f.v + (f.z.empty() ? "" : " && !" + f.z) + ";");      // only TRUE selects
        }
        flush(body);
        return true;
    }
    // a channel's value for row slot u (evaluated lazily under `guard`)
    Val chanVal(const ChanRef& c, int u, const std::string& guard)
    {
        if (c.kind == 0) return g.node(c.node, u, guard);
        const FusedBuildColumn& bc = D.joins[c.join].buildOutputs[c.col];
        const std::string us = std::to_string(u), js = std::to_string(c.join), id = "b" + js + "_" + std::to_string(c.col) + "_" + us;
        Val r;
        const std::string word = "pw" + js + "_" + us + "[" + std::to_string(bc.word) + "]";
        const std::string rawBits = bc.width == 8 ? word : "((" + word + " >> " + std::to_string(bc.shift) + ") & " + hex64(bc.width == 4 ? 0xffffffffULL : 0xffULL) + ")";
        r.v = "v" + id;
        if (bc.type == PG_TYPE_DOUBLE) { r.kind = 'd'; g.emit("const double " + r.v + " = __longlong_as_double((long long) " + rawBits + ");"); }
        else if (bc.type == PG_TYPE_BIGINT) { r.kind = 'i'; g.emit("const long long " + r.v + " = (long long) " + rawBits + ";"); }
        else if (bc.type == PG_TYPE_BOOLEAN) { r.kind = 'b'; g.emit("const bool " + r.v + " = " + rawBits + " != 0;"); }
        else if (bc.type == PG_TYPE_INTEGER || bc.type == PG_TYPE_DATE) { r.kind = 'i'; g.emit("const long long " + r.v + " = (long long) (int) (unsigned int) " + rawBits + ";"); }
        else { g.fail = "VARCHAR build column"; return r; }
        if (bc.nullWord >= 0) {
            r.z = "z" + id;
            g.emit("const bool " + r.z + " = ((pw" + js + "_" + us + "[" + std::to_string(bc.nullWord) + "] >> " + std::to_string(bc.nullShift) + ") & 0xff) != 0;");
        }
        return r;
    }
    // stage B: the probes.  withPayload = false: membership only (the selection pass of the page sink).  survivorsOnly: the rows are
    // known to match every join (emit pass): joins without build output are not probed again.
    bool stageJoins(std::ostringstream& body, bool withPayload, bool survivorsOnly = false)
    {
        for (size_t j = 0; j < D.joins.size(); j++) {
            const FusedJoinDesc& J = D.joins[j];
            if (survivorsOnly && J.strideWords == 0) continue;
            const std::string js = std::to_string(j);
            const ChanRef& keyRef = layouts[j][J.keyChannel];
            const bool membershipOnly = J.strideWords == 0 || !withPayload;
            // B1: keys, one bitmap word per row of all R rows
            for (int u = 0; u < R; u++) {
                const std::string us = std::to_string(u);
                g.emit("const bool in" + js + "_" + us + " = live" + us + ";");
                const Val k = chanVal(keyRef, u, "in" + js + "_" + us);
                if (!g.fail.empty()) { fail = g.fail; return false; }
                if (k.kind != 'i') { fail = "join key must be an integer channel"; return false; }
                g.emit("const unsigned long long ix" + js + "_" + us + " = (unsigned long long) (" + k.v + " - P.joins[" + js + "].D.minKey);");
                g.emit("const bool ir" + js + "_" + us + " = in" + js + "_" + us + (k.z.empty() ? "" : " && !" + k.z) + " && ix" + js + "_" + us + " < P.joins[" + js + "].D.range;");
                // one 8-byte bitmap word decides membership (L2 resident: range / 8 bytes); the rest of the 32-byte block, the rank directory
                // and the payload row are only touched by the rows that matched
                g.emit("const unsigned long long bw" + js + "_" + us + " = ir" + js + "_" + us + " ? __ldg(P.joins[" + js + "].D.bitmap + (ix" + js + "_" + us + " >> 6)) : 0ULL;");
            }
            flush(body);
            // B2: membership narrows `live`; matched rows fetch their rank (block popcounts + directory word) and payload row
            for (int u = 0; u < R; u++) {
                const std::string us = std::to_string(u), s = js + "_" + us;
                g.emit("const unsigned long long bt" + s + " = 1ULL << (ix" + s + " & 63);");
                g.emit("live" + us + " = ir" + s + " && (bw" + s + " & bt" + s + ") != 0;");
                if (!membershipOnly) {
                    std::string init = "unsigned long long pw" + s + "[" + std::to_string(J.strideWords) + "] = {";
                    for (int k = 0; k < J.strideWords; k++) init += k ? ", 0" : "0";
                    g.emit(init + "};");
                    std::string code = "if (live" + us + ") { const unsigned long long* bp = P.joins[" + js + "].D.bitmap + (ix" + s + " >> 8) * 4; const unsigned int wi = (unsigned int) (ix" + s + " >> 6) & 3u; "
                                       "unsigned int rk = __ldg(P.joins[" + js + "].D.prefix + (ix" + s + " >> 8)) + __popcll(bw" + s + " & (bt" + s + " - 1)); "
                                       "if (wi > 0) rk += __popcll(__ldg(bp)); if (wi > 1) rk += __popcll(__ldg(bp + 1)); if (wi > 2) rk += __popcll(__ldg(bp + 2)); "
                                       "const unsigned long long* pr = P.joins[" + js + "].rows + (unsigned long long) rk * " + std::to_string(J.strideWords) + "; ";
                    for (int k = 0; k < J.strideWords; k++) code += "pw" + s + "[" + std::to_string(k) + "] = __ldg(pr + " + std::to_string(k) + "); ";
                    g.emit(code + "}");
                }
            }
            flush(body);
        }
        return true;
    }
};

}  // namespace

std::string generateFusedAggSource(const FusedProgramDesc& D, FusedKernelInfo& info)
{
    info = FusedKernelInfo();
    const FpProgramDesc& P = D.fp;
    const int nNodes = (int) P.nodes.size();
    auto decline = [&](const std::string& why) { info.ok = false; info.reason = why; return std::string(); };
    if (D.accs.size() > 8) return decline("more than 8 sum columns");
    if (D.accs.empty() && !D.countRows) return decline("no aggregate");
    if (D.joins.size() > 2) return decline("more than two joins");
    if (D.W < 1 || D.W > 4) return decline("key width");

    Frag F(D);
    if (!F.buildLayouts()) return decline(F.fail);
    const std::vector<ChanRef>& fin = F.fin();
    std::vector<int> finalChannels;
    for (const auto& k : D.keys) { if (k.channel < 0 || k.channel >= (int) fin.size()) return decline("group-by channel out of range"); finalChannels.push_back(k.channel); }
    for (const auto& a : D.accs) { if (a.channel < 0 || a.channel >= (int) fin.size()) return decline("aggregate channel out of range"); finalChannels.push_back(a.channel); }
    F.planInputs(finalChannels, true);
    info.usedInputs = F.usedInputs; info.eagerInputs = F.eagerInputs;
    const int R = F.R;
    info.rowsPerThread = R; info.nv = (int) D.accs.size(); info.cnt = D.countRows;
    const int NV = (int) D.accs.size();
    const int THREADS = Frag::THREADS;
    Gen& g = F.g;
    const std::vector<bool>& eager = F.eager;

    std::ostringstream src;
    src << "// generated by presto_b200 fused_codegen: FilterAndProject -> LookupJoin* -> HashAggregation as one row program\n"
        << "#include \"fused_agg.cuh\"\nnamespace pg {\nstruct Prog {\n"
        // fragments with joins are latency bound on the probes' L2 loads: PG_FUSED_MIN_CTAS trades registers for resident warps
        << "    static constexpr int THREADS = " << THREADS << ", MIN_CTAS = " << (D.joins.empty() ? 2 : (getenv("PG_FUSED_MIN_CTAS") ? atoi(getenv("PG_FUSED_MIN_CTAS")) : 4)) << ", R = " << R << ", W = " << D.W << ", NV = " << NV << ";\n"
        << "    static constexpr bool CNT = " << (D.countRows ? "true" : "false") << ";\n"
        << "    static constexpr int NEAGER = " << info.eagerInputs.size() << ";\n";
    // ---- the 64-bit word the on-chip tiers identify a group by ----
    {
        bool keyNullable = false;     // a dry evaluation of the key channels tells whether any of them can be NULL on this page
        Gen dry(D); dry.R = 1; dry.eager = eager; dry.vals.assign(1, std::vector<Val>(nNodes)); dry.done.assign(1, std::vector<bool>(nNodes, false));
        for (const auto& k : D.keys) {
            const ChanRef& c = fin[k.channel];
            if (c.kind == 0) { if (!dry.node(c.node, 0, "g").z.empty()) keyNullable = true; }
            else if (D.joins[c.join].buildOutputs[c.col].nullWord >= 0) keyNullable = true;
        }
        int totalBits = 0;
        for (const auto& k : D.keys) totalBits += k.bits;
        const bool squeeze = D.W > 1 && !D.bigintPath && !keyNullable && totalBits <= 64;
        info.tiers = D.W == 1 || squeeze;
        info.queue = !D.joins.empty();
        info.queueBytes = info.queue ? (size_t) (8 * D.W + 8 * (D.accs.empty() ? 1 : (int) D.accs.size()) + 8) * 64 * (256 / 32) : 0;     // fused_agg.cuh QueueEntry x kQueueSlots x warps
        src << "    static constexpr bool QUEUE = " << (info.queue ? "true" : "false") << ";      // survivors of the probes are aggregated 32 at a time\n";
        src << "    static constexpr bool TIERKEY_CAN_BE_EMPTY = " << ((squeeze && totalBits == 64) ? "true" : "false") << ";\n";
        src << "    static constexpr bool TIERS = " << ((D.W == 1 || squeeze) ? "true" : "false") << ";\n";
        src << "    static __device__ __forceinline__ unsigned long long tierKey(const unsigned long long (&kw)[W])\n    {\n";
        if (D.W == 1 || !squeeze) src << "        return kw[0];\n";
        else {
            src << "        unsigned long long t = 0;\n";
            int off = 0;
            for (const auto& k : D.keys) {
                src << "        t |= ((kw[" << k.word << "] >> " << k.shift << ")" << (k.bits < 64 ? " & " + hex64((1ULL << k.bits) - 1) : std::string()) << ") << " << off << ";\n";
                off += k.bits;
            }
            src << "        return t;\n";
        }
        src << "    }\n    static __device__ __forceinline__ void expandKey(unsigned long long t, unsigned long long (&kw)[W])\n    {\n";
        for (int w = 0; w < D.W; w++) src << "        kw[" << w << "] = 0;\n";
        if (D.W == 1 || !squeeze) src << "        kw[0] = t;\n";
        else {
            int off = 0;
            for (const auto& k : D.keys) {
                src << "        kw[" << k.word << "] |= ((t >> " << off << ")" << (k.bits < 64 ? " & " + hex64((1ULL << k.bits) - 1) : std::string()) << ") << " << k.shift << ";\n";
                off += k.bits;
            }
        }
        src << "    }\n";
    }
    F.emitTileAndLoad(src, "FusedParams");
    info.tma = !F.eagerInputs.empty();
    if (info.tma) info.tmaStageBytes = F.emitTma(src, "FusedParams");

    // ---- process ----
    std::ostringstream body;
    auto flush = [&]() { F.flush(body); };
    auto chanVal = [&](const ChanRef& c, int u, const std::string& guard) { return F.chanVal(c, u, guard); };
    body << "        bool errOv = false, errDz = false;\n";
    F.stageValidity(body, true);
    if (!F.stageFilter(body)) return decline(F.fail);
    if (!F.stageJoins(body, true)) return decline(F.fail);
    // stage C: aggregate inputs and key fields (C1), key packing + hand-over to the aggregation (C2).
    //   without joins: C1 of all R rows first (their lazy loads in flight together), then C2 under `if (go)`;
    //   with joins (survivor queue): per row slot, everything inside `if (any lane of the warp still holds a row)` — after a selective
    //   probe most warps skip the whole stage.
    std::vector<std::vector<Val>> keyVals(R), accVals(R);
    auto emitC1 = [&](int u) -> bool {
        const std::string us = std::to_string(u);
        for (const auto& a : D.accs) {
            Val v = chanVal(fin[a.channel], u, "go" + us);
            if (!g.fail.empty()) return false;
            if (!v.z.empty()) { g.fail = "nullable aggregate input"; return false; }
            if (a.fromBigint ? v.kind != 'i' : v.kind != 'd') { g.fail = "aggregate input type"; return false; }
            accVals[u].push_back(v);
        }
        for (const auto& k : D.keys) {
            Val v = chanVal(fin[k.channel], u, "go" + us);
            if (!g.fail.empty()) return false;
            if (v.kind == 's') {
                // VARCHAR(n <= 7) key: [length byte][bytes] exactly as group_table.cuh packKey packs it
                const std::string id = "k" + std::to_string(&k - &D.keys[0]) + "_" + us, cs = std::to_string(v.channel);
                g.emit("unsigned long long s" + id + " = 0;");
                const std::string offs = F.eagerOffsets[v.channel]
                    ? "const int o = t.o" + cs + "[" + std::to_string(u / 2) + "][" + std::to_string(u % 2) + "]; int len = t.o" + cs + "[" + std::to_string(u / 2) + "][" + std::to_string(u % 2 + 1) + "] - o; "
                    : "const int o = P.offsets[" + cs + "][row" + us + "]; int len = P.offsets[" + cs + "][row" + us + " + 1] - o; ";
                g.emit("if (go" + us + (v.z.empty() ? "" : " && !" + v.z) + ") { " + offs +
                       "if (len > " + std::to_string(k.maxLen) + ") { errLen = true; len = " + std::to_string(k.maxLen) + "; } "
                       "const uint8_t* b = (const uint8_t*) P.values[" + cs + "] + o; s" + id + " = (unsigned long long) (unsigned int) len; "
                       "for (int i = 0; i < len && i < 7; i++) s" + id + " |= (unsigned long long) b[i] << (8 * (i + 1)); }");
                v.v = "s" + id;
            }
            keyVals[u].push_back(v);
        }
        return true;
    };
    auto emitC2 = [&](int u, std::ostringstream& body) -> bool {
        const std::string us = std::to_string(u);
        body << (info.queue ? "        {\n" : "        if (go" + us + ") {\n");
        body << "            unsigned long long kw[" << D.W << "]; int special = 0; unsigned int spell = 0;\n";
        for (int w = 0; w < D.W; w++) body << "            kw[" << w << "] = 0;\n";
        if (D.bigintPath) {
            const Val& v = keyVals[u][0];
            if (v.kind != 'i') { g.fail = "BIGINT key expected"; return false; }
            body << "            kw[0] = (unsigned long long) " << v.v << ";\n";
            if (!v.z.empty()) body << "            if (" << v.z << ") { special = 1; kw[0] = 0; } else\n";
            body << "            if (kw[0] == kEmptyKey) special = 2;\n";
        }
        else {
            body << "            unsigned int nullMask = 0;\n";
            for (size_t k = 0; k < D.keys.size(); k++) {
                const FusedKeyField& f = D.keys[k];
                const Val& v = keyVals[u][k];
                std::string bits;
                switch (f.type) {
                    case PG_TYPE_BIGINT: if (v.kind != 'i') { g.fail = "key type"; return false; } bits = "(unsigned long long) " + v.v; break;
                    case PG_TYPE_INTEGER: case PG_TYPE_DATE: if (v.kind != 'i') { g.fail = "key type"; return false; } bits = "(unsigned long long) (unsigned int) " + v.v; break;
                    case PG_TYPE_BOOLEAN: if (v.kind != 'b') { g.fail = "key type"; return false; } bits = "(" + v.v + " ? 1ULL : 0ULL)"; break;
                    case PG_TYPE_VARCHAR: if (v.kind != 's') { g.fail = "VARCHAR keys must be input columns passed through"; return false; } bits = v.v; break;
                    case PG_TYPE_DOUBLE: {
                        if (v.kind != 'd') { g.fail = "key type"; return false; }
                        body << "            unsigned long long d" << k << " = (unsigned long long) canonicalDoubleBits(__double_as_longlong(" << v.v << "));\n"
                             << "            if (d" << k << " == 0x8000000000000000ULL) { d" << k << " = 0; spell |= " << (1u << k) << "u; }\n";
                        bits = "d" + std::to_string(k);
                        break;
                    }
                    default: g.fail = "group-by key type"; return false;
                }
                if (!v.z.empty()) body << "            if (" << v.z << ") nullMask |= " << (1u << k) << "u; else\n";
                body << "            kw[" << f.word << "] |= (" << bits << ") << " << f.shift << ";\n";
            }
            body << "            kw[" << D.nullWord << "] |= (unsigned long long) nullMask << " << D.nullShift << ";\n";
            if (D.W == 1) body << "            if (kw[0] == kEmptyKey) special = 2;\n";
        }
        body << "            unsigned long long vals[" << (NV > 0 ? NV : 1) << "];\n";
        if (NV == 0) body << "            vals[0] = 0;\n";
        for (int c = 0; c < NV; c++) {
            const Val& v = accVals[u][c];
            if (D.accs[c].fromBigint) body << "            vals[" << c << "] = (unsigned long long) __double_as_longlong(__ll2double_rn(" << v.v << "));\n";
            else body << "            vals[" << c << "] = (unsigned long long) __double_as_longlong(" << v.v << ");\n";
        }
        if (info.queue) body << "            sink.push(m" << us << ", go" << us << ", kw, special, spell, vals, (unsigned int) row" << us << ");\n        }\n";
        else body << "            sink.add(kw, special, spell, vals, (unsigned int) row" << us << ");\n        }\n";
        return true;
    };
    if (!info.queue) {
        for (int u = 0; u < R; u++) { g.emit("const bool go" + std::to_string(u) + " = live" + std::to_string(u) + ";"); if (!emitC1(u)) return decline(g.fail); }
        flush();
        for (int u = 0; u < R; u++) if (!emitC2(u, body)) return decline(g.fail);
    }
    else {
        for (int u = 0; u < R; u++) {
            const std::string us = std::to_string(u);
            body << "        const bool go" << us << " = live" << us << ";\n        const unsigned int m" << us << " = __ballot_sync(0xffffffffu, go" << us << ");\n        if (m" << us << ") {\n";
            if (!emitC1(u)) return decline(g.fail);
            flush();
            if (!emitC2(u, body)) return decline(g.fail);
            body << "        }\n";
        }
    }
    if (!g.fail.empty()) return decline(g.fail);

    src << "    template <class Sink>\n    static __device__ __forceinline__ void process(const FusedParams& P, const Tile& t, long long base, Sink& sink)\n    {\n"
        << "        bool errLen = false;\n" << body.str()
        << "        if (errOv) atomicOr(P.fpError, 1u);\n        if (errDz) atomicOr(P.fpError, 2u);\n        if (errLen) atomicOr(P.fpError, 4u);\n"
        << "    }\n};\n}  // namespace pg\n"
        << "extern \"C\" __global__ void __launch_bounds__(pg::Prog::THREADS, pg::Prog::MIN_CTAS) fused_kernel(const __grid_constant__ pg::FusedParams P) { pg::fusedAggBody<pg::Prog>(P); }\n";
    if (info.tma)
        src << "extern \"C\" __global__ void __launch_bounds__(pg::Prog::THREADS + 32, pg::Prog::MIN_CTAS) fused_kernel_tma(const __grid_constant__ pg::FusedParams P) { pg::fusedAggBodyTma<pg::Prog>(P); }\n";
    info.ok = true;
    return src.str();
}

}  // namespace pg

namespace pg {

std::string generateFusedPageSource(const FusedProgramDesc& D, FusedPageInfo& info)
{
    info = FusedPageInfo();
    const FpProgramDesc& P = D.fp;
    auto decline = [&](const std::string& why) { info.ok = false; info.reason = why; return std::string(); };
    if (D.joins.size() > 2) return decline("more than two joins");
    const int THREADS = Frag::THREADS;

    // ---- selection pass: filter + membership probes ----
    Frag A(D);
    if (!A.buildLayouts()) return decline(A.fail);
    const size_t nOut = A.fin().size();
    if (nOut == 0 || nOut > 32) return decline("output channel count");
    std::vector<int> finalChannels;
    for (size_t c = 0; c < nOut; c++) finalChannels.push_back((int) c);
    A.planInputs(finalChannels, false);
    info.usedInputs = A.usedInputs; info.eagerInputs = A.eagerInputs; info.rowsPerThread = A.R;
    const int R = A.R;
    std::ostringstream src;
    src << "// generated by presto_b200 fused_codegen: FilterAndProject -> LookupJoin* materialised as a page (selection pass + emit pass)\n"
        << "#include \"fused_agg.cuh\"\nnamespace pg {\nstruct Prog {\n"
        << "    static constexpr int THREADS = " << THREADS << ", MIN_CTAS = 2, R = " << R << ";\n";
    A.emitTileAndLoad(src, "FusedPageParams");
    {
        std::ostringstream body;
        body << "        bool errOv = false, errDz = false;\n        const int lane = threadIdx.x & 31;\n";
        A.stageValidity(body, false);
        if (!A.stageFilter(body)) return decline(A.fail);
        if (!A.stageJoins(body, false)) return decline(A.fail);
        for (int p = 0; p < R / 2; p++) {
            body << "        {\n            const unsigned int m0 = __ballot_sync(0xffffffffu, live" << 2 * p << "), m1 = __ballot_sync(0xffffffffu, live" << 2 * p + 1 << ");\n"
                 << "            const long long unit = (base + " << p * 2 * THREADS << " + 64 * (long long) (threadIdx.x >> 5)) >> 6;\n"
                 << "            if (lane == 0 && unit * 64 < P.n) { P.masks[2 * unit] = m0; P.masks[2 * unit + 1] = m1; P.counts[unit] = __popc(m0) + __popc(m1); }\n        }\n";
        }
        src << "    static __device__ __forceinline__ void select(const FusedPageParams& P, const Tile& t, long long base)\n    {\n" << body.str()
            << "        if (errOv) atomicOr(P.fpError, 1u);\n        if (errDz) atomicOr(P.fpError, 2u);\n    }\n";
    }
    // ---- emit pass: survivors only; nothing is loaded ahead of time ----
    Frag B(D);
    B.buildLayouts();
    B.planInputs(finalChannels, false);
    B.R = R;
    std::fill(B.eager.begin(), B.eager.end(), false);
    B.resetGen();
    {
        std::ostringstream body;
        body << "        bool errOv = false, errDz = false;\n        const int lane = threadIdx.x & 31;\n        const unsigned int lt = (1u << lane) - 1;\n";
        for (int p = 0; p < R / 2; p++) {
            const std::string e = std::to_string(2 * p), o = std::to_string(2 * p + 1), ps = std::to_string(p);
            body << "        const long long unit" << ps << " = (base + " << p * 2 * THREADS << " + 64 * (long long) (threadIdx.x >> 5)) >> 6;\n"
                 << "        unsigned int ma" << ps << " = 0, mb" << ps << " = 0, off" << ps << " = 0;\n"
                 << "        if (unit" << ps << " * 64 < P.n) { const uint2 mm = *(const uint2*) (P.masks + 2 * unit" << ps << "); ma" << ps << " = mm.x; mb" << ps << " = mm.y; off" << ps << " = P.unitOffsets[unit" << ps << "]; }\n"
                 << "        const long long row" << e << " = base + " << p * 2 * THREADS << " + 2 * (long long) threadIdx.x, row" << o << " = row" << e << " + 1;\n"
                 << "        bool live" << e << " = (ma" << ps << " >> lane) & 1u, live" << o << " = (mb" << ps << " >> lane) & 1u;\n"
                 << "        const unsigned int dest" << e << " = off" << ps << " + __popc(ma" << ps << " & lt) + __popc(mb" << ps << " & lt), dest" << o << " = dest" << e << " + (live" << e << " ? 1u : 0u);\n";
        }
        // everything below runs under `if (any survivor in this warp's units)`: most warps of a selective fragment skip it
        std::string any;
        for (int p = 0; p < R / 2; p++) any += (p ? " | " : "") + std::string("ma") + std::to_string(p) + " | mb" + std::to_string(p);
        body << "        if ((" << any << ") == 0) return;\n";
        if (!B.stageJoins(body, true, true)) return decline(B.fail);
        info.outputTypes.assign(nOut, 0); info.outputNullable.assign(nOut, false);
        std::vector<std::vector<Val>> outVals(R);
        for (int u = 0; u < R; u++) {
            const std::string us = std::to_string(u);
            B.g.emit("const bool go" + us + " = live" + us + ";");
            for (size_t c = 0; c < nOut; c++) {
                const ChanRef& ref = B.fin()[c];
                const int type = ref.kind == 0 ? P.nodes[ref.node].type : D.joins[ref.join].buildOutputs[ref.col].type;
                if (type == PG_TYPE_VARCHAR) return decline("VARCHAR output channel");
                const Val v = B.chanVal(ref, u, "go" + us);
                if (!B.g.fail.empty()) return decline(B.g.fail);
                const int w = widthOf(type);
                if ((w == 1) != (v.kind == 'b') || (type == PG_TYPE_DOUBLE) != (v.kind == 'd')) return decline("output channel type / representation mismatch");
                info.outputTypes[c] = type;
                if (!v.z.empty()) info.outputNullable[c] = true;
                outVals[u].push_back(v);
                const std::string cs = std::to_string(c);
                std::string bits = v.kind == 'd' ? "(unsigned long long) __double_as_longlong(" + v.v + ")" : v.kind == 'b' ? "(" + v.v + " ? 1u : 0u)" : "(unsigned long long) " + v.v;
                if (!v.z.empty()) bits = "(" + v.z + " ? 0 : " + bits + ")";
                const char* ty = w == 8 ? "unsigned long long" : w == 4 ? "unsigned int" : "unsigned char";
                B.g.emit("if (go" + us + ") ((" + ty + "*) P.outValues[" + cs + "])[dest" + us + "] = (" + ty + ") " + bits + ";");
            }
        }
        // nullability is a per-channel property: a channel that is nullable for one row slot is nullable for all (same expressions)
        for (int u = 0; u < R; u++) {
            const std::string us = std::to_string(u);
            for (size_t c = 0; c < nOut; c++) {
                if (!info.outputNullable[c]) continue;
                const Val& v = outVals[u][c];
                B.g.emit("if (go" + us + ") P.outNulls[" + std::to_string(c) + "][dest" + us + "] = " + (v.z.empty() ? std::string("0") : "(" + v.z + " ? 1 : 0)") + ";");
            }
        }
        B.flush(body);
        src << "    static __device__ __forceinline__ void emit(const FusedPageParams& P, long long base)\n    {\n" << body.str()
            << "        if (errOv) atomicOr(P.fpError, 1u);\n        if (errDz) atomicOr(P.fpError, 2u);\n    }\n";
    }
    src << "};\n}  // namespace pg\n"
        << "extern \"C\" __global__ void __launch_bounds__(pg::Prog::THREADS, pg::Prog::MIN_CTAS) fused_select(const __grid_constant__ pg::FusedPageParams P) { pg::fusedSelectBody<pg::Prog>(P); }\n"
        << "extern \"C\" __global__ void __launch_bounds__(pg::Prog::THREADS, pg::Prog::MIN_CTAS) fused_emit(const __grid_constant__ pg::FusedPageParams P) { pg::fusedEmitBody<pg::Prog>(P); }\n";
    info.ok = true;
    return src.str();
}

}  // namespace pg
