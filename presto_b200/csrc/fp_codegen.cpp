// presto_b200 — RowExpression → PTX (see fp_codegen.h).
#include "fp_codegen.h"
#include <cstdio>
#include <cstring>
#include <sstream>

namespace pg {
namespace {

struct NodeVal {
    char kind = 'i';        // 'i' int64 in a .b64 register, 'd' double bits in a .b64 register, 'b' boolean in a predicate
    std::string val;        // register holding the value
    std::string nul;        // predicate "is NULL", empty = never NULL
    // VARCHAR: kind 's' = a slice of an input channel (val: .b64 address of its first byte, aux: .b32 length); kind 'c' = the literal of
    // node `lit` (bytes known to the generator).  Only `slice = / <> literal` is lowered (VarcharOperators.equal / notEqual :54-65).
    std::string aux;
    int lit = -1;
};

struct Gen {
    const FpProgramDesc& P;
    std::ostringstream body;
    int n64 = 0, n32 = 0, np = 0;
    std::vector<NodeVal> vals;
    std::vector<bool> done;
    std::string fail;
    // environment of the row being generated
    std::string row;        // .b64 register: row index
    std::string guard;      // predicate: this lane's row is to be evaluated (loads are predicated on it)
    std::string errOverflow, errDivZero;   // accumulated error predicates
    std::vector<std::string> inValues, inNulls, inOffsets;   // per input channel: registers holding the global pointers ("" = not a parameter)

    explicit Gen(const FpProgramDesc& p) : P(p) {}
    std::string r64() { return "%rd" + std::to_string(n64++); }
    std::string r32() { return "%r" + std::to_string(n32++); }
    std::string pr() { return "%p" + std::to_string(np++); }
    void emit(const std::string& s) { body << "    " << s << "\n"; }
    static std::string hex64(uint64_t v) { char b[32]; snprintf(b, sizeof b, "0x%016llx", (unsigned long long) v); return b; }

    std::string predFalse() { std::string p = pr(); emit("setp.ne.u32 " + p + ", %lane, %lane;"); return p; }
    std::string predTrue() { std::string p = pr(); emit("setp.eq.u32 " + p + ", %lane, %lane;"); return p; }
    std::string notp(const std::string& x) { std::string p = pr(); emit("not.pred " + p + ", " + x + ";"); return p; }
    std::string orNull(const std::string& a, const std::string& b)
    {
        if (a.empty()) return b;
        if (b.empty()) return a;
        std::string p = pr(); emit("or.pred " + p + ", " + a + ", " + b + ";"); return p;
    }
    std::string asDouble(const NodeVal& v)
    {
        if (v.kind == 'd') return v.val;
        if (v.kind != 'i') { fail = "boolean operand in arithmetic"; return v.val; }
        std::string t = r64(); emit("cvt.rn.f64.s64 " + t + ", " + v.val + ";"); return t;
    }
    // error predicate `cond` counts only for rows that are evaluated and whose operands are not NULL
    void raise(std::string& acc, const std::string& cond, const std::string& nul)
    {
        std::string p = pr();
        emit("and.pred " + p + ", " + cond + ", " + guard + ";");
        if (!nul.empty()) emit("and.pred " + p + ", " + p + ", " + notp(nul) + ";");
        emit("or.pred " + acc + ", " + acc + ", " + p + ";");
    }

    // slice = literal: equal lengths, then byte by byte (loads only where the lengths agree: nothing is read past the slice)
    NodeVal compareSliceLiteral(int op, const NodeVal& s, const NodeVal& c)
    {
        NodeVal r; r.kind = 'b'; r.val = pr(); r.nul = s.nul;
        if (op != PG_EX_EQ && op != PG_EX_NE) { fail = "VARCHAR ordering comparison"; return r; }
        const std::string& bytes = P.literals[(size_t) c.lit];
        std::string pl = pr(), pld = pr();
        emit("setp.eq.s32 " + pl + ", " + s.aux + ", " + std::to_string(bytes.size()) + ";");
        emit("and.pred " + pld + ", " + pl + ", " + guard + ";");
        emit("mov.pred " + r.val + ", " + pl + ";");
        for (size_t i = 0; i < bytes.size(); i++) {
            std::string x = r32(), pb = pr();
            emit("mov.b32 " + x + ", 0;");
            emit("@" + pld + " ld.global.nc.u8 " + x + ", [" + s.val + "+" + std::to_string(i) + "];");
            emit("setp.eq.u32 " + pb + ", " + x + ", " + std::to_string((unsigned) (unsigned char) bytes[i]) + ";");
            emit("and.pred " + r.val + ", " + r.val + ", " + pb + ";");
        }
        if (op == PG_EX_NE) emit("not.pred " + r.val + ", " + r.val + ";");
        return r;
    }

    NodeVal compare(int op, const NodeVal& a, const NodeVal& b)
    {
        if (a.kind == 's' && b.kind == 'c') return compareSliceLiteral(op, a, b);
        if (a.kind == 'c' && b.kind == 's') return compareSliceLiteral(op, b, a);      // = and <> are symmetric
        if (a.kind == 's' || a.kind == 'c' || b.kind == 's' || b.kind == 'c') { NodeVal r; r.kind = 'b'; r.val = pr(); fail = "VARCHAR comparison other than slice = literal"; return r; }
        NodeVal r; r.kind = 'b'; r.val = pr(); r.nul = orNull(a.nul, b.nul);
        if (a.kind == 'b' || b.kind == 'b') {
            if (a.kind != 'b' || b.kind != 'b' || (op != PG_EX_EQ && op != PG_EX_NE)) { fail = "boolean comparison"; return r; }
            emit("xor.pred " + r.val + ", " + a.val + ", " + b.val + ";");
            if (op == PG_EX_EQ) emit("not.pred " + r.val + ", " + r.val + ";");
            return r;
        }
        const bool dbl = a.kind == 'd' || b.kind == 'd';
        const char* cc = op == PG_EX_EQ ? "eq" : op == PG_EX_NE ? (dbl ? "neu" : "ne") : op == PG_EX_LT ? "lt" : op == PG_EX_LE ? "le" : op == PG_EX_GT ? "gt" : "ge";
        if (dbl) { std::string x = asDouble(a), y = asDouble(b); emit(std::string("setp.") + cc + ".f64 " + r.val + ", " + x + ", " + y + ";"); }
        else emit(std::string("setp.") + cc + ".s64 " + r.val + ", " + a.val + ", " + b.val + ";");
        return r;
    }
    // SQL three-valued AND (isOr = false) / OR (isOr = true)
    NodeVal logic(bool isOr, const NodeVal& a, const NodeVal& b)
    {
        NodeVal r; r.kind = 'b'; r.val = pr();
        if (a.kind != 'b' || b.kind != 'b') { fail = "AND/OR over non-boolean"; return r; }
        emit(std::string(isOr ? "or" : "and") + ".pred " + r.val + ", " + a.val + ", " + b.val + ";");
        if (a.nul.empty() && b.nul.empty()) return r;
        // decided[x] = operand x alone decides the result: not NULL and (OR: true / AND: false)
        auto decided = [&](const NodeVal& x) {
            std::string p = pr();
            if (isOr) emit("mov.pred " + p + ", " + x.val + ";"); else emit("not.pred " + p + ", " + x.val + ";");
            if (!x.nul.empty()) emit("and.pred " + p + ", " + p + ", " + notp(x.nul) + ";");
            return p;
        };
        std::string da = decided(a), db = decided(b), any = pr(), nn = orNull(a.nul, b.nul);
        emit("or.pred " + any + ", " + da + ", " + db + ";");
        r.nul = pr();
        const std::string notAny = notp(any);
        emit("and.pred " + r.nul + ", " + nn + ", " + notAny + ";");
        // value when decided: OR → true, AND → false; value bits of NULL operands are ignored through `any`
        std::string v2 = pr();
        if (isOr) emit("or.pred " + v2 + ", " + r.val + ", " + any + ";"); else emit("and.pred " + v2 + ", " + r.val + ", " + notAny + ";");
        r.val = v2;
        return r;
    }

    NodeVal hashOf(const NodeVal& a, int type)
    {
        NodeVal r; r.kind = 'i'; r.val = r64();
        if (a.kind == 'b') emit("selp.b64 " + r.val + ", 1231, 1237, " + a.val + ";");
        else {
            std::string v = a.val;
            if (type == PG_TYPE_DOUBLE) {   // TypeUtils.doubleHashCode: -0.0 → +0.0, then Double.doubleToLongBits (canonical NaN)
                std::string p = pr(), c = r64(), z = pr(), c2 = r64();
                emit("setp.nan.f64 " + p + ", " + a.val + ", " + a.val + ";");
                emit("selp.b64 " + c + ", 0x7ff8000000000000, " + a.val + ", " + p + ";");
                emit("setp.eq.b64 " + z + ", " + c + ", 0x8000000000000000;");
                emit("selp.b64 " + c2 + ", 0, " + c + ", " + z + ";");
                v = c2;
            }
            std::string t = r64(), x = r64(), y = r64();
            emit("mul.lo.u64 " + t + ", " + v + ", 0xC2B2AE3D27D4EB4F;");
            emit("shl.b64 " + x + ", " + t + ", 31;");
            emit("shr.u64 " + y + ", " + t + ", 33;");
            emit("or.b64 " + t + ", " + x + ", " + y + ";");
            emit("mul.lo.u64 " + r.val + ", " + t + ", 0x9E3779B185EBCA87;");
        }
        if (!a.nul.empty()) emit("selp.b64 " + r.val + ", 0, " + r.val + ", " + a.nul + ";");   // TypeUtils.hashPosition: NULL → 0
        return r;
    }

    const NodeVal& node(int k)
    {
        if (done[k]) return vals[k];
        const pg_expr_node& nd = P.nodes[k];
        NodeVal r;
        auto A = [&]() -> const NodeVal& { return node(nd.a); };
        auto B = [&]() -> const NodeVal& { return node(nd.b); };
        switch (nd.op) {
            case PG_EX_INPUT: {
                const int ch = nd.a, t = P.inputTypes[ch];
                if (t == PG_TYPE_VARCHAR) {       // offsets[row], offsets[row + 1] → address of the slice, its length
                    std::string ao = r64(), o0 = r32(), o1 = r32(), o64 = r64();
                    emit("mad.lo.u64 " + ao + ", " + row + ", 4, " + inOffsets[ch] + ";");
                    emit("mov.b32 " + o0 + ", 0;");
                    emit("mov.b32 " + o1 + ", 0;");
                    emit("@" + guard + " ld.global.nc.b32 " + o0 + ", [" + ao + "];");
                    emit("@" + guard + " ld.global.nc.b32 " + o1 + ", [" + ao + "+4];");
                    r.kind = 's'; r.aux = r32(); r.val = r64();
                    emit("sub.s32 " + r.aux + ", " + o1 + ", " + o0 + ";");
                    emit("cvt.u64.u32 " + o64 + ", " + o0 + ";");
                    emit("add.u64 " + r.val + ", " + inValues[ch] + ", " + o64 + ";");
                    if (P.inputHasNulls[ch]) {
                        std::string an = r64(), x = r32(); r.nul = pr();
                        emit("add.u64 " + an + ", " + inNulls[ch] + ", " + row + ";");
                        emit("mov.b32 " + x + ", 0;");
                        emit("@" + guard + " ld.global.nc.u8 " + x + ", [" + an + "];");
                        emit("setp.ne.u32 " + r.nul + ", " + x + ", 0;");
                    }
                    break;
                }
                const int w = (t == PG_TYPE_BIGINT || t == PG_TYPE_DOUBLE) ? 8 : (t == PG_TYPE_BOOLEAN ? 1 : 4);
                std::string addr = r64();
                emit("mad.lo.u64 " + addr + ", " + row + ", " + std::to_string(w) + ", " + inValues[ch] + ";");
                if (w == 8) {
                    r.kind = t == PG_TYPE_DOUBLE ? 'd' : 'i'; r.val = r64();
                    emit("mov.b64 " + r.val + ", 0;");
                    emit("@" + guard + " ld.global.nc.L1::no_allocate.b64 " + r.val + ", [" + addr + "];");
                }
                else if (w == 4) {
                    std::string x = r32(); r.kind = 'i'; r.val = r64();
                    emit("mov.b32 " + x + ", 0;");
                    emit("@" + guard + " ld.global.nc.L1::no_allocate.b32 " + x + ", [" + addr + "];");
                    emit("cvt.s64.s32 " + r.val + ", " + x + ";");
                }
                else {
                    std::string x = r32(); r.kind = 'b'; r.val = pr();
                    emit("mov.b32 " + x + ", 0;");
                    emit("@" + guard + " ld.global.nc.u8 " + x + ", [" + addr + "];");
                    emit("setp.ne.u32 " + r.val + ", " + x + ", 0;");
                }
                if (P.inputHasNulls[ch]) {
                    std::string an = r64(), x = r32(); r.nul = pr();
                    emit("add.u64 " + an + ", " + inNulls[ch] + ", " + row + ";");
                    emit("mov.b32 " + x + ", 0;");
                    emit("@" + guard + " ld.global.nc.u8 " + x + ", [" + an + "];");
                    emit("setp.ne.u32 " + r.nul + ", " + x + ", 0;");
                }
                break;
            }
            case PG_EX_CONST_BIGINT: r.kind = 'i'; r.val = r64(); emit("mov.b64 " + r.val + ", " + hex64((uint64_t) nd.imm_i) + ";"); break;
            case PG_EX_CONST_DOUBLE: { uint64_t b; memcpy(&b, &nd.imm_d, 8); r.kind = 'd'; r.val = r64(); emit("mov.b64 " + r.val + ", " + hex64(b) + ";"); break; }
            case PG_EX_CONST_BOOLEAN: r.kind = 'b'; r.val = nd.imm_i ? predTrue() : predFalse(); break;
            case PG_EX_CONST_VARCHAR:
                if ((size_t) k >= P.literals.size() || P.literals[(size_t) k].size() > 32) { fail = "VARCHAR literal longer than 32 bytes"; break; }
                r.kind = 'c'; r.lit = k;
                break;
            case PG_EX_CONST_NULL:
                r.nul = predTrue();
                if (nd.type == PG_TYPE_BOOLEAN) { r.kind = 'b'; r.val = predFalse(); }
                else { r.kind = nd.type == PG_TYPE_DOUBLE ? 'd' : 'i'; r.val = r64(); emit("mov.b64 " + r.val + ", 0;"); }
                break;
            case PG_EX_ADD: case PG_EX_SUB: case PG_EX_MUL: case PG_EX_DIV: {
                const NodeVal a = A(), b = B();
                r.nul = orNull(a.nul, b.nul); r.val = r64();
                if (nd.type == PG_TYPE_DOUBLE) {
                    r.kind = 'd';
                    std::string x = asDouble(a), y = asDouble(b);
                    const char* op = nd.op == PG_EX_ADD ? "add" : nd.op == PG_EX_SUB ? "sub" : nd.op == PG_EX_MUL ? "mul" : "div";
                    emit(std::string(op) + ".rn.f64 " + r.val + ", " + x + ", " + y + ";");   // explicit .rn: never contracted into an FMA
                }
                else {
                    if (a.kind != 'i' || b.kind != 'i') { fail = "integer arithmetic over non-integer operands"; break; }
                    r.kind = 'i';
                    if (nd.op == PG_EX_ADD || nd.op == PG_EX_SUB) {
                        std::string t1 = r64(), t2 = r64(), p = pr();
                        emit(std::string(nd.op == PG_EX_ADD ? "add" : "sub") + ".s64 " + r.val + ", " + a.val + ", " + b.val + ";");
                        emit("xor.b64 " + t1 + ", " + a.val + ", " + r.val + ";");
                        emit("xor.b64 " + t2 + ", " + (nd.op == PG_EX_ADD ? b.val + ", " + r.val : a.val + ", " + b.val) + ";");
                        emit("and.b64 " + t1 + ", " + t1 + ", " + t2 + ";");
                        emit("setp.lt.s64 " + p + ", " + t1 + ", 0;");
                        raise(errOverflow, p, r.nul);
                    }
                    else if (nd.op == PG_EX_MUL) {
                        std::string hi = r64(), sg = r64(), p = pr();
                        emit("mul.lo.s64 " + r.val + ", " + a.val + ", " + b.val + ";");
                        emit("mul.hi.s64 " + hi + ", " + a.val + ", " + b.val + ";");
                        emit("shr.s64 " + sg + ", " + r.val + ", 63;");
                        emit("setp.ne.s64 " + p + ", " + hi + ", " + sg + ";");
                        raise(errOverflow, p, r.nul);
                    }
                    else {
                        std::string pz = pr(), pm = pr(), pm2 = pr(), d = r64();
                        emit("setp.eq.s64 " + pz + ", " + b.val + ", 0;");
                        raise(errDivZero, pz, r.nul);
                        emit("setp.eq.b64 " + pm + ", " + a.val + ", 0x8000000000000000;");
                        emit("setp.eq.s64 " + pm2 + ", " + b.val + ", -1;");
                        emit("and.pred " + pm + ", " + pm + ", " + pm2 + ";");
                        raise(errOverflow, pm, r.nul);                                    // BigintOperators.divide :106-117: MIN / -1 → NUMERIC_VALUE_OUT_OF_RANGE
                        emit("or.pred " + pm + ", " + pm + ", " + pz + ";");
                        emit("selp.b64 " + d + ", 1, " + b.val + ", " + pm + ";");     // x / 0 and MIN / -1 are flagged; the division itself must not trap
                        emit("div.s64 " + r.val + ", " + a.val + ", " + d + ";");
                        emit("selp.b64 " + r.val + ", 0, " + r.val + ", " + pz + ";");
                    }
                }
                break;
            }
            case PG_EX_NEG: {
                const NodeVal a = A();
                r.nul = a.nul; r.val = r64(); r.kind = a.kind;
                if (a.kind == 'd') emit("xor.b64 " + r.val + ", " + a.val + ", 0x8000000000000000;");
                else if (a.kind == 'i') {
                    std::string p = pr();
                    emit("setp.eq.b64 " + p + ", " + a.val + ", 0x8000000000000000;");
                    raise(errOverflow, p, r.nul);
                    emit("neg.s64 " + r.val + ", " + a.val + ";");
                }
                else fail = "negation of a boolean";
                break;
            }
            case PG_EX_EQ: case PG_EX_NE: case PG_EX_LT: case PG_EX_LE: case PG_EX_GT: case PG_EX_GE: r = compare(nd.op, A(), B()); break;
            case PG_EX_BETWEEN: {
                const NodeVal x = A(), lo = B(), hi = node(nd.c);
                r = logic(false, compare(PG_EX_GE, x, lo), compare(PG_EX_LE, x, hi));
                break;
            }
            case PG_EX_AND: r = logic(false, A(), B()); break;
            case PG_EX_OR: r = logic(true, A(), B()); break;
            case PG_EX_NOT: {
                const NodeVal a = A();
                if (a.kind != 'b') { fail = "NOT over non-boolean"; break; }
                r.kind = 'b'; r.val = pr(); r.nul = a.nul;
                emit("not.pred " + r.val + ", " + a.val + ";");
                break;
            }
            case PG_EX_IS_NULL: { const NodeVal a = A(); if (a.kind == 'c') { fail = "IS NULL of a literal"; break; } r.kind = 'b'; r.val = a.nul.empty() ? predFalse() : a.nul; break; }
            case PG_EX_CAST_DOUBLE: { const NodeVal a = A(); r.kind = 'd'; r.nul = a.nul; r.val = asDouble(a); break; }
            case PG_EX_HASH: if (P.nodes[nd.a].type == PG_TYPE_VARCHAR) { fail = "hash of a VARCHAR value"; break; } r = hashOf(A(), P.nodes[nd.a].type); break;
            case PG_EX_COMBINE_HASH: {
                const NodeVal a = A(), b = B();
                if (a.kind != 'i' || b.kind != 'i') { fail = "combine_hash over non-integers"; break; }
                r.kind = 'i'; r.val = r64();
                emit("mad.lo.u64 " + r.val + ", " + a.val + ", 31, " + b.val + ";");
                break;
            }
            default: fail = "expression op " + std::to_string(nd.op) + " has no PTX lowering"; break;
        }
        vals[k] = r; done[k] = true;
        return vals[k];
    }
};

std::string prologue(Gen& g, const std::vector<int>& usedInputs, bool filterKernel)
{
    // parameters → registers; common index arithmetic
    std::ostringstream s;
    auto ldp = [&](const std::string& reg, const std::string& param, bool pointer) {
        s << "    ld.param.u64 " << reg << ", [" << param << "];\n";
        if (pointer) s << "    cvta.to.global.u64 " << reg << ", " << reg << ";\n";
    };
    ldp("%n", "p_n", false); ldp("%units", "p_units", false);
    for (int ch : usedInputs) {
        ldp(g.inValues[ch], "p_in" + std::to_string(ch) + "_v", true);
        if (!g.inOffsets[ch].empty()) ldp(g.inOffsets[ch], "p_in" + std::to_string(ch) + "_o", true);
        if (!g.inNulls[ch].empty()) ldp(g.inNulls[ch], "p_in" + std::to_string(ch) + "_n", true);
    }
    (void) filterKernel;
    s << "    mov.u32 %xtid, %tid.x;\n    mov.u32 %xntid, %ntid.x;\n    mov.u32 %xctaid, %ctaid.x;\n    mov.u32 %xnctaid, %nctaid.x;\n"
      << "    mad.lo.u32 %gtid, %xctaid, %xntid, %xtid;\n    shr.u32 %warp, %gtid, 5;\n    and.b32 %lane, %xtid, 31;\n"
      << "    mul.lo.u32 %nthr, %xnctaid, %xntid;\n    shr.u32 %nwarps, %nthr, 5;\n"
      << "    cvt.u64.u32 %unit, %warp;\n    cvt.u64.u32 %nwarps64, %nwarps;\n    cvt.u64.u32 %lane64, %lane;\n"
      << "    mov.u32 %lt, %lanemask_lt;\n    setp.eq.u32 %plane0, %lane, 0;\n"
      << "    setp.ne.u32 %perr_ov, %lane, %lane;\n    setp.ne.u32 %perr_dz, %lane, %lane;\n";
    return s.str();
}

// Rounds of 32 rows are generated kJitUnroll times per loop iteration (= one whole 256-row unit) with separate registers, as one
// straight-line block: the driver's scheduler then issues the loads of all rounds before their uses (memory-level parallelism of
// kJitUnroll x columns per lane).  Measured on Q1's 7-column program: 1 round 55 %, 4 rounds 62.6 %, 8 rounds 71.7 % of the HBM roofline.
constexpr int kJitUnroll = 8;
std::string fixedRegs()
{
    std::ostringstream s;
    s << "    .reg .b64 %n, %units, %unit, %nwarps64, %lane64, %begin, %end, %base, %sel, %off, %cnt, %errp, %w64, %a64;\n"
      << "    .reg .b32 %xtid, %xntid, %xctaid, %xnctaid, %gtid, %warp, %lane, %nthr, %nwarps, %lt, %cur, %t0, %t1, %count;\n"
      << "    .reg .pred %plane0, %perr_ov, %perr_dz, %pdone, %pdr;\n"
      << "    .reg .b64 %row<" << kJitUnroll << ">, %dest64_<" << kJitUnroll << ">;\n"
      << "    .reg .b32 %mask<" << kJitUnroll << ">;\n"
      << "    .reg .pred %pin<" << kJitUnroll << ">, %pact<" << kJitUnroll << ">, %pbit<" << kJitUnroll << ">, %psel<" << kJitUnroll << ">, %prd<" << kJitUnroll << ">;\n";
    return s.str();
}

}  // namespace

std::string generateFilterProjectPtx(const FpProgramDesc& P, FpKernelSignature& sig)
{
    sig.ok = false;
    const int nNodes = (int) P.nodes.size();
    // which inputs are referenced at all
    std::vector<bool> used(P.inputTypes.size(), false);
    {   // inputs reachable from the filter or a projection (VARCHAR input references that only pass through are not among them)
        std::vector<int> roots(P.projections.begin(), P.projections.end());
        if (P.filterNode >= 0) roots.push_back(P.filterNode);
        std::vector<bool> seen(nNodes, false);
        while (!roots.empty()) {
            const int k = roots.back(); roots.pop_back();
            if (k < 0 || k >= nNodes || seen[k]) continue;
            seen[k] = true;
            if (P.nodes[k].op == PG_EX_INPUT) used[P.nodes[k].a] = true;
            else for (int o : {P.nodes[k].a, P.nodes[k].b, P.nodes[k].c}) roots.push_back(o);
        }
    }
    sig.usedInputs.clear();
    for (size_t c = 0; c < used.size(); c++) if (used[c]) sig.usedInputs.push_back((int) c);
    // VARCHAR inputs: lowered only when every use is `slice = literal` / `slice <> literal` (the generator fails on anything else below);
    // programs that merely pass VARCHAR channels through never list them here
    for (int ch : sig.usedInputs) if (P.inputTypes[ch] == PG_TYPE_VARCHAR && P.literals.empty()) { sig.reason = "VARCHAR input"; return ""; }
    for (int k = 0; k < nNodes; k++) {
        const pg_expr_node& nd = P.nodes[k];
        const bool leaf = nd.op == PG_EX_INPUT || (nd.op >= PG_EX_CONST_BIGINT && nd.op <= PG_EX_CONST_VARCHAR);
        if (leaf || nd.op == PG_EX_EQ || nd.op == PG_EX_NE || nd.op == PG_EX_IS_NULL) continue;
        for (int o : {nd.a, nd.b, nd.c}) if (o >= 0 && o < nNodes && P.nodes[o].type == PG_TYPE_VARCHAR) { sig.reason = "VARCHAR operand outside = / <>"; return ""; }
    }

    std::ostringstream ptx;
    ptx << "//\n// generated by presto_b200 fp_codegen: RowExpression to PTX\n//\n.version 8.7\n.target sm_100a\n.address_size 64\n\n";

    auto paramList = [&](Gen& g, std::ostringstream& s, bool filterKernel) {
        s << "    .param .u64 p_n,\n    .param .u64 p_units";
        for (int ch : sig.usedInputs) {
            s << ",\n    .param .u64 p_in" << ch << "_v";
            if (P.inputTypes[ch] == PG_TYPE_VARCHAR) s << ",\n    .param .u64 p_in" << ch << "_o";
            if (P.inputHasNulls[ch]) s << ",\n    .param .u64 p_in" << ch << "_n";
        }
        if (filterKernel) s << ",\n    .param .u64 p_sel,\n    .param .u64 p_cnt,\n    .param .u64 p_err";
        else {
            for (size_t j = 0; j < P.projections.size(); j++) {
                s << ",\n    .param .u64 p_out" << j << "_v";
                if (sig.projectionNullable[j]) s << ",\n    .param .u64 p_out" << j << "_n";
            }
            s << ",\n    .param .u64 p_sel,\n    .param .u64 p_off,\n    .param .u64 p_err";
        }
        (void) g;
    };
    auto newGen = [&](Gen& g) {
        g.vals.assign(nNodes, NodeVal()); g.done.assign(nNodes, false);
        g.inValues.assign(P.inputTypes.size(), ""); g.inNulls.assign(P.inputTypes.size(), ""); g.inOffsets.assign(P.inputTypes.size(), "");
        for (int ch : sig.usedInputs) {
            g.inValues[ch] = "%in" + std::to_string(ch) + "v";
            if (P.inputTypes[ch] == PG_TYPE_VARCHAR) g.inOffsets[ch] = "%in" + std::to_string(ch) + "o";
            if (P.inputHasNulls[ch]) g.inNulls[ch] = "%in" + std::to_string(ch) + "n";
        }
        g.row = "%row0"; g.errOverflow = "%perr_ov"; g.errDivZero = "%perr_dz";
    };
    auto inputRegDecl = [&]() {
        std::ostringstream s;
        for (int ch : sig.usedInputs) {
            s << "    .reg .b64 %in" << ch << "v;\n";
            if (P.inputTypes[ch] == PG_TYPE_VARCHAR) s << "    .reg .b64 %in" << ch << "o;\n";
            if (P.inputHasNulls[ch]) s << "    .reg .b64 %in" << ch << "n;\n";
        }
        return s.str();
    };
    auto regDecl = [&](const Gen& g) {
        std::ostringstream s;
        s << fixedRegs() << inputRegDecl();
        if (g.n64) s << "    .reg .b64 %rd<" << g.n64 << ">;\n";
        if (g.n32) s << "    .reg .b32 %r<" << g.n32 << ">;\n";
        if (g.np) s << "    .reg .pred %p<" << g.np << ">;\n";
        return s.str();
    };

    // INPUT nodes below a set of roots, in node order
    auto inputNodesOf = [&](std::vector<int> roots) {
        std::vector<bool> seen(nNodes, false);
        while (!roots.empty()) {
            const int k = roots.back(); roots.pop_back();
            if (k < 0 || k >= nNodes || seen[k]) continue;
            seen[k] = true;
            if (P.nodes[k].op != PG_EX_INPUT) for (int o : {P.nodes[k].a, P.nodes[k].b, P.nodes[k].c}) roots.push_back(o);
        }
        std::vector<int> in;
        for (int k = 0; k < nNodes; k++) if (seen[k] && P.nodes[k].op == PG_EX_INPUT) in.push_back(k);
        return in;
    };
    // one evaluation context per unrolled round; all rounds' input loads are generated before any round's arithmetic and stores
    struct RoundCtx { std::vector<NodeVal> vals; std::vector<bool> done; };
    auto enter = [&](Gen& g, std::vector<RoundCtx>& ctx, int u, const std::string& guard) {
        g.vals = ctx[u].vals; g.done = ctx[u].done; g.row = "%row" + std::to_string(u); g.guard = guard + std::to_string(u);
    };
    auto leave = [&](Gen& g, std::vector<RoundCtx>& ctx, int u) { ctx[u].vals = g.vals; ctx[u].done = g.done; };

    // ---- nullability of the projections (needs one dry generation pass of the projection sub-DAG) ----
    {
        Gen dry(P); newGen(dry); dry.guard = "%pact0";
        sig.projectionNullable.assign(P.projections.size(), false);
        for (size_t j = 0; j < P.projections.size(); j++) sig.projectionNullable[j] = !dry.node(P.projections[j]).nul.empty();
        if (!dry.fail.empty()) { sig.reason = dry.fail; return ""; }
    }

    // ---- fp_filter: selection bitmask + per-unit counts ----
    if (P.filterNode >= 0) {
        Gen g(P); newGen(g);
        const std::string U = std::to_string(kJitUnroll);
        std::vector<RoundCtx> ctx(kJitUnroll);
        for (auto& c : ctx) { c.vals.assign(nNodes, NodeVal()); c.done.assign(nNodes, false); }
        const std::vector<int> filterInputs = inputNodesOf({P.filterNode});
        for (int u = 0; u < kJitUnroll; u++) { enter(g, ctx, u, "%pin"); for (int k : filterInputs) g.node(k); leave(g, ctx, u); }
        for (int u = 0; u < kJitUnroll; u++) {
            const std::string su = std::to_string(u);
            enter(g, ctx, u, "%pin");
            const NodeVal f = g.node(P.filterNode);
            if (!g.fail.empty() || f.kind != 'b') { sig.reason = g.fail.empty() ? "filter is not boolean" : g.fail; return ""; }
            g.emit("and.pred %psel" + su + ", %pin" + su + ", " + f.val + ";");
            if (!f.nul.empty()) g.emit("and.pred %psel" + su + ", %psel" + su + ", " + g.notp(f.nul) + ";");      // only TRUE selects
        }
        std::ostringstream k;
        k << ".visible .entry fp_filter(\n"; paramList(g, k, true); k << "\n)\n{\n" << regDecl(g) << prologue(g, sig.usedInputs, true)
          << "    ld.param.u64 %sel, [p_sel];\n    cvta.to.global.u64 %sel, %sel;\n    ld.param.u64 %cnt, [p_cnt];\n    cvta.to.global.u64 %cnt, %cnt;\n"
          << "    ld.param.u64 %errp, [p_err];\n    cvta.to.global.u64 %errp, %errp;\n"
          << "F_UNIT:\n    setp.ge.u64 %pdone, %unit, %units;\n    @%pdone bra F_EXIT;\n"
          << "    shl.b64 %begin, %unit, 8;\n    add.u64 %end, %begin, " << kJitUnitRows << ";\n    min.u64 %end, %end, %n;\n    mov.u64 %base, %begin;\n    mov.u32 %count, 0;\n"
          << "F_ROUND:\n    setp.ge.u64 %pdr, %base, %end;\n    @%pdr bra F_NEXT;\n";
        for (int u = 0; u < kJitUnroll; u++)
            k << "    add.u64 %row" << u << ", %base, %lane64;\n    add.u64 %row" << u << ", %row" << u << ", " << 32 * u << ";\n    setp.lt.u64 %pin" << u << ", %row" << u << ", %end;\n";
        k << g.body.str();
        for (int u = 0; u < kJitUnroll; u++)
            k << "    vote.sync.ballot.b32 %mask" << u << ", %psel" << u << ", 0xffffffff;\n"
              << "    add.u64 %a64, %base, " << 32 * u << ";\n    setp.lt.u64 %prd" << u << ", %a64, %end;\n    and.pred %prd" << u << ", %prd" << u << ", %plane0;\n"
              << "    shr.u64 %w64, %a64, 5;\n    mad.lo.u64 %a64, %w64, 4, %sel;\n    @%prd" << u << " st.global.u32 [%a64], %mask" << u << ";\n"
              << "    popc.b32 %t0, %mask" << u << ";\n    add.u32 %count, %count, %t0;\n";
        k << "    add.u64 %base, %base, " << 32 * kJitUnroll << ";\n    bra F_ROUND;\n"
          << "F_NEXT:\n    mad.lo.u64 %a64, %unit, 4, %cnt;\n    @%plane0 st.global.u32 [%a64], %count;\n    add.u64 %unit, %unit, %nwarps64;\n    bra F_UNIT;\n"
          << "F_EXIT:\n    @%perr_ov red.global.or.b32 [%errp], 1;\n    @%perr_dz red.global.or.b32 [%errp], 2;\n    ret;\n}\n\n";
        ptx << k.str();
    }

    // ---- fp_project ----
    {
        Gen g(P); newGen(g);
        std::vector<RoundCtx> ctx(kJitUnroll);
        for (auto& c : ctx) { c.vals.assign(nNodes, NodeVal()); c.done.assign(nNodes, false); }
        const std::vector<int> projectInputs = inputNodesOf(P.projections);
        for (int u = 0; u < kJitUnroll; u++) { enter(g, ctx, u, "%pact"); for (int k : projectInputs) g.node(k); leave(g, ctx, u); }
        for (int u = 0; u < kJitUnroll && g.fail.empty(); u++) {
            const std::string su = std::to_string(u), pact = "%pact" + su, dest = "%dest64_" + su;
            enter(g, ctx, u, "%pact");
            for (size_t j = 0; j < P.projections.size(); j++) {
                const NodeVal v = g.node(P.projections[j]);
                if (!g.fail.empty()) break;
                const int t = P.nodes[P.projections[j]].type;
                const int w = (t == PG_TYPE_BIGINT || t == PG_TYPE_DOUBLE) ? 8 : (t == PG_TYPE_BOOLEAN ? 1 : 4);
                if ((w == 1) != (v.kind == 'b')) { g.fail = "projection type / representation mismatch"; break; }
                const std::string outv = "%out" + std::to_string(j) + "v", outn = "%out" + std::to_string(j) + "n";
                std::string addr = g.r64();
                g.emit("mad.lo.u64 " + addr + ", " + dest + ", " + std::to_string(w) + ", " + outv + ";");
                if (w == 8) {
                    std::string x = v.val;
                    if (!v.nul.empty()) { x = g.r64(); g.emit("selp.b64 " + x + ", 0, " + v.val + ", " + v.nul + ";"); }
                    g.emit("@" + pact + " st.global.b64 [" + addr + "], " + x + ";");
                }
                else if (w == 4) {
                    std::string x = g.r32();
                    g.emit("cvt.u32.u64 " + x + ", " + v.val + ";");
                    if (!v.nul.empty()) g.emit("selp.b32 " + x + ", 0, " + x + ", " + v.nul + ";");
                    g.emit("@" + pact + " st.global.b32 [" + addr + "], " + x + ";");
                }
                else {
                    std::string x = g.r32();
                    g.emit("selp.u32 " + x + ", 1, 0, " + v.val + ";");
                    if (!v.nul.empty()) g.emit("selp.b32 " + x + ", 0, " + x + ", " + v.nul + ";");
                    g.emit("@" + pact + " st.global.u8 [" + addr + "], " + x + ";");
                }
                if (sig.projectionNullable[j]) {
                    std::string an = g.r64(), x = g.r32();
                    g.emit("add.u64 " + an + ", " + outn + ", " + dest + ";");
                    g.emit("selp.u32 " + x + ", 1, 0, " + v.nul + ";");
                    g.emit("@" + pact + " st.global.u8 [" + an + "], " + x + ";");
                }
            }
        }
        if (!g.fail.empty()) { sig.reason = g.fail; return ""; }
        std::ostringstream k;
        k << ".visible .entry fp_project(\n"; paramList(g, k, false); k << "\n)\n{\n" << regDecl(g);
        for (size_t j = 0; j < P.projections.size(); j++) { k << "    .reg .b64 %out" << j << "v;\n"; if (sig.projectionNullable[j]) k << "    .reg .b64 %out" << j << "n;\n"; }
        k << prologue(g, sig.usedInputs, false);
        for (size_t j = 0; j < P.projections.size(); j++) {
            k << "    ld.param.u64 %out" << j << "v, [p_out" << j << "_v];\n    cvta.to.global.u64 %out" << j << "v, %out" << j << "v;\n";
            if (sig.projectionNullable[j]) k << "    ld.param.u64 %out" << j << "n, [p_out" << j << "_n];\n    cvta.to.global.u64 %out" << j << "n, %out" << j << "n;\n";
        }
        k << "    ld.param.u64 %sel, [p_sel];\n    ld.param.u64 %off, [p_off];\n    ld.param.u64 %errp, [p_err];\n    cvta.to.global.u64 %errp, %errp;\n";
        const bool hasFilter = P.filterNode >= 0;
        if (hasFilter) k << "    cvta.to.global.u64 %sel, %sel;\n    cvta.to.global.u64 %off, %off;\n";
        k << "P_UNIT:\n    setp.ge.u64 %pdone, %unit, %units;\n    @%pdone bra P_EXIT;\n"
          << "    shl.b64 %begin, %unit, 8;\n    add.u64 %end, %begin, " << kJitUnitRows << ";\n    min.u64 %end, %end, %n;\n    mov.u64 %base, %begin;\n";
        if (hasFilter) k << "    mad.lo.u64 %a64, %unit, 4, %off;\n    ld.global.u32 %cur, [%a64];\n";
        else k << "    cvt.u32.u64 %cur, %begin;\n";
        k << "P_ROUND:\n    setp.ge.u64 %pdr, %base, %end;\n    @%pdr bra P_NEXT;\n";
        for (int u = 0; u < kJitUnroll; u++) {
            k << "    add.u64 %row" << u << ", %base, %lane64;\n    add.u64 %row" << u << ", %row" << u << ", " << 32 * u << ";\n    setp.lt.u64 %pin" << u << ", %row" << u << ", %end;\n";
            if (hasFilter)
                k << "    add.u64 %a64, %base, " << 32 * u << ";\n    setp.lt.u64 %prd" << u << ", %a64, %end;\n    shr.u64 %w64, %a64, 5;\n    mad.lo.u64 %a64, %w64, 4, %sel;\n"
                  << "    mov.u32 %mask" << u << ", 0;\n    @%prd" << u << " ld.global.u32 %mask" << u << ", [%a64];\n";
            else k << "    vote.sync.ballot.b32 %mask" << u << ", %pin" << u << ", 0xffffffff;\n";
        }
        for (int u = 0; u < kJitUnroll; u++)
            k << "    shr.u32 %t0, %mask" << u << ", %lane;\n    and.b32 %t0, %t0, 1;\n    setp.ne.u32 %pbit" << u << ", %t0, 0;\n    and.pred %pact" << u << ", %pin" << u << ", %pbit" << u << ";\n"
              << "    and.b32 %t1, %mask" << u << ", %lt;\n    popc.b32 %t1, %t1;\n    add.u32 %t1, %cur, %t1;\n    cvt.u64.u32 %dest64_" << u << ", %t1;\n"
              << "    popc.b32 %t0, %mask" << u << ";\n    add.u32 %cur, %cur, %t0;\n";
        k << g.body.str()
          << "    add.u64 %base, %base, " << 32 * kJitUnroll << ";\n    bra P_ROUND;\n"
          << "P_NEXT:\n    add.u64 %unit, %unit, %nwarps64;\n    bra P_UNIT;\n"
          << "P_EXIT:\n    @%perr_ov red.global.or.b32 [%errp], 1;\n    @%perr_dz red.global.or.b32 [%errp], 2;\n    ret;\n}\n";
        ptx << k.str();
    }
    sig.ok = true;
    return ptx.str();
}

}  // namespace pg
