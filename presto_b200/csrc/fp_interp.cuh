// presto_b200 — the RowExpression interpreter's node evaluation, shared by the PageProcessor kernels (filter_project.cu) and the join
// filter function inside the probe kernels (hash_join.cu: JoinFilterFunction, MB/operator/JoinFilterFunction.java /
// MB/sql/gen/JoinFilterFunctionCompiler.java — there an expression over the build row at one position and the probe row at another).
#pragma once
#include "common.cuh"

namespace pg {

constexpr int kMaxNodes = 64;     // one bit per node in the null masks
constexpr int kMaxInputs = 32;

enum { FP_ERR_OVERFLOW = 1, FP_ERR_DIV_ZERO = 2 };

struct FpNode { int32_t op, type, a, b, c, isDouble; int64_t immI; int32_t len, pad; };   // immI: the double's bits for CONST_DOUBLE; CONST_VARCHAR: device address of the literal, len its length
struct FpInput { const void* values; const uint8_t* nulls; int32_t type; const int32_t* offsets; };   // offsets: VARCHAR channels

struct Val { uint64_t bits; bool isNull; };

// VARCHAR values: bits = byte offset (low 32) | length (high 32) relative to the producing node's base — the channel's bytes for an
// INPUT, the operator's literal pool for a CONST_VARCHAR (immI = device address of the literal, len = its length).
__device__ __forceinline__ const uint8_t* sliceBase(const FpNode* nodes, const FpInput* inputs, int node)
{
    const FpNode& nd = nodes[node];
    return nd.op == PG_EX_INPUT ? (const uint8_t*) inputs[nd.a].values : (const uint8_t*) (uintptr_t) nd.immI;
}
// Slice.compareTo: unsigned bytes over the common prefix, then the lengths (VarcharOperators.java:71-97)
__device__ __forceinline__ int compareSlices(const uint8_t* p, uint32_t lp, const uint8_t* q, uint32_t lq)
{
    const uint32_t m = lp < lq ? lp : lq;
    for (uint32_t i = 0; i < m; i++) { const int d = (int) p[i] - (int) q[i]; if (d) return d; }
    return lp < lq ? -1 : (lp > lq ? 1 : 0);
}
__device__ __forceinline__ int compareVarchar(const FpNode* nodes, const FpInput* inputs, int na, const Val& x, int nb, const Val& y)
{
    return compareSlices(sliceBase(nodes, inputs, na) + (uint32_t) x.bits, (uint32_t) (x.bits >> 32), sliceBase(nodes, inputs, nb) + (uint32_t) y.bits, (uint32_t) (y.bits >> 32));
}

__device__ __forceinline__ double asDouble(const Val& v, bool isDouble) { return isDouble ? __longlong_as_double((long long) v.bits) : (double) (long long) v.bits; }

// `nodes` / `inputs` / `error`: the program; an INPUT node reads input nd.a at `row`.
__device__ __forceinline__ Val evalNodeAt(const FpNode* nodes, const FpInput* inputs, unsigned int* error, const FpNode& nd, int64_t row, const Val& x, const Val& y, const Val& z)
{
    Val r{0, false};
    switch (nd.op) {
        case PG_EX_INPUT: {
            const FpInput& in = inputs[nd.a];
            if (in.nulls && in.nulls[row]) r.isNull = true;
            else switch (in.type) {
                case PG_TYPE_BIGINT: case PG_TYPE_DOUBLE: r.bits = ldStream64((const uint64_t*) in.values + row); break;
                case PG_TYPE_INTEGER: case PG_TYPE_DATE: r.bits = (uint64_t) (int64_t) (int32_t) ldStream32((const uint32_t*) in.values + row); break;
                case PG_TYPE_VARCHAR: { const uint32_t o = (uint32_t) in.offsets[row]; r.bits = (uint64_t) o | ((uint64_t) ((uint32_t) in.offsets[row + 1] - o) << 32); break; }
                default: r.bits = ((const uint8_t*) in.values)[row] != 0; break;
            }
            break;
        }
        case PG_EX_CONST_BIGINT: case PG_EX_CONST_BOOLEAN: case PG_EX_CONST_DOUBLE: r.bits = (uint64_t) nd.immI; break;
        case PG_EX_CONST_NULL: r.isNull = true; break;
        case PG_EX_CONST_VARCHAR: r.bits = (uint64_t) (uint32_t) nd.len << 32; break;
        case PG_EX_ADD: case PG_EX_SUB: case PG_EX_MUL: case PG_EX_DIV: {
            if (x.isNull || y.isNull) { r.isNull = true; break; }
            if (nd.isDouble) {
                const double p = asDouble(x, nodes[nd.a].isDouble), q = asDouble(y, nodes[nd.b].isDouble);
                const double v = nd.op == PG_EX_ADD ? p + q : nd.op == PG_EX_SUB ? p - q : nd.op == PG_EX_MUL ? p * q : p / q;
                r.bits = (uint64_t) __double_as_longlong(v);
            }
            else {
                const long long p = (long long) x.bits, q = (long long) y.bits;
                long long v = 0;
                if (nd.op == PG_EX_ADD) { v = (long long) ((uint64_t) p + (uint64_t) q); if (((p ^ v) & (q ^ v)) < 0) atomicOr(error, (unsigned) FP_ERR_OVERFLOW); }
                else if (nd.op == PG_EX_SUB) { v = (long long) ((uint64_t) p - (uint64_t) q); if (((p ^ q) & (p ^ v)) < 0) atomicOr(error, (unsigned) FP_ERR_OVERFLOW); }
                else if (nd.op == PG_EX_MUL) {
                    v = (long long) ((uint64_t) p * (uint64_t) q);
                    if (__mul64hi(p, q) != (v >> 63)) atomicOr(error, (unsigned) FP_ERR_OVERFLOW);
                }
                else {
                    if (q == 0) { atomicOr(error, (unsigned) FP_ERR_DIV_ZERO); v = 0; }
                    else if (p == LLONG_MIN && q == -1) { atomicOr(error, (unsigned) FP_ERR_OVERFLOW); v = 0; }   // BigintOperators.divide :106-117
                    else v = p / q;
                }
                r.bits = (uint64_t) v;
            }
            break;
        }
        case PG_EX_NEG:
            if (x.isNull) { r.isNull = true; break; }
            if (nd.isDouble) r.bits = x.bits ^ 0x8000000000000000ULL;
            else { if ((long long) x.bits == LLONG_MIN) atomicOr(error, (unsigned) FP_ERR_OVERFLOW); r.bits = (uint64_t) (-(long long) x.bits); }
            break;
        case PG_EX_EQ: case PG_EX_NE: case PG_EX_LT: case PG_EX_LE: case PG_EX_GT: case PG_EX_GE: {
            if (x.isNull || y.isNull) { r.isNull = true; break; }
            bool res;
            const bool da = nodes[nd.a].isDouble, db = nodes[nd.b].isDouble;
            if (nodes[nd.a].type == PG_TYPE_VARCHAR) {
                const int c = compareVarchar(nodes, inputs, nd.a, x, nd.b, y);
                res = nd.op == PG_EX_EQ ? c == 0 : nd.op == PG_EX_NE ? c != 0 : nd.op == PG_EX_LT ? c < 0 : nd.op == PG_EX_LE ? c <= 0 : nd.op == PG_EX_GT ? c > 0 : c >= 0;
            }
            else if (da || db) {
                const double p = asDouble(x, da), q = asDouble(y, db);
                res = nd.op == PG_EX_EQ ? p == q : nd.op == PG_EX_NE ? p != q : nd.op == PG_EX_LT ? p < q : nd.op == PG_EX_LE ? p <= q : nd.op == PG_EX_GT ? p > q : p >= q;
            }
            else {
                const long long p = (long long) x.bits, q = (long long) y.bits;
                res = nd.op == PG_EX_EQ ? p == q : nd.op == PG_EX_NE ? p != q : nd.op == PG_EX_LT ? p < q : nd.op == PG_EX_LE ? p <= q : nd.op == PG_EX_GT ? p > q : p >= q;
            }
            r.bits = res;
            break;
        }
        case PG_EX_BETWEEN: {
            const bool da = nodes[nd.a].isDouble, db = nodes[nd.b].isDouble, dc = nodes[nd.c].isDouble;
            const bool n1 = x.isNull || y.isNull, n2 = x.isNull || z.isNull;
            bool c1 = false, c2 = false;
            if (nodes[nd.a].type == PG_TYPE_VARCHAR) {     // VarcharOperators.between :103-105
                if (!n1) c1 = compareVarchar(nodes, inputs, nd.a, x, nd.b, y) >= 0;
                if (!n2) c2 = compareVarchar(nodes, inputs, nd.a, x, nd.c, z) <= 0;
            }
            else {
            if (!n1) c1 = (da || db) ? asDouble(x, da) >= asDouble(y, db) : (long long) x.bits >= (long long) y.bits;
            if (!n2) c2 = (da || dc) ? asDouble(x, da) <= asDouble(z, dc) : (long long) x.bits <= (long long) z.bits;
            }
            if ((!n1 && !c1) || (!n2 && !c2)) r.bits = 0;
            else if (n1 || n2) r.isNull = true;
            else r.bits = 1;
            break;
        }
        case PG_EX_AND:
            if ((!x.isNull && !x.bits) || (!y.isNull && !y.bits)) r.bits = 0;
            else if (x.isNull || y.isNull) r.isNull = true;
            else r.bits = 1;
            break;
        case PG_EX_OR:
            if ((!x.isNull && x.bits) || (!y.isNull && y.bits)) r.bits = 1;
            else if (x.isNull || y.isNull) r.isNull = true;
            else r.bits = 0;
            break;
        case PG_EX_NOT: if (x.isNull) r.isNull = true; else r.bits = !x.bits; break;
        case PG_EX_IS_NULL: r.bits = x.isNull ? 1 : 0; break;
        case PG_EX_CAST_DOUBLE: if (x.isNull) r.isNull = true; else r.bits = (uint64_t) __double_as_longlong(asDouble(x, nodes[nd.a].isDouble)); break;
        case PG_EX_HASH: {
            const int t = nodes[nd.a].type;
            if (x.isNull) r.bits = 0;
            else if (t == PG_TYPE_DOUBLE) r.bits = (uint64_t) hashLong(doubleHashBits((int64_t) x.bits));
            else if (t == PG_TYPE_BOOLEAN) r.bits = x.bits ? 1231 : 1237;
            else if (t == PG_TYPE_VARCHAR) r.bits = (uint64_t) xxHash64Bytes(sliceBase(nodes, inputs, nd.a) + (uint32_t) x.bits, (int64_t) (x.bits >> 32));
            else r.bits = (uint64_t) hashLong((int64_t) x.bits);
            break;
        }
        case PG_EX_COMBINE_HASH: r.bits = 31ULL * x.bits + y.bits; break;
        default: break;
    }
    return r;
}

}  // namespace pg
