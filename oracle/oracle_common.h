/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the prestodb/presto Java operators on the hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (presto_b200/, libprestogpu.so) never links, imports or calls it.
 *
 * Parity pinning: the Java reference cannot run here (no JVM), and the reference's tests contain no literal hash
 * vectors; the oracle is pinned by re-encoding the reference's behavioural tests (tests/test_oracle_*.py cite each
 * one) and by cross-checking XXH64 against the python `xxhash` package.  HashCommon.arraySize capacity and airlift
 * XxHash64 are third-party (fastutil 8.5.2, airlift slice 0.38): restated from their published algorithms —
 * "parity unpinned" for exact table capacity, which is not an operator result.
 *
 * Path abbreviations: MB = presto-main-base/src/main/java/com/facebook/presto/, PC = presto-common/src/main/java/com/facebook/presto/common/
 */
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include <memory>
#include "../include/presto_gpu.h"

namespace oracle {

struct OracleError : std::runtime_error {
    int code;
    OracleError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline int typeWidth(int type)
{
    switch (type) {
        case PG_TYPE_BIGINT: case PG_TYPE_DOUBLE: return 8;
        case PG_TYPE_INTEGER: case PG_TYPE_DATE: return 4;
        case PG_TYPE_BOOLEAN: return 1;
        case PG_TYPE_VARCHAR: return 0;
        case PG_TYPE_LONG_DECIMAL: return 16;
    }
    throw OracleError(PG_ERR_UNSUPPORTED, "unknown type");
}

// ---- hashing -----------------------------------------------------------------------------------------------------
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// PC/type/AbstractLongType.java:137-141 (same body for AbstractIntType.java:139-143 on the sign-extended int)
inline int64_t hashLong(int64_t value)
{
    return (int64_t) (rotl64((uint64_t) value * 0xC2B2AE3D27D4EB4FULL, 31) * 0x9E3779B185EBCA87ULL);
}

// fastutil HashCommon.murmurHash3(long); inlined verbatim in MB/operator/PagesHash.java:236-252
inline int64_t murmurHash3(int64_t h)
{
    uint64_t x = (uint64_t) h;
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return (int64_t) x;
}

// MB/operator/scalar/CombineHashFunction.java:26-31
inline int64_t combineHash(int64_t previous, int64_t value) { return (int64_t) (31ULL * (uint64_t) previous + (uint64_t) value); }

// MB/operator/HashGenerator.java:22-29
inline int32_t hashPartition(int64_t rawHash, int32_t partitionCount)
{
    uint32_t folded = (uint32_t) ((uint64_t) rawHash ^ ((uint64_t) rawHash >> 32));   // Long.hashCode
    return (int32_t) (((uint64_t) folded * (uint64_t) (uint32_t) partitionCount) >> 32);
}

// airlift slice 0.38 XxHash64.hash(Slice) == XXH64 seed 0 (published algorithm; cross-checked vs python xxhash in tests)
inline uint64_t xxRound(uint64_t acc, uint64_t input)
{
    acc += input * 0xC2B2AE3D27D4EB4FULL;
    acc = rotl64(acc, 31);
    return acc * 0x9E3779B185EBCA87ULL;
}
inline uint64_t xxMerge(uint64_t acc, uint64_t val)
{
    acc ^= xxRound(0, val);
    return acc * 0x9E3779B185EBCA87ULL + 0x85EBCA77C2B2AE63ULL;
}
inline int64_t xxHash64(const uint8_t* p, size_t len)
{
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        while (p + 32 <= end) {
            uint64_t a, b, c, d;
            memcpy(&a, p, 8); memcpy(&b, p + 8, 8); memcpy(&c, p + 16, 8); memcpy(&d, p + 24, 8);
            v1 = xxRound(v1, a); v2 = xxRound(v2, b); v3 = xxRound(v3, c); v4 = xxRound(v4, d);
            p += 32;
        }
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxMerge(h, v1); h = xxMerge(h, v2); h = xxMerge(h, v3); h = xxMerge(h, v4);
    }
    else {
        h = P5;
    }
    h += (uint64_t) len;
    while (p + 8 <= end) {
        uint64_t k; memcpy(&k, p, 8);
        h ^= xxRound(0, k);
        h = rotl64(h, 27) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= end) {
        uint32_t k; memcpy(&k, p, 4);
        h ^= (uint64_t) k * P1;
        h = rotl64(h, 23) * P2 + P3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t) (*p) * P5;
        h = rotl64(h, 11) * P1;
        p++;
    }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return (int64_t) h;
}

// LocalPartitionGenerator.getPartition(rawHash)  MB/operator/exchange/LocalPartitionGenerator.java:43-63:
//   processRawHash = (int) XxHash64.hash(Long.reverse(rawHash)); partition = processRawHash & (partitionCount - 1)
// airlift XxHash64.hash(long) restated as XXH64 (seed 0) of the value's 8 little-endian bytes (cross-checked vs python xxhash).
inline int32_t localPartition(int64_t rawHash, int32_t partitionCount)
{
    uint64_t x = (uint64_t) rawHash, r = 0;
    for (int i = 0; i < 64; i++) { r = (r << 1) | (x & 1); x >>= 1; }     // Long.reverse
    uint8_t le[8];
    for (int i = 0; i < 8; i++) le[i] = (uint8_t) (r >> (8 * i));
    return (int32_t) xxHash64(le, 8) & (partitionCount - 1);
}

// Double.doubleToLongBits: canonical NaN
inline int64_t doubleToLongBits(double d)
{
    if (d != d) return 0x7ff8000000000000LL;
    int64_t b; memcpy(&b, &d, 8); return b;
}

// TypeUtils.doubleHashCode  PC/type/TypeUtils.java:232-238 (DoubleType.DOUBLE = new DoubleType(true): DoubleType.java:37,92-100)
inline int64_t doubleHashCode(double d)
{
    if (d == 0) d = 0;      // value == -0 ? 0 : value
    return hashLong(doubleToLongBits(d));
}

// fastutil HashCommon.arraySize(expected, f): max(2, nextPowerOfTwo(ceil(expected / f))), > 2^30 → IllegalArgumentException
inline int32_t arraySize(int64_t expected, float f)
{
    int64_t need = (int64_t) std::ceil((double) expected / (double) f);
    int64_t s = 1;
    while (s < need) s <<= 1;
    if (s < 2) s = 2;
    if (s > (1LL << 30)) throw OracleError(PG_ERR_INSUFFICIENT_RESOURCES, "Too large (" + std::to_string(expected) + " expected elements with load factor 0.75)");
    return (int32_t) s;
}

// BigintGroupByHash.java:337-346 / MultiChannelGroupByHash.java:473-482
inline int32_t calculateMaxFill(int32_t hashSize)
{
    int32_t maxFill = (int32_t) std::ceil(hashSize * 0.75f);
    if (maxFill == hashSize) maxFill--;
    return maxFill;
}

// ---- Block view (PC/block/Block.java accessors over pg_block, all three encodings) --------------------------------
struct BlockView {
    const pg_block* b = nullptr;
    explicit BlockView(const pg_block* blk = nullptr) : b(blk) {}
    inline int64_t dictPos(int64_t pos) const
    {
        if (b->encoding == PG_ENC_FLAT) return pos;
        if (b->encoding == PG_ENC_RLE) return 0;
        return b->ids[pos];
    }
    inline bool mayHaveNull() const { return b->is_null != nullptr; }
    inline bool isNull(int64_t pos) const { return b->is_null != nullptr && b->is_null[dictPos(pos)] != 0; }
    inline int64_t getLong(int64_t pos) const
    {
        int64_t p = dictPos(pos);
        switch (b->type) {
            case PG_TYPE_BIGINT: case PG_TYPE_DOUBLE: return ((const int64_t*) b->values)[p];
            case PG_TYPE_INTEGER: case PG_TYPE_DATE: return ((const int32_t*) b->values)[p];
            case PG_TYPE_BOOLEAN: return ((const int8_t*) b->values)[p];
        }
        throw OracleError(PG_ERR_UNSUPPORTED, "getLong on variable width block");
    }
    inline double getDouble(int64_t pos) const { int64_t v = getLong(pos); double d; memcpy(&d, &v, 8); return d; }
    inline const int64_t* int128Ptr(int64_t pos) const { return (const int64_t*) b->values + 2 * dictPos(pos); }
    inline const uint8_t* slicePtr(int64_t pos) const { return (const uint8_t*) b->values + b->offsets[dictPos(pos)]; }
    inline int32_t sliceLength(int64_t pos) const { int64_t p = dictPos(pos); return b->offsets[p + 1] - b->offsets[p]; }
};

// PC/type/TypeUtils.java:167-173 hashPosition: NULL → 0, else type.hash
inline int64_t typeHashPosition(int type, const BlockView& v, int64_t pos)
{
    if (v.isNull(pos)) return 0;
    switch (type) {
        case PG_TYPE_BIGINT: case PG_TYPE_INTEGER: case PG_TYPE_DATE: return hashLong(v.getLong(pos));
        case PG_TYPE_DOUBLE: return doubleHashCode(v.getDouble(pos));   // PC/type/DoubleType.java:92-100 → TypeUtils.doubleHashCode
        case PG_TYPE_BOOLEAN: return v.getLong(pos) != 0 ? 1231 : 1237;              // PC/type/BooleanType.java:121-125
        case PG_TYPE_VARCHAR: return xxHash64(v.slicePtr(pos), (size_t) v.sliceLength(pos)); // AbstractVariableWidthBlock.java:102-106
    }
    throw OracleError(PG_ERR_UNSUPPORTED, "unknown type");
}

// ---- Owning column (BlockBuilder) ----------------------------------------------------------------------------------
struct Column {
    int type = 0;
    int64_t n = 0;
    std::vector<uint8_t> data;      // fixed width values, or varchar bytes
    std::vector<uint8_t> nulls;     // size n
    std::vector<int32_t> offsets;   // varchar: n+1
    bool hasNull = false;

    explicit Column(int t = 0) : type(t) { if (t == PG_TYPE_VARCHAR) offsets.push_back(0); }
    void appendNull()
    {
        int w = typeWidth(type);
        if (w) data.resize(data.size() + w, 0); else offsets.push_back(offsets.back());
        nulls.push_back(1); hasNull = true; n++;
    }
    void appendLong(int64_t v)
    {
        int w = typeWidth(type);
        size_t o = data.size(); data.resize(o + w);
        if (w == 8) memcpy(&data[o], &v, 8);
        else if (w == 4) { int32_t x = (int32_t) v; memcpy(&data[o], &x, 4); }
        else { data[o] = (uint8_t) v; }
        nulls.push_back(0); n++;
    }
    void appendDouble(double d) { int64_t v; memcpy(&v, &d, 8); appendLong(v); }
    void appendInt128(int64_t low, int64_t high)     // Int128ArrayBlockBuilder.writeLong x2
    {
        size_t o = data.size(); data.resize(o + 16);
        memcpy(&data[o], &low, 8); memcpy(&data[o + 8], &high, 8);
        nulls.push_back(0); n++;
    }
    void appendSlice(const uint8_t* p, int32_t len)
    {
        data.insert(data.end(), p, p + len);
        offsets.push_back((int32_t) data.size());
        nulls.push_back(0); n++;
    }
    // Type.appendTo(block, position, blockBuilder)
    void appendFrom(const BlockView& v, int64_t pos)
    {
        if (v.isNull(pos)) { appendNull(); return; }
        if (type == PG_TYPE_VARCHAR) appendSlice(v.slicePtr(pos), v.sliceLength(pos));
        else if (type == PG_TYPE_LONG_DECIMAL) { const int64_t* w = v.int128Ptr(pos); appendInt128(w[0], w[1]); }
        else appendLong(v.getLong(pos));
    }
    void appendFrom(const Column& c, int64_t pos)
    {
        if (c.nulls[pos]) { appendNull(); return; }
        if (type == PG_TYPE_VARCHAR) appendSlice(c.data.data() + c.offsets[pos], c.offsets[pos + 1] - c.offsets[pos]);
        else if (type == PG_TYPE_LONG_DECIMAL) { int64_t w[2]; memcpy(w, &c.data[pos * 16], 16); appendInt128(w[0], w[1]); }
        else appendLong(c.getLong(pos));
    }
    int64_t getLong(int64_t pos) const
    {
        int w = typeWidth(type);
        if (w == 8) { int64_t v; memcpy(&v, &data[pos * 8], 8); return v; }
        if (w == 4) { int32_t v; memcpy(&v, &data[pos * 4], 4); return v; }
        return (int8_t) data[pos];
    }
    pg_block asBlock() const
    {
        pg_block b{};
        b.type = type; b.encoding = PG_ENC_FLAT; b.position_count = n;
        b.values = data.data();
        b.is_null = hasNull ? nulls.data() : nullptr;
        b.offsets = type == PG_TYPE_VARCHAR ? offsets.data() : nullptr;
        return b;
    }
    void clear() { n = 0; data.clear(); nulls.clear(); hasNull = false; offsets.clear(); if (type == PG_TYPE_VARCHAR) offsets.push_back(0); }
};

struct OwnedPage {
    std::vector<Column> columns;
    std::vector<pg_block> blocks;
    int64_t positionCount = 0;
    void fill(pg_page* out)
    {
        blocks.resize(columns.size());
        for (size_t i = 0; i < columns.size(); i++) blocks[i] = columns[i].asBlock();
        out->position_count = positionCount;
        out->channel_count = (int32_t) columns.size();
        out->memory = PG_MEM_HOST;
        out->blocks = blocks.data();
    }
};

// Operator base: MB/operator/Operator.java:20-102
struct Operator {
    virtual ~Operator() = default;
    virtual bool needsInput() = 0;
    virtual void addInput(const pg_page* page) = 0;
    virtual bool getOutput(OwnedPage& out) = 0;   // false = null
    virtual void finish() = 0;
    virtual bool isFinished() = 0;
    virtual bool isBlocked() { return false; }    // Operator.isBlocked() (Operator.java:32)
    OwnedPage lastOutput;
    pg_operator_stats stats{};
};

}  // namespace oracle
