// presto_b200 — open-addressing group table shared by HashAggregationOperator and GroupByHash.
//
// Replaces BigintGroupByHash (MB/operator/BigintGroupByHash.java:44-437: LongBigArray values + IntBigArray groupIds) and
// MultiChannelGroupByHash (MultiChannelGroupByHash.java:56-872: groupAddressByHash / groupIdsByHash / rawHashByHashPosition +
// key pages) with ONE structure: every group key is packed into W 64-bit words (W = 1 for a single BIGINT key or any key
// tuple of < 64 bits, e.g. Q1's two varchar(1) columns), stored struct-of-arrays in HBM, linear probing on the reference's
// own slot function (murmurHash3 fmix64).  Group ids are NOT assigned at insert time (that would make them depend on thread
// scheduling): each slot records the smallest global row index that touched it (firstRow) and ids are the rank of firstRow —
// exactly the first-appearance order the reference's sequential loop produces (BigintGroupByHash.java:255-262).
//
//   W == 1 : slot claimed by a 64-bit CAS on the key itself (EMPTY sentinel; the one key equal to the sentinel and the
//            NULL key of the Bigint path live in two dedicated slots past the table: cap, cap+1)
//   W  > 1 : slot claimed by CAS on a state word (EMPTY → LOCKED), key words stored, then published as READY|tag; readers
//            that see LOCKED spin (independent thread scheduling, sm_70+).
#pragma once
#include "common.cuh"

namespace pg {

constexpr uint64_t kEmptyKey = 0x8000000000000000ULL;   // Long.MIN_VALUE: W==1 empty-slot sentinel
constexpr uint64_t kNoRow = 0xFFFFFFFFFFFFFFFFULL;
constexpr uint32_t kNoRow32 = 0xFFFFFFFFu;
constexpr uint32_t kStEmpty = 0, kStLocked = 1;
constexpr int kMaxKeyWords = 4;
// firstRow word of a slot = (global row index << 8) | spelling mask of that row: bit k set = DOUBLE key column k of the row was -0.0
// (DoubleType.equalTo groups -0.0 with +0.0, but the reference's key pages keep the bits of the FIRST row of the group:
// MultiChannelGroupByHash.addNewGroup :324-358 appends the current row).  atomicMin keeps the smallest row, hence its spelling.
constexpr int kFirstRowShift = 8;
__host__ __device__ __forceinline__ uint64_t firstRowWord(uint64_t globalRow, uint32_t spelling = 0) { return (globalRow << kFirstRowShift) | spelling; }
constexpr int kMaxKeyColumns = 8;

// counters[] indices
enum { CNT_GROUPS = 0, CNT_DEFERRED = 1, CNT_ERROR = 2, CNT_NEW = 3, CNT_FPERR = 4, CNT_COUNT = 8 };

struct GroupTable {          // device view (passed by value to kernels)
    uint64_t* keys;          // [W][capP]
    uint32_t* state;         // [capP]  (W > 1 only)
    uint64_t* firstRow;      // [capP]  kNoRow = never touched
    uint64_t* acc;           // [capP][nvPhys] accumulator bit patterns (double or int64)
    int32_t* gid;            // [capP]  GroupByHash standalone only (-1 = not yet assigned)
    uint32_t cap;            // power of two
    uint32_t mask;
    uint32_t capP;           // cap + 2 (special slots: cap = NULL key (Bigint path), cap + 1 = key == sentinel)
    uint32_t nvPhys;
    uint32_t rowClaimLimit;  // row-path claims stop here (rows are deferred); flush-path claims may go on to cap - 1
    unsigned int* counters;  // CNT_*
};

__device__ __forceinline__ uint32_t tagOf(uint64_t h) { return (uint32_t) (((h >> 40) << 2) | 2u); }

template <int W>
__device__ __forceinline__ uint64_t hashKeyWords(const uint64_t (&kw)[W])
{
    uint64_t h = murmur64(kw[0]);   // W == 1: exactly BigintGroupByHash.getHashPosition's murmurHash3(value)
#pragma unroll
    for (int w = 1; w < W; w++) h = murmur64(h ^ kw[w]);
    return h;
}

// One group-counter increment per inserting row, issued once per set of lanes that reach it together (warp-aggregated): with
// millions of new groups per page a per-row atomic on this single address is the insert path's bottleneck.
__device__ __forceinline__ unsigned int claimGroupNumber(unsigned int* counter)
{
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned int) __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

// ---- global table -----------------------------------------------------------------------------------------------------
// Returns the slot of the key; inserting it when absent and `insert`.  -1: absent and !insert.  -2: table at its claim limit
// (the caller defers the row).  `fromFlush` claims are exempt from the row claim limit.
template <int W>
__device__ __forceinline__ int tableFind(const GroupTable& t, const uint64_t (&kw)[W], uint64_t h, bool insert, bool fromFlush, bool* isNew = nullptr)
{
    uint32_t slot = (uint32_t) h & t.mask;
    if (W == 1) {
        unsigned long long* keys = (unsigned long long*) t.keys;
        for (uint32_t probe = 0; probe < t.cap; probe++) {
            uint64_t k = __ldcg(&keys[slot]);
            if (k == kw[0]) return (int) slot;
            if (k == kEmptyKey) {
                if (!insert) return -1;
                // once the table is at its claim limit every new key is deferred: find that out with a load, not with two atomics on
                // the one counter every deferred row of the page would otherwise hammer
                if (!fromFlush && __ldcg(&t.counters[CNT_GROUPS]) >= t.rowClaimLimit) return -2;
                unsigned int c = claimGroupNumber(&t.counters[CNT_GROUPS]);
                if (!fromFlush && c >= t.rowClaimLimit) { atomicSub(&t.counters[CNT_GROUPS], 1u); return -2; }
                uint64_t old = atomicCAS(&keys[slot], (unsigned long long) kEmptyKey, (unsigned long long) kw[0]);
                if (old == kEmptyKey) { if (isNew) *isNew = true; return (int) slot; }
                atomicSub(&t.counters[CNT_GROUPS], 1u);
                if (old == kw[0]) return (int) slot;
            }
            slot = (slot + 1) & t.mask;
        }
        return -2;
    }
    else {
        const uint32_t ready = tagOf(h);
        volatile uint32_t* state = t.state;
        volatile uint64_t* keys = t.keys;
        for (uint32_t probe = 0; probe < t.cap; probe++) {
            uint32_t st = state[slot];
            if (st == kStEmpty) {
                if (!insert) return -1;
                if (!fromFlush && __ldcg(&t.counters[CNT_GROUPS]) >= t.rowClaimLimit) return -2;
                unsigned int c = claimGroupNumber(&t.counters[CNT_GROUPS]);
                if (!fromFlush && c >= t.rowClaimLimit) { atomicSub(&t.counters[CNT_GROUPS], 1u); return -2; }
                uint32_t old = atomicCAS(&t.state[slot], kStEmpty, kStLocked);
                if (old == kStEmpty) {
#pragma unroll
                    for (int w = 0; w < W; w++) keys[(size_t) w * t.capP + slot] = kw[w];
                    __threadfence();
                    state[slot] = ready;
                    if (isNew) *isNew = true;
                    return (int) slot;
                }
                atomicSub(&t.counters[CNT_GROUPS], 1u);
                st = old;
            }
            while (st == kStLocked) st = state[slot];
            if (st == ready) {
                bool eq = true;
#pragma unroll
                for (int w = 0; w < W; w++) eq &= keys[(size_t) w * t.capP + slot] == kw[w];
                if (eq) return (int) slot;
            }
            slot = (slot + 1) & t.mask;
        }
        return -2;
    }
}

// ---- key packing plan ---------------------------------------------------------------------------------------------------
struct KeyField {
    const void* values;
    const uint8_t* nulls;
    const int32_t* offsets;   // VARCHAR
    int32_t type;
    int32_t word, shift, bits;
    int32_t maxLen;           // VARCHAR(n) keys: n (<= 7)
};

struct KeyPlan {
    KeyField f[kMaxKeyColumns];
    int32_t nKeys;
    int32_t bigintPath;        // single BIGINT key: word 0 is the raw value, NULL → special slot
    int32_t nullWord, nullShift;
    unsigned int* errorFlag;   // CNT_ERROR word
};

enum { ERR_BIGINT_OVERFLOW = 1, ERR_VARCHAR_TOO_LONG = 2, ERR_DECIMAL_OVERFLOW = 4 };

// Packs the key of `row`.  Returns 0 = ordinary key, 1 = NULL key of the Bigint path, 2 = key equal to the W==1 sentinel.
template <int W>
__device__ __forceinline__ int packKey(const KeyPlan& kp, int64_t row, uint64_t (&kw)[W], uint32_t* spelling = nullptr)
{
#pragma unroll
    for (int w = 0; w < W; w++) kw[w] = 0;
    if (kp.bigintPath) {
        const KeyField& f = kp.f[0];
        if (f.nulls && f.nulls[row]) return 1;
        kw[0] = ldStream64((const uint64_t*) f.values + row);
        return kw[0] == kEmptyKey ? 2 : 0;
    }
    uint32_t nullMask = 0;
    for (int k = 0; k < kp.nKeys; k++) {
        const KeyField& f = kp.f[k];
        uint64_t v = 0;
        if (f.nulls && f.nulls[row]) nullMask |= 1u << k;
        else {
            switch (f.type) {
                case PG_TYPE_BIGINT: v = ((const uint64_t*) f.values)[row]; break;
                case PG_TYPE_DOUBLE: {
                    int64_t b = canonicalDoubleBits(((const int64_t*) f.values)[row]);
                    v = (uint64_t) b;
                    if (b == (int64_t) 0x8000000000000000LL) { v = 0; if (spelling) *spelling |= 1u << k; }   // -0.0 groups with +0.0 (DoubleType.equalTo: a == b)
                    break;
                }
                case PG_TYPE_INTEGER: case PG_TYPE_DATE: v = ((const uint32_t*) f.values)[row]; break;
                case PG_TYPE_BOOLEAN: v = ((const uint8_t*) f.values)[row] != 0; break;
                case PG_TYPE_VARCHAR: {
                    int32_t o = f.offsets[row], len = f.offsets[row + 1] - o;
                    const uint8_t* p = (const uint8_t*) f.values + o;
                    if (len > f.maxLen) { atomicOr(kp.errorFlag, (unsigned int) ERR_VARCHAR_TOO_LONG); len = f.maxLen; }
                    v = (uint64_t) (uint32_t) len;
                    for (int i = 0; i < len && i < 7; i++) v |= (uint64_t) p[i] << (8 * (i + 1));
                    break;
                }
            }
        }
#pragma unroll
        for (int w = 0; w < W; w++) if (w == f.word) kw[w] |= v << f.shift;
    }
#pragma unroll
    for (int w = 0; w < W; w++) if (w == kp.nullWord) kw[w] |= (uint64_t) nullMask << kp.nullShift;
    return (W == 1 && kw[0] == kEmptyKey) ? 2 : 0;
}

}  // namespace pg
