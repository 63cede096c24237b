// presto_b200 — VarcharInterner: VARCHAR values of any length as 32-bit ids, on the device.
//
// The reference keeps variable-width group / join keys as slices in its key pages and compares them byte by byte after a hash
// match (MultiChannelGroupByHash.java:293-358 putIfAbsent → positionEqualsCurrentRow; PagesHashStrategy.positionEqualsRow for
// joins).  The GPU tables work on fixed-width packed keys (group_table.cuh), so a VARCHAR channel that does not fit a packed
// field is first replaced by the id of its value: equal slices ⇔ equal ids, ids are private to one interner (one operator /
// one LookupSource), never visible in a result.  The bytes live once in the interner's pool; output key columns are gathered
// from it.  Ids follow insertion, which depends on thread scheduling — every consumer orders its results by something else
// (first-row ranks for group ids, build positions for join rows).
#pragma once
#include "common.cuh"

namespace pg {

struct InternView {                 // device view (kernels)
    unsigned long long* slots;      // [cap]  0 = empty, 1 = locked, (hash >> 32) << 32 | (id + 2) = ready: the tag spares a probe the entry of most other values
    uint32_t cap, mask;
    uint64_t* entryMeta;            // [entryCap] (pool offset / 8) << 31 | length: values start 8-byte aligned in the pool
    uint64_t* entryHash;            // [entryCap] XXH64 of the bytes = VarcharType.hash of the value
    uint8_t* pool;
    unsigned long long* counters;   // [0] entries, [1] pool bytes in use
};
__host__ __device__ __forceinline__ uint32_t internLengthOf(uint64_t meta) { return (uint32_t) (meta & 0x7fffffffULL); }
__host__ __device__ __forceinline__ uint64_t internOffsetOf(uint64_t meta) { return (meta >> 31) << 3; }

class VarcharInterner {
    DeviceBuffer slots, entryMeta, entryHash, pool, counters, dictionaryIds;
    PinnedBuffer host;
    uint32_t cap = 0;
    uint64_t entryCap = 0, poolCap = 0;
    uint64_t nEntries = 0, poolUsed = 0;     // as of the last synchronisation
    void ensure(Operator* op, uint64_t extraEntries, uint64_t extraBytes);
    void readCounters(Operator* op);

public:
    // `accounting`: the device-bytes counter of whoever owns the interner (an operator's stats, a LookupSource); `op` lends its stream
    explicit VarcharInterner(Operator* op, int64_t* accounting);
    InternView view() const;
    uint64_t entries() const { return nEntries; }
    // ids of rows [0, n) of a staged FLAT VARCHAR column into idsOut (device, int32 per row).  insert = false only looks values up:
    // an absent value gets -1 (read-only: probes of one LookupSource may look values up concurrently on their own streams).  NULL rows
    // get 0 — the caller keeps the null map.  Runs on op->stream; inserting synchronises it.
    void intern(Operator* op, const DeviceColumn& col, int64_t n, int32_t* idsOut, bool insert);
    // The same for a column the stager kept encoded (DeviceColumn::encoding): the dictionary's entries are interned once, every row takes
    // the id of its entry — MultiChannelGroupByHash's dictionary path (:543-552, 628-714: work per dictionary entry, not per position).
    // nullsOut (device, one byte per row, may be null when the dictionary has no null map) receives the rows' null flags.
    void internEncoded(Operator* op, const DeviceColumn& col, int64_t n, int32_t* idsOut, uint8_t* nullsOut, bool insert);
    // output side: lengths[g] = length of value ids[g] (0 where nulls[g]); then, given exclusive offsets of those lengths, the bytes.
    void gatherLengths(Operator* op, const int32_t* ids, const uint8_t* nulls, int64_t g, int32_t* lengthsOut);
    void gatherBytes(Operator* op, const int32_t* ids, const uint8_t* nulls, const int32_t* offsets, int64_t g, uint8_t* bytesOut);
};

}  // namespace pg
