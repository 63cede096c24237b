/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 *
 * GroupByHash (Bigint / MultiChannel), grouped accumulators and HashAggregationOperator restated from:
 *   MB/operator/BigintGroupByHash.java:79-102,222-346,361-434
 *   MB/operator/MultiChannelGroupByHash.java:94-156,293-482
 *   MB/operator/GroupByHash.java:46-59
 *   MB/operator/aggregation/{DoubleSumAggregation.java:33-61, AverageAggregations.java:34-67, CountAggregation.java:34-50,
 *                            LongSumAggregation.java:33-56}, loop shape AccumulatorCompiler.java:462-566
 *   MB/operator/aggregation/builder/InMemoryHashAggregationBuilder.java:204-220,326-352,400-403
 *   MB/operator/HashAggregationOperator.java:403-501
 */
#include "oracle_common.h"

namespace oracle {

// ======================================================================================================================
// GroupByHash interface (MB/operator/GroupByHash.java:61-89)
// ======================================================================================================================
struct GroupByHash {
    virtual ~GroupByHash() = default;
    virtual int32_t getGroupCount() const = 0;
    virtual void addPage(const pg_page* page) = 0;
    virtual void getGroupIds(const pg_page* page, int64_t* out) = 0;
    virtual bool contains(int64_t position, const pg_page* page) = 0;
    virtual void appendValuesTo(int32_t groupId, std::vector<Column>& out, int offset) = 0;
    virtual std::vector<int> getTypes() const = 0;
    virtual int32_t getCapacity() const = 0;
    virtual int64_t getHashCollisions() const = 0;
};

// ----------------------------------------------------------------------------------------------------------------------
// BigintGroupByHash
// ----------------------------------------------------------------------------------------------------------------------
struct BigintGroupByHash : GroupByHash {
    int hashChannel;
    bool outputRawHash;
    int32_t hashCapacity, maxFill, mask;
    std::vector<int64_t> values;        // LongBigArray values
    std::vector<int32_t> groupIds;      // IntBigArray groupIds (-1 empty)
    int32_t nullGroupId = -1;
    std::vector<int64_t> valuesByGroupId;
    int32_t nextGroupId = 0;
    int64_t hashCollisions = 0;

    BigintGroupByHash(int hashChannel_, bool outputRawHash_, int expectedSize)   // :79-102
        : hashChannel(hashChannel_), outputRawHash(outputRawHash_)
    {
        if (hashChannel < 0) throw OracleError(PG_ERR_INVALID_ARGUMENT, "hashChannel must be at least zero");
        if (expectedSize <= 0) throw OracleError(PG_ERR_INVALID_ARGUMENT, "expectedSize must be greater than zero");
        hashCapacity = arraySize(expectedSize, 0.75f);
        maxFill = calculateMaxFill(hashCapacity);
        mask = hashCapacity - 1;
        values.assign(hashCapacity, 0);
        groupIds.assign(hashCapacity, -1);
        valuesByGroupId.assign(hashCapacity, 0);
    }

    static int64_t getHashPosition(int64_t rawHash, int32_t mask) { return murmurHash3(rawHash) & mask; }   // :332-335
    bool needRehash() const { return nextGroupId >= maxFill; }                                              // :327-330

    int32_t putIfAbsent(int64_t position, const BlockView& block)   // :222-253
    {
        if (block.isNull(position)) {
            if (nullGroupId < 0) nullGroupId = nextGroupId++;
            return nullGroupId;
        }
        int64_t value = block.getLong(position);
        int64_t hashPosition = getHashPosition(value, mask);
        while (true) {
            int32_t groupId = groupIds[hashPosition];
            if (groupId == -1) break;
            if (value == values[hashPosition]) return groupId;
            hashPosition = (hashPosition + 1) & mask;
            hashCollisions++;
        }
        return addNewGroup(hashPosition, value);
    }

    int32_t addNewGroup(int64_t hashPosition, int64_t value)   // :255-269
    {
        int32_t groupId = nextGroupId++;
        values[hashPosition] = value;
        if ((size_t) groupId >= valuesByGroupId.size()) valuesByGroupId.resize((size_t) groupId + 1);
        valuesByGroupId[groupId] = value;
        groupIds[hashPosition] = groupId;
        if (needRehash()) tryRehash();
        return groupId;
    }

    bool tryRehash()   // :271-325
    {
        int64_t newCapacityLong = hashCapacity * 2LL;
        if (newCapacityLong > INT32_MAX) throw OracleError(PG_ERR_INSUFFICIENT_RESOURCES, "Size of hash table cannot exceed 1 billion entries");
        int32_t newCapacity = (int32_t) newCapacityLong;
        int32_t newMask = newCapacity - 1;
        std::vector<int64_t> newValues(newCapacity, 0);
        std::vector<int32_t> newGroupIds(newCapacity, -1);
        for (int32_t groupId = 0; groupId < nextGroupId; groupId++) {
            if (groupId == nullGroupId) continue;
            int64_t value = valuesByGroupId[groupId];
            int64_t hashPosition = getHashPosition(value, newMask);
            while (newGroupIds[hashPosition] != -1) {
                hashPosition = (hashPosition + 1) & newMask;
                hashCollisions++;
            }
            newValues[hashPosition] = value;
            newGroupIds[hashPosition] = groupId;
        }
        mask = newMask;
        hashCapacity = newCapacity;
        maxFill = calculateMaxFill(hashCapacity);
        values.swap(newValues);
        groupIds.swap(newGroupIds);
        if (valuesByGroupId.size() < (size_t) maxFill) valuesByGroupId.resize(maxFill);
        return true;
    }

    int32_t getGroupCount() const override { return nextGroupId; }
    int32_t getCapacity() const override { return hashCapacity; }
    int64_t getHashCollisions() const override { return hashCollisions; }
    std::vector<int> getTypes() const override
    {
        return outputRawHash ? std::vector<int>{PG_TYPE_BIGINT, PG_TYPE_BIGINT} : std::vector<int>{PG_TYPE_BIGINT};
    }

    void addPage(const pg_page* page) override   // AddPageWork.process :361-380 (memory never blocks here)
    {
        BlockView block(&page->blocks[hashChannel]);
        if (needRehash()) tryRehash();
        for (int64_t p = 0; p < page->position_count; p++) putIfAbsent(p, block);
    }
    void getGroupIds(const pg_page* page, int64_t* out) override   // GetGroupIdsWork.process :406-426
    {
        BlockView block(&page->blocks[hashChannel]);
        if (needRehash()) tryRehash();
        for (int64_t p = 0; p < page->position_count; p++) out[p] = putIfAbsent(p, block);
    }
    bool contains(int64_t position, const pg_page* page) override   // :184-207
    {
        BlockView block(&page->blocks[hashChannel]);
        if (block.isNull(position)) return nullGroupId >= 0;
        int64_t value = block.getLong(position);
        int64_t hashPosition = getHashPosition(value, mask);
        while (true) {
            int32_t groupId = groupIds[hashPosition];
            if (groupId == -1) return false;
            if (value == values[hashPosition]) return true;
            hashPosition = (hashPosition + 1) & mask;
        }
    }
    void appendValuesTo(int32_t groupId, std::vector<Column>& out, int offset) override   // :139-159
    {
        if (groupId == nullGroupId) out[offset].appendNull(); else out[offset].appendLong(valuesByGroupId[groupId]);
        if (outputRawHash) {
            if (groupId == nullGroupId) out[offset + 1].appendLong(0);
            else out[offset + 1].appendLong(hashLong(valuesByGroupId[groupId]));
        }
    }
};

// ----------------------------------------------------------------------------------------------------------------------
// MultiChannelGroupByHash.  Key rows live in `keyStore` (the reference's channelBuilders / PageBuilder pages); the
// synthetic address (pageIndex<<32|position, SyntheticAddress.java:28-42) degenerates to the row index in keyStore
// because the oracle never rolls the page builder over — addresses are not observable.
// ----------------------------------------------------------------------------------------------------------------------
struct MultiChannelGroupByHash : GroupByHash {
    std::vector<int> hashTypes;
    std::vector<int> channels;
    int inputHashChannel;        // -1 absent
    std::vector<int> types;      // hashTypes [+ BIGINT]
    std::vector<Column> keyStore;
    int32_t hashCapacity, maxFill, mask;
    std::vector<int64_t> groupAddressByHash;   // -1 empty
    std::vector<int32_t> groupIdsByHash;
    std::vector<int8_t> rawHashByHashPosition;
    std::vector<int64_t> groupAddressByGroupId;
    int32_t nextGroupId = 0;
    int64_t hashCollisions = 0;

    MultiChannelGroupByHash(const std::vector<int>& hashTypes_, const std::vector<int>& hashChannels_, int inputHashChannel_, int expectedSize)   // :94-156
        : hashTypes(hashTypes_), channels(hashChannels_), inputHashChannel(inputHashChannel_)
    {
        if (hashTypes.size() != channels.size()) throw OracleError(PG_ERR_INVALID_ARGUMENT, "hashTypes and hashChannels have different sizes");
        if (expectedSize <= 0) throw OracleError(PG_ERR_INVALID_ARGUMENT, "expectedSize must be greater than zero");
        types = hashTypes;
        if (inputHashChannel >= 0) types.push_back(PG_TYPE_BIGINT);
        for (int t : types) keyStore.emplace_back(t);
        hashCapacity = arraySize(expectedSize, 0.75f);
        maxFill = calculateMaxFill(hashCapacity);
        mask = hashCapacity - 1;
        groupAddressByHash.assign(hashCapacity, -1);
        rawHashByHashPosition.assign(hashCapacity, 0);
        groupIdsByHash.assign(hashCapacity, 0);
    }

    // PrecomputedHashGenerator.java:31-35 / InterpretedHashGenerator.java:70-79
    int64_t hashPositionOfRow(int64_t position, const pg_page* page) const
    {
        if (inputHashChannel >= 0) return BlockView(&page->blocks[inputHashChannel]).getLong(position);
        int64_t result = 0;
        for (size_t i = 0; i < hashTypes.size(); i++)
            result = combineHash(result, typeHashPosition(hashTypes[i], BlockView(&page->blocks[channels[i]]), position));
        return result;
    }
    // hashStrategy.hashPosition(blockIndex, position) over the key store (SimplePagesHashStrategy.java:110-123)
    int64_t hashPositionOfAddress(int64_t address) const
    {
        if (inputHashChannel >= 0) return keyStore[hashTypes.size()].getLong(address);
        int64_t result = 0;
        for (size_t i = 0; i < hashTypes.size(); i++) {
            pg_block b = keyStore[i].asBlock();
            result = result * 31 + typeHashPosition(hashTypes[i], BlockView(&b), address);
        }
        return result;
    }

    // positionNotDistinctFromRow (SimplePagesHashStrategy.java:203-221): null == null, else value equality
    bool positionNotDistinctFromRow(int64_t address, int64_t position, const pg_page* page) const
    {
        for (size_t i = 0; i < hashTypes.size(); i++) {
            const Column& left = keyStore[i];
            BlockView right(&page->blocks[channels[i]]);
            bool leftNull = left.nulls[address] != 0, rightNull = right.isNull(position);
            if (leftNull != rightNull) return false;
            if (leftNull) continue;
            if (hashTypes[i] == PG_TYPE_VARCHAR) {
                int32_t ll = left.offsets[address + 1] - left.offsets[address], rl = right.sliceLength(position);
                if (ll != rl || memcmp(left.data.data() + left.offsets[address], right.slicePtr(position), (size_t) ll) != 0) return false;
            }
            else if (hashTypes[i] == PG_TYPE_DOUBLE) {
                // DoubleComparisonOperators.DoubleDistinctFromOperator :105-140 → notEqual → TypeUtils.doubleEquals (TypeUtils.java:223-230;
                // use-new-nan-definition defaults to true, FunctionsConfig.java:39): +0 == -0, NaN == NaN
                double a, b; int64_t la = left.getLong(address), lb = right.getLong(position);
                memcpy(&a, &la, 8); memcpy(&b, &lb, 8);
                if (!(a == b || doubleToLongBits(a) == doubleToLongBits(b))) return false;
            }
            else if (left.getLong(address) != right.getLong(position)) return false;
        }
        return true;
    }

    bool positionNotDistinctFromCurrentRow(int64_t address, int32_t hashPosition, int64_t position, const pg_page* page, int8_t rawHash) const   // :460-466
    {
        if (rawHashByHashPosition[hashPosition] != rawHash) return false;
        return positionNotDistinctFromRow(address, position, page);
    }

    bool needRehash() const { return nextGroupId >= maxFill; }

    int32_t putIfAbsent(int64_t position, const pg_page* page)   // :293-322
    {
        int64_t rawHash = hashPositionOfRow(position, page);
        int32_t hashPosition = (int32_t) (murmurHash3(rawHash) & mask);
        int32_t groupId = -1;
        while (groupAddressByHash[hashPosition] != -1) {
            if (positionNotDistinctFromCurrentRow(groupAddressByHash[hashPosition], hashPosition, position, page, (int8_t) rawHash)) {
                groupId = groupIdsByHash[hashPosition];
                break;
            }
            hashPosition = (hashPosition + 1) & mask;
            hashCollisions++;
        }
        if (groupId < 0) groupId = addNewGroup(hashPosition, position, page, rawHash);
        return groupId;
    }

    int32_t addNewGroup(int32_t hashPosition, int64_t position, const pg_page* page, int64_t rawHash)   // :324-358
    {
        for (size_t i = 0; i < channels.size(); i++) keyStore[i].appendFrom(BlockView(&page->blocks[channels[i]]), position);
        if (inputHashChannel >= 0) keyStore[channels.size()].appendLong(rawHash);
        int64_t address = keyStore[0].n - 1;
        int32_t groupId = nextGroupId++;
        groupAddressByHash[hashPosition] = address;
        rawHashByHashPosition[hashPosition] = (int8_t) rawHash;
        groupIdsByHash[hashPosition] = groupId;
        if ((size_t) groupId >= groupAddressByGroupId.size()) groupAddressByGroupId.resize((size_t) groupId + 1);
        groupAddressByGroupId[groupId] = address;
        if (needRehash()) tryRehash();
        return groupId;
    }

    bool tryRehash()   // :380-443
    {
        int64_t newCapacityLong = hashCapacity * 2LL;
        if (newCapacityLong > INT32_MAX) throw OracleError(PG_ERR_INSUFFICIENT_RESOURCES, "Size of hash table cannot exceed 1 billion entries");
        int32_t newCapacity = (int32_t) newCapacityLong;
        int32_t newMask = newCapacity - 1;
        std::vector<int64_t> newKey(newCapacity, -1);
        std::vector<int8_t> rawHashes(newCapacity, 0);
        std::vector<int32_t> newValue(newCapacity, 0);
        int32_t oldIndex = 0;
        for (int32_t groupId = 0; groupId < nextGroupId; groupId++) {
            while (groupAddressByHash[oldIndex] == -1) oldIndex++;
            int64_t address = groupAddressByHash[oldIndex];
            int64_t rawHash = hashPositionOfAddress(address);
            int32_t pos = (int32_t) (murmurHash3(rawHash) & newMask);
            while (newKey[pos] != -1) {
                pos = (pos + 1) & newMask;
                hashCollisions++;
            }
            newKey[pos] = address;
            rawHashes[pos] = (int8_t) rawHash;
            newValue[pos] = groupIdsByHash[oldIndex];
            oldIndex++;
        }
        mask = newMask;
        hashCapacity = newCapacity;
        maxFill = calculateMaxFill(newCapacity);
        groupAddressByHash.swap(newKey);
        rawHashByHashPosition.swap(rawHashes);
        groupIdsByHash.swap(newValue);
        return true;
    }

    int32_t getGroupCount() const override { return nextGroupId; }
    int32_t getCapacity() const override { return hashCapacity; }
    int64_t getHashCollisions() const override { return hashCollisions; }
    std::vector<int> getTypes() const override { return types; }

    // AddNonDictionaryPageWork / GetNonDictionaryGroupIdsWork (:587-626,716-763).  The dictionary and RLE work classes
    // (:628-714,765-865) are result-equivalent optimisations (one probe per dictionary entry / per page).
    void addPage(const pg_page* page) override
    {
        if (needRehash()) tryRehash();
        for (int64_t p = 0; p < page->position_count; p++) putIfAbsent(p, page);
    }
    void getGroupIds(const pg_page* page, int64_t* out) override
    {
        if (needRehash()) tryRehash();
        for (int64_t p = 0; p < page->position_count; p++) out[p] = putIfAbsent(p, page);
    }
    bool contains(int64_t position, const pg_page* page) override   // :262-285
    {
        int64_t rawHash = hashPositionOfRow(position, page);
        int32_t hashPosition = (int32_t) (murmurHash3(rawHash) & mask);
        while (groupAddressByHash[hashPosition] != -1) {
            if (positionNotDistinctFromCurrentRow(groupAddressByHash[hashPosition], hashPosition, position, page, (int8_t) rawHash)) return true;
            hashPosition = (hashPosition + 1) & mask;
        }
        return false;
    }
    void appendValuesTo(int32_t groupId, std::vector<Column>& out, int offset) override   // :199-206 → hashStrategy.appendTo
    {
        int64_t address = groupAddressByGroupId[groupId];
        for (size_t i = 0; i < types.size(); i++) out[offset + i].appendFrom(keyStore[i], address);
    }
};

// GroupByHash.createGroupByHash (MB/operator/GroupByHash.java:46-59)
static std::unique_ptr<GroupByHash> createGroupByHash(const std::vector<int>& hashTypes, const std::vector<int>& hashChannels, int inputHashChannel, int expectedSize)
{
    if (hashTypes.size() == 1 && hashTypes[0] == PG_TYPE_BIGINT && hashChannels.size() == 1)
        return std::make_unique<BigintGroupByHash>(hashChannels[0], inputHashChannel >= 0, expectedSize);
    return std::make_unique<MultiChannelGroupByHash>(hashTypes, hashChannels, inputHashChannel, expectedSize);
}

// ======================================================================================================================
// Grouped accumulators: struct-of-arrays state indexed by group id (StateCompiler.generateGroupedStateClass), one pass
// over the page per aggregate in row order (AccumulatorCompiler.java:462-566).
// ======================================================================================================================
struct GroupedAccumulator {
    pg_aggregate spec;
    int step;
    std::vector<double> doubles;    // NullableDoubleState.double / LongAndDoubleState.double
    std::vector<int64_t> longs;     // LongState / NullableLongState.long / LongAndDoubleState.long
    std::vector<uint8_t> isNull;    // Nullable*State.isNull (initial true)

    void ensureCapacity(int64_t groupCount)
    {
        if ((int64_t) doubles.size() < groupCount) {
            doubles.resize(groupCount, 0.0);
            longs.resize(groupCount, 0);
            isNull.resize(groupCount, 1);
            if (spec.function == PG_AGG_SUM_SHORT_DECIMAL || spec.function == PG_AGG_SUM_LONG_DECIMAL) { decimals.resize(groupCount); overflows.resize(groupCount, 0); }
        }
    }

    static int64_t addExact(int64_t a, int64_t b)   // MB/type/BigintOperators.java:68-78
    {
        int64_t r;
        if (__builtin_add_overflow(a, b, &r))
            throw OracleError(PG_ERR_NUMERIC_OVERFLOW, "bigint addition overflow: " + std::to_string(a) + " + " + std::to_string(b));
        return r;
    }

    // TypeUtils.doubleCompare (PC/type/TypeUtils.java:240-261): NaN is larger than everything, +0 == -0
    static int doubleCompare(double a, double b)
    {
        if (a < b) return -1;
        if (a > b) return 1;
        if (a == b) return 0;
        const int64_t x = doubleToLongBits(a), y = doubleToLongBits(b);
        return x < y ? -1 : (x > y ? 1 : 0);
    }
    // AbstractMinMaxAggregationFunction.compareAndUpdateState :256-287 with the type's LESS_THAN / GREATER_THAN operator
    // (DoubleComparisonOperators.java: doubleCompare(left, right) < 0 / > 0; BigintOperators: plain < / >)
    void minMaxLong(int64_t g, int64_t value, bool isMin)
    {
        if (isNull[g]) { isNull[g] = 0; longs[g] = value; return; }
        if (isMin ? value < longs[g] : value > longs[g]) longs[g] = value;
    }
    void minMaxDouble(int64_t g, double value, bool isMin)
    {
        if (isNull[g]) { isNull[g] = 0; doubles[g] = value; return; }
        const int c = doubleCompare(value, doubles[g]);
        if (isMin ? c < 0 : c > 0) doubles[g] = value;
    }

    // ---- DecimalSumAggregation: LongDecimalWithOverflowState = {Slice longDecimal (null until the first input), long overflow} ----
    // 128-bit unscaled decimals in UnscaledDecimal128Arithmetic's sign-magnitude form: low = long 0, high = long 1 (bit 63 = sign).
    struct Decimal128 { uint64_t low = 0, high = 0; };     // `high` carries the sign bit
    static constexpr uint64_t kSignMask = 0x8000000000000000ULL;
    std::vector<Decimal128> decimals;
    std::vector<int64_t> overflows;
    static Decimal128 unscaledDecimal(int64_t v)           // UnscaledDecimal128Arithmetic.unscaledDecimal(long) :170-183
    {
        Decimal128 d;
        if (v < 0) { d.low = (uint64_t) 0 - (uint64_t) v; d.high = kSignMask; } else d.low = (uint64_t) v;
        return d;
    }
    // UnscaledDecimal128Arithmetic.addWithOverflow :312-337 (addUnsignedReturnOverflow :371-407, subtractUnsigned :409-447,
    // compareAbsolute :693-708): sign-magnitude addition; the magnitude keeps 127 bits, what spills over is returned as +-1.
    static int64_t addWithOverflow(const Decimal128& l, const Decimal128& r, Decimal128& out)
    {
        const bool ln = (l.high & kSignMask) != 0, rn = (r.high & kSignMask) != 0;
        const unsigned __int128 lm = ((unsigned __int128) (l.high & ~kSignMask) << 64) | l.low, rm = ((unsigned __int128) (r.high & ~kSignMask) << 64) | r.low;
        int64_t overflow = 0;
        unsigned __int128 m; bool negative;
        if (ln == rn) {
            m = lm + rm;                                            // < 2^128: no wrap
            overflow = (int64_t) (uint64_t) (m >> 127);             // :401 `intermediateResult >> 31` of the top int
            m &= ~((unsigned __int128) 1 << 127);
            negative = ln;
            if (ln) overflow = -overflow;
        }
        else if (lm > rm) { m = lm - rm; negative = ln; }
        else if (lm < rm) { m = rm - lm; negative = !ln; }
        else { m = 0; negative = false; }                           // setToZero
        out.low = (uint64_t) m; out.high = (uint64_t) (m >> 64) | (negative ? kSignMask : 0);
        return overflow;
    }
    void decimalInput(int64_t g, const Decimal128& v)               // DecimalSumAggregation.inputShortDecimal / inputLongDecimal :145-163
    {
        if (isNull[g]) { isNull[g] = 0; decimals[g] = Decimal128(); }
        overflows[g] += addWithOverflow(decimals[g], v, decimals[g]);
    }
    void decimalCombine(int64_t g, int64_t otherOverflow, const Decimal128& other)     // DecimalSumAggregation.combine :165-177
    {
        int64_t toAdd = otherOverflow;
        if (isNull[g]) { isNull[g] = 0; decimals[g] = other; }
        else toAdd += addWithOverflow(decimals[g], other, decimals[g]);
        overflows[g] += toAdd;
    }
    // UnscaledDecimal128Arithmetic.exceedsOrEqualTenToThirtyEight :1671-1686 (getLong(decimal, 1) is the high long without its sign)
    static bool exceedsOrEqualTenToThirtyEight(const Decimal128& d)
    {
        const int64_t high = (int64_t) (d.high & ~kSignMask);
        if (high >= 0 && high < 0x4b3b4ca85a86c47aLL) return false;
        if (high != 0x4b3b4ca85a86c47aLL) return true;
        const int64_t low = (int64_t) d.low;
        return low < 0 || low >= 0x098a224000000000LL;
    }

    // Number of input channels the FINAL step consumes for this aggregate (our intermediate layout):
    // sum → 1 (nullable value), count → 1, avg → 2 (sum double, count bigint)
    static int intermediateChannels(int fn) { return (fn == PG_AGG_AVG_DOUBLE || fn == PG_AGG_AVG_BIGINT) ? 2 : 1; }

    void addInput(const int64_t* groupIds, const pg_page* page)
    {
        int64_t n = page->position_count;
        BlockView mask;
        bool hasMask = spec.mask_channel >= 0;
        if (hasMask) mask = BlockView(&page->blocks[spec.mask_channel]);
        BlockView in;
        bool hasInput = spec.input_channel >= 0;
        if (hasInput) in = BlockView(&page->blocks[spec.input_channel]);
        for (int64_t p = 0; p < n; p++) {
            if (hasMask && (mask.isNull(p) || mask.getLong(p) == 0)) continue;
            if (hasInput && in.isNull(p)) continue;
            int64_t g = groupIds[p];
            switch (spec.function) {
                case PG_AGG_COUNT_STAR: case PG_AGG_COUNT_COLUMN:
                    longs[g] = longs[g] + 1; break;
                case PG_AGG_SUM_DOUBLE:   // DoubleSumAggregation.sum :33-42
                    if (isNull[g]) { isNull[g] = 0; doubles[g] = in.getDouble(p); } else doubles[g] = doubles[g] + in.getDouble(p);
                    break;
                case PG_AGG_SUM_BIGINT:   // LongSumAggregation.sum :33-38
                    isNull[g] = 0; longs[g] = addExact(longs[g], in.getLong(p)); break;
                case PG_AGG_AVG_DOUBLE:   // AverageAggregations.input(double) :41-46
                    longs[g] = longs[g] + 1; doubles[g] = doubles[g] + in.getDouble(p); break;
                case PG_AGG_AVG_BIGINT:   // AverageAggregations.input(long) :34-39
                    longs[g] = longs[g] + 1; doubles[g] = doubles[g] + (double) in.getLong(p); break;
                case PG_AGG_SUM_SHORT_DECIMAL: decimalInput(g, unscaledDecimal(in.getLong(p))); break;
                case PG_AGG_SUM_LONG_DECIMAL: { const int64_t* w = in.int128Ptr(p); Decimal128 d; d.low = (uint64_t) w[0]; d.high = (uint64_t) w[1]; decimalInput(g, d); break; }
                case PG_AGG_MIN_BIGINT: case PG_AGG_MAX_BIGINT: minMaxLong(g, in.getLong(p), spec.function == PG_AGG_MIN_BIGINT); break;      // AbstractMinMaxAggregationFunction.input :199-202
                case PG_AGG_MIN_DOUBLE: case PG_AGG_MAX_DOUBLE: minMaxDouble(g, in.getDouble(p), spec.function == PG_AGG_MIN_DOUBLE); break;  // :194-197
                default: throw OracleError(PG_ERR_UNSUPPORTED, "aggregate function not in the oracle");
            }
        }
    }

    // FINAL step: combine functions (DoubleSumAggregation.combine :45-55, AverageAggregations.combine :48-53,
    // CountAggregation.combine :41-44, LongSumAggregation.combine :40-50) over intermediate columns starting at `channel`.
    void addIntermediate(const int64_t* groupIds, const pg_page* page, int channel)
    {
        int64_t n = page->position_count;
        BlockView a(&page->blocks[channel]);
        for (int64_t p = 0; p < n; p++) {
            int64_t g = groupIds[p];
            switch (spec.function) {
                case PG_AGG_COUNT_STAR: case PG_AGG_COUNT_COLUMN:
                    longs[g] += a.getLong(p); break;
                case PG_AGG_SUM_DOUBLE:
                    if (a.isNull(p)) break;
                    if (isNull[g]) { isNull[g] = 0; doubles[g] = a.getDouble(p); } else doubles[g] += a.getDouble(p);
                    break;
                case PG_AGG_SUM_BIGINT:
                    if (a.isNull(p)) break;
                    if (isNull[g]) { isNull[g] = 0; longs[g] = a.getLong(p); } else longs[g] = addExact(longs[g], a.getLong(p));
                    break;
                case PG_AGG_AVG_DOUBLE: case PG_AGG_AVG_BIGINT: {
                    BlockView c(&page->blocks[channel + 1]);
                    longs[g] += c.getLong(p); doubles[g] += a.getDouble(p); break;
                }
                case PG_AGG_SUM_SHORT_DECIMAL: case PG_AGG_SUM_LONG_DECIMAL: {     // LongDecimalWithOverflowStateSerializer.deserialize :52-66
                    if (a.isNull(p)) break;
                    if (a.sliceLength(p) != 24) throw OracleError(PG_ERR_INVALID_STATE, "Unexpected serialized state size: " + std::to_string(a.sliceLength(p)));
                    int64_t w[3]; memcpy(w, a.slicePtr(p), 24);
                    Decimal128 d; d.low = (uint64_t) w[1]; d.high = (uint64_t) w[2];
                    decimalCombine(g, w[0], d);
                    break;
                }
                case PG_AGG_MIN_BIGINT: case PG_AGG_MAX_BIGINT:      // AbstractMinMaxAggregationFunction.combine :225-228 (NULL intermediate states are skipped)
                    if (!a.isNull(p)) minMaxLong(g, a.getLong(p), spec.function == PG_AGG_MIN_BIGINT);
                    break;
                case PG_AGG_MIN_DOUBLE: case PG_AGG_MAX_DOUBLE:
                    if (!a.isNull(p)) minMaxDouble(g, a.getDouble(p), spec.function == PG_AGG_MIN_DOUBLE);
                    break;
                default: throw OracleError(PG_ERR_UNSUPPORTED, "aggregate function not in the oracle");
            }
        }
    }

    void outputTypes(std::vector<int>& out) const
    {
        bool partial = step == PG_STEP_PARTIAL;
        switch (spec.function) {
            case PG_AGG_COUNT_STAR: case PG_AGG_COUNT_COLUMN: case PG_AGG_SUM_BIGINT: case PG_AGG_MIN_BIGINT: case PG_AGG_MAX_BIGINT: out.push_back(PG_TYPE_BIGINT); break;
            case PG_AGG_SUM_DOUBLE: case PG_AGG_MIN_DOUBLE: case PG_AGG_MAX_DOUBLE: out.push_back(PG_TYPE_DOUBLE); break;
            case PG_AGG_AVG_DOUBLE: case PG_AGG_AVG_BIGINT:
                out.push_back(PG_TYPE_DOUBLE); if (partial) out.push_back(PG_TYPE_BIGINT); break;
            case PG_AGG_SUM_SHORT_DECIMAL: case PG_AGG_SUM_LONG_DECIMAL: out.push_back(partial ? PG_TYPE_VARCHAR : PG_TYPE_LONG_DECIMAL); break;
            default: throw OracleError(PG_ERR_UNSUPPORTED, "aggregate function not in the oracle");
        }
    }

    void evaluate(int32_t g, std::vector<Column>& out, int& offset) const
    {
        bool partial = step == PG_STEP_PARTIAL;
        switch (spec.function) {
            case PG_AGG_COUNT_STAR: case PG_AGG_COUNT_COLUMN: out[offset++].appendLong(longs[g]); break;
            case PG_AGG_SUM_DOUBLE: case PG_AGG_MIN_DOUBLE: case PG_AGG_MAX_DOUBLE: if (isNull[g]) out[offset++].appendNull(); else out[offset++].appendDouble(doubles[g]); break;
            case PG_AGG_SUM_BIGINT: case PG_AGG_MIN_BIGINT: case PG_AGG_MAX_BIGINT: if (isNull[g]) out[offset++].appendNull(); else out[offset++].appendLong(longs[g]); break;
            case PG_AGG_AVG_DOUBLE: case PG_AGG_AVG_BIGINT:
                if (partial) { out[offset++].appendDouble(doubles[g]); out[offset++].appendLong(longs[g]); }
                else if (longs[g] == 0) out[offset++].appendNull();                       // AverageAggregations.output :55-66
                else out[offset++].appendDouble(doubles[g] / (double) longs[g]);
                break;
            case PG_AGG_SUM_SHORT_DECIMAL: case PG_AGG_SUM_LONG_DECIMAL:
                if (isNull[g]) { out[offset++].appendNull(); break; }
                if (partial) {      // LongDecimalWithOverflowStateSerializer.serialize :37-49
                    const int64_t w[3] = {overflows[g], (int64_t) decimals[g].low, (int64_t) decimals[g].high};
                    out[offset++].appendSlice((const uint8_t*) w, 24);
                }
                else {              // DecimalSumAggregation.outputLongDecimal :179-193
                    if (overflows[g] != 0 || exceedsOrEqualTenToThirtyEight(decimals[g])) throw OracleError(PG_ERR_NUMERIC_OVERFLOW, "Decimal overflow");
                    out[offset++].appendInt128((int64_t) decimals[g].low, (int64_t) decimals[g].high);
                }
                break;
            default: throw OracleError(PG_ERR_UNSUPPORTED, "aggregate function not in the oracle");
        }
    }
};

// ======================================================================================================================
// HashAggregationOperator (in-memory builder, no spill, no partial flush)
// ======================================================================================================================
struct HashAggregationOperator : Operator {
    std::unique_ptr<GroupByHash> groupByHash;
    std::vector<GroupedAccumulator> aggregators;
    int step;
    bool finishing = false, finished = false, outputStarted = false;
    int32_t nextOutputGroup = 0;
    int64_t maxOutputRows;
    std::vector<int64_t> groupIdScratch;
    int nGroupChannelsIn;   // FINAL: intermediate channels start after the group channels (+hash)

    explicit HashAggregationOperator(const pg_hash_aggregation_config* c)
    {
        std::vector<int> types(c->group_types, c->group_types + c->n_group_channels);
        std::vector<int> chans(c->group_channels, c->group_channels + c->n_group_channels);
        groupByHash = createGroupByHash(types, chans, c->hash_channel, c->expected_groups);
        step = c->step;
        for (int i = 0; i < c->n_aggregates; i++) {
            GroupedAccumulator a; a.spec = c->aggregates[i]; a.step = step;
            aggregators.push_back(a);
        }
        maxOutputRows = c->options.max_output_rows > 0 ? c->options.max_output_rows : (1 << 20);
        nGroupChannelsIn = 0;
    }

    bool needsInput() override { return !finishing && !outputStarted; }   // HashAggregationOperator.java:403-410
    void addInput(const pg_page* page) override                              // :413-431 → InMemoryHashAggregationBuilder.processPage :204-220
    {
        if (!needsInput()) throw OracleError(PG_ERR_INVALID_STATE, "Operator is already finishing");
        stats.input_positions += page->position_count;
        groupIdScratch.resize((size_t) page->position_count);
        groupByHash->getGroupIds(page, groupIdScratch.data());
        int64_t groupCount = groupByHash->getGroupCount();
        int channel = 0;
        for (auto& a : aggregators) {
            a.ensureCapacity(groupCount);
            if (step == PG_STEP_FINAL) {
                // FINAL input channel = aggregate.input_channel is the first intermediate column of this aggregate
                a.addIntermediate(groupIdScratch.data(), page, a.spec.input_channel);
            }
            else a.addInput(groupIdScratch.data(), page);
            channel++;
        }
    }
    void finish() override { finishing = true; }
    bool isFinished() override { return finished; }
    bool getOutput(OwnedPage& out) override   // :451-501 + buildResult :326-352 (ascending group id)
    {
        if (finished || !finishing) return false;
        outputStarted = true;
        int32_t groupCount = groupByHash->getGroupCount();
        if (nextOutputGroup >= groupCount) { finished = true; return false; }
        std::vector<int> types = groupByHash->getTypes();
        for (auto& a : aggregators) { a.ensureCapacity(groupCount); a.outputTypes(types); }
        out.columns.clear();
        for (int t : types) out.columns.emplace_back(t);
        int keyCols = (int) groupByHash->getTypes().size();
        int64_t emitted = 0;
        while (emitted < maxOutputRows && nextOutputGroup < groupCount) {
            int32_t g = nextOutputGroup++;
            groupByHash->appendValuesTo(g, out.columns, 0);
            int offset = keyCols;
            for (auto& a : aggregators) a.evaluate(g, out.columns, offset);
            emitted++;
        }
        out.positionCount = emitted;
        stats.output_positions += emitted;
        stats.group_count = groupCount;
        if (nextOutputGroup >= groupCount) finished = true;
        return true;
    }
};

Operator* createHashAggregation(const pg_hash_aggregation_config* c) { return new HashAggregationOperator(c); }

// Standalone GroupByHash handle for or_group_by_hash_*
struct GroupByHashHandle {
    std::unique_ptr<GroupByHash> h;
    OwnedPage keys;
};

GroupByHashHandle* createGroupByHashHandle(const pg_group_by_hash_config* c)
{
    auto* r = new GroupByHashHandle();
    std::vector<int> types(c->types, c->types + c->n_channels);
    std::vector<int> chans(c->channels, c->channels + c->n_channels);
    r->h = createGroupByHash(types, chans, c->hash_channel, c->expected_size);
    return r;
}

}  // namespace oracle

using namespace oracle;

extern "C" {

#define OR_TRY try {
#define OR_CATCH } catch (const OracleError& e) { or_set_error(e.what()); return e.code; } catch (const std::exception& e) { or_set_error(e.what()); return PG_ERR_INVALID_STATE; } return PG_OK;

void or_set_error(const char* msg);

int32_t or_group_by_hash_create(const pg_group_by_hash_config* config, void** out) { OR_TRY *out = createGroupByHashHandle(config); OR_CATCH }
int32_t or_group_by_hash_add_page(void* h, const pg_page* page) { OR_TRY ((GroupByHashHandle*) h)->h->addPage(page); OR_CATCH }
int32_t or_group_by_hash_get_group_ids(void* h, const pg_page* page, int64_t* out) { OR_TRY ((GroupByHashHandle*) h)->h->getGroupIds(page, out); OR_CATCH }
int32_t or_group_by_hash_get_group_count(void* h, int64_t* out) { OR_TRY *out = ((GroupByHashHandle*) h)->h->getGroupCount(); OR_CATCH }
int32_t or_group_by_hash_get_capacity(void* h, int64_t* out) { OR_TRY *out = ((GroupByHashHandle*) h)->h->getCapacity(); OR_CATCH }
int32_t or_group_by_hash_contains(void* h, const pg_page* page, uint8_t* out)
{
    OR_TRY
    for (int64_t p = 0; p < page->position_count; p++) out[p] = ((GroupByHashHandle*) h)->h->contains(p, page) ? 1 : 0;
    OR_CATCH
}
int32_t or_group_by_hash_get_keys(void* hh, pg_page* out)
{
    OR_TRY
    auto* h = (GroupByHashHandle*) hh;
    h->keys.columns.clear();
    for (int t : h->h->getTypes()) h->keys.columns.emplace_back(t);
    int32_t n = h->h->getGroupCount();
    for (int32_t g = 0; g < n; g++) h->h->appendValuesTo(g, h->keys.columns, 0);
    h->keys.positionCount = n;
    h->keys.fill(out);
    OR_CATCH
}
int32_t or_group_by_hash_destroy(void* h) { delete (GroupByHashHandle*) h; return PG_OK; }

}  // extern "C"
