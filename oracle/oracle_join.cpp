/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 *
 * HashBuilderOperator → PagesIndex → PagesHash (+ArrayPositionLinks) → JoinHash → LookupJoinOperator restated from:
 *   MB/operator/PagesIndex.java:202-226          (addPage: keep blocks, append synthetic addresses)
 *   MB/operator/PagesHash.java:55-132,169-181    (build loop / getAddressIndex)
 *   MB/operator/ArrayPositionLinks.java:38-100   (link / next)
 *   MB/operator/JoinHash.java:81-124             (getJoinPosition / getNextJoinPosition)
 *   MB/operator/JoinProbe.java:74-119            (null probe keys never match)
 *   MB/operator/LookupJoinOperator.java:481-505,567-616 (processProbe / joinCurrentPosition / outerJoinCurrentPosition)
 *   MB/operator/LookupJoinPageBuilder.java:75-137 (appendRow / appendNullForBuild)
 *   MB/operator/SimplePagesHashStrategy.java:110-175 (hashPosition = 31*h + typeHash; equality ignore nulls)
 */
#include "oracle_common.h"

namespace oracle {

// The build side after finish(): the reference's LookupSource (JoinHash over PagesHash).
struct LookupSource {
    // PagesIndex: one flat column per build channel (the reference keeps a list of blocks per channel plus
    // valueAddresses[i] = pageIndex<<32|position; concatenating the pages makes address == position).
    std::vector<int> types;
    std::vector<Column> channels;
    std::vector<int> hashChannels;
    std::vector<int> outputChannels;
    int precomputedHashChannel = -1;
    int64_t positionCount = 0;

    // PagesHash
    int32_t mask = 0;
    std::vector<int32_t> key;
    std::vector<int8_t> positionToHashes;
    std::vector<int32_t> positionLinks;    // ArrayPositionLinks (-1 end)
    bool linksEmpty = true;
    int64_t hashCollisions = 0;
    int64_t positionIsNullCount = 0;

    int64_t hashPosition(int64_t position) const   // SimplePagesHashStrategy.hashPosition :110-123
    {
        if (precomputedHashChannel >= 0) return channels[precomputedHashChannel].getLong(position);
        int64_t result = 0;
        for (int hc : hashChannels) {
            int64_t h = channels[hc].nulls[position] ? 0 : valueHash(hc, position);
            result = result * 31 + h;
        }
        return result;
    }
    // type.hash of a build value: AbstractLongType.hash for the integer types, XxHash64 of the slice for VARCHAR
    // (AbstractVariableWidthBlock.java:102-106)
    int64_t valueHash(int hc, int64_t position) const
    {
        const Column& c = channels[hc];
        if (c.type == PG_TYPE_VARCHAR) return xxHash64(c.data.data() + c.offsets[position], (size_t) (c.offsets[position + 1] - c.offsets[position]));
        return hashLong(c.getLong(position));
    }
    // type.equalTo: slices compare by length and bytes (AbstractVariableWidthBlock.equals :108-117)
    static bool slicesEqual(const uint8_t* a, int32_t la, const uint8_t* b, int32_t lb) { return la == lb && (la == 0 || memcmp(a, b, (size_t) la) == 0); }
    bool isPositionNull(int64_t position) const
    {
        for (int hc : hashChannels) if (channels[hc].nulls[position]) return true;
        return false;
    }
    bool positionEqualsPositionIgnoreNulls(int64_t l, int64_t r) const
    {
        for (int hc : hashChannels) {
            const Column& c = channels[hc];
            if (c.type == PG_TYPE_VARCHAR) {
                if (!slicesEqual(c.data.data() + c.offsets[l], c.offsets[l + 1] - c.offsets[l], c.data.data() + c.offsets[r], c.offsets[r + 1] - c.offsets[r])) return false;
            }
            else if (c.getLong(l) != c.getLong(r)) return false;
        }
        return true;
    }
    bool positionEqualsRowIgnoreNulls(int64_t l, int64_t rightPosition, const std::vector<BlockView>& probeHashBlocks) const
    {
        for (size_t i = 0; i < hashChannels.size(); i++) {
            const Column& c = channels[hashChannels[i]];
            if (c.type == PG_TYPE_VARCHAR) {
                if (!slicesEqual(c.data.data() + c.offsets[l], c.offsets[l + 1] - c.offsets[l], probeHashBlocks[i].slicePtr(rightPosition), probeHashBlocks[i].sliceLength(rightPosition))) return false;
            }
            else if (c.getLong(l) != probeHashBlocks[i].getLong(rightPosition)) return false;
        }
        return true;
    }

    void build()   // PagesHash ctor :55-132
    {
        int32_t hashSize = arraySize(positionCount, 0.75f);
        mask = hashSize - 1;
        key.assign(hashSize, -1);
        positionToHashes.assign((size_t) positionCount, 0);
        positionLinks.assign((size_t) positionCount, -1);
        int64_t positionsInStep = std::min<int64_t>(positionCount + 1, (128 * 1024) / 32);   // CACHE_SIZE / Integer.SIZE :76
        std::vector<int64_t> positionToFullHashes((size_t) positionsInStep);
        for (int64_t step = 0; step * positionsInStep < positionCount; step++) {
            int64_t stepBegin = step * positionsInStep;
            int64_t stepEnd = std::min((step + 1) * positionsInStep, positionCount);
            int64_t stepSize = stepEnd - stepBegin;
            for (int64_t position = 0; position < stepSize; position++) {
                int64_t realPosition = position + stepBegin;
                int64_t hash = hashPosition(realPosition);
                positionToFullHashes[position] = hash;
                positionToHashes[realPosition] = (int8_t) hash;
            }
            for (int64_t position = 0; position < stepSize; position++) {
                int32_t realPosition = (int32_t) (position + stepBegin);
                if (isPositionNull(realPosition)) { ++positionIsNullCount; continue; }
                int64_t hash = positionToFullHashes[position];
                int32_t pos = (int32_t) (murmurHash3(hash) & mask);
                while (key[pos] != -1) {
                    int32_t currentKey = key[pos];
                    if (((int8_t) hash) == positionToHashes[currentKey] && positionEqualsPositionIgnoreNulls(currentKey, realPosition)) {
                        // positionLinks.link(realPosition, currentKey): ArrayPositionLinks.java:45-50
                        positionLinks[realPosition] = currentKey;
                        linksEmpty = false;
                        break;
                    }
                    pos = (pos + 1) & mask;
                    hashCollisions++;
                }
                key[pos] = realPosition;
            }
        }
    }

    int32_t getAddressIndex(int64_t rightPosition, const std::vector<BlockView>& probeHashBlocks, int64_t rawHash) const   // :169-181
    {
        int32_t pos = (int32_t) (murmurHash3(rawHash) & mask);
        while (key[pos] != -1) {
            if (positionToHashes[key[pos]] == (int8_t) rawHash && positionEqualsRowIgnoreNulls(key[pos], rightPosition, probeHashBlocks)) return key[pos];
            pos = (pos + 1) & mask;
        }
        return -1;
    }
};

// oracle_exchange_project.cpp: the RowExpression interpreter evaluating a JoinFilterFunction
struct JoinFilterEvaluator;
JoinFilterEvaluator* createJoinFilter(const pg_filter_project_config* c, int buildChannels);
void destroyJoinFilter(JoinFilterEvaluator* e);
bool joinFilterEligible(JoinFilterEvaluator* e, const pg_page* buildThenProbe, int64_t buildPosition, int64_t probePosition);

struct JoinBridge {
    JoinFilterEvaluator* filter = nullptr;        // JoinHash.filterFunction (handed over by the build side's factory)
    ~JoinBridge() { if (filter) destroyJoinFilter(filter); }
    std::unique_ptr<LookupSource> lookupSource;   // set when the build finishes (the future completes)
    // OuterLookupSource.OuterPositionTracker :165-203 (boolean[] visitedPositions) + the outerPositionsFuture of
    // PartitionedLookupSourceFactory: complete once the probe factory is closed and every probe operator is done
    bool outer = false, noMoreProbes = false;
    int liveProbes = 0;
    std::vector<uint8_t> visitedPositions;
    bool outerPositionsReady() const { return lookupSource != nullptr && noMoreProbes && liveProbes == 0; }
};

struct HashBuilderOperator : Operator {
    JoinBridge* bridge;
    std::unique_ptr<LookupSource> ls;
    bool finishing = false;

    HashBuilderOperator(const pg_hash_builder_config* c, JoinBridge* b) : bridge(b)
    {
        ls = std::make_unique<LookupSource>();
        ls->types.assign(c->types, c->types + c->n_channels);
        for (int t : ls->types) ls->channels.emplace_back(t);
        ls->hashChannels.assign(c->hash_channels, c->hash_channels + c->n_hash_channels);
        ls->outputChannels.assign(c->output_channels, c->output_channels + c->n_output_channels);
        ls->precomputedHashChannel = c->precomputed_hash_channel;
        if (c->filter && c->filter->filter_node >= 0) bridge->filter = createJoinFilter(c->filter, c->n_channels);
        for (int hc : ls->hashChannels) {
            int t = ls->types[hc];
            if (t != PG_TYPE_BIGINT && t != PG_TYPE_INTEGER && t != PG_TYPE_DATE && t != PG_TYPE_VARCHAR) throw OracleError(PG_ERR_UNSUPPORTED, "join key type");
        }
    }
    bool needsInput() override { return !finishing; }
    void addInput(const pg_page* page) override   // HashBuilderOperator.addInput :333-349 → PagesIndex.addPage :202-226
    {
        if (finishing) throw OracleError(PG_ERR_INVALID_STATE, "Operator is already finishing");
        for (size_t ch = 0; ch < ls->channels.size(); ch++) {
            BlockView v(&page->blocks[ch]);
            for (int64_t p = 0; p < page->position_count; p++) ls->channels[ch].appendFrom(v, p);
        }
        ls->positionCount += page->position_count;
        stats.input_positions += page->position_count;
    }
    void finish() override   // finishInput :519-537 → buildLookupSource :642-651
    {
        if (finishing) return;
        finishing = true;
        ls->build();
        stats.group_count = ls->positionCount;
        if (bridge->outer) bridge->visitedPositions.assign((size_t) ls->positionCount, 0);
        bridge->lookupSource = std::move(ls);
    }
    bool isFinished() override { return finishing; }
    bool getOutput(OwnedPage&) override { return false; }
};

struct LookupJoinOperator : Operator {
    JoinBridge* bridge;
    int joinType;
    std::vector<int> types, hashChannels, outputChannels;
    int precomputedHashChannel;
    bool finishing = false;
    bool hasPending = false;
    OwnedPage pending;

    bool tracked = false;
    ~LookupJoinOperator() override { if (tracked) bridge->liveProbes--; }
    LookupJoinOperator(const pg_lookup_join_config* c, JoinBridge* b) : bridge(b)
    {
        if ((c->join_type == PG_JOIN_LOOKUP_OUTER || c->join_type == PG_JOIN_FULL_OUTER) && !b->outer)
            throw OracleError(PG_ERR_INVALID_ARGUMENT, "LOOKUP_OUTER / FULL_OUTER joins need an outer bridge");
        // LookupJoinOperators.java:30-39: the probe of LOOKUP_OUTER joins like INNER, of FULL_OUTER like PROBE_OUTER
        joinType = (c->join_type == PG_JOIN_PROBE_OUTER || c->join_type == PG_JOIN_FULL_OUTER) ? PG_JOIN_PROBE_OUTER : PG_JOIN_INNER;
        if (b->outer) { tracked = true; b->liveProbes++; }
        types.assign(c->types, c->types + c->n_channels);
        hashChannels.assign(c->hash_channels, c->hash_channels + c->n_hash_channels);
        outputChannels.assign(c->output_channels, c->output_channels + c->n_output_channels);
        precomputedHashChannel = c->precomputed_hash_channel;
    }
    bool needsInput() override { return !finishing && !hasPending && bridge->lookupSource != nullptr; }   // LookupJoinOperator.java:198-210
    bool isBlocked() override { return bridge->lookupSource == nullptr; }      // the lookupSourceProviderFuture
    void addInput(const pg_page* page) override
    {
        if (!needsInput()) throw OracleError(PG_ERR_INVALID_STATE, "lookup join does not need input");
        const LookupSource& ls = *bridge->lookupSource;
        stats.input_positions += page->position_count;
        pending.columns.clear();
        for (int oc : outputChannels) pending.columns.emplace_back(types[oc]);
        for (int oc : ls.outputChannels) pending.columns.emplace_back(ls.types[oc]);
        std::vector<BlockView> probeHash;
        for (int hc : hashChannels) probeHash.emplace_back(&page->blocks[hc]);
        std::vector<BlockView> probeOut;
        for (int oc : outputChannels) probeOut.emplace_back(&page->blocks[oc]);
        BlockView hashBlock;
        if (precomputedHashChannel >= 0) hashBlock = BlockView(&page->blocks[precomputedHashChannel]);
        int64_t rows = 0;
        size_t nProbeOut = outputChannels.size();
        // the filter function's view: build channels, then all probe channels (createJoinSourcesLayout)
        std::vector<pg_block> filterBlocks;
        pg_page filterPage{};
        if (bridge->filter) {
            for (const Column& c : ls.channels) filterBlocks.push_back(c.asBlock());
            for (int ch = 0; ch < page->channel_count; ch++) filterBlocks.push_back(page->blocks[ch]);
            filterPage.blocks = filterBlocks.data(); filterPage.channel_count = (int32_t) filterBlocks.size(); filterPage.position_count = page->position_count;
        }
        for (int64_t position = 0; position < page->position_count; position++) {   // processProbe :481-505
            // JoinProbe.getCurrentJoinPosition :79-90
            int64_t joinPosition;
            bool rowNull = false;
            for (auto& b : probeHash) if (b.isNull(position)) { rowNull = true; break; }
            if (rowNull) joinPosition = -1;
            else {
                int64_t rawHash;
                if (precomputedHashChannel >= 0) rawHash = hashBlock.getLong(position);
                else {   // SimplePagesHashStrategy.hashRow :125-136
                    rawHash = 0;
                    for (auto& b : probeHash) rawHash = rawHash * 31 + (b.b->type == PG_TYPE_VARCHAR ? xxHash64(b.slicePtr(position), (size_t) b.sliceLength(position)) : hashLong(b.getLong(position)));
                }
                joinPosition = ls.getAddressIndex(position, probeHash, rawHash);
            }
            bool producedRow = false;
            while (joinPosition >= 0) {   // joinCurrentPosition :567-586
                if (bridge->filter && !joinFilterEligible(bridge->filter, &filterPage, joinPosition, position)) {   // JoinHash.isJoinPositionEligible :113-118
                    joinPosition = ls.linksEmpty ? -1 : ls.positionLinks[joinPosition];
                    continue;
                }
                producedRow = true;
                for (size_t i = 0; i < nProbeOut; i++) pending.columns[i].appendFrom(probeOut[i], position);
                for (size_t i = 0; i < ls.outputChannels.size(); i++) pending.columns[nProbeOut + i].appendFrom(ls.channels[ls.outputChannels[i]], joinPosition);
                if (bridge->outer) bridge->visitedPositions[joinPosition] = 1;     // OuterLookupSource.appendTo :100-105
                rows++;
                joinPosition = ls.linksEmpty ? -1 : ls.positionLinks[joinPosition];   // JoinHash.getNextJoinPosition :105-111
            }
            if (!producedRow && joinType == PG_JOIN_PROBE_OUTER) {   // outerJoinCurrentPosition :608-616 → appendNullForBuild
                for (size_t i = 0; i < nProbeOut; i++) pending.columns[i].appendFrom(probeOut[i], position);
                for (size_t i = 0; i < ls.outputChannels.size(); i++) pending.columns[nProbeOut + i].appendNull();
                rows++;
            }
        }
        pending.positionCount = rows;
        hasPending = rows > 0;
        stats.output_positions += rows;
    }
    bool getOutput(OwnedPage& out) override
    {
        if (!hasPending) return false;
        out.columns.swap(pending.columns);
        out.positionCount = pending.positionCount;
        hasPending = false;
        return true;
    }
    void finish() override { finishing = true; if (tracked) { tracked = false; bridge->liveProbes--; } }
    bool isFinished() override { return finishing && !hasPending; }
};

// LookupOuterOperator (MB/operator/LookupOuterOperator.java:109-225) over SharedLookupOuterPositionIterator.appendToNext
// (OuterLookupSource.java:139-151): unvisited build positions in ascending order, NULL probe channels + build output channels.
struct LookupOuterOperator : Operator {
    JoinBridge* bridge;
    std::vector<int> probeTypes;
    int64_t currentPosition = 0, maxRows;
    bool closed = false;
    LookupOuterOperator(const pg_lookup_outer_config* c, JoinBridge* b) : bridge(b)
    {
        if (!b->outer) throw OracleError(PG_ERR_INVALID_ARGUMENT, "the bridge was not created for outer tracking");
        probeTypes.assign(c->probe_output_types, c->probe_output_types + c->n_probe_output_channels);
        maxRows = c->options.max_output_rows > 0 ? c->options.max_output_rows : (int64_t(1) << 20);
    }
    bool needsInput() override { return false; }
    void addInput(const pg_page*) override { throw OracleError(PG_ERR_UNSUPPORTED, "LookupOuterOperator is a source operator"); }
    bool isBlocked() override { return !closed && !bridge->outerPositionsReady(); }
    void finish() override { closed = true; }
    bool isFinished() override { return closed; }
    bool getOutput(OwnedPage& out) override
    {
        if (closed || !bridge->outerPositionsReady()) return false;
        const LookupSource& ls = *bridge->lookupSource;
        out.columns.clear();
        for (int t : probeTypes) out.columns.emplace_back(t);
        for (int oc : ls.outputChannels) out.columns.emplace_back(ls.types[oc]);
        int64_t rows = 0;
        while (rows < maxRows && currentPosition < (int64_t) bridge->visitedPositions.size()) {      // pageBuilder.isFull() stands in for maxRows
            if (!bridge->visitedPositions[currentPosition]) {
                for (size_t i = 0; i < probeTypes.size(); i++) out.columns[i].appendNull();
                for (size_t i = 0; i < ls.outputChannels.size(); i++) out.columns[probeTypes.size() + i].appendFrom(ls.channels[ls.outputChannels[i]], currentPosition);
                rows++;
            }
            currentPosition++;
        }
        if (currentPosition >= (int64_t) bridge->visitedPositions.size()) closed = true;      // outputPositionsFinished -> close()
        out.positionCount = rows;
        stats.output_positions += rows;
        return rows > 0;
    }
};

// HashSemiJoinOperator.addInput (MB/operator/HashSemiJoinOperator.java:142-192).  The reference probes a ChannelSet (a GroupByHash
// over the set channel, ChannelSet.java:33-77); membership is the same question the PagesHash above answers, so the set is the
// LookupSource a HashBuilderOperator built from the set channel: isEmpty() = no rows at all (NULL is a group of its own in the
// GroupByHash), containsNull() = some set row was NULL.
struct SemiJoinOperator : Operator {
    JoinBridge* bridge;
    std::vector<int> types;
    int probeChannel;
    bool finishing = false, hasPending = false;
    OwnedPage pending;

    SemiJoinOperator(const pg_semi_join_config* c, JoinBridge* b) : bridge(b)
    {
        types.assign(c->types, c->types + c->n_channels);
        probeChannel = c->probe_join_channel;
    }
    bool needsInput() override { return !finishing && !hasPending && bridge->lookupSource != nullptr; }
    bool isBlocked() override { return bridge->lookupSource == nullptr; }
    void addInput(const pg_page* page) override
    {
        if (!needsInput()) throw OracleError(PG_ERR_INVALID_STATE, "semi join does not need input (set not built, finishing, or output pending)");
        const LookupSource& set = *bridge->lookupSource;
        const bool setEmpty = set.positionCount == 0, setHasNull = set.positionIsNullCount > 0;
        stats.input_positions += page->position_count;
        pending.columns.clear();
        for (int t : types) pending.columns.emplace_back(t);
        pending.columns.emplace_back(PG_TYPE_BOOLEAN);
        std::vector<BlockView> in;
        for (size_t ch = 0; ch < types.size(); ch++) in.emplace_back(&page->blocks[ch]);
        std::vector<BlockView> probe{BlockView(&page->blocks[probeChannel])};
        Column& flag = pending.columns.back();
        for (int64_t position = 0; position < page->position_count; position++) {
            for (size_t ch = 0; ch < types.size(); ch++) pending.columns[ch].appendFrom(in[ch], position);      // page.appendColumn keeps the input
            if (probe[0].isNull(position)) {                                  // :161-168
                if (setEmpty) flag.appendLong(0); else flag.appendNull();
            }
            else {
                const int64_t rawHash = probe[0].b->type == PG_TYPE_VARCHAR ? xxHash64(probe[0].slicePtr(position), (size_t) probe[0].sliceLength(position))
                                                                            : hashLong(probe[0].getLong(position));
                const bool contains = !setEmpty && set.getAddressIndex(position, probe, rawHash) >= 0;   // :170-177
                if (!contains && setHasNull) flag.appendNull();               // :178-183
                else flag.appendLong(contains ? 1 : 0);
            }
        }
        pending.positionCount = page->position_count;
        hasPending = page->position_count > 0;
        stats.output_positions += page->position_count;
    }
    bool getOutput(OwnedPage& out) override
    {
        if (!hasPending) return false;
        out.columns.swap(pending.columns);
        out.positionCount = pending.positionCount;
        hasPending = false;
        return true;
    }
    void finish() override { finishing = true; }
    bool isFinished() override { return finishing && !hasPending; }
};

JoinBridge* createJoinBridge() { return new JoinBridge(); }
Operator* createSemiJoin(const pg_semi_join_config* c, JoinBridge* b) { return new SemiJoinOperator(c, b); }
Operator* createHashBuilder(const pg_hash_builder_config* c, JoinBridge* b) { return new HashBuilderOperator(c, b); }
Operator* createLookupJoin(const pg_lookup_join_config* c, JoinBridge* b) { return new LookupJoinOperator(c, b); }
Operator* createLookupOuter(const pg_lookup_outer_config* c, JoinBridge* b) { return new LookupOuterOperator(c, b); }
void joinBridgeSetOuter(JoinBridge* b) { if (b->lookupSource) throw OracleError(PG_ERR_INVALID_STATE, "the build side has already finished"); b->outer = true; }
void joinBridgeNoMoreProbes(JoinBridge* b) { b->noMoreProbes = true; }
int64_t joinBridgePositionCount(JoinBridge* b) { return b->lookupSource ? b->lookupSource->positionCount : -1; }
void destroyJoinBridge(JoinBridge* b) { delete b; }

}  // namespace oracle
