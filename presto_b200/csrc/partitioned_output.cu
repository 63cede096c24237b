// presto_b200 — PartitionedOutputOperator (row repartition) on sm_100a.
//
// Replaces (MB = presto-main-base/src/main/java/com/facebook/presto/):
//   MB/operator/repartition/OptimizedPartitionedOutputOperator.java:487-557  partitionPage: partition function per row →
//        per-partition position lists (row order), replicate-any-row / null-channel replication
//   …:664-716 PartitionBuffer.appendData + LongArrayBlockEncodingBuffer.java:154-175  copy values by position list
//   MB/operator/HashGenerator.java:22-29 getPartition, InterpretedHashGenerator.java:70-79, PrecomputedHashGenerator.java:31-35
//
// GPU shape: a stable multisplit.  Every warp owns a contiguous run of rows ("unit").
//   pass 1  partition id per row (exact HashGenerator.getPartition) → ids[] (2 B/row) + per-unit histogram in shared memory
//   scan    exclusive sum over (partition-major, unit-minor) histograms = start of every unit's run inside every partition
//   pass 2  each warp walks its rows in order, ranks lanes of the same partition with match/popc, scatters every column
// The result is all partitions back to back, rows of a partition in input order = what the reference's position lists give.
// Across GPUs each partition slab is then the send buffer of one all-to-all peer (presto_b200/exchange.py).
#include "common.cuh"
#include <cub/device/device_scan.cuh>
#include <algorithm>

namespace pg {

constexpr int kMaxPartChannels = 8;
constexpr int kMaxScatterColumns = 32;
constexpr int kRowsPerUnit = 1024;       // rows owned by one warp (their output positions live in shared memory during pass 2)
constexpr int kPartWarps = 8;

struct HashCol { const void* values; const uint8_t* nulls; const int32_t* offsets; int32_t type; };   // offsets: VARCHAR

struct PartitionIdParams {
    HashCol cols[kMaxPartChannels];
    int32_t nCols;
    const int64_t* precomputed;          // $hashvalue channel or nullptr
    const uint8_t* nullChannelNulls;     // rows whose null channel is NULL go to every partition
    int32_t replicateRow0;
    int32_t partitionCount;
    int32_t localFunction;               // 1: LocalPartitionGenerator instead of HashGenerator.getPartition; 2: RoundRobinBucketFunction
    int32_t bucketCount;                 // the function's range (= partitionCount without a bucket table)
    const int32_t* bucketToPartition;    // BucketPartitionFunction's table, or nullptr = identity
    uint32_t roundRobinBase;             // RoundRobinBucketFunction's counter at row 0 of this page
    int64_t n;
    int64_t units;
    uint16_t* ids;                       // [n]  partition, or partitionCount = "every partition"
    uint32_t* hist;                      // [partitionCount][units]
};

__device__ __forceinline__ int64_t typeHashOf(const HashCol& c, int64_t row)
{
    if (c.nulls && c.nulls[row]) return 0;   // TypeUtils.hashPosition: NULL_HASH_CODE
    switch (c.type) {
        case PG_TYPE_BIGINT: return hashLong(((const int64_t*) c.values)[row]);
        case PG_TYPE_DOUBLE: return hashLong(doubleHashBits(((const int64_t*) c.values)[row]));
        case PG_TYPE_INTEGER: case PG_TYPE_DATE: return hashLong((int64_t) ((const int32_t*) c.values)[row]);
        case PG_TYPE_VARCHAR: {     // VarcharType.hash → XxHash64.hash(slice)  (AbstractVariableWidthBlock.java:102-106)
            const int32_t o = c.offsets[row];
            return xxHash64Bytes((const uint8_t*) c.values + o, (int64_t) (c.offsets[row + 1] - o));
        }
        default: return ((const uint8_t*) c.values)[row] ? 1231 : 1237;   // BOOLEAN
    }
}

__global__ void __launch_bounds__(kPartWarps * 32) partitionIdKernel(const __grid_constant__ PartitionIdParams P)
{
    extern __shared__ uint32_t shist[];   // [warps][partitionCount]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PC = P.partitionCount;
    uint32_t* h = shist + warp * PC;
    for (int64_t unit = (int64_t) blockIdx.x * kPartWarps + warp; unit < P.units; unit += (int64_t) gridDim.x * kPartWarps) {
        for (int p = lane; p < PC; p += 32) h[p] = 0;
        __syncwarp();
        const int64_t begin = unit * kRowsPerUnit, end = min(begin + (int64_t) kRowsPerUnit, P.n);
        for (int64_t base = begin; base < end; base += 32) {
            const int64_t row = base + lane;
            int part = -1 - lane;
            if (row < end) {
                if ((P.replicateRow0 && row == 0) || (P.nullChannelNulls && P.nullChannelNulls[row])) part = PC;
                else {
                    int64_t rawHash = 0;
                    if (P.localFunction == 2) part = (int) (((P.roundRobinBase + (uint32_t) row) & 0x7fffffffu) % (uint32_t) P.bucketCount);   // counter & 0x7fffffff, :297-301
                    else {
                        if (P.precomputed) rawHash = P.precomputed[row];
                        else   // InterpretedHashGenerator.hashPosition: CombineHashFunction.getHash(result, typeHash)
                            for (int c = 0; c < P.nCols; c++) rawHash = (int64_t) (31ULL * (uint64_t) rawHash + (uint64_t) typeHashOf(P.cols[c], row));
                        part = P.localFunction ? localPartition(rawHash, P.bucketCount) : hashPartition(rawHash, P.bucketCount);
                    }
                    if (P.bucketToPartition) part = P.bucketToPartition[part];     // BucketPartitionFunction.getPartition :49-53
                }
                P.ids[row] = (uint16_t) part;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, part);
            if (part >= 0 && part < PC && lane == __ffs(peers) - 1) h[part] += __popc(peers);
            const unsigned rep = __ballot_sync(0xffffffffu, part == PC);
            if (rep) { __syncwarp(); for (int p = lane; p < PC; p += 32) h[p] += __popc(rep); }
            __syncwarp();
        }
        for (int p = lane; p < PC; p += 32) P.hist[(size_t) p * P.units + unit] = h[p];
        __syncwarp();
    }
}

struct ScatterColumn { const void* src; const uint8_t* srcNulls; void* dst; uint8_t* dstNulls; int32_t width; };
struct ScatterParams {
    ScatterColumn cols[kMaxScatterColumns];
    int32_t nCols;
    int32_t partitionCount;
    int64_t n, units;
    const uint16_t* ids;
    const uint32_t* base;    // scanned hist
    // fused exchange: partition p's rows go to per-partition destination columns (peer GPU memory) instead of one local slab
    const uint64_t* remoteValues;   // [nCols][partitionCount] destination column base addresses, or nullptr
    const uint64_t* remoteNulls;    // [nCols][partitionCount] destination null maps (0 = none)
    const uint32_t* adjust;         // [partitionCount] destination row offset - start of the partition in the local layout (mod 2^32)
};

template <typename T>
__device__ __forceinline__ void permuteColumnRemote(const T* __restrict__ src, const uint64_t* ptrs, const uint32_t* dest, const uint16_t* partOf,
                                                    int64_t begin, int count, int lane)
{
    // the scatter IS the transfer: every store lands in the destination GPU's receive column (NVLink P2P)
    int i = lane;
    for (; i + 7 * 32 < count; i += 8 * 32) {
        T v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = src[begin + i + k * 32];
#pragma unroll
        for (int k = 0; k < 8; k++) ((T*) ptrs[partOf[i + k * 32]])[dest[i + k * 32]] = v[k];
    }
    for (; i < count; i += 32) ((T*) ptrs[partOf[i]])[dest[i]] = src[begin + i];
}

template <typename T>
__device__ __forceinline__ void permuteColumn(const T* __restrict__ src, T* __restrict__ dst, const uint32_t* dest, int64_t begin, int count, int lane)
{
    // the unit's destinations are already in shared memory: every load below is independent → 8 requests in flight per lane
    int i = lane;
    for (; i + 7 * 32 < count; i += 8 * 32) {
        T v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = src[begin + i + k * 32];
#pragma unroll
        for (int k = 0; k < 8; k++) dst[dest[i + k * 32]] = v[k];
    }
    for (; i < count; i += 32) dst[dest[i]] = src[begin + i];
}

// Pass 2.  Phase A: the warp walks its unit in row order and turns partition ids into output positions (stable: rank inside the
// partition = rows of the same partition before it) kept in shared memory.  Phase B: one column at a time is permuted through them.
// Rows that go to EVERY partition (replicate-any-row / NULL partition key) are rare: they are ranked in phase A and copied in a
// separate epilogue so the common path stays branch-free.
__global__ void __launch_bounds__(kPartWarps * 32) partitionScatterKernel(const __grid_constant__ ScatterParams P)
{
    extern __shared__ uint32_t smemU32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PC = P.partitionCount;
    uint32_t* cur = smemU32 + warp * PC;                                   // [warps][PC]
    uint32_t* dest = smemU32 + kPartWarps * PC + warp * kRowsPerUnit;      // [warps][kRowsPerUnit]
    uint16_t* partOf = (uint16_t*) (smemU32 + kPartWarps * PC + kPartWarps * kRowsPerUnit) + warp * kRowsPerUnit;   // [warps][kRowsPerUnit]
    const bool remote = P.remoteValues != nullptr;
    const unsigned ltMask = (1u << lane) - 1;
    for (int64_t unit = (int64_t) blockIdx.x * kPartWarps + warp; unit < P.units; unit += (int64_t) gridDim.x * kPartWarps) {
        for (int p = lane; p < PC; p += 32) cur[p] = P.base[(size_t) p * P.units + unit] + (remote ? P.adjust[p] : 0u);
        __syncwarp();
        const int64_t begin = unit * kRowsPerUnit, end = min(begin + (int64_t) kRowsPerUnit, P.n);
        const int count = (int) (end - begin);
        bool anyReplicated = false;
        for (int off = 0; off < count; off += 32) {
            const int i = off + lane;
            const int part = i < count ? (int) P.ids[begin + i] : -1 - lane;
            const unsigned peers = __match_any_sync(0xffffffffu, part);
            const unsigned rep = __ballot_sync(0xffffffffu, part == PC);
            if (rep == 0) {
                uint32_t d = 0;
                if (part >= 0) d = cur[part] + __popc(peers & ltMask);
                __syncwarp();
                if (part >= 0 && lane == __ffs(peers) - 1) cur[part] += __popc(peers);
                if (i < count) { dest[i] = d; if (remote) partOf[i] = (uint16_t) part; }
            }
            else {
                // rows of this round that go to every partition: rank inside partition q counts its own rows and the replicated ones
                anyReplicated = true;
                uint32_t d = 0xFFFFFFFFu;
                for (int q = 0; q < PC; q++) {
                    const unsigned members = __ballot_sync(0xffffffffu, part == q) | rep;
                    if (part == q) d = cur[q] + __popc(members & ltMask);
                    __syncwarp();
                    if (lane == 0) cur[q] += __popc(members);
                    __syncwarp();
                }
                if (i < count) dest[i] = d;      // 0xFFFFFFFF marks a replicated row: copied by the epilogue
            }
            __syncwarp();
        }
        if (remote) {
            // (replicated rows are rejected on the host in this mode)
            for (int c = 0; c < P.nCols; c++) {
                const ScatterColumn& sc = P.cols[c];
                const uint64_t* ptrs = P.remoteValues + (size_t) c * PC;
                if (sc.width == 8) permuteColumnRemote((const uint64_t*) sc.src, ptrs, dest, partOf, begin, count, lane);
                else if (sc.width == 4) permuteColumnRemote((const uint32_t*) sc.src, ptrs, dest, partOf, begin, count, lane);
                else permuteColumnRemote((const uint8_t*) sc.src, ptrs, dest, partOf, begin, count, lane);
                const uint64_t* nptrs = P.remoteNulls + (size_t) c * PC;
                if (nptrs[0]) {      // a null map travels for this channel (all destinations have one)
                    if (sc.srcNulls) permuteColumnRemote(sc.srcNulls, nptrs, dest, partOf, begin, count, lane);
                    else for (int i = lane; i < count; i += 32) ((uint8_t*) nptrs[partOf[i]])[dest[i]] = 0;
                }
            }
        }
        else if (!anyReplicated) {
            for (int c = 0; c < P.nCols; c++) {
                const ScatterColumn& sc = P.cols[c];
                if (sc.width == 8) permuteColumn((const uint64_t*) sc.src, (uint64_t*) sc.dst, dest, begin, count, lane);
                else if (sc.width == 4) permuteColumn((const uint32_t*) sc.src, (uint32_t*) sc.dst, dest, begin, count, lane);
                else permuteColumn((const uint8_t*) sc.src, (uint8_t*) sc.dst, dest, begin, count, lane);
                if (sc.dstNulls) {
                    if (sc.srcNulls) permuteColumn(sc.srcNulls, sc.dstNulls, dest, begin, count, lane);
                    else for (int i = lane; i < count; i += 32) sc.dstNulls[dest[i]] = 0;
                }
            }
        }
        else {
            // slow path for units that contain replicated rows: recompute every row's positions round by round
            for (int p = lane; p < PC; p += 32) cur[p] = P.base[(size_t) p * P.units + unit];
            __syncwarp();
            for (int off = 0; off < count; off += 32) {
                const int i = off + lane;
                const int part = i < count ? (int) P.ids[begin + i] : -1 - lane;
                const unsigned rep = __ballot_sync(0xffffffffu, part == PC);
                for (int q = 0; q < PC; q++) {
                    const unsigned members = __ballot_sync(0xffffffffu, part == q) | rep;
                    const bool mine = part == q || part == PC;
                    const uint32_t d = cur[q] + __popc(members & ltMask);
                    __syncwarp();
                    if (lane == 0) cur[q] += __popc(members);
                    if (mine) {
                        const int64_t row = begin + i;
                        for (int c = 0; c < P.nCols; c++) {
                            const ScatterColumn& sc = P.cols[c];
                            if (sc.width == 8) ((uint64_t*) sc.dst)[d] = ((const uint64_t*) sc.src)[row];
                            else if (sc.width == 4) ((uint32_t*) sc.dst)[d] = ((const uint32_t*) sc.src)[row];
                            else ((uint8_t*) sc.dst)[d] = ((const uint8_t*) sc.src)[row];
                            if (sc.dstNulls) sc.dstNulls[d] = sc.srcNulls ? sc.srcNulls[row] : 0;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    }
}

// Pass 2 of the fused exchange.  Stores to a peer GPU travel as NVLink write packets: 8-byte stores that hop between destinations
// every few lanes become 32-byte packets (measured ~300–400 GB/s per GPU).  Here the rows of a unit are first ordered by destination
// in shared memory (stable, the ranks of phase A), so that consecutive lanes store consecutive addresses of ONE destination — 256
// contiguous bytes per warp instruction.
constexpr int kRemoteWarps = 4;
__global__ void __launch_bounds__(kRemoteWarps * 32) partitionScatterStagedKernel(const __grid_constant__ ScatterParams P)
{
    extern __shared__ __align__(16) unsigned char remoteSmem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PC = P.partitionCount;
    const unsigned ltMask = (1u << lane) - 1;
    unsigned char* mine = remoteSmem + (size_t) warp * (8192 + 3 * 2048);
    uint64_t* stage = (uint64_t*) mine;                       // [kRowsPerUnit] values of one column in destination order
    uint32_t* dest = (uint32_t*) mine;                        // phase A only (aliases stage)
    uint16_t* sortedPos = (uint16_t*) (mine + 8192);          // row i of the unit → position in destination order
    uint16_t* partSorted = sortedPos + kRowsPerUnit;          // position in destination order → partition
    uint16_t* partOf = partSorted + kRowsPerUnit;             // row i → partition
    uint32_t* cur = (uint32_t*) (remoteSmem + (size_t) kRemoteWarps * (8192 + 3 * 2048)) + (size_t) warp * 3 * PC;
    uint32_t* unitBase = cur + PC;                            // destination row of the unit's first row of partition p
    uint32_t* localStart = unitBase + PC;                     // first position of partition p in the unit's destination order
    for (int64_t unit = (int64_t) blockIdx.x * kRemoteWarps + warp; unit < P.units; unit += (int64_t) gridDim.x * kRemoteWarps) {
        for (int p = lane; p < PC; p += 32) { const uint32_t b = P.base[(size_t) p * P.units + unit] + P.adjust[p]; cur[p] = b; unitBase[p] = b; }
        __syncwarp();
        const int64_t begin = unit * kRowsPerUnit, end = min(begin + (int64_t) kRowsPerUnit, P.n);
        const int count = (int) (end - begin);
        // phase A: stable ranks (as in partitionScatterKernel)
        for (int off = 0; off < count; off += 32) {
            const int i = off + lane;
            const int part = i < count ? (int) P.ids[begin + i] : -1 - lane;
            const unsigned peers = __match_any_sync(0xffffffffu, part);
            uint32_t d = 0;
            if (part >= 0) d = cur[part] + __popc(peers & ltMask);
            __syncwarp();
            if (part >= 0 && lane == __ffs(peers) - 1) cur[part] += __popc(peers);
            if (i < count) { dest[i] = d; partOf[i] = (uint16_t) part; }
            __syncwarp();
        }
        // positions of the partitions in the unit's destination order: exclusive scan of the unit's per-partition counts
        uint32_t running = 0;
        for (int p0 = 0; p0 < PC; p0 += 32) {
            const int p = p0 + lane;
            const uint32_t c = p < PC ? cur[p] - unitBase[p] : 0;
            uint32_t inc = c;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, dlt); if (lane >= dlt) inc += o; }
            if (p < PC) localStart[p] = running + inc - c;
            running += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncwarp();
        for (int i = lane; i < count; i += 32) {
            const int p = partOf[i];
            const uint32_t sp = localStart[p] + (dest[i] - unitBase[p]);
            sortedPos[i] = (uint16_t) sp;
            partSorted[sp] = (uint16_t) p;
        }
        __syncwarp();       // dest[] is dead from here on: its storage becomes the staging buffer
        // phase B: one column at a time through the staging buffer
        auto moveColumn = [&](auto typeTag, const void* srcV, const uint64_t* ptrs, bool zeros) {
            using T = decltype(typeTag);
            const T* src = (const T*) srcV;
            T* st = (T*) stage;
            int i = lane;
            if (zeros) { for (; i < count; i += 32) st[i] = (T) 0; }
            else {
                for (; i + 7 * 32 < count; i += 8 * 32) {
                    T v[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = src[begin + i + k * 32];
#pragma unroll
                    for (int k = 0; k < 8; k++) st[sortedPos[i + k * 32]] = v[k];
                }
                for (; i < count; i += 32) st[sortedPos[i]] = src[begin + i];
            }
            __syncwarp();
            for (int j = lane; j < count; j += 32) {
                const int p = partSorted[j];
                ((T*) ptrs[p])[unitBase[p] + ((uint32_t) j - localStart[p])] = st[j];
            }
            __syncwarp();
        };
        for (int c = 0; c < P.nCols; c++) {
            const ScatterColumn& sc = P.cols[c];
            const uint64_t* ptrs = P.remoteValues + (size_t) c * PC;
            if (sc.width == 8) moveColumn((uint64_t) 0, sc.src, ptrs, false);
            else if (sc.width == 4) moveColumn((uint32_t) 0, sc.src, ptrs, false);
            else moveColumn((uint8_t) 0, sc.src, ptrs, false);
            const uint64_t* nptrs = P.remoteNulls + (size_t) c * PC;
            if (nptrs[0]) moveColumn((uint8_t) 0, sc.srcNulls, nptrs, sc.srcNulls == nullptr);
        }
    }
}

// ---- VARCHAR channels (VariableWidthBlockEncodingBuffer.java:1-293: per partition the selected positions' offsets and slice bytes are
// appended) ride on the fixed-width machinery: a slice's length and its start offset are scattered as two synthetic 4-byte columns (so
// stable order, replicated rows and NULL maps need nothing new); a scan of the scattered lengths gives the output offsets and one
// pass copies the bytes.
__global__ void varcharLengthsKernel(const int32_t* offsets, const uint8_t* nulls, int64_t n, int32_t* lens)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        lens[i] = (nulls && nulls[i]) ? 0 : offsets[i + 1] - offsets[i];
}
__global__ void varcharGatherBytesKernel(const uint8_t* inBytes, const int32_t* starts, const int32_t* outOffsets, int64_t total, uint8_t* outBytes)
{
    for (int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; j < total; j += (int64_t) gridDim.x * blockDim.x) {
        const int32_t to = outOffsets[j], len = outOffsets[j + 1] - to, from = starts[j];
        for (int32_t i = 0; i < len; i++) outBytes[to + i] = inBytes[from + i];
    }
}

class PartitionedOutputOp : public Operator {
    std::vector<int> types, partitionChannels, staged;
    int partitionCount, hashChannel, nullChannel, partitionFunction;
    int bucketCount = 0; DeviceBuffer bucketTable; uint32_t roundRobinCounter = 0;
    bool replicatesAnyRow, hasAnyRowBeenReplicated = false;
    bool finishing = false, hasPending = false;
    PageStager stager;
    DeviceBuffer ids, hist, base, scanTemp;
    std::vector<DeviceBuffer> outValues, outNulls;
    std::vector<PinnedBuffer> hostValues, hostNulls;
    std::vector<pg_block> outBlocks;
    std::vector<int64_t> offsets;
    int64_t pendingRows = 0;
    // VARCHAR channels: scattered lengths / starts, output offsets and bytes
    std::vector<DeviceBuffer> vcLens, vcOutLens, vcOutStarts, vcOutOffsets, vcOutBytes;
    std::vector<PinnedBuffer> vcHostBytes, vcHostOffsets;
    DeviceBuffer vcScanTemp;
    bool hasVarchar = false;

public:
    explicit PartitionedOutputOp(const pg_partitioned_output_config* c) : Operator(c->options), stager(this)
    {
        PG_CHECK_ARG(c->partition_count >= 1 && c->partition_count <= 4096, "partition count must be in [1, 4096]");
        PG_CHECK_ARG(c->n_channels >= 1 && c->n_channels <= kMaxScatterColumns, "1..32 channels");
        PG_CHECK_ARG(c->n_partition_channels <= kMaxPartChannels, "at most 8 partition channels");
        types.assign(c->types, c->types + c->n_channels);
        partitionChannels.assign(c->partition_channels, c->partition_channels + c->n_partition_channels);
        partitionCount = c->partition_count; hashChannel = c->precomputed_hash_channel; nullChannel = c->null_channel;
        replicatesAnyRow = c->replicates_any_row != 0;
        partitionFunction = c->partition_function;
        PG_CHECK_ARG(partitionFunction == PG_PARTITION_HASH_GENERATOR || partitionFunction == PG_PARTITION_LOCAL || partitionFunction == PG_PARTITION_ROUND_ROBIN, "unknown partition function");
        bucketCount = partitionCount;
        if (c->bucket_to_partition) {      // BucketPartitionFunction.java:31-36
            PG_CHECK_ARG(c->bucket_count >= 1 && c->bucket_count <= (1 << 20), "bucket_count out of range");
            bucketCount = c->bucket_count;
            std::vector<int32_t> table(c->bucket_to_partition, c->bucket_to_partition + bucketCount);
            for (int32_t p : table) PG_CHECK_ARG(p >= 0 && p < partitionCount, "bucket_to_partition entry outside [0, partition_count)");
            bucketTable.accounting = &stats.device_bytes;
            bucketTable.reserve((size_t) bucketCount * 4);
            PG_CUDA(cudaMemcpyAsync(bucketTable.ptr, table.data(), (size_t) bucketCount * 4, cudaMemcpyHostToDevice, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
        }
        if (partitionFunction == PG_PARTITION_LOCAL)
            PG_CHECK_ARG((bucketCount & (bucketCount - 1)) == 0, "partitionCount must be a power of 2");   // LocalPartitionGenerator.java:33
        if (partitionFunction == PG_PARTITION_ROUND_ROBIN && c->null_channel >= 0)
            throw Error(PG_ERR_UNSUPPORTED, "round-robin partitioning with a null channel (replicated rows do not advance the counter) is not on the GPU path");
        int internalColumns = 0;
        for (int t : types) { if (t == PG_TYPE_VARCHAR) { hasVarchar = true; internalColumns += 2; } else internalColumns++; }
        PG_CHECK_ARG(internalColumns <= kMaxScatterColumns, "too many channels (a VARCHAR channel counts twice)");
        for (int ch : partitionChannels) {
            PG_CHECK_ARG(ch >= 0 && ch < c->n_channels, "partition channel out of range");
        }
        PG_CHECK_ARG(hashChannel < c->n_channels && nullChannel < c->n_channels, "channel out of range");
        PG_CHECK_ARG(hashChannel >= 0 || !partitionChannels.empty() || partitionCount == 1 || partitionFunction == PG_PARTITION_ROUND_ROBIN, "no partition channels");
        if (hashChannel >= 0) PG_CHECK_ARG(types[hashChannel] == PG_TYPE_BIGINT, "precomputed hash channel must be BIGINT");
        for (int i = 0; i < c->n_channels; i++) staged.push_back(i);
        offsets.assign(partitionCount + 1, 0);
        for (DeviceBuffer* d : {&ids, &hist, &base, &scanTemp}) d->accounting = &stats.device_bytes;
        const size_t smem = (size_t) kPartWarps * partitionCount * 4;
        PG_CUDA(cudaFuncSetAttribute(partitionIdKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        PG_CUDA(cudaFuncSetAttribute(partitionScatterKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (smem + (size_t) kPartWarps * kRowsPerUnit * 6)));
    }

    bool needsInput() override { return !finishing && !hasPending; }

    // ---- pass 1 + scan: partition ids, histograms, partition boundaries (state kept for the scatter) ----
    std::vector<DeviceColumn> curCols; int curSlot = -1; int64_t curN = 0, curUnits = 0; bool prepared = false;
    DeviceBuffer remoteTable;

    // ranks (4 B per row of a unit) and, for the fused exchange only, the partition of each row (2 B)
    size_t scatterSmem(bool remote) const { return (size_t) kPartWarps * partitionCount * 4 + (size_t) kPartWarps * kRowsPerUnit * (remote ? 6 : 4); }
    int launchGrid() const { return (int) std::min<int64_t>(divUp(curUnits, kPartWarps), (int64_t) numSMs * 8); }

    void prepare(const pg_page* page)
    {
        const int64_t n = page->position_count;
        PG_CHECK_ARG(n < (int64_t(1) << 31), "page exceeds 2^31 positions");
        PG_CHECK_ARG(page->channel_count == (int) types.size(), "page channel count differs from the operator's configuration");
        curSlot = stager.stage(page, staged, curCols);
        for (size_t i = 0; i < curCols.size(); i++) PG_CHECK_ARG(curCols[i].type == types[i], "channel type differs from the operator's configuration");
        curN = n; curUnits = divUp(n, kRowsPerUnit);
        ids.reserve((size_t) n * 2);
        const size_t histN = (size_t) partitionCount * curUnits + 1;
        hist.reserve(histN * 4); base.reserve(histN * 4);
        PartitionIdParams A{};
        A.nCols = (int) partitionChannels.size();
        for (int i = 0; i < A.nCols; i++) { const DeviceColumn& c = curCols[partitionChannels[i]]; A.cols[i].values = c.values; A.cols[i].nulls = c.nulls; A.cols[i].offsets = c.offsets; A.cols[i].type = c.type; }
        A.precomputed = hashChannel >= 0 ? (const int64_t*) curCols[hashChannel].values : nullptr;
        A.nullChannelNulls = nullChannel >= 0 ? curCols[nullChannel].nulls : nullptr;
        A.replicateRow0 = (replicatesAnyRow && !hasAnyRowBeenReplicated) ? 1 : 0;
        hasAnyRowBeenReplicated = hasAnyRowBeenReplicated || replicatesAnyRow;
        A.bucketCount = bucketCount; A.bucketToPartition = bucketTable.as<int32_t>();
        // RoundRobinBucketFunction keeps counting across pages; a replicated row 0 / NULL-channel row does not call the function (:504-530)
        // (only rows that reach partitionFunction.getPartition advance it: not the replicated row 0, :504-512)
        const uint32_t skipped = A.replicateRow0 ? 1u : 0u;
        A.roundRobinBase = roundRobinCounter - skipped;
        if (partitionFunction == PG_PARTITION_ROUND_ROBIN) roundRobinCounter = (uint32_t) ((roundRobinCounter + (uint64_t) n - skipped) & 0x7fffffffu);
        A.partitionCount = partitionCount; A.localFunction = partitionFunction == PG_PARTITION_LOCAL ? 1 : (partitionFunction == PG_PARTITION_ROUND_ROBIN ? 2 : 0); A.n = n; A.units = curUnits; A.ids = ids.as<uint16_t>(); A.hist = hist.as<uint32_t>();
        PG_CUDA(cudaMemsetAsync(hist.as<uint32_t>() + histN - 1, 0, 4, stream));
        const size_t smem = (size_t) kPartWarps * partitionCount * 4;
        PG_LAUNCH(this, partitionIdKernel, launchGrid(), kPartWarps * 32, smem, A);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, hist.as<uint32_t>(), base.as<uint32_t>(), (int) histN, stream);
        scanTemp.reserve(tb);
        cub::DeviceScan::ExclusiveSum(scanTemp.ptr, tb, hist.as<uint32_t>(), base.as<uint32_t>(), (int) histN, stream);
        stats.kernel_launches += 2;
        // partition boundaries: base[p * units] and the grand total
        std::vector<uint32_t> hostOffsets(partitionCount + 1);
        PG_CUDA(cudaMemcpy2DAsync(hostOffsets.data(), 4, base.ptr, (size_t) curUnits * 4, 4, partitionCount, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaMemcpyAsync(&hostOffsets[partitionCount], base.as<uint32_t>() + histN - 1, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        stats.d2h_bytes += (partitionCount + 1) * 4;
        for (int p = 0; p <= partitionCount; p++) offsets[p] = hostOffsets[p];
        prepared = true;
    }

    void addInput(const pg_page* page) override
    {
        PG_CHECK_STATE(needsInput() && !prepared, "partitioned output does not need input (finishing, output pending or a count awaiting its scatter)");
        const int64_t n = page->position_count;
        stats.input_positions += n;
        if (n == 0) return;
        prepare(page);
        prepared = false;
        const std::vector<DeviceColumn>& cols = curCols;
        const int64_t total = offsets[partitionCount];
        const size_t nCols = types.size();
        if (outValues.size() < nCols) { outValues.resize(nCols); outNulls.resize(nCols); for (auto& d : outValues) d.accounting = &stats.device_bytes; for (auto& d : outNulls) d.accounting = &stats.device_bytes; }
        ScatterParams S{};
        S.partitionCount = partitionCount; S.n = n; S.units = curUnits; S.ids = ids.as<uint16_t>(); S.base = base.as<uint32_t>();
        outBlocks.assign(nCols, pg_block{});
        const size_t nVc = (size_t) std::count(types.begin(), types.end(), (int) PG_TYPE_VARCHAR);
        if (vcLens.size() < nVc) {
            vcLens.resize(nVc); vcOutLens.resize(nVc); vcOutStarts.resize(nVc); vcOutOffsets.resize(nVc); vcOutBytes.resize(nVc);
            for (auto* vec : {&vcLens, &vcOutLens, &vcOutStarts, &vcOutOffsets, &vcOutBytes}) for (auto& d : *vec) d.accounting = &stats.device_bytes;
        }
        std::vector<uint8_t*> dstNullsOf(nCols, nullptr);
        int k = 0;
        size_t v = 0;
        for (size_t i = 0; i < nCols; i++) {
            if (types[i] == PG_TYPE_VARCHAR) {
                PG_CHECK_ARG(cols[i].offsets != nullptr, "VARCHAR block without offsets");
                vcLens[v].reserve((size_t) n * 4); vcOutLens[v].reserve((size_t) (total + 1) * 4); vcOutStarts[v].reserve((size_t) total * 4);
                PG_LAUNCH(this, varcharLengthsKernel, (int) std::min<int64_t>(divUp(n, 256), 4096), 256, 0, cols[i].offsets, cols[i].nulls, n, vcLens[v].as<int32_t>());
                PG_CUDA(cudaMemsetAsync(vcOutLens[v].as<int32_t>() + total, 0, 4, stream));
                ScatterColumn& sl = S.cols[k++];
                sl.width = 4; sl.src = vcLens[v].ptr; sl.srcNulls = cols[i].nulls; sl.dst = vcOutLens[v].ptr; sl.dstNulls = nullptr;
                if (cols[i].nulls) { outNulls[i].reserve(total); sl.dstNulls = outNulls[i].as<uint8_t>(); }
                dstNullsOf[i] = sl.dstNulls;
                ScatterColumn& ss = S.cols[k++];
                ss.width = 4; ss.src = cols[i].offsets; ss.srcNulls = nullptr; ss.dst = vcOutStarts[v].ptr; ss.dstNulls = nullptr;
                v++;
                continue;
            }
            ScatterColumn& sc = S.cols[k++];
            sc.width = typeWidth(types[i]); sc.src = cols[i].values; sc.srcNulls = cols[i].nulls;
            outValues[i].reserve((size_t) total * sc.width);
            sc.dst = outValues[i].ptr; sc.dstNulls = nullptr;
            if (cols[i].nulls) { outNulls[i].reserve(total); sc.dstNulls = outNulls[i].as<uint8_t>(); }
            dstNullsOf[i] = sc.dstNulls;
        }
        S.nCols = k;
        PG_LAUNCH(this, partitionScatterKernel, launchGrid(), kPartWarps * 32, scatterSmem(false), S);
        // VARCHAR: scattered lengths → output offsets → bytes
        std::vector<int32_t> vcTotalBytes(nVc, 0);
        v = 0;
        for (size_t i = 0; i < nCols; i++) {
            if (types[i] != PG_TYPE_VARCHAR) continue;
            vcOutOffsets[v].reserve((size_t) (total + 1) * 4);
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, vcOutLens[v].as<int32_t>(), vcOutOffsets[v].as<int32_t>(), (int) (total + 1), stream);
            vcScanTemp.reserve(tb);
            cub::DeviceScan::ExclusiveSum(vcScanTemp.ptr, tb, vcOutLens[v].as<int32_t>(), vcOutOffsets[v].as<int32_t>(), (int) (total + 1), stream);
            stats.kernel_launches += 2;
            PG_CUDA(cudaMemcpyAsync(&vcTotalBytes[v], vcOutOffsets[v].as<int32_t>() + total, 4, cudaMemcpyDeviceToHost, stream));
            v++;
        }
        if (nVc) {
            PG_CUDA(cudaStreamSynchronize(stream));
            v = 0;
            for (size_t i = 0; i < nCols; i++) {
                if (types[i] != PG_TYPE_VARCHAR) continue;
                vcOutBytes[v].reserve((size_t) vcTotalBytes[v] + 16);
                PG_LAUNCH(this, varcharGatherBytesKernel, (int) std::min<int64_t>(divUp(total, 256), 4096), 256, 0, (const uint8_t*) cols[i].values, vcOutStarts[v].as<int32_t>(),
                          vcOutOffsets[v].as<int32_t>(), total, vcOutBytes[v].as<uint8_t>());
                v++;
            }
        }
        stager.markCompute(curSlot);
        const bool toHost = options.output_memory == PG_MEM_HOST;
        if (toHost && hostValues.size() < nCols) { hostValues.resize(nCols); hostNulls.resize(nCols); }
        if (toHost && vcHostBytes.size() < nVc) { vcHostBytes.resize(nVc); vcHostOffsets.resize(nVc); }
        v = 0;
        for (size_t i = 0; i < nCols; i++) {
            pg_block& b = outBlocks[i];
            b.type = types[i]; b.encoding = PG_ENC_FLAT; b.position_count = total;
            const bool isVc = types[i] == PG_TYPE_VARCHAR;
            const size_t vbytes = isVc ? (size_t) vcTotalBytes[v] : (size_t) total * typeWidth(types[i]);
            const void* dValues = isVc ? vcOutBytes[v].ptr : outValues[i].ptr;
            if (toHost) {
                PinnedBuffer& hv = isVc ? vcHostBytes[v] : hostValues[i];
                hv.reserve(std::max<size_t>(vbytes, 16));
                if (vbytes) PG_CUDA(cudaMemcpyAsync(hv.ptr, dValues, vbytes, cudaMemcpyDeviceToHost, stream));
                b.values = hv.ptr; b.is_null = nullptr;
                stats.d2h_bytes += (int64_t) vbytes;
                if (isVc) {
                    vcHostOffsets[v].reserve((size_t) (total + 1) * 4);
                    PG_CUDA(cudaMemcpyAsync(vcHostOffsets[v].ptr, vcOutOffsets[v].ptr, (size_t) (total + 1) * 4, cudaMemcpyDeviceToHost, stream));
                    b.offsets = vcHostOffsets[v].as<int32_t>();
                }
                if (dstNullsOf[i]) { hostNulls[i].reserve(total); PG_CUDA(cudaMemcpyAsync(hostNulls[i].ptr, outNulls[i].ptr, total, cudaMemcpyDeviceToHost, stream)); b.is_null = hostNulls[i].as<uint8_t>(); stats.d2h_bytes += total; }
            }
            else { b.values = dValues; b.is_null = dstNullsOf[i]; if (isVc) b.offsets = vcOutOffsets[v].as<int32_t>(); }
            stats.output_bytes += (int64_t) vbytes;
            if (isVc) v++;
        }
        PG_CUDA(cudaStreamSynchronize(stream));
        pendingRows = total;
        hasPending = true;
        stats.output_positions += total;
    }

    // ---- fused repartition + exchange: pass 1 here, pass 2 in scatterTo() straight into (peer) destination columns ----
    void count(const pg_page* page, int64_t* outCounts)
    {
        PG_CHECK_STATE(!finishing && !hasPending && !prepared, "partitioned output: count in the wrong state");
        PG_CHECK_ARG(!replicatesAnyRow && nullChannel < 0, "replicated rows are not supported by the fused exchange");
        if (hasVarchar) throw Error(PG_ERR_UNSUPPORTED, "VARCHAR channels are not supported by the fused (peer-memory) exchange");
        stats.input_positions += page->position_count;
        if (page->position_count == 0) { for (int p = 0; p < partitionCount; p++) outCounts[p] = 0; curN = 0; prepared = true; return; }
        prepare(page);
        for (int p = 0; p < partitionCount; p++) outCounts[p] = offsets[p + 1] - offsets[p];
    }

    void scatterTo(void* const* dstValues, uint8_t* const* dstNulls, const int64_t* dstRowOffsets)
    {
        PG_CHECK_STATE(prepared, "scatter_to without a preceding count");
        prepared = false;
        if (curN == 0) return;
        const size_t nCols = types.size();
        // destination tables: [nCols][P] value pointers, [nCols][P] null pointers, [P] row adjustments
        const size_t words = nCols * partitionCount;
        std::vector<uint64_t> table(2 * words + (size_t) partitionCount);
        for (size_t c = 0; c < nCols; c++)
            for (int p = 0; p < partitionCount; p++) {
                table[c * partitionCount + p] = (uint64_t) (uintptr_t) dstValues[(size_t) p * nCols + c];
                table[words + c * partitionCount + p] = dstNulls ? (uint64_t) (uintptr_t) dstNulls[(size_t) p * nCols + c] : 0;
            }
        uint32_t* adj = (uint32_t*) &table[2 * words];
        for (int p = 0; p < partitionCount; p++) adj[p] = (uint32_t) ((uint64_t) dstRowOffsets[p] - (uint64_t) offsets[p]);
        remoteTable.accounting = &stats.device_bytes;
        remoteTable.reserve(table.size() * 8);
        PG_CUDA(cudaMemcpyAsync(remoteTable.ptr, table.data(), table.size() * 8, cudaMemcpyHostToDevice, stream));
        ScatterParams S{};
        S.nCols = (int) nCols; S.partitionCount = partitionCount; S.n = curN; S.units = curUnits; S.ids = ids.as<uint16_t>(); S.base = base.as<uint32_t>();
        S.remoteValues = remoteTable.as<uint64_t>(); S.remoteNulls = remoteTable.as<uint64_t>() + words; S.adjust = (const uint32_t*) (remoteTable.as<uint64_t>() + 2 * words);
        for (size_t i = 0; i < nCols; i++) { ScatterColumn& sc = S.cols[i]; sc.width = typeWidth(types[i]); sc.src = curCols[i].values; sc.srcNulls = curCols[i].nulls; sc.dst = nullptr; sc.dstNulls = nullptr; }
        // destination-ordered staging in shared memory → contiguous peer stores (PG_REMOTE_SCATTER=direct keeps the unstaged variant)
        const char* rs = getenv("PG_REMOTE_SCATTER");
        const size_t smem = (size_t) kRemoteWarps * (8192 + 3 * 2048) + (size_t) kRemoteWarps * 3 * partitionCount * 4;
        if ((rs && !strcmp(rs, "direct")) || smem > 200 * 1024) PG_LAUNCH(this, partitionScatterKernel, launchGrid(), kPartWarps * 32, scatterSmem(true), S);
        else {
            PG_CUDA(cudaFuncSetAttribute(partitionScatterStagedKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            const int grid = (int) std::min<int64_t>(divUp(curUnits, kRemoteWarps), (int64_t) numSMs * 8);
            PG_LAUNCH(this, partitionScatterStagedKernel, grid, kRemoteWarps * 32, smem, S);
        }
        stager.markCompute(curSlot);
        PG_CUDA(cudaStreamSynchronize(stream));
        stats.output_positions += curN;
        stats.output_bytes += curN * 8 * (int64_t) nCols;
    }

    bool getOutput(pg_page* out) override
    {
        if (!hasPending) return false;
        out->position_count = pendingRows; out->channel_count = (int32_t) outBlocks.size(); out->memory = options.output_memory; out->blocks = outBlocks.data();
        hasPending = false;
        return true;
    }
    void finish() override { finishing = true; }
    bool isFinished() override { return finishing && !hasPending; }
    void getOffsets(int64_t* out) const { for (size_t i = 0; i < offsets.size(); i++) out[i] = offsets[i]; }
};

}  // namespace pg

extern "C" {
int32_t pg_partitioned_output_create(const pg_partitioned_output_config* config, pg_operator** out)
{
    try { *out = (pg_operator*) new pg::PartitionedOutputOp(config); }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
int32_t pg_partitioned_output_count(pg_operator* op, const pg_page* page, int64_t* outCounts)
{
    try {
        auto* p = dynamic_cast<pg::PartitionedOutputOp*>((pg::Operator*) op);
        if (!p) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a partitioned output operator");
        p->bind();
        p->count(page, outCounts);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
int32_t pg_partitioned_output_scatter_to(pg_operator* op, void* const* dstValues, uint8_t* const* dstNulls, const int64_t* dstRowOffsets)
{
    try {
        auto* p = dynamic_cast<pg::PartitionedOutputOp*>((pg::Operator*) op);
        if (!p) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a partitioned output operator");
        p->bind();
        p->scatterTo(dstValues, dstNulls, dstRowOffsets);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
int32_t pg_partitioned_output_get_offsets(pg_operator* op, int64_t* out)
{
    try {
        auto* p = dynamic_cast<pg::PartitionedOutputOp*>((pg::Operator*) op);
        if (!p) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a partitioned output operator");
        p->getOffsets(out);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    return PG_OK;
}
}
