// presto_b200 — SerializedPage wire format on sm_100a.
//
// Replaces (SPI = presto-spi/src/main/java/com/facebook/presto/spi/page/, PC = presto-common/src/main/java/com/facebook/presto/common/block/):
//   SPI/PagesSerde.java:67-156          serialize / deserialize / wrapSlice (no compressor, no cipher)
//   SPI/PagesSerdeUtil.java:45-120      writeRawPage / readRawPage, page metadata (21 bytes), CRC32 checksum
//   PC/{LongArray,IntArray,ByteArray}BlockEncoding.java, PC/VariableWidthBlockEncoding.java:35-75, PC/RunLengthBlockEncoding.java:30-55,
//   PC/DictionaryBlockEncoding.java:34-80 (write: the compacted dictionary, DictionaryBlock.java:584-628; read), PC/EncoderUtil.java:31-100 (null bits, MSB first)
//
// The wire is a byte stream without alignment: a LONG_ARRAY's values typically start at offset 44.  Everything that touches it goes
// through one primitive, copyBytesKernel, that moves a byte range between arbitrarily aligned addresses with aligned 16-byte stores
// (five aligned 32-bit loads + four funnel shifts per store).  Columns with NULLs are compacted / expanded through an aligned
// temporary with per-unit non-null counts + scan + ballot ranks; null maps are packed / unpacked eight rows per byte; the CRC32
// is computed in 4 KB chunks on the device and combined on the host with the zero-extension operator (zlib's crc32_combine idea).
// The host only writes / parses the fixed fields (a few dozen bytes per channel).
#include "common.cuh"
#include "crc32_host.h"
#include "lz4_codec.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <atomic>
#include <cstring>
#include <memory>
#include <random>
#include <string>

namespace pg {

// ---- byte copy between arbitrarily aligned device addresses --------------------------------------------------------------------
__device__ __forceinline__ uint32_t loadSrcWord(const uint8_t* srcAligned, int64_t wordIndex) { return ((const uint32_t*) srcAligned)[wordIndex]; }

__global__ void __launch_bounds__(256) copyBytesKernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int64_t bytes)
{
    // head: bytes up to the first 16-byte boundary of dst; body: aligned 16-byte stores; tail: the rest
    const int64_t head = min(bytes, (int64_t) ((16 - ((uintptr_t) dst & 15)) & 15));
    const int64_t chunks = (bytes - head) >> 4;
    const int64_t tailStart = head + (chunks << 4);
    const int64_t gtid = blockIdx.x * (int64_t) blockDim.x + threadIdx.x, nthr = (int64_t) gridDim.x * blockDim.x;
    if (gtid < head) dst[gtid] = src[gtid];
    if (gtid < bytes - tailStart) dst[tailStart + gtid] = src[tailStart + gtid];
    const uint8_t* s0 = src + head;                              // source of the first aligned chunk
    const uint8_t* sAligned = (const uint8_t*) ((uintptr_t) s0 & ~(uintptr_t) 3);
    const unsigned shift = (unsigned) ((uintptr_t) s0 & 3) * 8;  // bits to drop from the first word
    uint4* d16 = (uint4*) (dst + head);
    for (int64_t c = gtid; c < chunks; c += nthr) {
        const int64_t w = c * 4;
        const uint32_t a = loadSrcWord(sAligned, w), b = loadSrcWord(sAligned, w + 1), cc = loadSrcWord(sAligned, w + 2), d = loadSrcWord(sAligned, w + 3);
        uint4 v;
        if (shift == 0) v = make_uint4(a, b, cc, d);
        else {
            const uint32_t e = loadSrcWord(sAligned, w + 4);     // (the word holding the chunk's last bytes: inside the source range)
            v = make_uint4(__funnelshift_r(a, b, shift), __funnelshift_r(b, cc, shift), __funnelshift_r(cc, d, shift), __funnelshift_r(d, e, shift));
        }
        d16[c] = v;
    }
}

// ---- null maps: byte per row <-> bit per row, MSB first (EncoderUtil.encodeNullsAsBits :31-63 / decodeNullBits :68-100) ---------
__global__ void packNullBitsKernel(const uint8_t* __restrict__ isNull, int64_t n, uint8_t* __restrict__ dst)
{
    const int64_t outBytes = (n + 7) >> 3;
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < outBytes; i += (int64_t) gridDim.x * blockDim.x) {
        uint8_t v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { const int64_t p = i * 8 + k; if (p < n && isNull[p]) v |= (uint8_t) (0x80u >> k); }
        dst[i] = v;
    }
}
__global__ void unpackNullBitsKernel(const uint8_t* __restrict__ bits, int64_t n, uint8_t* __restrict__ isNull)
{
    for (int64_t p = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; p < n; p += (int64_t) gridDim.x * blockDim.x)
        isNull[p] = (bits[p >> 3] >> (7 - (p & 7))) & 1;
}

// ---- compaction / expansion of the non-null values ------------------------------------------------------------------------------
constexpr int kSerdeUnit = 1024;      // rows per warp unit
__global__ void nonNullCountKernel(const uint8_t* __restrict__ isNull, int64_t n, int64_t units, uint32_t* __restrict__ counts)
{
    const int lane = threadIdx.x & 31;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warpId; u < units; u += nWarps) {
        uint32_t c = 0;
        for (int64_t p = u * kSerdeUnit + lane; p < min(n, (u + 1) * kSerdeUnit); p += 32) c += isNull[p] ? 0 : 1;
        for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
        if (lane == 0) counts[u] = c;
    }
}
// dir = 0: compact[rank] = values[row] for non-null rows;  dir = 1: values[row] = non-null ? compact[rank] : 0
template <typename T>
__global__ void nonNullMoveKernel(const uint8_t* __restrict__ isNull, int64_t n, int64_t units, const uint32_t* __restrict__ unitOffsets,
                                  T* __restrict__ values, T* __restrict__ compact, int dir)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warpId; u < units; u += nWarps) {
        uint32_t cursor = unitOffsets[u];
        const int64_t end = min(n, (u + 1) * kSerdeUnit);
        for (int64_t base = u * kSerdeUnit; base < end; base += 32) {
            const int64_t p = base + lane;
            const bool live = p < end && !isNull[p];
            const unsigned m = __ballot_sync(0xffffffffu, live);
            const uint32_t r = cursor + __popc(m & lt);
            if (dir == 0) { if (live) compact[r] = values[p]; }
            else if (p < end) values[p] = live ? compact[r] : (T) 0;
            cursor += __popc(m);
        }
    }
}

// VARIABLE_WIDTH offsets: wire[p] = offsets[p + 1] - offsets[0] (ends, relative to the first position)  <->  offsets[p + 1] = wire[p]
__global__ void varcharEndsKernel(const int32_t* __restrict__ offsets, int64_t n, int32_t* __restrict__ ends)
{
    const int32_t first = n ? offsets[0] : 0;
    for (int64_t p = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; p < n; p += (int64_t) gridDim.x * blockDim.x) ends[p] = offsets[p + 1] - first;
}

// ---- CRC-32 (java.util.zip.CRC32 = CRC-32/ISO-HDLC) per 4 KB chunk ---------------------------------------------------------------
constexpr int kCrcChunk = 4096;
__global__ void __launch_bounds__(128) crcChunksKernel(const uint8_t* __restrict__ data, int64_t bytes, int64_t chunks, uint32_t* __restrict__ crcs)
{
    __shared__ uint32_t table[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = (uint32_t) i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        table[i] = c;
    }
    __syncthreads();
    for (int64_t ch = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; ch < chunks; ch += (int64_t) gridDim.x * blockDim.x) {
        const uint8_t* p = data + ch * kCrcChunk;
        const int64_t len = min((int64_t) kCrcChunk, bytes - ch * kCrcChunk);
        uint32_t c = 0xFFFFFFFFu;
        for (int64_t i = 0; i < len; i++) c = table[(c ^ p[i]) & 0xff] ^ (c >> 8);
        crcs[ch] = ~c;
    }
}

// ---- DictionaryBlock.compact (PC/block/DictionaryBlock.java:584-628) on the device: the referenced dictionary entries, renumbered in
// order of first reference ---------------------------------------------------------------------------------------------------------
// ---- LZ4 block codec (lz4_codec.cuh): one thread per 8 KB segment finds / writes its sequences, one warp decodes a block ----
constexpr int kLz4Threads = 64;
__global__ void __launch_bounds__(kLz4Threads) lz4FindKernel(const uint8_t* __restrict__ src, int64_t n, int64_t nSegs, uint16_t* tables, lz4::Match* matches, lz4::SegmentInfo* segs)
{
    const int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x, T = (int64_t) gridDim.x * blockDim.x;
    uint16_t* table = tables + t * lz4::kHashEntries;
    for (int64_t s = t; s < nSegs; s += T) lz4::findMatches(src, n, s, table, matches + s * lz4::kMaxMatchesPerSegment, segs[s]);
}
__global__ void lz4PlanKernel(int64_t n, lz4::SegmentInfo* segs, int64_t nSegs, lz4::BlockInfo* block)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) lz4::planBlock(n, segs, nSegs, *block);
}
__global__ void __launch_bounds__(kLz4Threads) lz4EmitKernel(const uint8_t* __restrict__ src, int64_t n, int64_t nSegs, const lz4::Match* matches, const lz4::SegmentInfo* segs,
                                                             const lz4::BlockInfo* block, uint8_t* dst)
{
    const int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x, T = (int64_t) gridDim.x * blockDim.x;
    for (int64_t s = t; s < nSegs; s += T) lz4::emitSegment(src, n, s, nSegs, matches + s * lz4::kMaxMatchesPerSegment, segs[s], *block, dst);
}
__global__ void lz4DecodeKernel(const uint8_t* src, int64_t srcLen, uint8_t* dst, int64_t dstCap, long long* result)
{
    const long long r = lz4::decodeBlock<true>(src, srcLen, dst, dstCap, (int) threadIdx.x, 32);
    if (threadIdx.x == 0) *result = r;
}

__global__ void dictFillKernel(uint32_t* p, uint32_t v, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) p[i] = v;
}
__global__ void dictFirstUseKernel(const int32_t* __restrict__ ids, int64_t n, int64_t dictSize, uint32_t* firstUse, unsigned int* flags /* [0]=bad id */)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        const int32_t id = ids[i];
        if (id < 0 || id >= dictSize) { atomicOr(flags, 1u); continue; }
        if ((uint32_t) i < firstUse[id]) atomicMin(&firstUse[id], (uint32_t) i);
    }
}
__global__ void dictCountUsedKernel(const uint32_t* firstUse, int64_t dictSize, unsigned int* used, uint32_t* iota)
{
    unsigned int mine = 0;
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < dictSize; i += (int64_t) gridDim.x * blockDim.x) { mine += firstUse[i] != 0xFFFFFFFFu; iota[i] = (uint32_t) i; }
    if (mine) atomicAdd(used, mine);
}
__global__ void dictRemapKernel(const uint32_t* order, int64_t used, int32_t* remap)
{
    for (int64_t r = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; r < used; r += (int64_t) gridDim.x * blockDim.x) remap[order[r]] = (int32_t) r;
}
__global__ void dictApplyRemapKernel(const int32_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ remap, int32_t* __restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) out[i] = remap[ids[i]];
}

// DictionaryId.randomDictionaryId (PC/block/DictionaryId.java:29-32): this process's random node id + a sequence number; a reader keys
// its per-dictionary caches on it, so every written dictionary gets its own
static void nextDictionaryId(int64_t (&id)[3])
{
    static const std::pair<int64_t, int64_t> node = [] { std::random_device rd; auto r = [&] { return ((uint64_t) rd() << 32) | rd(); }; return std::make_pair((int64_t) r(), (int64_t) r()); }();
    static std::atomic<int64_t> sequence{0};
    id[0] = node.first; id[1] = node.second; id[2] = sequence.fetch_add(1);
}

// ---- the serde -----------------------------------------------------------------------------------------------------------------
static const char* flatEncodingName(int type)
{
    switch (type) {
        case PG_TYPE_BIGINT: case PG_TYPE_DOUBLE: return "LONG_ARRAY";
        case PG_TYPE_INTEGER: case PG_TYPE_DATE: return "INT_ARRAY";
        case PG_TYPE_BOOLEAN: return "BYTE_ARRAY";
        case PG_TYPE_VARCHAR: return "VARIABLE_WIDTH";
    }
    throw Error(PG_ERR_UNSUPPORTED, "type has no wire encoding on the GPU path");
}

struct SerdeBase : Operator {         // borrows the Operator plumbing (device, streams, stats); not driven through pg_operator_*
    explicit SerdeBase(const pg_operator_options& o) : Operator(o) {}
    bool needsInput() override { return false; }
    void addInput(const pg_page*) override {}
    bool getOutput(pg_page*) override { return false; }
    void finish() override {}
    bool isFinished() override { return true; }
};

class PagesSerde : public SerdeBase {
    bool checksumEnabled;
    int compression;
    DeviceBuffer lzTables, lzMatches, lzSegs, lzBlock, lzWire, lzRaw;      // LZ4: per-thread hash tables, match records, segment / block plans, the compressed wire, a decoded payload
    PageStager stager;
    Crc32Host crcHost;
    DeviceBuffer wire, crcs, unitCounts, unitOffsets, scanTemp, dWireIn;
    std::vector<DeviceBuffer> temps;                 // aligned temporaries (compacted values, varchar ends) of one call
    PinnedBuffer hostWire, hostScratch;
    DeviceBuffer multiWire; PinnedBuffer multiHost;     // serializeRanges: the SerializedPages of one call back to back
    // deserialized page storage
    std::vector<DeviceBuffer> outValues, outNulls, outOffsets, outIds;
    std::vector<PinnedBuffer> hostValues, hostNulls, hostOffsets, hostIds;
    std::vector<pg_block> outBlocks;

    int gridFor(int64_t items, int perThread = 1) const { return (int) std::max<int64_t>(1, std::min<int64_t>(divUp(items, 256 * (int64_t) perThread), (int64_t) numSMs * 16)); }

    void copyBytes(uint8_t* dst, const uint8_t* src, int64_t bytes)
    {
        if (bytes <= 0) return;
        PG_LAUNCH(this, copyBytesKernel, gridFor(bytes, 16), 256, 0, dst, src, bytes);
    }

    // number of non-null rows; leaves the exclusive scan of the per-unit counts in unitOffsets
    int64_t scanNonNull(const uint8_t* isNull, int64_t n)
    {
        const int64_t units = divUp(n, kSerdeUnit);
        unitCounts.reserve((size_t) (units + 1) * 4); unitOffsets.reserve((size_t) (units + 1) * 4);
        PG_CUDA(cudaMemsetAsync(unitCounts.as<uint32_t>() + units, 0, 4, stream));
        PG_LAUNCH(this, nonNullCountKernel, gridFor(units * 32), 256, 0, isNull, n, units, unitCounts.as<uint32_t>());
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, unitCounts.as<uint32_t>(), unitOffsets.as<uint32_t>(), (int) (units + 1), stream);
        scanTemp.reserve(tb);
        cub::DeviceScan::ExclusiveSum(scanTemp.ptr, tb, unitCounts.as<uint32_t>(), unitOffsets.as<uint32_t>(), (int) (units + 1), stream);
        stats.kernel_launches += 2;
        uint32_t total = 0;
        PG_CUDA(cudaMemcpyAsync(&total, unitOffsets.as<uint32_t>() + units, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        return total;
    }
    template <typename T> void moveNonNullT(const uint8_t* isNull, int64_t n, void* values, void* compact, int dir)
    {
        const int64_t units = divUp(n, kSerdeUnit);
        PG_LAUNCH(this, nonNullMoveKernel<T>, gridFor(units * 32), 256, 0, isNull, n, units, unitOffsets.as<uint32_t>(), (T*) values, (T*) compact, dir);
    }
    void moveNonNull(int width, const uint8_t* isNull, int64_t n, void* values, void* compact, int dir)
    {
        if (width == 8) moveNonNullT<uint64_t>(isNull, n, values, compact, dir);
        else if (width == 4) moveNonNullT<uint32_t>(isNull, n, values, compact, dir);
        else moveNonNullT<uint8_t>(isNull, n, values, compact, dir);
    }

    // CRC32 of `bytes` device bytes, then of the 9 trailer bytes (PagesSerdeUtil.computeSerializedPageChecksum :109-120)
    int64_t checksumOf(const uint8_t* dData, int64_t bytes, uint8_t markers, int32_t positionCount, int32_t uncompressedSize)
    {
        uint32_t crc = 0;
        if (bytes > 0) {
            const int64_t chunks = divUp(bytes, kCrcChunk);
            crcs.reserve((size_t) chunks * 4);
            PG_LAUNCH(this, crcChunksKernel, (int) std::min<int64_t>(divUp(chunks, 128), (int64_t) numSMs * 16), 128, 0, dData, bytes, chunks, crcs.as<uint32_t>());
            std::vector<uint32_t> h((size_t) chunks);
            PG_CUDA(cudaMemcpyAsync(h.data(), crcs.ptr, (size_t) chunks * 4, cudaMemcpyDeviceToHost, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
            stats.d2h_bytes += chunks * 4;
            uint32_t opFull[32], opLast[32];
            const int64_t lastLen = bytes - (chunks - 1) * kCrcChunk;
            if (chunks > 1) Crc32Host::zeroOperator(opFull, kCrcChunk);
            Crc32Host::zeroOperator(opLast, (uint64_t) lastLen);
            crc = h[0];
            for (int64_t c = 1; c < chunks; c++) crc = Crc32Host::combine(crc, h[(size_t) c], c == chunks - 1 ? opLast : opFull);
        }
        uint8_t trailer[9];
        trailer[0] = markers; memcpy(trailer + 1, &positionCount, 4); memcpy(trailer + 5, &uncompressedSize, 4);
        crc = crcHost.update(crc, trailer, 9);
        return (int64_t) (uint64_t) crc;
    }

public:
    explicit PagesSerde(const pg_pages_serde_config* c) : SerdeBase(c->options), checksumEnabled(c->checksum_enabled != 0), compression(c->compression), stager(this)
    {
        PG_CHECK_ARG(compression == PG_COMPRESSION_NONE || compression == PG_COMPRESSION_LZ4, "unknown compression");
        for (DeviceBuffer* d : {&wire, &crcs, &unitCounts, &unitOffsets, &scanTemp, &dWireIn, &lzTables, &lzMatches, &lzSegs, &lzBlock, &lzWire, &lzRaw}) d->accounting = &stats.device_bytes;
    }

    // PagesSerde.wrapSlice :118-143: the raw page W[21, 21 + rawSize) as one LZ4 block, kept when compressed / raw <= MINIMUM_COMPRESSION_RATIO
    // (0.8).  Returns the compressed size and leaves the block at lzWire + 21, or -1 when the page stays uncompressed.
    int64_t tryCompress(const uint8_t* W, int64_t rawSize)
    {
        using namespace lz4;
        const int64_t nSegs = segmentsOf(rawSize);
        const int threads = (int) std::min<int64_t>(divUp(nSegs, kLz4Threads) * kLz4Threads, (int64_t) numSMs * 2 * kLz4Threads);
        const int grid = threads / kLz4Threads;
        lzTables.reserve((size_t) threads * kHashEntries * sizeof(uint16_t));
        lzMatches.reserve((size_t) nSegs * kMaxMatchesPerSegment * sizeof(Match));
        lzSegs.reserve((size_t) nSegs * sizeof(SegmentInfo));
        lzBlock.reserve(sizeof(BlockInfo));
        const uint8_t* raw = W + PG_SERIALIZED_PAGE_HEADER_BYTES;
        PG_LAUNCH(this, lz4FindKernel, grid, kLz4Threads, 0, raw, rawSize, nSegs, lzTables.as<uint16_t>(), lzMatches.as<Match>(), lzSegs.as<SegmentInfo>());
        PG_LAUNCH(this, lz4PlanKernel, 1, 32, 0, rawSize, lzSegs.as<SegmentInfo>(), nSegs, lzBlock.as<BlockInfo>());
        BlockInfo block{};
        PG_CUDA(cudaMemcpyAsync(&block, lzBlock.ptr, sizeof(block), cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        stats.d2h_bytes += (int64_t) sizeof(block);
        if ((double) block.compressedBytes / (double) rawSize > 0.8) return -1;
        lzWire.reserve((size_t) PG_SERIALIZED_PAGE_HEADER_BYTES + block.compressedBytes + 32);
        PG_LAUNCH(this, lz4EmitKernel, grid, kLz4Threads, 0, raw, rawSize, nSegs, lzMatches.as<Match>(), lzSegs.as<SegmentInfo>(), lzBlock.as<BlockInfo>(),
                  lzWire.as<uint8_t>() + PG_SERIALIZED_PAGE_HEADER_BYTES);
        return block.compressedBytes;
    }

    // ================================================================== serialize ==================================================================
    struct ChannelPlan;
    std::vector<std::unique_ptr<PageStager>> dictStagers;      // staging areas of the dictionaries of one call

    // DictionaryBlockEncoding.writeBlock :34-54 writes dictionaryBlock.compact(): bring ids + dictionary to the device, find the
    // referenced entries and their order of first reference, renumber.  Leaves the (compacted) dictionary in P.values / nulls /
    // offsets (P.n entries) and the ids to write in P.ids.
    void planDictionary(const pg_page* page, int c, ChannelPlan& P);

    struct ChannelPlan {
        bool rle = false; int32_t run = 0;          // RLE: run length; the value block is planned like a 1-row flat block
        bool dict = false; const int32_t* ids = nullptr; int64_t idsAt = 0;     // DICTIONARY: `run` = position count, the (compacted) dictionary is planned like a flat block of n entries
        int type = 0; int width = 0; int64_t n = 0;
        const void* values = nullptr; const uint8_t* nulls = nullptr; const int32_t* offsets = nullptr;     // device
        bool mayHaveNull = false;                   // Block.mayHaveNull(): a null map exists (even for 0 rows)
        int64_t nonNull = 0, varcharBytes = 0; int32_t firstOffset = 0;
        int64_t fixedAt = 0, endsAt = 0, nullFlagAt = 0, bitsAt = 0, totalAt = 0, valuesAt = 0, end = 0;
    };

    void serialize(const pg_page* page, const uint8_t** bytes, int64_t* size)
    {
        dictStagers.clear();
        const int64_t n = page->position_count;
        PG_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31), "page exceeds 2^31 positions");
        const int nch = page->channel_count;
        temps.clear();
        // ---- channels to device, flat (RLE blocks keep their single value) ----
        std::vector<int> chans; for (int c = 0; c < nch; c++) chans.push_back(c);
        std::vector<ChannelPlan> plan((size_t) nch);
        // RLE channels are staged as their 1-row value block: build a shadow page whose RLE blocks are 1-row flat blocks is not possible
        // inside one pg_page (blocks share position_count), so RLE channels are staged separately, one page each.
        std::vector<pg_block> flatBlocks; std::vector<int> flatOf((size_t) nch, -1);
        for (int c = 0; c < nch; c++) {
            const pg_block& b = page->blocks[c];
            plan[c].type = b.type; plan[c].width = typeWidth(b.type); plan[c].mayHaveNull = b.is_null != nullptr;
            if (b.encoding == PG_ENC_RLE) { plan[c].rle = true; plan[c].run = (int32_t) b.position_count; plan[c].n = 1; }
            else if (b.encoding == PG_ENC_DICTIONARY) { plan[c].dict = true; plan[c].run = (int32_t) n; plan[c].n = b.dictionary_size; }
            else { plan[c].n = n; flatOf[c] = (int) flatBlocks.size(); flatBlocks.push_back(b); }
        }
        std::vector<DeviceColumn> cols; int slot = -1;
        if (!flatBlocks.empty() && n > 0) {
            pg_page shadow{}; shadow.position_count = n; shadow.channel_count = (int32_t) flatBlocks.size(); shadow.memory = page->memory; shadow.blocks = flatBlocks.data();
            std::vector<int> fch; for (size_t i = 0; i < flatBlocks.size(); i++) fch.push_back((int) i);
            slot = stager.stage(&shadow, fch, cols);
        }
        std::vector<DeviceBuffer> rleStore;
        for (int c = 0; c < nch; c++) {
            ChannelPlan& P = plan[c];
            if (P.rle) {
                // the single value: bring it to the device as a 1-row column (a few bytes)
                const pg_block& b = page->blocks[c];
                const bool host = page->memory == PG_MEM_HOST;
                auto upload = [&](const void* p, size_t bytesN) -> const void* {
                    rleStore.emplace_back(); DeviceBuffer& d = rleStore.back(); d.reserve(std::max<size_t>(bytesN, 16));
                    if (bytesN) PG_CUDA(cudaMemcpyAsync(d.ptr, p, bytesN, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, stream));
                    return d.ptr;
                };
                if (b.type == PG_TYPE_VARCHAR) {
                    int32_t off[2] = {0, 0};
                    PG_CUDA(cudaMemcpyAsync(off, b.offsets, 8, host ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost, stream));
                    PG_CUDA(cudaStreamSynchronize(stream));
                    P.offsets = (const int32_t*) upload(b.offsets, 8);
                    P.values = upload(b.values, (size_t) off[1]);
                }
                else P.values = upload(b.values, (size_t) P.width);
                P.nulls = b.is_null ? (const uint8_t*) upload(b.is_null, 1) : nullptr;
            }
            else if (P.dict) planDictionary(page, c, P);
            else if (n > 0) { const DeviceColumn& dc = cols[(size_t) flatOf[c]]; P.values = dc.values; P.nulls = dc.nulls; P.offsets = dc.offsets; }
        }
        // ---- sizes that depend on the data: non-null counts, varchar byte totals ----
        for (int c = 0; c < nch; c++) {
            ChannelPlan& P = plan[c];
            P.nonNull = P.n;
            if (P.type == PG_TYPE_VARCHAR) {
                if (P.n > 0) {
                    int32_t first = 0, last = 0;
                    PG_CUDA(cudaMemcpyAsync(&first, P.offsets, 4, cudaMemcpyDeviceToHost, stream));
                    PG_CUDA(cudaMemcpyAsync(&last, P.offsets + P.n, 4, cudaMemcpyDeviceToHost, stream));
                    PG_CUDA(cudaStreamSynchronize(stream));
                    P.firstOffset = first; P.varcharBytes = (int64_t) last - first;
                }
            }
            else if (P.nulls && P.n > 0) {
                P.nonNull = scanNonNull(P.nulls, P.n);
                // compact now (the scan result in unitOffsets is reused by the next channel)
                temps.emplace_back(); DeviceBuffer& t = temps.back(); t.reserve(std::max<size_t>((size_t) P.nonNull * P.width, 16));
                moveNonNull(P.width, P.nulls, P.n, const_cast<void*>(P.values), t.ptr, 0);
                P.values = t.ptr;
            }
        }
        // ---- layout ----
        int64_t at = PG_SERIALIZED_PAGE_HEADER_BYTES + 4;          // page metadata, channel count
        for (int c = 0; c < nch; c++) {
            ChannelPlan& P = plan[c];
            P.fixedAt = at;
            if (P.rle) at += 4 + 3 + 4;                            // "RLE", run length
            if (P.dict) at += 4 + 10 + 4;                          // "DICTIONARY", position count
            at += 4 + (int64_t) strlen(flatEncodingName(P.type)) + 4;   // name, position count
            if (P.type == PG_TYPE_VARCHAR) { P.endsAt = at; at += 4 * P.n; }
            P.nullFlagAt = at; at += 1;
            P.bitsAt = at; if (P.mayHaveNull) at += (P.n + 7) / 8;
            if (P.type == PG_TYPE_VARCHAR) { P.totalAt = at; at += 4; P.valuesAt = at; at += P.varcharBytes; }
            else { P.valuesAt = at; at += P.nonNull * P.width; }
            if (P.dict) { P.idsAt = at; at += 4 * (int64_t) P.run + 24; }     // ids, dictionary instance id
            P.end = at;
        }
        const int64_t rawTotal = at;
        PG_CHECK_ARG(rawTotal < (int64_t(1) << 31), "serialized page exceeds 2 GB (SerializedPage sizes are 32-bit)");
        wire.reserve((size_t) rawTotal + 32);
        uint8_t* W = wire.as<uint8_t>();
        // ---- fixed fields: assembled on the host, copied as one patch per channel ----
        hostScratch.reserve((size_t) (64 + 64 * nch));
        uint8_t* hs = hostScratch.as<uint8_t>(); size_t used = 0;
        auto patch = [&](int64_t where, const std::string& data) {
            memcpy(hs + used, data.data(), data.size());
            PG_CUDA(cudaMemcpyAsync(W + where, hs + used, data.size(), cudaMemcpyHostToDevice, stream));
            used += data.size();
        };
        auto le32 = [](int32_t v) { return std::string((const char*) &v, 4); };
        auto named = [&](const char* s) { return le32((int32_t) strlen(s)) + s; };
        hostScratch.reserve((size_t) (64 + 160 * nch));
        hs = hostScratch.as<uint8_t>();
        for (int c = 0; c < nch; c++) {
            const ChannelPlan& P = plan[c];
            std::string head;
            if (P.rle) head += named("RLE") + le32(P.run);
            if (P.dict) {
                head += named("DICTIONARY") + le32(P.run);
                int64_t id[3]; nextDictionaryId(id);
                patch(P.idsAt + 4 * (int64_t) P.run, std::string((const char*) id, 24));
            }
            head += named(flatEncodingName(P.type)) + le32((int32_t) P.n);
            patch(P.fixedAt, head);
            patch(P.nullFlagAt, std::string(1, P.mayHaveNull ? '\1' : '\0'));
            if (P.type == PG_TYPE_VARCHAR) patch(P.totalAt, le32((int32_t) P.varcharBytes));
        }
        // ---- data ----
        for (int c = 0; c < nch; c++) {
            const ChannelPlan& P = plan[c];
            if (P.dict) copyBytes(W + P.idsAt, (const uint8_t*) P.ids, 4 * (int64_t) P.run);
            if (P.n == 0) continue;
            if (P.nulls) PG_LAUNCH(this, packNullBitsKernel, gridFor((P.n + 7) / 8), 256, 0, P.nulls, P.n, W + P.bitsAt);
            if (P.type == PG_TYPE_VARCHAR) {
                temps.emplace_back(); DeviceBuffer& t = temps.back(); t.reserve(std::max<size_t>((size_t) P.n * 4, 16));
                PG_LAUNCH(this, varcharEndsKernel, gridFor(P.n), 256, 0, P.offsets, P.n, t.as<int32_t>());
                copyBytes(W + P.endsAt, t.as<uint8_t>(), P.n * 4);
                copyBytes(W + P.valuesAt, (const uint8_t*) P.values + P.firstOffset, P.varcharBytes);
            }
            else copyBytes(W + P.valuesAt, (const uint8_t*) P.values, P.nonNull * P.width);
        }
        if (slot >= 0) stager.markCompute(slot);
        // ---- page metadata (PagesSerdeUtil.writeSerializedPageMetadata :70-77) ----
        const int32_t rawSize = (int32_t) (rawTotal - PG_SERIALIZED_PAGE_HEADER_BYTES);
        {
            std::string cc = le32(nch);
            patch(PG_SERIALIZED_PAGE_HEADER_BYTES, cc);
        }
        uint8_t markers = 0; int64_t checksum = 0;
        int32_t payload = rawSize;                              // bytes after the 21-byte metadata
        int64_t total = rawTotal;
        if (compression == PG_COMPRESSION_LZ4) {
            const int64_t z = tryCompress(W, rawSize);
            if (z >= 0) { markers |= PG_PAGE_MARKER_COMPRESSED; payload = (int32_t) z; W = lzWire.as<uint8_t>(); total = PG_SERIALIZED_PAGE_HEADER_BYTES + z; }
        }
        if (checksumEnabled) {                                  // over the slice that travels (compressed or not), the markers, the counts
            markers |= PG_PAGE_MARKER_CHECKSUMMED;
            checksum = checksumOf(W + PG_SERIALIZED_PAGE_HEADER_BYTES, payload, markers, (int32_t) n, rawSize);
        }
        std::string meta = le32((int32_t) n) + std::string(1, (char) markers) + le32(rawSize) + le32(payload) + std::string((const char*) &checksum, 8);
        patch(0, meta);
        stats.input_positions += n; stats.output_bytes += total;
        if (options.output_memory == PG_MEM_HOST) {
            hostWire.reserve((size_t) total + 16);
            PG_CUDA(cudaMemcpyAsync(hostWire.ptr, W, (size_t) total, cudaMemcpyDeviceToHost, stream));
            stats.d2h_bytes += total;
            *bytes = hostWire.as<uint8_t>();
        }
        else *bytes = W;
        PG_CUDA(cudaStreamSynchronize(stream));
        temps.clear(); dictStagers.clear();
        *size = total;
    }

    // One SerializedPage per row range of `page` (range r = rows [rowOffsets[r], rowOffsets[r + 1])), back to back in one buffer: the
    // pages OptimizedPartitionedOutputOperator.flush :774-797 hands to OutputBuffer.enqueue, one per partition, from the partitioned
    // page (pg_partitioned_output_create's output + pg_partitioned_output_get_offsets).  byteOffsets gets nRanges + 1 entries.
    void serializeRanges(const pg_page* page, const int64_t* rowOffsets, int32_t nRanges, const uint8_t** bytes, int64_t* byteOffsets)
    {
        PG_CHECK_ARG(nRanges >= 0 && rowOffsets != nullptr, "row offsets are missing");
        const bool toHost = options.output_memory == PG_MEM_HOST;
        std::vector<std::vector<uint8_t>> hostParts;
        int64_t at = 0;
        byteOffsets[0] = 0;
        size_t capacity = 0;
        for (int32_t r = 0; r < nRanges; r++) {
            const int64_t lo = rowOffsets[r], hi = rowOffsets[r + 1];
            PG_CHECK_ARG(lo >= 0 && lo <= hi && hi <= page->position_count, "row offsets must be ascending and inside the page");
            std::vector<pg_block> blocks(page->blocks, page->blocks + page->channel_count);
            for (pg_block& b : blocks) {
                if (b.encoding == PG_ENC_FLAT) {
                    const int w = typeWidth(b.type);
                    if (w) b.values = (const char*) b.values + lo * w; else if (b.offsets) b.offsets += lo;
                    if (b.is_null) b.is_null += lo;
                }
                else if (b.encoding == PG_ENC_DICTIONARY) b.ids += lo;
                b.position_count = hi - lo;
            }
            pg_page part = *page; part.blocks = blocks.data(); part.position_count = hi - lo;
            const uint8_t* one = nullptr; int64_t size = 0;
            serialize(&part, &one, &size);
            if (toHost) {
                if ((size_t) (at + size) > capacity) {
                    // grow, keeping what is there (PinnedBuffer::reserve does not preserve contents)
                    std::vector<uint8_t> keep((const uint8_t*) multiHost.ptr, (const uint8_t*) multiHost.ptr + at);
                    capacity = std::max<size_t>((size_t) (at + size) * 2, 4096);
                    multiHost.reserve(capacity);
                    if (at) memcpy(multiHost.ptr, keep.data(), (size_t) at);
                }
                memcpy((uint8_t*) multiHost.ptr + at, one, (size_t) size);
            }
            else {
                if ((size_t) (at + size) > capacity) {
                    DeviceBuffer bigger; bigger.accounting = &stats.device_bytes;
                    capacity = std::max<size_t>((size_t) (at + size) * 2, 4096);
                    bigger.reserve(capacity);
                    if (at) PG_CUDA(cudaMemcpyAsync(bigger.ptr, multiWire.ptr, (size_t) at, cudaMemcpyDeviceToDevice, stream));
                    PG_CUDA(cudaStreamSynchronize(stream));
                    multiWire = std::move(bigger);
                }
                PG_CUDA(cudaMemcpyAsync((uint8_t*) multiWire.ptr + at, one, (size_t) size, cudaMemcpyDeviceToDevice, stream));
                PG_CUDA(cudaStreamSynchronize(stream));
            }
            at += size;
            byteOffsets[r + 1] = at;
        }
        *bytes = toHost ? multiHost.as<uint8_t>() : multiWire.as<uint8_t>();
    }

    // ================================================================= deserialize =================================================================
    // a small window of the wire on the host (the wire itself may be device memory)
    struct WireReader {
        PagesSerde* self; const uint8_t* hostBytes; const uint8_t* devBytes; int64_t size; int64_t at = 0;
        void need(int64_t n) const { if (at + n > size) throw Error(PG_ERR_INVALID_ARGUMENT, "serialized page is truncated"); }
        uint8_t window[256]; int64_t winAt = -1, winLen = 0;      // device wire: the fixed fields of a channel arrive in one small copy
        void fetch(void* out, int64_t n)
        {
            need(n);
            if (hostBytes) memcpy(out, hostBytes + at, (size_t) n);
            else {
                if (n > (int64_t) sizeof(window)) throw Error(PG_ERR_INVALID_ARGUMENT, "bad block encoding name");
                if (winAt < 0 || at < winAt || at + n > winAt + winLen) {
                    winAt = at; winLen = std::min<int64_t>((int64_t) sizeof(window), size - at);
                    PG_CUDA(cudaMemcpyAsync(window, devBytes + at, (size_t) winLen, cudaMemcpyDeviceToHost, self->stream));
                    PG_CUDA(cudaStreamSynchronize(self->stream));
                    self->stats.d2h_bytes += winLen;
                }
                memcpy(out, window + (at - winAt), (size_t) n);
            }
            at += n;
        }
        int32_t i32() { int32_t v; fetch(&v, 4); return v; }
        int64_t i64() { int64_t v; fetch(&v, 8); return v; }
        uint8_t u8() { uint8_t v; fetch(&v, 1); return v; }
        std::string name() { const int32_t n = i32(); if (n < 0 || n > 64) throw Error(PG_ERR_INVALID_ARGUMENT, "bad block encoding name"); std::string s((size_t) n, '\0'); fetch(&s[0], n); return s; }
        void skip(int64_t n) { need(n); at += n; }
    };

    struct FlatPlan { int type = 0, width = 0; int64_t n = 0; bool hasNulls = false; int64_t bitsAt = 0, endsAt = 0, valuesAt = 0, valueBytes = 0; };

    // parses one flat block whose encoding name has been consumed; all data stays on the wire
    FlatPlan parseFlat(WireReader& r, const std::string& enc, int type)
    {
        if (enc != flatEncodingName(type)) throw Error(PG_ERR_INVALID_ARGUMENT, "block encoding " + enc + " does not fit the channel type");
        FlatPlan F; F.type = type; F.width = typeWidth(type);
        F.n = r.i32();
        PG_CHECK_ARG(F.n >= 0, "negative position count");
        if (type == PG_TYPE_VARCHAR) { F.endsAt = r.at; r.skip(4 * F.n); }
        F.hasNulls = r.u8() != 0;
        if (F.hasNulls) { F.bitsAt = r.at; r.skip((F.n + 7) / 8); }
        if (type == PG_TYPE_VARCHAR) { F.valueBytes = r.i32(); PG_CHECK_ARG(F.valueBytes >= 0, "negative slice length"); F.valuesAt = r.at; r.skip(F.valueBytes); }
        else F.valuesAt = r.at;      // length depends on the non-null count: resolved by materialize()
        return F;
    }

    // expands one flat block from the device copy of the wire into aligned device storage; returns the wire position after its values
    int64_t materialize(const uint8_t* dW, int64_t wireSize, const FlatPlan& F, DeviceBuffer& values, DeviceBuffer& nulls, DeviceBuffer& offsets)
    {
        const int64_t n = F.n;
        if (F.hasNulls) { nulls.reserve(std::max<size_t>((size_t) n, 16)); if (n) PG_LAUNCH(this, unpackNullBitsKernel, gridFor(n), 256, 0, dW + F.bitsAt, n, nulls.as<uint8_t>()); }
        if (F.type == PG_TYPE_VARCHAR) {
            offsets.reserve((size_t) (n + 1) * 4 + 16);
            PG_CUDA(cudaMemsetAsync(offsets.ptr, 0, 4, stream));
            copyBytes(offsets.as<uint8_t>() + 4, dW + F.endsAt, 4 * n);
            values.reserve(std::max<size_t>((size_t) F.valueBytes, 16));
            copyBytes(values.as<uint8_t>(), dW + F.valuesAt, F.valueBytes);
            return F.valuesAt + F.valueBytes;
        }
        values.reserve(std::max<size_t>((size_t) n * F.width, 16));
        int64_t nonNull = n;
        if (F.hasNulls && n > 0) nonNull = scanNonNull(nulls.as<uint8_t>(), n);
        const int64_t bytes = nonNull * F.width;
        if (F.valuesAt + bytes > wireSize) throw Error(PG_ERR_INVALID_ARGUMENT, "serialized page is truncated");
        if (!F.hasNulls) copyBytes(values.as<uint8_t>(), dW + F.valuesAt, bytes);
        else if (n > 0) {
            temps.emplace_back(); DeviceBuffer& t = temps.back(); t.reserve(std::max<size_t>((size_t) bytes, 16));
            copyBytes(t.as<uint8_t>(), dW + F.valuesAt, bytes);
            moveNonNull(F.width, nulls.as<uint8_t>(), n, values.ptr, t.ptr, 1);
        }
        return F.valuesAt + bytes;
    }

    void deserialize(const uint8_t* bytes, int64_t size, int32_t memory, const int32_t* types, int32_t nTypes, pg_page* out)
    {
        PG_CHECK_ARG(bytes != nullptr && size >= PG_SERIALIZED_PAGE_HEADER_BYTES, "serialized page is truncated");
        PG_CHECK_ARG(memory == PG_MEM_HOST || memory == PG_MEM_DEVICE, "bad memory kind");
        temps.clear();
        const bool host = memory == PG_MEM_HOST;
        // the wire on the device (host bytes: one H2D copy — the ingest)
        const uint8_t* dW = bytes;
        if (host) {
            dWireIn.reserve((size_t) size + 32);
            PG_CUDA(cudaMemcpyAsync(dWireIn.ptr, bytes, (size_t) size, cudaMemcpyHostToDevice, stream));
            stats.h2d_bytes += size;
            dW = dWireIn.as<uint8_t>();
        }
        WireReader r{this, host ? bytes : nullptr, host ? nullptr : bytes, size};
        const int32_t positionCount = r.i32(); const uint8_t markers = r.u8();          // PagesSerdeUtil.readSerializedPage :79-88
        const int32_t uncompressed = r.i32(), sz = r.i32(); const int64_t checksum = r.i64();
        if (markers & PG_PAGE_MARKER_ENCRYPTED) throw Error(PG_ERR_UNSUPPORTED, "encrypted pages are not supported on the GPU path");
        PG_CHECK_ARG(sz >= 1 && (int64_t) sz + PG_SERIALIZED_PAGE_HEADER_BYTES <= size, "serialized page is truncated");
        int64_t wireSize = (int64_t) sz + PG_SERIALIZED_PAGE_HEADER_BYTES;
        r.size = wireSize;
        if ((markers & PG_PAGE_MARKER_CHECKSUMMED) && checksumOf(dW + PG_SERIALIZED_PAGE_HEADER_BYTES, sz, markers, positionCount, uncompressed) != checksum)
            throw Error(PG_ERR_INVALID_ARGUMENT, "serialized page checksum mismatch");
        if (markers & PG_PAGE_MARKER_COMPRESSED) {              // PagesSerde.deserialize :95-107: one LZ4 block → uncompressedSize bytes
            PG_CHECK_ARG(uncompressed >= 4, "compressed page with a bad uncompressed size");
            lzRaw.reserve((size_t) PG_SERIALIZED_PAGE_HEADER_BYTES + (size_t) uncompressed + 32);
            lzBlock.reserve(sizeof(lz4::BlockInfo));
            // one warp decodes the block (lz4::decodeBlock<true>): sequences are inherently sequential — about 600 cycles each on a single warp
            // (measured: 24 MB/s on a 64 MB page of 8-byte sequences; a variant with the input staged and the history kept in a shared-memory
            // ring measured the same, the time is the dependent instruction chain, not memory) — pages decode in parallel with each other
            PG_LAUNCH(this, lz4DecodeKernel, 1, 32, 0, dW + PG_SERIALIZED_PAGE_HEADER_BYTES, (int64_t) sz, lzRaw.as<uint8_t>() + PG_SERIALIZED_PAGE_HEADER_BYTES, (int64_t) uncompressed,
                      lzBlock.as<long long>());
            long long decoded = -1;
            PG_CUDA(cudaMemcpyAsync(&decoded, lzBlock.ptr, 8, cudaMemcpyDeviceToHost, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
            stats.d2h_bytes += 8;
            if (decoded != uncompressed) throw Error(PG_ERR_INVALID_ARGUMENT, "page size changed after decompression (malformed LZ4 block)");     // :104 checkState
            // the rest of the call reads the decoded raw page: on the device, behind a metadata-sized gap so that positions keep their meaning
            dW = lzRaw.as<uint8_t>();
            wireSize = (int64_t) uncompressed + PG_SERIALIZED_PAGE_HEADER_BYTES;
            const int64_t at = r.at;
            r = WireReader{this, nullptr, dW, wireSize};
            r.at = at;
        }
        PG_CHECK_ARG(wireSize >= PG_SERIALIZED_PAGE_HEADER_BYTES + 4, "serialized page is truncated");
        const int32_t channels = r.i32();                                               // readRawPage :53-62
        PG_CHECK_ARG(channels == nTypes, "channel count of the page differs from the types given");
        if ((int) outValues.size() < channels) {
            outValues.resize(channels); outNulls.resize(channels); outOffsets.resize(channels); outIds.resize(channels);
            for (auto* v : {&outValues, &outNulls, &outOffsets, &outIds}) for (auto& d : *v) d.accounting = &stats.device_bytes;
        }
        outBlocks.assign((size_t) channels, pg_block{});
        std::vector<FlatPlan> flats((size_t) channels);
        for (int ch = 0; ch < channels; ch++) {
            pg_block& b = outBlocks[(size_t) ch];
            const int type = types[ch];
            b.type = type; b.encoding = PG_ENC_FLAT; b.dictionary_size = 0;
            std::string enc = r.name();
            int64_t idsAt = -1;
            if (enc == "RLE") {                                   // RunLengthBlockEncoding.readBlock :43-53
                const int32_t run = r.i32();
                PG_CHECK_ARG(run >= 0, "negative run length");
                enc = r.name();
                b.encoding = PG_ENC_RLE; b.position_count = run; b.dictionary_size = 1;
            }
            else if (enc == "DICTIONARY") {                       // DictionaryBlockEncoding.readBlock :57-80
                const int32_t count = r.i32();
                PG_CHECK_ARG(count >= 0, "negative position count");
                enc = r.name();
                b.encoding = PG_ENC_DICTIONARY; b.position_count = count;
            }
            FlatPlan F = parseFlat(r, enc, type);
            // positions after the values depend on the non-null count
            const int64_t after = materialize(dW, wireSize, F, outValues[ch], outNulls[ch], outOffsets[ch]);
            r.at = after;
            if (b.encoding == PG_ENC_RLE) PG_CHECK_ARG(F.n == 1, "RLE value block must hold one position");
            if (b.encoding == PG_ENC_DICTIONARY) {
                b.dictionary_size = F.n;
                idsAt = r.at; r.skip(4 * b.position_count);
                r.skip(24);                                       // dictionary instance id
                outIds[ch].reserve(std::max<size_t>((size_t) b.position_count * 4, 16));
                copyBytes(outIds[ch].as<uint8_t>(), dW + idsAt, 4 * b.position_count);
            }
            if (b.encoding == PG_ENC_FLAT) b.position_count = F.n;
            flats[(size_t) ch] = F;
        }
        // ---- hand the blocks out (device pointers, or pinned host copies) ----
        const bool toHost = options.output_memory == PG_MEM_HOST;
        if (toHost && (int) hostValues.size() < channels) { hostValues.resize(channels); hostNulls.resize(channels); hostOffsets.resize(channels); hostIds.resize(channels); }
        for (int ch = 0; ch < channels; ch++) {
            pg_block& b = outBlocks[(size_t) ch]; const FlatPlan& F = flats[(size_t) ch];
            const size_t vbytes = F.type == PG_TYPE_VARCHAR ? (size_t) F.valueBytes : (size_t) F.n * F.width;
            auto give = [&](DeviceBuffer& d, PinnedBuffer* h, size_t nbytes) -> const void* {
                if (!toHost) return d.ptr;
                h->reserve(std::max<size_t>(nbytes, 16));
                if (nbytes) { PG_CUDA(cudaMemcpyAsync(h->ptr, d.ptr, nbytes, cudaMemcpyDeviceToHost, stream)); stats.d2h_bytes += (int64_t) nbytes; }
                return h->ptr;
            };
            b.values = give(outValues[ch], toHost ? &hostValues[ch] : nullptr, vbytes);
            b.is_null = F.hasNulls ? (const uint8_t*) give(outNulls[ch], toHost ? &hostNulls[ch] : nullptr, (size_t) F.n) : nullptr;
            b.offsets = F.type == PG_TYPE_VARCHAR ? (const int32_t*) give(outOffsets[ch], toHost ? &hostOffsets[ch] : nullptr, (size_t) (F.n + 1) * 4) : nullptr;
            b.ids = b.encoding == PG_ENC_DICTIONARY ? (const int32_t*) give(outIds[ch], toHost ? &hostIds[ch] : nullptr, (size_t) b.position_count * 4) : nullptr;
        }
        PG_CUDA(cudaStreamSynchronize(stream));
        temps.clear();
        out->position_count = positionCount; out->channel_count = channels; out->memory = options.output_memory; out->blocks = outBlocks.data();
        stats.output_positions += positionCount;
    }
};

void PagesSerde::planDictionary(const pg_page* page, int c, ChannelPlan& P)
{
    const pg_block& b = page->blocks[c];
    const int64_t n = b.position_count, D = b.dictionary_size;
    PG_CHECK_ARG(D >= 0 && D < (int64_t(1) << 31) && b.ids != nullptr, "DICTIONARY block without ids");
    const bool host = page->memory == PG_MEM_HOST;
    temps.reserve(temps.size() + 16);        // the references taken below must survive the emplace_backs that follow them
    // ids on the device
    const int32_t* ids = b.ids;
    if (host) {
        temps.emplace_back(); DeviceBuffer& t = temps.back(); t.reserve(std::max<size_t>((size_t) n * 4, 16));
        PG_CUDA(cudaMemcpyAsync(t.ptr, b.ids, (size_t) n * 4, cudaMemcpyHostToDevice, stream));
        stats.h2d_bytes += n * 4;
        ids = t.as<int32_t>();
    }
    // the dictionary itself as a flat column of D entries
    auto stageFlat = [&](const pg_block& blk, int64_t rows, int memory, DeviceColumn& out) {
        dictStagers.emplace_back(new PageStager(this));
        pg_block one = blk; one.position_count = rows;
        pg_page shadow{}; shadow.position_count = rows; shadow.channel_count = 1; shadow.memory = memory; shadow.blocks = &one;
        std::vector<DeviceColumn> cols;
        if (rows > 0) { dictStagers.back()->stage(&shadow, std::vector<int>{0}, cols); out = cols[0]; }
        else { out = DeviceColumn(); out.type = blk.type; }
    };
    pg_block flatDict = b; flatDict.encoding = PG_ENC_FLAT; flatDict.ids = nullptr; flatDict.dictionary_size = 0;
    DeviceColumn dict;
    stageFlat(flatDict, D, page->memory, dict);
    // first reference of every entry
    temps.emplace_back(); DeviceBuffer& firstUse = temps.back(); firstUse.reserve(std::max<size_t>((size_t) D * 4, 16));
    temps.emplace_back(); DeviceBuffer& iota = temps.back(); iota.reserve(std::max<size_t>((size_t) D * 4, 16));
    temps.emplace_back(); DeviceBuffer& flags = temps.back(); flags.reserve(16);
    PG_CUDA(cudaMemsetAsync(flags.ptr, 0, 8, stream));
    if (D > 0) PG_LAUNCH(this, dictFillKernel, gridFor(D), 256, 0, firstUse.as<uint32_t>(), 0xFFFFFFFFu, D);
    PG_LAUNCH(this, dictFirstUseKernel, gridFor(n), 256, 0, ids, n, D, firstUse.as<uint32_t>(), flags.as<unsigned int>());
    if (D > 0) PG_LAUNCH(this, dictCountUsedKernel, gridFor(D), 256, 0, firstUse.as<uint32_t>(), D, flags.as<unsigned int>() + 1, iota.as<uint32_t>());
    unsigned int h[2] = {0, 0};
    PG_CUDA(cudaMemcpyAsync(h, flags.ptr, 8, cudaMemcpyDeviceToHost, stream));
    PG_CUDA(cudaStreamSynchronize(stream));
    stats.d2h_bytes += 8;
    if (h[0]) throw Error(PG_ERR_INVALID_ARGUMENT, "reference to a non-existent key");      // DictionaryBlock.compact :615-617
    const int64_t used = h[1];
    P.ids = ids;
    if (used == D) { P.values = dict.values; P.nulls = dict.nulls; P.offsets = dict.offsets; P.n = D; P.mayHaveNull = b.is_null != nullptr; return; }      // entire dictionary is referenced :607-609
    // order of first reference → new numbering
    temps.emplace_back(); DeviceBuffer& keysOut = temps.back(); keysOut.reserve((size_t) D * 4);
    temps.emplace_back(); DeviceBuffer& order = temps.back(); order.reserve((size_t) D * 4);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, firstUse.as<uint32_t>(), keysOut.as<uint32_t>(), iota.as<uint32_t>(), order.as<uint32_t>(), (int) D, 0, 32, stream);
    temps.emplace_back(); DeviceBuffer& sortTemp = temps.back(); sortTemp.reserve(tb);
    cub::DeviceRadixSort::SortPairs(sortTemp.ptr, tb, firstUse.as<uint32_t>(), keysOut.as<uint32_t>(), iota.as<uint32_t>(), order.as<uint32_t>(), (int) D, 0, 32, stream);
    stats.kernel_launches += 3;
    temps.emplace_back(); DeviceBuffer& remap = temps.back(); remap.reserve((size_t) D * 4);
    temps.emplace_back(); DeviceBuffer& newIds = temps.back(); newIds.reserve(std::max<size_t>((size_t) n * 4, 16));
    PG_LAUNCH(this, dictRemapKernel, gridFor(used), 256, 0, order.as<uint32_t>(), used, remap.as<int32_t>());
    PG_LAUNCH(this, dictApplyRemapKernel, gridFor(n), 256, 0, ids, n, remap.as<int32_t>(), newIds.as<int32_t>());
    P.ids = newIds.as<int32_t>();
    // dictionary.copyPositions(referenced entries in that order): a DICTIONARY view over the staged dictionary, flattened by the stager
    pg_block view{};
    view.type = b.type; view.encoding = PG_ENC_DICTIONARY; view.position_count = used; view.values = dict.values; view.is_null = dict.nulls;
    view.offsets = dict.offsets; view.ids = (const int32_t*) order.ptr; view.dictionary_size = D;
    DeviceColumn compacted;
    stageFlat(view, used, PG_MEM_DEVICE, compacted);
    P.values = compacted.values; P.nulls = compacted.nulls; P.offsets = compacted.offsets; P.n = used;
    P.mayHaveNull = b.is_null != nullptr;      // copyPositions keeps a null array when its source has one (LongArrayBlock.copyPositions)
}

}  // namespace pg

#define S_TRY try {
#define S_CATCH } catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; } catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; } return PG_OK;

extern "C" {
int32_t pg_pages_serde_create(const pg_pages_serde_config* config, pg_pages_serde** out)
{
    S_TRY PG_CHECK_ARG(config && out, "null argument"); *out = (pg_pages_serde*) new pg::PagesSerde(config); S_CATCH
}
int32_t pg_pages_serde_serialize(pg_pages_serde* serde, const pg_page* page, const uint8_t** bytes, int64_t* size)
{
    S_TRY PG_CHECK_ARG(serde && page && bytes && size, "null argument"); auto* s = (pg::PagesSerde*) serde; s->bind(); s->serialize(page, bytes, size); S_CATCH
}
int32_t pg_pages_serde_serialize_ranges(pg_pages_serde* serde, const pg_page* page, const int64_t* row_offsets, int32_t n_ranges, const uint8_t** bytes, int64_t* byte_offsets)
{
    S_TRY PG_CHECK_ARG(serde && page && bytes && byte_offsets, "null argument"); auto* s = (pg::PagesSerde*) serde; s->bind(); s->serializeRanges(page, row_offsets, n_ranges, bytes, byte_offsets); S_CATCH
}
int32_t pg_pages_serde_deserialize(pg_pages_serde* serde, const uint8_t* bytes, int64_t size, int32_t memory, const int32_t* types, int32_t n_types, pg_page* out)
{
    S_TRY PG_CHECK_ARG(serde && out, "null argument"); auto* s = (pg::PagesSerde*) serde; s->bind(); s->deserialize(bytes, size, memory, types, n_types, out); S_CATCH
}
int32_t pg_pages_serde_destroy(pg_pages_serde* serde)
{
    S_TRY if (serde) { auto* s = (pg::PagesSerde*) serde; s->bind(); delete s; } S_CATCH
}
}
