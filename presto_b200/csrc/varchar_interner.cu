// presto_b200 — VarcharInterner (varchar_interner.h): open-addressing table of slices keyed by XXH64, values stored once in a byte
// pool.  One thread per row: rows are short strings (group / join keys), the work per row is a hash over the bytes, one or two
// L2-resident slot reads and one byte comparison — the page's bytes are read once from HBM (hash) and once more from L2 (compare).
#include "varchar_interner.h"

namespace pg {

constexpr unsigned long long kInEmpty = 0, kInLocked = 1;
constexpr int64_t kInternChunk = int64_t(1) << 22;     // rows between two capacity checks (worst case: every row a new value)

// `pool8`: the value's first byte in the pool, 8-byte aligned.  Pool lines may have been filled by another SM moments ago: they are read at
// L2 (ld.global.cg), a word at a time; the row's slice starts anywhere, its words are assembled from byte loads that hit L1.
__device__ __forceinline__ bool sliceEqualsPool(const uint8_t* pool8, const uint8_t* p, uint32_t len)
{
    uint32_t i = 0;
    for (; i + 8 <= len; i += 8) if (__ldcg((const unsigned long long*) (pool8 + i)) != loadLe64(p + i)) return false;
    if (i < len) {
        const uint32_t rest = len - i;
        uint64_t v = 0;
        for (uint32_t k = 0; k < rest; k++) v |= (uint64_t) p[i + k] << (8 * k);
        if ((__ldcg((const unsigned long long*) (pool8 + i)) & ((1ULL << (8 * rest)) - 1)) != v) return false;      // bytes past the value are padding
    }
    return true;
}

template <bool INSERT>
__global__ void __launch_bounds__(256) internKernel(InternView T, const uint8_t* bytes, const int32_t* offsets, const uint8_t* nulls, int64_t n, int32_t* idsOut)
{
    for (int64_t row = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; row < n; row += (int64_t) gridDim.x * blockDim.x) {
        if (nulls && nulls[row]) { idsOut[row] = 0; continue; }
        const int32_t o = offsets[row];
        const uint32_t len = (uint32_t) (offsets[row + 1] - o);
        const uint8_t* p = bytes + o;
        const uint64_t h = (uint64_t) xxHash64Bytes(p, (int64_t) len);
        const unsigned long long tag = (h >> 32) << 32;
        uint32_t slot = (uint32_t) h & T.mask;
        volatile unsigned long long* slots = T.slots;
        int32_t result = -1;
        for (uint32_t probe = 0; probe < T.cap; probe++) {
            unsigned long long st = slots[slot];
            if (st == kInEmpty) {
                if (!INSERT) break;
                const unsigned long long old = atomicCAS(&T.slots[slot], (unsigned long long) kInEmpty, (unsigned long long) kInLocked);
                if (old == kInEmpty) {
                    const uint32_t id = (uint32_t) atomicAdd(&T.counters[0], 1ULL);
                    const uint64_t off = atomicAdd(&T.counters[1], (unsigned long long) ((len + 7u) & ~7u));      // values start 8-byte aligned
                    for (uint32_t i = 0; i < len; i++) T.pool[off + i] = p[i];
                    T.entryMeta[id] = ((off >> 3) << 31) | len; T.entryHash[id] = h;
                    __threadfence();
                    slots[slot] = tag | (id + 2);          // publish: readers that saw LOCKED have been spinning on this word
                    result = (int32_t) id;
                    break;
                }
                st = old;
            }
            while (st == kInLocked) st = slots[slot];
            if ((st & 0xffffffff00000000ULL) == tag) {
                const uint32_t id = (uint32_t) st - 2;
                const uint64_t meta = __ldcg(&T.entryMeta[id]);
                if (internLengthOf(meta) == len && sliceEqualsPool(T.pool + internOffsetOf(meta), p, len)) { result = (int32_t) id; break; }
            }
            slot = (slot + 1) & T.mask;
        }
        idsOut[row] = result;
    }
}

__global__ void internRehashKernel(InternView T, uint64_t nEntries)
{
    for (uint64_t id = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; id < nEntries; id += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t h = T.entryHash[id];
        uint32_t slot = (uint32_t) h & T.mask;
        while (atomicCAS(&T.slots[slot], kInEmpty, ((h >> 32) << 32) | (id + 2)) != kInEmpty) slot = (slot + 1) & T.mask;
    }
}

__global__ void internLengthsKernel(InternView T, const int32_t* ids, const uint8_t* nulls, int64_t g, int32_t* lengthsOut)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < g; i += (int64_t) gridDim.x * blockDim.x)
        lengthsOut[i] = (nulls && nulls[i]) ? 0 : (int32_t) internLengthOf(T.entryMeta[ids[i]]);
}

// one warp per value: lanes stride over its bytes (values of a few bytes cost one partly filled transaction, long ones stay coalesced)
__global__ void internBytesKernel(InternView T, const int32_t* ids, const uint8_t* nulls, const int32_t* offsets, int64_t g, uint8_t* bytesOut)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5, warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < g; i += warps) {
        if (nulls && nulls[i]) continue;
        const int32_t id = ids[i];
        const uint64_t meta = T.entryMeta[id];
        const uint8_t* src = T.pool + internOffsetOf(meta);
        const uint32_t len = internLengthOf(meta);
        uint8_t* dst = bytesOut + offsets[i];
        for (uint32_t b = lane; b < len; b += 32) dst[b] = src[b];
    }
}

VarcharInterner::VarcharInterner(Operator* op, int64_t* accounting)
{
    for (DeviceBuffer* b : {&slots, &entryMeta, &entryHash, &pool, &counters}) b->accounting = accounting;
    counters.reserve(2 * sizeof(unsigned long long));
    PG_CUDA(cudaMemsetAsync(counters.ptr, 0, 2 * sizeof(unsigned long long), op->stream));
    host.reserve(64);
}

InternView VarcharInterner::view() const
{
    InternView v{};
    v.slots = slots.as<unsigned long long>(); v.cap = cap; v.mask = cap ? cap - 1 : 0;
    v.entryMeta = entryMeta.as<uint64_t>(); v.entryHash = entryHash.as<uint64_t>();
    v.pool = pool.as<uint8_t>(); v.counters = counters.as<unsigned long long>();
    return v;
}

void VarcharInterner::readCounters(Operator* op)
{
    PG_CUDA(cudaMemcpyAsync(host.ptr, counters.ptr, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, op->stream));
    PG_CUDA(cudaStreamSynchronize(op->stream));
    nEntries = host.as<unsigned long long>()[0]; poolUsed = host.as<unsigned long long>()[1];
    op->stats.d2h_bytes += 16;
}

// Room for `extraEntries` more values of `extraBytes` bytes in total, whatever the page holds (called between kernels; the stream is
// synchronised before an outgrown buffer goes back to the allocator).
void VarcharInterner::ensure(Operator* op, uint64_t extraEntries, uint64_t extraBytes)
{
    cudaStream_t stream = op->stream;
    const uint64_t needEntries = nEntries + extraEntries, needBytes = poolUsed + extraBytes + 8 * extraEntries + 8;     // every value is padded to 8 bytes
    if (needEntries > (1ULL << 30)) throw Error(PG_ERR_INSUFFICIENT_RESOURCES, "more than 2^30 distinct VARCHAR key values (the table's slots are indexed with 32 bits)");
    auto regrow = [&](DeviceBuffer& buf, size_t newBytes, size_t keepBytes) {
        DeviceBuffer bigger; bigger.accounting = buf.accounting;
        bigger.reserve(newBytes);
        if (keepBytes) PG_CUDA(cudaMemcpyAsync(bigger.ptr, buf.ptr, keepBytes, cudaMemcpyDeviceToDevice, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        buf = std::move(bigger);
    };
    if (needEntries > entryCap) {
        const uint64_t newCap = std::max<uint64_t>(std::max<uint64_t>(entryCap * 2, needEntries), 1024);
        regrow(entryMeta, newCap * 8, nEntries * 8);
        regrow(entryHash, newCap * 8, nEntries * 8);
        entryCap = newCap;
    }
    if (needBytes > poolCap) {
        const uint64_t newCap = std::max<uint64_t>(std::max<uint64_t>(poolCap * 2, needBytes), 4096);
        regrow(pool, newCap, poolUsed);
        poolCap = newCap;
    }
    if (needEntries * 2 > cap) {
        uint64_t newCap = std::max<uint32_t>(cap, 1024);
        while (newCap < needEntries * 2) newCap <<= 1;
        DeviceBuffer bigger; bigger.accounting = slots.accounting;
        bigger.reserve(newCap * 8);
        PG_CUDA(cudaMemsetAsync(bigger.ptr, 0, newCap * 8, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        slots = std::move(bigger);
        cap = (uint32_t) newCap;
        if (nEntries) PG_LAUNCH(op, internRehashKernel, (int) std::min<int64_t>(divUp((int64_t) nEntries, 256), 2048), 256, 0, view(), nEntries);
    }
}

void VarcharInterner::intern(Operator* op, const DeviceColumn& col, int64_t n, int32_t* idsOut, bool insert)
{
    PG_CHECK_ARG(col.type == PG_TYPE_VARCHAR && col.offsets != nullptr, "VARCHAR key channel without offsets");
    if (n <= 0) return;
    cudaStream_t stream = op->stream;
    if (!insert) {      // read-only (cap == 0: nothing was ever interned, every value is absent)
        PG_LAUNCH(op, internKernel<false>, (int) std::min<int64_t>(divUp(n, 256), (int64_t) op->numSMs * 8), 256, 0, view(), (const uint8_t*) col.values, col.offsets,
                  col.nulls, n, idsOut);
        return;
    }
    for (int64_t begin = 0; begin < n; begin += kInternChunk) {
        const int64_t m = std::min(kInternChunk, n - begin);
        // what the table holds now, and how many bytes this range can add (two offsets of the page)
        int32_t* ends = host.as<int32_t>() + 8;
        PG_CUDA(cudaMemcpyAsync(ends, col.offsets + begin, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaMemcpyAsync(ends + 1, col.offsets + begin + m, 4, cudaMemcpyDeviceToHost, stream));
        readCounters(op);
        PG_CHECK_ARG(ends[1] >= ends[0], "VARCHAR offsets must not decrease");
        ensure(op, (uint64_t) m, (uint64_t) (ends[1] - ends[0]));
        PG_LAUNCH(op, internKernel<true>, (int) std::min<int64_t>(divUp(m, 256), (int64_t) op->numSMs * 8), 256, 0, view(), (const uint8_t*) col.values,
                  col.offsets + begin, col.nulls ? col.nulls + begin : nullptr, m, idsOut + begin);
    }
    readCounters(op);
}

__global__ void internMapDictionaryKernel(const int32_t* __restrict__ entryIds, const uint8_t* __restrict__ entryNulls, const int32_t* __restrict__ ids, int64_t n,
                                          int32_t* __restrict__ idsOut, uint8_t* __restrict__ nullsOut)
{
    for (int64_t row = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; row < n; row += (int64_t) gridDim.x * blockDim.x) {
        const int32_t e = ids ? ids[row] : 0;        // RLE: the single entry
        idsOut[row] = entryIds[e];
        if (nullsOut) nullsOut[row] = entryNulls ? entryNulls[e] : 0;
    }
}

void VarcharInterner::internEncoded(Operator* op, const DeviceColumn& col, int64_t n, int32_t* idsOut, uint8_t* nullsOut, bool insert)
{
    PG_CHECK_ARG(col.encoding != PG_ENC_FLAT && col.entries >= 1 && (col.encoding == PG_ENC_RLE || col.ids != nullptr), "not an encoded column");
    PG_CHECK_ARG(nullsOut != nullptr || col.nulls == nullptr, "a dictionary with NULL entries needs a per-row null map");
    if (n <= 0) return;
    dictionaryIds.accounting = slots.accounting;
    dictionaryIds.reserve((size_t) col.entries * 4 + 16);
    DeviceColumn dictionary = col;
    dictionary.encoding = PG_ENC_FLAT; dictionary.ids = nullptr; dictionary.n = col.entries;
    intern(op, dictionary, col.entries, dictionaryIds.as<int32_t>(), insert);
    PG_LAUNCH(op, internMapDictionaryKernel, (int) std::min<int64_t>(divUp(n, 256), (int64_t) op->numSMs * 8), 256, 0, dictionaryIds.as<int32_t>(), col.nulls, col.ids, n,
              idsOut, nullsOut);
}

void VarcharInterner::gatherLengths(Operator* op, const int32_t* ids, const uint8_t* nulls, int64_t g, int32_t* lengthsOut)
{
    if (g <= 0) return;
    PG_LAUNCH(op, internLengthsKernel, (int) std::min<int64_t>(divUp(g, 256), 2048), 256, 0, view(), ids, nulls, g, lengthsOut);
}

void VarcharInterner::gatherBytes(Operator* op, const int32_t* ids, const uint8_t* nulls, const int32_t* offsets, int64_t g, uint8_t* bytesOut)
{
    if (g <= 0) return;
    PG_LAUNCH(op, internBytesKernel, (int) std::min<int64_t>(divUp(g * 32, 256), 4096), 256, 0, view(), ids, nulls, offsets, g, bytesOut);
}

}  // namespace pg
