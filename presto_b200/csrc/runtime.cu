// presto_b200 — runtime: device management, operator base, page staging, scalar hash helpers.
#include "common.cuh"
#include "crc32_host.h"
#include "lz4_codec.cuh"
#include <cub/device/device_scan.cuh>
#include <map>
#include <mutex>

namespace pg {

static thread_local std::string g_lastError;
void setLastError(const std::string& msg) { g_lastError = msg; }
const char* lastError() { return g_lastError.c_str(); }

// ---- caching allocator ----------------------------------------------------------------------------------------------------
namespace {
struct BlockCache {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void*>> freeBlocks;   // (device, size class) → blocks
    size_t cachedBytes = 0;
    size_t limitBytes;
    bool pinned;
    BlockCache(size_t limit, bool pin) : limitBytes(limit), pinned(pin) {}
    static size_t sizeClass(size_t bytes)
    {
        if (bytes < 512) return 512;
        int top = 63 - __builtin_clzll((unsigned long long) bytes);
        size_t step = size_t(1) << (top - 2);           // 3 significant bits
        return (bytes + step - 1) & ~(step - 1);
    }
    void* alloc(size_t bytes, size_t* granted)
    {
        int dev = 0;
        if (!pinned) PG_CUDA(cudaGetDevice(&dev));
        const size_t cls = sizeClass(bytes);
        {
            std::lock_guard<std::mutex> l(mu);
            auto it = freeBlocks.find({dev, cls});
            if (it != freeBlocks.end() && !it->second.empty()) {
                void* p = it->second.back(); it->second.pop_back(); cachedBytes -= cls; *granted = cls; return p;
            }
        }
        void* p = nullptr;
        cudaError_t e = pinned ? cudaHostAlloc(&p, cls, cudaHostAllocDefault) : cudaMalloc(&p, cls);
        if (e == cudaErrorMemoryAllocation) {   // give cached blocks back to the driver and retry once
            cudaGetLastError();
            trim();
            e = pinned ? cudaHostAlloc(&p, cls, cudaHostAllocDefault) : cudaMalloc(&p, cls);
        }
        PG_CUDA(e);
        *granted = cls;
        return p;
    }
    void free(void* p, size_t granted)
    {
        int dev = 0;
        if (!pinned) cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> l(mu);
            if (cachedBytes + granted <= limitBytes) { freeBlocks[{dev, granted}].push_back(p); cachedBytes += granted; return; }
        }
        if (pinned) cudaFreeHost(p); else cudaFree(p);
    }
    void trim()
    {
        std::lock_guard<std::mutex> l(mu);
        int cur = 0; cudaGetDevice(&cur);
        for (auto& kv : freeBlocks) {
            if (!pinned) cudaSetDevice(kv.first.first);
            for (void* p : kv.second) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
            kv.second.clear();
        }
        if (!pinned) cudaSetDevice(cur);
        cachedBytes = 0;
    }
};
BlockCache& deviceCache() { static BlockCache* c = new BlockCache(size_t(48) << 30, false); return *c; }
BlockCache& pinnedCache() { static BlockCache* c = new BlockCache(size_t(8) << 30, true); return *c; }
}  // namespace

void* cachedDeviceAlloc(size_t bytes, size_t* granted) { return deviceCache().alloc(bytes, granted); }
void cachedDeviceFree(void* p, size_t granted) { deviceCache().free(p, granted); }
void* cachedPinnedAlloc(size_t bytes, size_t* granted) { return pinnedCache().alloc(bytes, granted); }
void cachedPinnedFree(void* p, size_t granted) { pinnedCache().free(p, granted); }
void trimAllocationCaches() { deviceCache().trim(); pinnedCache().trim(); }

// ---- Operator base ----------------------------------------------------------------------------------------------------
Operator::Operator(const pg_operator_options& o) : options(o)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1)
        throw Error(PG_ERR_NO_DEVICE, "no CUDA device visible: presto_b200 operators only run on the GPU (no CPU fallback)");
    PG_CHECK_ARG(o.device >= 0 && o.device < count, "device ordinal out of range");
    PG_CUDA(cudaSetDevice(o.device));
    if (o.cuda_stream) { stream = (cudaStream_t) o.cuda_stream; ownsStream = false; }
    else { PG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)); ownsStream = true; }
    PG_CUDA(cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking));
    PG_CUDA(cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, o.device));
}

Operator::~Operator()
{
    cudaSetDevice(options.device);
    if (stream) cudaStreamSynchronize(stream);
    if (copyStream) { cudaStreamSynchronize(copyStream); cudaStreamDestroy(copyStream); }
    if (ownsStream && stream) cudaStreamDestroy(stream);
}

// ---- flatten kernels (DictionaryBlock / RunLengthEncodedBlock → flat) ---------------------------------------------------
template <typename T>
__global__ void flattenKernel(const T* __restrict__ dict, const uint8_t* __restrict__ dictNulls, const int32_t* __restrict__ ids,
                              int64_t n, T* __restrict__ out, uint8_t* __restrict__ outNulls)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        int32_t id = ids ? ids[i] : 0;
        out[i] = dict[id];
        if (outNulls) outNulls[i] = dictNulls[id];
    }
}

// VARCHAR dictionaries (DictionaryBlock over a VariableWidthBlock — how low-cardinality strings usually arrive): row lengths, then
// (after an inclusive scan of the lengths into the flat offsets) the bytes of every row
__global__ void flattenVarcharLengthsKernel(const int32_t* __restrict__ dictOffsets, const int32_t* __restrict__ ids, int64_t n, int32_t* __restrict__ outOffsets)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        const int32_t id = ids ? ids[i] : 0;
        outOffsets[i + 1] = dictOffsets[id + 1] - dictOffsets[id];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) outOffsets[0] = 0;
}
__global__ void flattenVarcharBytesKernel(const uint8_t* __restrict__ dictBytes, const int32_t* __restrict__ dictOffsets, const uint8_t* __restrict__ dictNulls,
                                          const int32_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ outOffsets, uint8_t* __restrict__ outBytes,
                                          uint8_t* __restrict__ outNulls)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        const int32_t id = ids ? ids[i] : 0;
        const int32_t src = dictOffsets[id], len = dictOffsets[id + 1] - src, dst = outOffsets[i];
        for (int32_t k = 0; k < len; k++) outBytes[dst + k] = dictBytes[src + k];
        if (outNulls) outNulls[i] = dictNulls[id];
    }
}

PageStager::PageStager(Operator* o) : op(o)
{
    for (auto& s : slots) {
        PG_CUDA(cudaEventCreateWithFlags(&s.copyDone, cudaEventDisableTiming));
        PG_CUDA(cudaEventCreateWithFlags(&s.computeDone, cudaEventDisableTiming));
    }
}

PageStager::~PageStager()
{
    for (auto& s : slots) {
        if (s.copyDone) cudaEventDestroy(s.copyDone);
        if (s.computeDone) cudaEventDestroy(s.computeDone);
    }
}

void PageStager::markCompute(int s)
{
    PG_CUDA(cudaEventRecord(slots[s].computeDone, op->stream));
    slots[s].computePending = true;
}

bool PageStager::stageSmall(const pg_page* page, const std::vector<int>& channels, std::vector<DeviceColumn>& cols, Slot& slot)
{
    const size_t nch = channels.size();
    if (page->memory != PG_MEM_HOST || page->position_count <= 0 || nch == 0) return false;
    struct Piece { const void* src; size_t bytes, at; };
    std::vector<Piece> pieces;
    pieces.reserve(nch * 3);
    size_t total = 0;
    auto add = [&](const void* src, size_t bytes) { pieces.push_back({src, bytes, total}); total += (bytes + 15) & ~size_t(15); return pieces.back().at; };
    struct Where { size_t values = 0, nulls = 0, offsets = 0, first = 0; bool hasNulls = false; };
    std::vector<Where> where(nch);
    const int64_t n = page->position_count;
    for (size_t i = 0; i < nch; i++) {
        const int ch = channels[i];
        if (ch < 0 || ch >= page->channel_count) return false;       // the general path reports it
        const pg_block& b = page->blocks[ch];
        if (b.encoding != PG_ENC_FLAT || b.position_count != n) return false;
        Where& w = where[i];
        if (b.type == PG_TYPE_VARCHAR) {
            if (!b.offsets) return false;
            w.first = (size_t) b.offsets[0];
            w.values = add((const char*) b.values + w.first, (size_t) b.offsets[n] - w.first);
            w.offsets = add(b.offsets, (size_t) (n + 1) * 4);
        }
        else w.values = add(b.values, (size_t) n * typeWidth(b.type));
        if (b.is_null) { w.hasNulls = true; w.nulls = add(b.is_null, (size_t) n); }
        if (total > kSmallPageBytes) return false;
    }
    // the arena of this slot may still be travelling (the previous page staged through it did not wait for its copy)
    if (slot.arenaCopyPending) { PG_CUDA(cudaEventSynchronize(slot.copyDone)); slot.arenaCopyPending = false; }
    if (slot.computePending) { PG_CUDA(cudaStreamWaitEvent(op->copyStream, slot.computeDone, 0)); slot.computePending = false; }
    slot.arena.accounting = &op->stats.device_bytes;
    slot.arena.reserve(total + 16);
    slot.arenaHost.reserve(total + 16);
    for (const Piece& p : pieces) if (p.bytes) memcpy((char*) slot.arenaHost.ptr + p.at, p.src, p.bytes);
    PG_CUDA(cudaMemcpyAsync(slot.arena.ptr, slot.arenaHost.ptr, total, cudaMemcpyHostToDevice, op->copyStream));
    PG_CUDA(cudaEventRecord(slot.copyDone, op->copyStream));
    PG_CUDA(cudaStreamWaitEvent(op->stream, slot.copyDone, 0));
    slot.arenaCopyPending = true;            // the page's own memory is no longer referenced: no wait here
    cols.assign(nch, DeviceColumn());
    const char* base = (const char*) slot.arena.ptr;
    for (size_t i = 0; i < nch; i++) {
        const pg_block& b = page->blocks[channels[i]];
        const Where& w = where[i];
        DeviceColumn& c = cols[i];
        c.type = b.type; c.n = n;
        c.values = base + w.values - w.first;
        c.offsets = b.type == PG_TYPE_VARCHAR ? (const int32_t*) (base + w.offsets) : nullptr;
        c.nulls = w.hasNulls ? (const uint8_t*) (base + w.nulls) : nullptr;
        const int width = typeWidth(b.type);
        op->stats.input_bytes += n * (width ? width : 4) + (w.hasNulls ? n : 0);
    }
    op->stats.h2d_bytes += (int64_t) total;
    return true;
}

int PageStager::stage(const pg_page* page, const std::vector<int>& channels, std::vector<DeviceColumn>& cols)
{
    int s = next;
    next = (next + 1) % kSlots;
    Slot& slot = slots[s];
    if (stageSmall(page, channels, cols, slot)) return s;
    size_t nch = channels.size();
    if (slot.values.size() < nch) {
        slot.values.resize(nch); slot.nulls.resize(nch); slot.offsets.resize(nch); slot.scratch.resize(nch);
        for (size_t i = 0; i < nch; i++)
            for (DeviceBuffer* b : {&slot.values[i], &slot.nulls[i], &slot.offsets[i], &slot.scratch[i]}) b->accounting = &op->stats.device_bytes;
    }
    cols.assign(nch, DeviceColumn());
    bool host = page->memory == PG_MEM_HOST;
    // copies into this slot must not overtake kernels still reading it
    cudaStream_t cs = host ? op->copyStream : op->stream;
    if (host && slot.computePending) { PG_CUDA(cudaStreamWaitEvent(cs, slot.computeDone, 0)); slot.computePending = false; }
    bool copied = false;
    for (size_t i = 0; i < nch; i++) {
        int ch = channels[i];
        PG_CHECK_ARG(ch >= 0 && ch < page->channel_count, "channel out of range for the input page");
        const pg_block& b = page->blocks[ch];
        PG_CHECK_ARG(b.position_count == page->position_count, "block position count differs from the page's");
        DeviceColumn& c = cols[i];
        c.type = b.type; c.n = b.position_count;
        int w = typeWidth(b.type);
        bool encoded = b.encoding != PG_ENC_FLAT;
        PG_CHECK_ARG(b.encoding == PG_ENC_FLAT || b.encoding == PG_ENC_DICTIONARY || b.encoding == PG_ENC_RLE, "unknown block encoding");
        if (encoded && b.type == PG_TYPE_LONG_DECIMAL) throw Error(PG_ERR_UNSUPPORTED, "dictionary / RLE encoded Int128 blocks are not on the GPU path");
        int64_t entries = encoded ? (b.encoding == PG_ENC_RLE ? 1 : b.dictionary_size) : b.position_count;
        const void* dValues = b.values; const uint8_t* dNulls = b.is_null; const int32_t* dOffsets = b.offsets; const int32_t* dIds = b.ids;
        if (host && c.n > 0) {
            if (b.type == PG_TYPE_VARCHAR) {
                PG_CHECK_ARG(b.offsets != nullptr, "VARCHAR block without offsets");
                size_t obytes = (size_t) (entries + 1) * 4;
                // only the bytes this block's offsets cover travel (a sub-range of a larger block shares the byte slice: offsets stay
                // absolute, so the device base pointer is moved back by the first offset instead of rebasing every offset)
                const size_t first = encoded ? 0 : (size_t) b.offsets[0];
                size_t vbytes = (size_t) b.offsets[entries] - first;
                // encoded: the dictionary goes to the scratch buffer (bytes, then offsets, then ids); the flat column is built below
                DeviceBuffer& vb = encoded ? slot.scratch[i] : slot.values[i];
                const size_t offAt = (vbytes + 15) & ~size_t(15), idsAt = (offAt + obytes + 15) & ~size_t(15);
                vb.reserve(encoded ? idsAt + (size_t) c.n * 4 + 16 : vbytes + 16);
                if (vbytes) PG_CUDA(cudaMemcpyAsync(vb.ptr, (const char*) b.values + first, vbytes, cudaMemcpyHostToDevice, cs));
                op->stats.h2d_bytes += (int64_t) (obytes + vbytes);
                dValues = (const char*) vb.ptr - first;
                if (!encoded) {
                    slot.offsets[i].reserve(obytes);
                    PG_CUDA(cudaMemcpyAsync(slot.offsets[i].ptr, b.offsets, obytes, cudaMemcpyHostToDevice, cs));
                    dOffsets = slot.offsets[i].as<int32_t>();
                }
                else {
                    PG_CUDA(cudaMemcpyAsync((char*) vb.ptr + offAt, b.offsets, obytes, cudaMemcpyHostToDevice, cs));
                    dOffsets = (const int32_t*) ((char*) vb.ptr + offAt);
                    if (b.encoding == PG_ENC_DICTIONARY) {
                        PG_CUDA(cudaMemcpyAsync((char*) vb.ptr + idsAt, b.ids, (size_t) c.n * 4, cudaMemcpyHostToDevice, cs));
                        op->stats.h2d_bytes += c.n * 4;
                        dIds = (const int32_t*) ((char*) vb.ptr + idsAt);
                    }
                }
            }
            else {
                size_t vbytes = (size_t) entries * w;
                DeviceBuffer& vb = encoded ? slot.scratch[i] : slot.values[i];
                vb.reserve(vbytes + (encoded ? (size_t) c.n * 4 + 16 : 0));
                PG_CUDA(cudaMemcpyAsync(vb.ptr, b.values, vbytes, cudaMemcpyHostToDevice, cs));
                op->stats.h2d_bytes += (int64_t) vbytes;
                dValues = vb.ptr;
                if (b.encoding == PG_ENC_DICTIONARY) {
                    size_t off = (vbytes + 15) & ~size_t(15);
                    PG_CUDA(cudaMemcpyAsync((char*) vb.ptr + off, b.ids, (size_t) c.n * 4, cudaMemcpyHostToDevice, cs));
                    op->stats.h2d_bytes += c.n * 4;
                    dIds = (const int32_t*) ((char*) vb.ptr + off);
                }
            }
            if (b.is_null) {
                DeviceBuffer& nb = slot.nulls[i];
                // encoded: dictionary nulls first, flat nulls after
                nb.reserve((size_t) entries + (encoded ? (size_t) c.n + 16 : 0));
                PG_CUDA(cudaMemcpyAsync(nb.ptr, b.is_null, (size_t) entries, cudaMemcpyHostToDevice, cs));
                op->stats.h2d_bytes += entries;
                dNulls = nb.as<uint8_t>();
            }
            copied = true;
        }
        if (encoded && c.n > 0 && b.type == PG_TYPE_VARCHAR && std::find(keepEncoded.begin(), keepEncoded.end(), ch) != keepEncoded.end()) {
            PG_CHECK_ARG(dOffsets != nullptr, "VARCHAR block without offsets");
            c.encoding = b.encoding; c.entries = entries; c.ids = b.encoding == PG_ENC_DICTIONARY ? dIds : nullptr;
            c.values = dValues; c.nulls = dNulls; c.offsets = dOffsets;
            op->stats.input_bytes += c.n * 4 + (b.is_null ? c.n : 0);
            continue;
        }
        if (encoded && c.n > 0 && b.type == PG_TYPE_VARCHAR) {
            // lengths → inclusive scan → offsets; then the bytes (and nulls) row by row
            PG_CHECK_ARG(dOffsets != nullptr, "VARCHAR block without offsets");
            const int32_t* ids = b.encoding == PG_ENC_DICTIONARY ? dIds : nullptr;
            slot.offsets[i].reserve((size_t) (c.n + 1) * 4);
            int32_t* outOff = slot.offsets[i].as<int32_t>();
            const int grid = (int) std::min<int64_t>(divUp(c.n, 256), 4096);
            flattenVarcharLengthsKernel<<<grid, 256, 0, cs>>>(dOffsets, ids, c.n, outOff);
            size_t tb = 0;
            cub::DeviceScan::InclusiveSum(nullptr, tb, outOff + 1, outOff + 1, (int) c.n, cs);
            DeviceBuffer scanTmp; scanTmp.reserve(tb);
            cub::DeviceScan::InclusiveSum(scanTmp.ptr, tb, outOff + 1, outOff + 1, (int) c.n, cs);
            int32_t totalBytes = 0;
            PG_CUDA(cudaMemcpyAsync(&totalBytes, outOff + c.n, 4, cudaMemcpyDeviceToHost, cs));
            PG_CUDA(cudaStreamSynchronize(cs));
            slot.values[i].reserve((size_t) totalBytes + 16);
            uint8_t* outNulls = nullptr;
            if (dNulls) {
                if (host) outNulls = slot.nulls[i].as<uint8_t>() + (((size_t) entries + 15) & ~size_t(15));
                else { slot.nulls[i].reserve((size_t) c.n); outNulls = slot.nulls[i].as<uint8_t>(); }
            }
            flattenVarcharBytesKernel<<<grid, 256, 0, cs>>>((const uint8_t*) dValues, dOffsets, dNulls, ids, c.n, outOff, slot.values[i].as<uint8_t>(), outNulls);
            op->stats.kernel_launches += 4;
            PG_CUDA(cudaGetLastError());
            PG_CUDA(cudaStreamSynchronize(cs));      // scanTmp dies here
            dValues = slot.values[i].ptr; dNulls = outNulls; dOffsets = outOff;
        }
        else if (encoded && c.n > 0) {
            // flatten on the stream the source data lands on
            slot.values[i].reserve((size_t) c.n * w);
            uint8_t* outNulls = nullptr;
            if (dNulls) {
                if (host) outNulls = slot.nulls[i].as<uint8_t>() + (((size_t) entries + 15) & ~size_t(15));
                else { slot.nulls[i].reserve((size_t) c.n); outNulls = slot.nulls[i].as<uint8_t>(); }
            }
            int grid = (int) std::min<int64_t>(divUp(c.n, 256), 4096);
            const int32_t* ids = b.encoding == PG_ENC_DICTIONARY ? dIds : nullptr;
            if (w == 8) flattenKernel<uint64_t><<<grid, 256, 0, cs>>>((const uint64_t*) dValues, dNulls, ids, c.n, slot.values[i].as<uint64_t>(), outNulls);
            else if (w == 4) flattenKernel<uint32_t><<<grid, 256, 0, cs>>>((const uint32_t*) dValues, dNulls, ids, c.n, slot.values[i].as<uint32_t>(), outNulls);
            else flattenKernel<uint8_t><<<grid, 256, 0, cs>>>((const uint8_t*) dValues, dNulls, ids, c.n, slot.values[i].as<uint8_t>(), outNulls);
            op->stats.kernel_launches++;
            PG_CUDA(cudaGetLastError());
            dValues = slot.values[i].ptr; dNulls = outNulls;
        }
        c.values = dValues; c.nulls = dNulls; c.offsets = dOffsets;
        op->stats.input_bytes += c.n * (w ? w : 4) + (b.is_null ? c.n : 0);
    }
    if (host && copied) {
        PG_CUDA(cudaEventRecord(slot.copyDone, cs));
        PG_CUDA(cudaStreamWaitEvent(op->stream, slot.copyDone, 0));
        // the page's host memory is only borrowed for the duration of addInput
        PG_CUDA(cudaEventSynchronize(slot.copyDone));
    }
    return s;
}

}  // namespace pg

// ======================================================================================================================
// C ABI: library / device / scalar helpers
// ======================================================================================================================
using namespace pg;

#define PG_TRY try {
#define PG_CATCH                                                                                          \
    }                                                                                                     \
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }                             \
    catch (const std::bad_alloc&) { pg::setLastError("host out of memory"); return PG_ERR_OUT_OF_MEMORY; } \
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }          \
    return PG_OK;

extern "C" {

const char* pg_version(void) { return "presto_b200 0.1.0 sm_100a"; }
const char* pg_last_error(void) { return pg::lastError(); }

int32_t pg_device_count(int32_t* count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        *count = 0;
        pg::setLastError(std::string("no CUDA device visible: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return PG_ERR_NO_DEVICE;
    }
    *count = n;
    return PG_OK;
}
int32_t pg_set_device(int32_t device) { PG_TRY PG_CUDA(cudaSetDevice(device)); PG_CATCH }
int32_t pg_host_alloc(size_t bytes, void** out) { PG_TRY PG_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault)); PG_CATCH }
int32_t pg_host_free(void* p) { PG_TRY PG_CUDA(cudaFreeHost(p)); PG_CATCH }
int32_t pg_device_alloc(size_t bytes, void** out) { PG_TRY PG_CUDA(cudaMalloc(out, bytes)); PG_CATCH }
int32_t pg_device_free(void* p) { PG_TRY PG_CUDA(cudaFree(p)); PG_CATCH }
int32_t pg_memcpy_h2d(void* dst, const void* src, size_t bytes) { PG_TRY PG_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); PG_CATCH }
int32_t pg_memcpy_d2h(void* dst, const void* src, size_t bytes) { PG_TRY PG_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); PG_CATCH }
int32_t pg_device_synchronize(void) { PG_TRY PG_CUDA(cudaDeviceSynchronize()); PG_CATCH }
int64_t pg_hash_crc32_chunked(const void* data, size_t len, size_t chunk) { return (int64_t) (uint64_t) pg::crc32ChunkedHost((const uint8_t*) data, len, chunk ? chunk : 4096); }
int32_t pg_release_cached_memory(void) { PG_TRY PG_CUDA(cudaDeviceSynchronize()); pg::trimAllocationCaches(); PG_CATCH }
int32_t pg_ipc_export(const void* devicePtr, uint8_t* handle)
{
    PG_TRY
    static_assert(sizeof(cudaIpcMemHandle_t) == PG_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    PG_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(devicePtr)));
    memcpy(handle, &h, sizeof(h));
    PG_CATCH
}
int32_t pg_ipc_import(const uint8_t* handle, void** out)
{
    PG_TRY
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    PG_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    PG_CATCH
}
int32_t pg_ipc_close(void* imported) { PG_TRY PG_CUDA(cudaIpcCloseMemHandle(imported)); PG_CATCH }

int64_t pg_hash_bigint(int64_t value) { return pg::hashLong(value); }
int64_t pg_hash_combine(int64_t previous, int64_t h) { return (int64_t) (31ULL * (uint64_t) previous + (uint64_t) h); }
int64_t pg_hash_murmur3_fmix64(int64_t x) { return (int64_t) pg::murmur64((uint64_t) x); }
int32_t pg_hash_partition(int64_t rawHash, int32_t partitionCount) { return pg::hashPartition(rawHash, partitionCount); }
int64_t pg_hash_double(double value) { int64_t b; memcpy(&b, &value, 8); return pg::hashLong(pg::doubleHashBits(b)); }
int32_t pg_hash_local_partition(int64_t rawHash, int32_t partitionCount)
{
    if (partitionCount <= 0 || (partitionCount & (partitionCount - 1)) != 0) return -1;   // checkArgument(Integer.bitCount(partitionCount) == 1)
    return pg::localPartition(rawHash, partitionCount);
}

int64_t pg_lz4_compress_bound(int64_t n) { return pg::lz4::compressBound(n); }
int64_t pg_lz4_compress_host(const void* src, int64_t n, void* dst, int64_t capacity)
{
    using namespace pg::lz4;
    if (n < 0 || capacity < compressBound(n) || n >= (int64_t(1) << 31)) return -1;
    const int64_t nSegs = segmentsOf(n);
    std::vector<SegmentInfo> segs((size_t) nSegs);
    std::vector<Match> matches((size_t) nSegs * kMaxMatchesPerSegment);
    std::vector<uint16_t> table(kHashEntries);
    for (int64_t s = 0; s < nSegs; s++) findMatches((const uint8_t*) src, n, s, table.data(), matches.data() + s * kMaxMatchesPerSegment, segs[(size_t) s]);
    BlockInfo block;
    planBlock(n, segs.data(), nSegs, block);
    for (int64_t s = 0; s < nSegs; s++) emitSegment((const uint8_t*) src, n, s, nSegs, matches.data() + s * kMaxMatchesPerSegment, segs[(size_t) s], block, (uint8_t*) dst);
    return block.compressedBytes;
}
int64_t pg_lz4_decompress_host(const void* src, int64_t n, void* dst, int64_t capacity)
{
    if (n < 0 || capacity < 0) return -1;
    return pg::lz4::decodeBlock<false>((const uint8_t*) src, n, (uint8_t*) dst, capacity, 0, 1);
}

int64_t pg_hash_xxhash64(const void* data, size_t len)
{
    // XXH64 seed 0 (airlift XxHash64.hash(Slice)): the very function the kernels hash VARCHAR slices with (common.cuh), so the CPU
    // tests that cross-check this entry point against the xxhash package pin the device code too
    return pg::xxHash64Bytes((const uint8_t*) data, (int64_t) len);
}

// ---- operator protocol ------------------------------------------------------------------------------------------------
int32_t pg_operator_needs_input(pg_operator* op, int32_t* out) { PG_TRY auto* o = (Operator*) op; o->bind(); *out = o->needsInput() ? 1 : 0; PG_CATCH }
int32_t pg_operator_is_blocked(pg_operator* op, int32_t* out) { PG_TRY auto* o = (Operator*) op; *out = o->isBlocked() ? 1 : 0; PG_CATCH }
int32_t pg_operator_add_input(pg_operator* op, const pg_page* page)
{
    PG_TRY
    auto* o = (Operator*) op;
    o->bind();
    PG_CHECK_ARG(page != nullptr && page->position_count >= 0, "bad page");
    o->addInput(page);
    PG_CATCH
}
int32_t pg_operator_get_output(pg_operator* op, pg_page* out)
{
    PG_TRY
    auto* o = (Operator*) op;
    o->bind();
    if (!o->getOutput(out)) { out->position_count = 0; out->channel_count = 0; out->memory = o->options.output_memory; out->blocks = nullptr; }
    PG_CATCH
}
int32_t pg_operator_finish(pg_operator* op) { PG_TRY auto* o = (Operator*) op; o->bind(); o->finish(); PG_CATCH }
int32_t pg_operator_is_finished(pg_operator* op, int32_t* out) { PG_TRY auto* o = (Operator*) op; o->bind(); *out = o->isFinished() ? 1 : 0; PG_CATCH }
int32_t pg_operator_get_stats(pg_operator* op, pg_operator_stats* out)
{
    PG_TRY
    auto* o = (Operator*) op;
    o->bind();
    o->syncStats();
    *out = o->stats;
    PG_CATCH
}
int32_t pg_operator_destroy(pg_operator* op)
{
    PG_TRY
    auto* o = (Operator*) op;
    if (o) { cudaSetDevice(o->options.device); if (o->stream) cudaStreamSynchronize(o->stream); if (o->copyStream) cudaStreamSynchronize(o->copyStream); delete o; }
    PG_CATCH
}

}  // extern "C"
