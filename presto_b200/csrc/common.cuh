// presto_b200 — common host/device infrastructure for libprestogpu.so (sm_100a only).
//
// The device helpers at the bottom (and group_table.cuh, agg_tiers.cuh, dense_index.cuh, fused_agg.cuh on top of them) are also
// compiled at run time by NVRTC for fused pipeline kernels (jit_runtime.h): under __CUDACC_RTC__ the host half is skipped and the
// fixed-width integer types come from the prelude below.
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#define LLONG_MIN (-9223372036854775807LL - 1)
#define LLONG_MAX 9223372036854775807LL
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <memory>
#endif
#include "../../include/presto_gpu.h"

namespace pg {
#ifndef __CUDACC_RTC__

// ---- errors ---------------------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void setLastError(const std::string& msg);

#define PG_CUDA(expr)                                                                                             \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess) {                                                                                  \
            int _code = (_e == cudaErrorMemoryAllocation) ? PG_ERR_OUT_OF_MEMORY                                  \
                      : (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? PG_ERR_NO_DEVICE : PG_ERR_CUDA; \
            throw pg::Error(_code, std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
        }                                                                                                         \
    } while (0)

#define PG_CHECK_ARG(cond, msg) do { if (!(cond)) throw pg::Error(PG_ERR_INVALID_ARGUMENT, msg); } while (0)
#define PG_CHECK_STATE(cond, msg) do { if (!(cond)) throw pg::Error(PG_ERR_INVALID_STATE, msg); } while (0)

inline int typeWidth(int type)
{
    switch (type) {
        case PG_TYPE_BIGINT: case PG_TYPE_DOUBLE: return 8;
        case PG_TYPE_INTEGER: case PG_TYPE_DATE: return 4;
        case PG_TYPE_BOOLEAN: return 1;
        case PG_TYPE_VARCHAR: return 0;
        case PG_TYPE_LONG_DECIMAL: return 16;
    }
    throw Error(PG_ERR_UNSUPPORTED, "unknown pg_type " + std::to_string(type));
}

inline int64_t divUp(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- caching allocator ------------------------------------------------------------------------------------------------
// Operators come and go per query; cudaMalloc / cudaHostAlloc cost milliseconds, so freed blocks are kept per device in size
// classes (3 significant bits, <= 12.5% slack) and handed out again.  Callers only release a block after the stream that used it
// has been synchronised (every operator synchronises before it returns to the host).
void* cachedDeviceAlloc(size_t bytes, size_t* granted);
void cachedDeviceFree(void* p, size_t granted);
void* cachedPinnedAlloc(size_t bytes, size_t* granted);
void cachedPinnedFree(void* p, size_t granted);
void trimAllocationCaches();

// ---- device buffer (grow-only) -------------------------------------------------------------------------------------
struct DeviceBuffer {
    void* ptr = nullptr;
    size_t capacity = 0;
    int64_t* accounting = nullptr;   // operator's device_bytes counter
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    DeviceBuffer(DeviceBuffer&& o) noexcept { *this = std::move(o); }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept
    {
        if (this != &o) { release(); ptr = o.ptr; capacity = o.capacity; accounting = o.accounting; o.ptr = nullptr; o.capacity = 0; }
        return *this;
    }
    ~DeviceBuffer() { release(); }
    void release()
    {
        if (ptr) { cachedDeviceFree(ptr, capacity); if (accounting) *accounting -= (int64_t) capacity; }
        ptr = nullptr; capacity = 0;
    }
    // Ensure at least `bytes`; contents are NOT preserved when it grows.
    void reserve(size_t bytes)
    {
        if (bytes <= capacity) return;
        release();
        ptr = cachedDeviceAlloc(bytes, &capacity);
        if (accounting) *accounting += (int64_t) capacity;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

struct PinnedBuffer {
    void* ptr = nullptr;
    size_t capacity = 0;
    PinnedBuffer() = default;
    PinnedBuffer(const PinnedBuffer&) = delete;
    PinnedBuffer& operator=(const PinnedBuffer&) = delete;
    PinnedBuffer(PinnedBuffer&& o) noexcept { ptr = o.ptr; capacity = o.capacity; o.ptr = nullptr; o.capacity = 0; }
    ~PinnedBuffer() { if (ptr) cachedPinnedFree(ptr, capacity); }
    void reserve(size_t bytes)
    {
        if (bytes <= capacity) return;
        if (ptr) cachedPinnedFree(ptr, capacity);
        ptr = nullptr; capacity = 0;
        ptr = cachedPinnedAlloc(bytes, &capacity);
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// ---- a flat column resident in HBM ---------------------------------------------------------------------------------
struct DeviceColumn {
    int type = 0;
    int64_t n = 0;
    const void* values = nullptr;      // fixed width values, or varchar bytes
    const uint8_t* nulls = nullptr;    // byte per row or nullptr
    const int32_t* offsets = nullptr;  // varchar
    // Only for VARCHAR channels the operator asked the stager to keep encoded (PageStager::keepEncoded): values / nulls / offsets then
    // describe the `entries` dictionary entries (1 for RLE) and row r holds entry ids[r] (RLE: entry 0 for every row).
    int encoding = PG_ENC_FLAT;
    const int32_t* ids = nullptr;
    int64_t entries = 0;
};

// ---- Operator base: MB/operator/Operator.java:20-102 ----------------------------------------------------------------
struct Operator {
    pg_operator_options options{};
    pg_operator_stats stats{};
    cudaStream_t stream = nullptr;       // compute stream
    cudaStream_t copyStream = nullptr;   // H2D staging stream
    bool ownsStream = false;
    int numSMs = 148;

    explicit Operator(const pg_operator_options& o);
    virtual ~Operator();
    virtual bool needsInput() = 0;
    virtual void addInput(const pg_page* page) = 0;
    virtual bool getOutput(pg_page* out) = 0;     // false → Java null
    virtual void finish() = 0;
    virtual bool isFinished() = 0;
    virtual bool isBlocked() { return false; }    // Operator.isBlocked(): waiting on another Driver (a LookupSource under construction)
    virtual void syncStats() {}
    void bind() const { PG_CUDA(cudaSetDevice(options.device)); }
    int64_t maxOutputRows() const { return options.max_output_rows > 0 ? options.max_output_rows : (int64_t(1) << 20); }
};

// ---- page staging: host (or device, or dictionary/RLE) blocks → flat DeviceColumns ------------------------------------
// Double-buffered so the H2D copy of page i+1 overlaps the kernels of page i (copy stream / compute stream + events).
struct PageStager {
    static constexpr int kSlots = 2;
    Operator* op;
    struct Slot {
        std::vector<DeviceBuffer> values, nulls, offsets, scratch;
        DeviceBuffer arena;              // small host pages: every column of the page in one buffer, one copy
        PinnedBuffer arenaHost;
        cudaEvent_t copyDone = nullptr, computeDone = nullptr;
        bool computePending = false, arenaCopyPending = false;
    } slots[kSlots];
    int next = 0;
    // VARCHAR channels that stay DictionaryBlock / RunLengthEncodedBlock on the device instead of being flattened (DeviceColumn::encoding):
    // for consumers that work on the dictionary itself (interned group keys: MultiChannelGroupByHash's dictionary path :543-552, 628-714)
    std::vector<int> keepEncoded;

    explicit PageStager(Operator* o);
    ~PageStager();
    // Stage the listed channels of `page`.  Returns the slot used; cols[i] describes channel channels[i].
    // On return host memory of the page is no longer referenced; the compute stream has been made to wait for the copies.
    int stage(const pg_page* page, const std::vector<int>& channels, std::vector<DeviceColumn>& cols);
    // Host pages of FLAT blocks up to kSmallPageBytes (the reference's pages are ≤ 1 MB: PageProcessor MAX_PAGE_SIZE_IN_BYTES) are
    // gathered into one pinned arena and cross PCIe in ONE copy instead of two or three per column.
    static constexpr size_t kSmallPageBytes = 1 << 20;
    bool stageSmall(const pg_page* page, const std::vector<int>& channels, std::vector<DeviceColumn>& cols, Slot& slot);
    // Call after enqueueing the kernels that read slot `s` on op->stream.
    void markCompute(int s);
};

// kernel launch bookkeeping
#define PG_LAUNCH(op, kernel, grid, block, smem, ...)                                         \
    do {                                                                                      \
        kernel<<<(grid), (block), (smem), (op)->stream>>>(__VA_ARGS__);                       \
        (op)->stats.kernel_launches++;                                                        \
        PG_CUDA(cudaGetLastError());                                                          \
    } while (0)

#endif  // !__CUDACC_RTC__

// ---- device helpers --------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// AbstractLongType.hash  PC/type/AbstractLongType.java:137-141
__host__ __device__ __forceinline__ int64_t hashLong(int64_t value)
{
    uint64_t v = (uint64_t) value * 0xC2B2AE3D27D4EB4FULL;
    v = (v << 31) | (v >> 33);
    return (int64_t) (v * 0x9E3779B185EBCA87ULL);
}
// fastutil HashCommon.murmurHash3(long) == PagesHash.java:236-252
__host__ __device__ __forceinline__ uint64_t murmur64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
// HashGenerator.getPartition  MB/operator/HashGenerator.java:22-29
__host__ __device__ __forceinline__ int32_t hashPartition(int64_t rawHash, int32_t partitionCount)
{
    uint32_t folded = (uint32_t) ((uint64_t) rawHash ^ ((uint64_t) rawHash >> 32));
    return (int32_t) (((uint64_t) folded * (uint64_t) (uint32_t) partitionCount) >> 32);
}
// airlift XxHash64.hash(long) = XXH64 (seed 0) of the value's 8 little-endian bytes
__host__ __device__ __forceinline__ int64_t xxHash64Long(uint64_t value)
{
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    uint64_t h = P5 + 8;
    uint64_t k = value * P2; k = (k << 31) | (k >> 33); k *= P1;
    h ^= k; h = ((h << 27) | (h >> 37)) * P1 + P4;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return (int64_t) h;
}
__host__ __device__ __forceinline__ uint64_t reverseBits64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
#endif
}
// LocalPartitionGenerator.getPartition(rawHash)  MB/operator/exchange/LocalPartitionGenerator.java:43-63 (partitionCount: power of two)
__host__ __device__ __forceinline__ int32_t localPartition(int64_t rawHash, int32_t partitionCount)
{
    return (int32_t) xxHash64Long(reverseBits64((uint64_t) rawHash)) & (partitionCount - 1);
}
// Double.doubleToLongBits (canonical NaN)
__host__ __device__ __forceinline__ int64_t canonicalDoubleBits(int64_t bits)
{
    return ((bits & 0x7fffffffffffffffLL) > 0x7ff0000000000000LL) ? 0x7ff8000000000000LL : bits;
}
// TypeUtils.doubleHashCode (PC/type/TypeUtils.java:232-238; DoubleType.DOUBLE = new DoubleType(true), DoubleType.java:37,92-100):
// -0.0 hashes like +0.0, every NaN like the canonical NaN
__host__ __device__ __forceinline__ int64_t doubleHashBits(int64_t bits)
{
    const int64_t c = canonicalDoubleBits(bits);
    return c == (int64_t) 0x8000000000000000LL ? 0 : c;
}
// XXH64 (seed 0) of a short (< 32 byte) slice held in registers: `lo` = bytes 0..7 little endian, len <= 8
__host__ __device__ inline int64_t xxHash64Short(uint64_t lo, int len)
{
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    uint64_t h = P5 + (uint64_t) len;
    int pos = 0;
    if (len >= 8) {
        uint64_t k = lo * P2; k = (k << 31) | (k >> 33); k *= P1;
        h ^= k; h = ((h << 27) | (h >> 37)) * P1 + P4; pos = 8;
    }
    if (len - pos >= 4) {
        uint64_t k = (lo >> (8 * pos)) & 0xffffffffULL;
        h ^= k * P1; h = ((h << 23) | (h >> 41)) * P2 + P3; pos += 4;
    }
    while (pos < len) {
        uint64_t b = (lo >> (8 * pos)) & 0xff;
        h ^= b * P5; h = ((h << 11) | (h >> 53)) * P1; pos++;
    }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return (int64_t) h;
}

// XXH64 (seed 0) of an arbitrary byte slice = VarcharType.hash (AbstractVariableWidthBlock.java:102-106 → airlift XxHash64.hash(Slice)).
// Bytes are fetched one at a time (slices start at any alignment); the published XXH64 layout: 32-byte stripes into four lanes, then
// 8- / 4- / 1-byte tails, then the avalanche.
__host__ __device__ __forceinline__ uint64_t loadLe64(const uint8_t* p)
{
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v |= (uint64_t) p[i] << (8 * i);
    return v;
}
__host__ __device__ inline int64_t xxHash64Bytes(const uint8_t* p, int64_t len)
{
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    int64_t pos = 0;
    uint64_t h;
    if (len >= 32) {
        uint64_t v[4] = {P1 + P2, P2, 0, 0 - P1};
        for (; pos + 32 <= len; pos += 32) {
#pragma unroll
            for (int l = 0; l < 4; l++) {
                uint64_t a = v[l] + loadLe64(p + pos + 8 * l) * P2;
                v[l] = ((a << 31) | (a >> 33)) * P1;
            }
        }
        h = ((v[0] << 1) | (v[0] >> 63)) + ((v[1] << 7) | (v[1] >> 57)) + ((v[2] << 12) | (v[2] >> 52)) + ((v[3] << 18) | (v[3] >> 46));
#pragma unroll
        for (int l = 0; l < 4; l++) {
            uint64_t k = v[l] * P2; k = ((k << 31) | (k >> 33)) * P1;
            h = (h ^ k) * P1 + P4;
        }
    }
    else h = P5;
    h += (uint64_t) len;
    for (; pos + 8 <= len; pos += 8) {
        uint64_t k = loadLe64(p + pos) * P2; k = ((k << 31) | (k >> 33)) * P1;
        h ^= k; h = ((h << 27) | (h >> 37)) * P1 + P4;
    }
    if (pos + 4 <= len) {
        const uint64_t k = (uint64_t) p[pos] | ((uint64_t) p[pos + 1] << 8) | ((uint64_t) p[pos + 2] << 16) | ((uint64_t) p[pos + 3] << 24);
        h ^= k * P1; h = ((h << 23) | (h >> 41)) * P2 + P3; pos += 4;
    }
    for (; pos < len; pos++) { h ^= (uint64_t) p[pos] * P5; h = ((h << 11) | (h >> 53)) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return (int64_t) h;
}

// streaming 128-bit load (read-once data: do not allocate in L1)
__device__ __forceinline__ ulonglong2 ldStream128(const void* p)
{
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t ldStream64(const void* p)
{
    uint64_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldStream32(const void* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src)
{
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t) v, src), hi = __shfl_sync(0xffffffffu, (uint32_t) (v >> 32), src);
    return ((uint64_t) hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shflXor64(uint64_t v, int m)
{
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t) v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t) (v >> 32), m);
    return ((uint64_t) hi << 32) | lo;
}
#endif

}  // namespace pg
