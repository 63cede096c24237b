// presto_b200 — the on-chip tiers of the hash aggregation (shared by the precompiled kernels in hash_agg.cu and the kernels
// NVRTC compiles for fused pipelines, fused_agg.cuh): shared-memory carve-up, the CTA's shared hash table, shared-space reductions.
#pragma once
#include "group_table.cuh"

namespace pg {

constexpr int kAggThreads = 256;
constexpr int kSharedProbes = 8;
constexpr int kWarpAggMaxGroups = 8;
constexpr int kWarpMaxGroups = 40;       // up to here: per-warp accumulator sets (47 / 41 G rows/s at 16 / 32 groups vs 39 / 34 on tiers 0/1; beyond, the
                                         // 512-slot table of this geometry fills up and the 4096-slot table per SM wins: 44+ G rows/s)
constexpr int kMidMinGroups = 40;        // from here on the CTA-per-SM shared table beats tier 0 + a small tier-1 table (see launchFastT)
constexpr int kMaxSubsetPasses = 1;      // mid mode (47 G rows/s at 1000 groups) loses to direct mode (~28 G rows/s flat) from the second hash-subset pass on

template <int MAXV> struct RowsPerThread { static constexpr int value = MAXV <= 4 ? 4 : 2; };   // always even: rows are loaded in pairs
// Tier 0: the first kTier0Groups groups a CTA claims get per-thread private accumulators in shared memory.
constexpr int kTier0Columns = 40;                  // per-thread private accumulator columns of one CTA (Q1: 6 groups x 6 columns = 36)
template <int MAXV> struct Tier0Groups { static constexpr int value = kTier0Columns / MAXV; };
constexpr int kTier0Bytes = kTier0Columns * kAggThreads * 8;   // K0 * MAXV <= 40 columns of 256 doubles = 80 KB

// shared-memory carve-up, same arithmetic on host and device
struct SharedLayout {
    uint32_t keys, acc, first, state, sid, misc, tier0, total;
    __host__ __device__ SharedLayout(uint32_t S, int W, int nAccAlloc, bool withTier0) { plan(S, W, nAccAlloc, withTier0 ? (uint32_t) kTier0Bytes : 0u); }
    // explicit tier-0 size (fused kernels size it for the groups the page actually has, to leave room for a TMA staging ring)
    __host__ __device__ SharedLayout(uint32_t S, int W, int nAccAlloc, uint32_t tier0Bytes, int) { plan(S, W, nAccAlloc, tier0Bytes); }
    __host__ __device__ void plan(uint32_t S, int W, int nAccAlloc, uint32_t tier0Bytes)
    {
        const uint32_t S2 = S + 2;
        uint32_t o = 0;
        keys = o; o += 8 * W * S2;
        acc = o; o += 8 * nAccAlloc * S2;
        first = o; o += 4 * S2;
        state = o; o += (W > 1 ? 4 * S2 : 0);
        sid = o; o += (S2 + 15) & ~15u;
        misc = o; o += 1280;  // claimed count, next tier-0 / warp-table id, slot of each id (<= 254)
        o = (o + 15) & ~15u;
        tier0 = o; o += tier0Bytes;
        total = o;
    }
};

// ---- shared-memory table ---------------------------------------------------------------------------------------------
template <int W>
struct SharedTable {
    volatile uint64_t* keys;    // [W][S2]
    volatile uint32_t* state;   // [S2] (W > 1)
    uint64_t* acc;              // [nAcc][S2]
    uint32_t* first;            // [S2]
    volatile uint8_t* sid;      // [S2] tier-0 id of the slot's group, 0xFF = none
    uint32_t* count;            // misc[0] claimed slots, misc[1] next tier-0 id, misc[2..] slot of tier-0 id
    uint32_t S, S2, sLimit;
    int tier0Groups;

    __device__ __forceinline__ void claimed(uint32_t s) const
    {
        atomicAdd(count, 1u);
        if (tier0Groups) {
            uint32_t id = atomicAdd(count + 1, 1u);
            // sid[s] is published last and read without a barrier by other threads (compute-sanitizer racecheck flags it): a reader that
            // still sees 0xFF sends its row to tier 1 (st.acc of the same slot), one that sees the id uses the private / per-warp
            // accumulators — both are folded into the slot before the flush, so either outcome is correct
            if (id < (uint32_t) tier0Groups) { count[2 + id] = s; __threadfence_block(); sid[s] = (uint8_t) id; }
        }
    }

    __device__ __forceinline__ int find(const uint64_t (&kw)[W], uint64_t h) const
    {
        uint32_t s = (uint32_t) (h >> 32) & (S - 1);
        if (W == 1) {
#pragma unroll 1
            for (int p = 0; p < kSharedProbes; p++) {
                uint64_t k = keys[s];
                if (k == kw[0]) return (int) s;
                if (k == kEmptyKey) {
                    if (*(volatile uint32_t*) count >= sLimit) return -1;
                    uint64_t old = atomicCAS((unsigned long long*) &keys[s], (unsigned long long) kEmptyKey, (unsigned long long) kw[0]);
                    if (old == kEmptyKey) { claimed(s); return (int) s; }
                    if (old == kw[0]) return (int) s;
                }
                s = (s + 1) & (S - 1);
            }
            return -1;
        }
        else {
            const uint32_t ready = tagOf(h);
#pragma unroll 1
            for (int p = 0; p < kSharedProbes; p++) {
                uint32_t st = state[s];
                if (st == kStEmpty) {
                    if (*(volatile uint32_t*) count >= sLimit) return -1;
                    uint32_t old = atomicCAS((uint32_t*) &state[s], kStEmpty, kStLocked);
                    if (old == kStEmpty) {
#pragma unroll
                        for (int w = 0; w < W; w++) keys[w * S2 + s] = kw[w];
                        __threadfence_block();
                        state[s] = ready;
                        claimed(s);
                        return (int) s;
                    }
                    st = old;
                }
                while (st == kStLocked) st = state[s];
                if (st == ready) {
                    bool eq = true;
#pragma unroll
                    for (int w = 0; w < W; w++) eq &= keys[w * S2 + s] == kw[w];
                    if (eq) return (int) s;
                }
                s = (s + 1) & (S - 1);
            }
            return -1;
        }
    }
};

template <int NV, bool CNT> struct FastGeom {
    static constexpr int NC = NV + (CNT ? 1 : 0);       // accumulator columns
    static constexpr int K0 = (kTier0Columns / NC) > 31 ? 31 : (kTier0Columns / NC);   // tier-0 groups (K0 * NC <= 40 columns of 256 x 8 B = 80 KB; ids live in a 32-bit mask)
};

// Rows that are not tier 0: the CTA's shared table (direct shared atomics) or the HBM table.  Kept out of line so the hot loop stays small.
__device__ __forceinline__ void sharedAddF64(uint64_t* addr, uint64_t bits)
{
    const unsigned a = (unsigned) __cvta_generic_to_shared(addr);
    asm volatile("red.shared.add.f64 [%0], %1;" :: "r"(a), "d"(__longlong_as_double((long long) bits)) : "memory");
}
__device__ __forceinline__ void sharedAddU64(uint64_t* addr, uint64_t v)
{
    const unsigned a = (unsigned) __cvta_generic_to_shared(addr);
    asm volatile("red.shared.add.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void sharedMinU32(uint32_t* addr, uint32_t v)
{
    const unsigned a = (unsigned) __cvta_generic_to_shared(addr);
    asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}


}  // namespace pg
