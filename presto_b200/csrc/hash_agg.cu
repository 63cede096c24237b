// presto_b200 — HashAggregationOperator on sm_100a.
//
// Replaces (MB = presto-main-base/src/main/java/com/facebook/presto/):
//   MB/operator/HashAggregationOperator.java:403-501          operator shell
//   MB/operator/aggregation/builder/InMemoryHashAggregationBuilder.java:204-220  processPage = getGroupIds + one pass per aggregate
//   MB/operator/BigintGroupByHash.java:222-269 / MultiChannelGroupByHash.java:293-358  putIfAbsent / addNewGroup
//   MB/operator/aggregation/AccumulatorCompiler.java:462-566  grouped accumulator input loops
// with ONE fused pass: each row's packed key is found-or-claimed in a per-CTA shared-memory table (warp-aggregated with
// match/shfl so hot groups cost one shared atomic per warp, not per row) that falls through to the HBM table; accumulators
// are indexed by slot, never by group id, so group ids never have to be materialised per row.  Group ids = rank of each
// slot's first-touch row, assigned at output time (InMemoryHashAggregationBuilder.buildResult order :326-352,400-403).
//
// Algorithmic HBM bytes per row = key bytes + 8 B per aggregated column; everything else is on-chip.
//
// Where a row's accumulators live depends on how many groups the page has (learned from a probe chunk of the first page):
//   <= 10 groups      tier 0: per-thread private accumulators in shared memory, no atomics            (hashAggFastKernel)
//   11 .. 40          warp mode: per-warp accumulator sets, same-group lanes resolved in rounds         (hashAggFastKernel, warpGroups > 0)
//   .. ~3000          one 4096-slot shared table per SM, shared-memory reductions                       (hashAggFastKernel<.., MID = true>)
//   more              direct mode: native f64 REDs on the HBM table, one L2-resident slot range per pass (hashAggDirectKernel)
// Keys that are not one BIGINT / two 32-bit columns, NULLs, masks, bigint sums: hashAggKernel<W, MAXV> (tiers 0 / 1 / 2).
#include "common.cuh"
#include "group_table.cuh"
#include "varchar_interner.h"
#include "expr_literals.h"
#include "key_layout.h"
#include "agg_tiers.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <functional>
#include <map>
#include "fused_agg.cuh"
#include "jit_runtime.h"
#include "operators_internal.h"

namespace pg {

// accumulator column kinds
enum { K_F64 = 0, K_F64_FROM_I64 = 1, K_LO32 = 2, K_HI32 = 3, K_COUNT = 4, K_I64_ADD = 5, K_MIN_F64 = 6, K_MAX_F64 = 7, K_MIN_I64 = 8, K_MAX_I64 = 9,
       // sum(decimal): the 128-bit value as four 32-bit limbs, each summed in its own 64-bit accumulator (exact for 2^31 rows per group;
       // limb 3 is signed).  K_DEC_*: a PG_TYPE_LONG_DECIMAL input (sign-magnitude, 16 B); K_DST_*: a serialized state {overflow, low,
       // high} (24 B inside a VARBINARY block), i.e. overflow * 2^127 + value.  Short decimals are plain longs: K_LO32 / K_HI32.
       K_DEC_L0 = 10, K_DEC_L1 = 11, K_DEC_L2 = 12, K_DEC_L3 = 13, K_DST_L0 = 14, K_DST_L1 = 15, K_DST_L2 = 16, K_DST_L3 = 17 };
__host__ __device__ inline bool kindIsDecimal(int k) { return k >= K_DEC_L0 && k <= K_DEC_L3; }
__host__ __device__ inline bool kindIsDecimalState(int k) { return k >= K_DST_L0 && k <= K_DST_L3; }
__host__ __device__ inline bool kindIsFloat(int k) { return k == K_F64 || k == K_F64_FROM_I64; }
// min / max (MB/operator/aggregation/AbstractMinMaxAggregationFunction.java: state = value if it compares better) are ONE combine rule
// here: every value is mapped to an unsigned word whose order is the type's order (`ordered`), MIN stores its complement, and the
// accumulator takes the unsigned maximum.  The all-zero word every accumulator starts from is the identity of that rule, so table
// initialisation, growth (rehash) and the tier folding need nothing special.  DOUBLE: NaN is canonical and sorts above +Infinity
// (TypeUtils.doubleCompare: the new NaN definition); -0.0 sorts below +0.0 (the reference keeps whichever zero came first).
__host__ __device__ inline bool kindIsMax(int k) { return k >= K_MIN_F64 && k <= K_MAX_I64; }
__host__ __device__ inline uint64_t orderedOfI64(uint64_t v) { return v ^ 0x8000000000000000ULL; }
__host__ __device__ inline uint64_t orderedOfF64(uint64_t bits)
{
    bits = (uint64_t) canonicalDoubleBits((int64_t) bits);
    return (bits >> 63) ? ~bits : (bits ^ 0x8000000000000000ULL);
}
__host__ __device__ inline uint64_t i64OfOrdered(uint64_t o) { return o ^ 0x8000000000000000ULL; }
__host__ __device__ inline uint64_t f64OfOrdered(uint64_t o) { return (o >> 63) ? (o ^ 0x8000000000000000ULL) : ~o; }
// combine rule of an accumulator column: 0 integer add, 1 double add, 2 unsigned max
enum { OP_ADD_I64 = 0, OP_ADD_F64 = 1, OP_MAX_U64 = 2 };

struct AccInput {
    const uint64_t* values;     // 8-byte input column (DOUBLE bits or BIGINT); nullptr for counts
    const uint8_t* nulls;
    const uint8_t* mask;        // BOOLEAN mask column or nullptr
    const uint8_t* maskNulls;
    const int32_t* offsets;     // K_DST_*: offsets of the VARBINARY block holding the states (`values` = its bytes)
    int32_t kind;
    int32_t phys;               // physical accumulator column
};

template <int MAXV>
struct AggLaunch {
    GroupTable t;
    KeyPlan kp;
    AccInput acc[MAXV];
    int32_t nAcc;
    uint32_t floatMask;          // bit c: accumulator c adds doubles
    uint32_t maxMask;            // bit c: accumulator c takes the unsigned maximum (min / max aggregates)
    int64_t n;                   // rows in this launch
    uint64_t rowBase;            // global index of row 0 (first-appearance bookkeeping)
    const uint32_t* bitmapIn;    // when set: only rows whose bit is set are processed (deferred-row pass)
    uint32_t* bitmapOut;         // rows that could not claim a slot (table at its limit)
    uint32_t S;                  // shared table slots (power of two); 0 = no shared table
    uint32_t sLimit;
    int32_t nAccAlloc;           // accumulator columns the shared table was sized for (>= nAcc)
    int32_t vec;                 // every 8-byte column pointer is 16-byte aligned: load row pairs with one 128-bit request
};

__device__ __forceinline__ void accAddShared(uint64_t* addr, uint64_t bits, int op)
{
    if (op == OP_ADD_F64) atomicAdd((double*) addr, __longlong_as_double((long long) bits));
    else if (op == OP_MAX_U64) atomicMax((unsigned long long*) addr, (unsigned long long) bits);
    else atomicAdd((unsigned long long*) addr, (unsigned long long) bits);
}
__device__ __forceinline__ void accAddGlobal(uint64_t* addr, uint64_t bits, int op)
{
    if (op == OP_ADD_F64) atomicAdd((double*) addr, __longlong_as_double((long long) bits));   // REDG.E.ADD.F64
    else if (op == OP_MAX_U64) atomicMax((unsigned long long*) addr, (unsigned long long) bits);
    else atomicAdd((unsigned long long*) addr, (unsigned long long) bits);
}
__device__ __forceinline__ uint64_t combineAcc(uint64_t a, uint64_t b, int op)
{
    if (op == OP_ADD_F64) return (uint64_t) __double_as_longlong(__longlong_as_double((long long) a) + __longlong_as_double((long long) b));
    if (op == OP_MAX_U64) return a > b ? a : b;
    return a + b;
}
__device__ __forceinline__ uint64_t contribOf(int kind, uint64_t raw)
{
    switch (kind) {
        case K_COUNT: return 1;
        case K_F64: case K_I64_ADD: return raw;
        case K_F64_FROM_I64: return (uint64_t) __double_as_longlong((double) (long long) raw);
        case K_LO32: return raw & 0xffffffffULL;
        case K_MAX_I64: return orderedOfI64(raw);
        case K_MIN_I64: return ~orderedOfI64(raw);
        case K_MAX_F64: return orderedOfF64(raw);
        case K_MIN_F64: return ~orderedOfF64(raw);
        default: return (uint64_t) ((long long) raw >> 32);   // K_HI32
    }
}

// Limb `kind - base` of a 128-bit two's complement value (limbs 0..2 unsigned, limb 3 signed).
__device__ __forceinline__ uint64_t limbOf(__int128 v, int limb)
{
    return limb == 3 ? (uint64_t) (int64_t) (v >> 96) : (uint64_t) ((unsigned __int128) v >> (32 * limb)) & 0xffffffffULL;
}
// sign-magnitude (UnscaledDecimal128Arithmetic) -> two's complement
__device__ __forceinline__ __int128 decimalOfSignMagnitude(uint64_t low, uint64_t high)
{
    const __int128 m = (__int128) (((unsigned __int128) (high & 0x7fffffffffffffffULL) << 64) | low);
    return (high >> 63) ? -m : m;
}
__device__ __forceinline__ uint64_t decimalContrib(const AccInput& a, int64_t row)
{
    if (kindIsDecimal(a.kind)) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(a.values + 2 * row);
        return limbOf(decimalOfSignMagnitude(t.x, t.y), a.kind - K_DEC_L0);
    }
    // a state: {overflow, low, high}, value = overflow * 2^127 + (low, high); the overflow lands in limb 3 as overflow << 31
    const unsigned char* p = (const unsigned char*) a.values + a.offsets[row];
    uint64_t w[3];
    for (int i = 0; i < 3; i++) { uint64_t x = 0; for (int b = 0; b < 8; b++) x |= (uint64_t) p[8 * i + b] << (8 * b); w[i] = x; }
    const uint64_t limb = limbOf(decimalOfSignMagnitude(w[1], w[2]), a.kind - K_DST_L0);
    return a.kind == K_DST_L3 ? limb + ((uint64_t) (int64_t) w[0] << 31) : limb;
}

// ---- the fused kernel ---------------------------------------------------------------------------------------------------
// Three tiers, hottest first:
//   tier 0  the first K0 groups a CTA claims: per-thread private accumulators in shared memory — plain LDS/DADD/STS, no atomics,
//           no shuffles (TPC-H Q1 has 4 groups: every row ends here)
//   tier 1  the CTA's shared hash table (up to sLimit groups): warp-aggregated shared atomics
//   tier 2  the HBM table: find-or-claim + REDG per accumulator (high cardinality)
template <int W, int MAXV>
__global__ void __launch_bounds__(kAggThreads) hashAggKernel(const __grid_constant__ AggLaunch<MAXV> P)
{
    constexpr int R = RowsPerThread<MAXV>::value;
    constexpr int K0 = Tier0Groups<MAXV>::value;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nAcc = P.nAcc;
    const GroupTable& T = P.t;
    const bool useShared = P.S != 0;
    auto opOf = [&](int c) { return ((P.maxMask >> c) & 1u) ? OP_MAX_U64 : (((P.floatMask >> c) & 1u) ? OP_ADD_F64 : OP_ADD_I64); };

    SharedTable<W> st;
    st.S = P.S; st.S2 = P.S + 2; st.sLimit = P.sLimit; st.tier0Groups = useShared ? K0 : 0;
    const SharedLayout lay(P.S, W, P.nAccAlloc, true);
    st.keys = (volatile uint64_t*) (smemRaw + lay.keys);
    st.acc = (uint64_t*) (smemRaw + lay.acc);
    st.first = (uint32_t*) (smemRaw + lay.first);
    st.state = (volatile uint32_t*) (smemRaw + lay.state);
    st.sid = (volatile uint8_t*) (smemRaw + lay.sid);
    st.count = (uint32_t*) (smemRaw + lay.misc);
    uint64_t* tier0 = (uint64_t*) (smemRaw + lay.tier0);     // [K0 * MAXV][256]
    if (useShared) {
        for (uint32_t i = tid; i < W * st.S2; i += kAggThreads) st.keys[i] = kEmptyKey;
        for (uint32_t i = tid; i < (uint32_t) P.nAccAlloc * st.S2; i += kAggThreads) st.acc[i] = 0;
        for (uint32_t i = tid; i < st.S2; i += kAggThreads) { st.first[i] = kNoRow32; st.sid[i] = 0xFF; if (W > 1) st.state[i] = kStEmpty; }
        if (tid < 64) st.count[tid] = 0;
#pragma unroll
        for (int i = 0; i < kTier0Columns; i++) tier0[i * kAggThreads + tid] = 0;
        __syncthreads();
    }

    bool warpBypass = !useShared;
    uint32_t warpHits = 0, warpRows = 0;
    uint32_t seen0 = 0;                       // tier-0 ids this thread has already recorded a first row for
    constexpr int TILE = kAggThreads * R;
    const int64_t nTiles = (P.n + TILE - 1) / TILE;

    for (int64_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        uint64_t kw[R][W];
        uint64_t contrib[R][MAXV];
        uint32_t addMask[R];
        int special[R];
        uint32_t spell[R];            // -0.0 spelling mask of the row's DOUBLE key columns (rare: such rows go straight to the HBM table)
        bool valid[R];
        uint32_t rowIdxOf[R];
        // ---- load phase: every load of the tile is issued before any is consumed; 8-byte columns come in as 128-bit pairs ----
#pragma unroll
        for (int j = 0; j < R / 2; j++) {
            const int64_t row0 = tile * TILE + (int64_t) j * (2 * kAggThreads) + 2 * tid;
            bool v0 = row0 < P.n, v1 = row0 + 1 < P.n;
            if (P.bitmapIn) {
                if (v0) v0 = (P.bitmapIn[row0 >> 5] >> (row0 & 31)) & 1u;
                if (v1) v1 = (P.bitmapIn[(row0 + 1) >> 5] >> ((row0 + 1) & 31)) & 1u;
            }
            const bool pair = P.vec && v0 && v1;
            valid[2 * j] = v0; valid[2 * j + 1] = v1;
            rowIdxOf[2 * j] = (uint32_t) row0; rowIdxOf[2 * j + 1] = (uint32_t) (row0 + 1);
            addMask[2 * j] = 0; addMask[2 * j + 1] = 0;
            special[2 * j] = 0; special[2 * j + 1] = 0;
            spell[2 * j] = 0; spell[2 * j + 1] = 0;
#pragma unroll
            for (int w = 0; w < W; w++) { kw[2 * j][w] = 0; kw[2 * j + 1][w] = 0; }
            // keys
            if (P.kp.bigintPath) {
                const KeyField& f = P.kp.f[0];
                uint64_t a = 0, b = 0;
                if (pair) { const ulonglong2 t = ldStream128((const uint64_t*) f.values + row0); a = t.x; b = t.y; }
                else { if (v0) a = ldStream64((const uint64_t*) f.values + row0); if (v1) b = ldStream64((const uint64_t*) f.values + row0 + 1); }
                kw[2 * j][0] = a; kw[2 * j + 1][0] = b;
                if (v0) special[2 * j] = (f.nulls && f.nulls[row0]) ? 1 : (a == kEmptyKey ? 2 : 0);
                if (v1) special[2 * j + 1] = (f.nulls && f.nulls[row0 + 1]) ? 1 : (b == kEmptyKey ? 2 : 0);
            }
            else {
                if (v0) special[2 * j] = packKey<W>(P.kp, row0, kw[2 * j], &spell[2 * j]);
                if (v1) special[2 * j + 1] = packKey<W>(P.kp, row0 + 1, kw[2 * j + 1], &spell[2 * j + 1]);
            }
            // accumulator inputs
#pragma unroll
            for (int c = 0; c < MAXV; c++) {
                contrib[2 * j][c] = 0; contrib[2 * j + 1][c] = 0;
                if (c < nAcc) {
                    const AccInput& a = P.acc[c];
                    bool add0 = v0, add1 = v1;
                    if (a.mask) {
                        if (add0) add0 = a.mask[row0] != 0 && !(a.maskNulls && a.maskNulls[row0]);
                        if (add1) add1 = a.mask[row0 + 1] != 0 && !(a.maskNulls && a.maskNulls[row0 + 1]);
                    }
                    if (a.nulls) { if (add0) add0 = a.nulls[row0] == 0; if (add1) add1 = a.nulls[row0 + 1] == 0; }
                    uint64_t r0 = 0, r1 = 0;
                    if (a.kind >= K_DEC_L0) {      // sum(decimal) over 16-byte values / serialized states: rows whose input is NULL hold no bytes
                        if (add0) contrib[2 * j][c] = decimalContrib(a, row0);
                        if (add1) contrib[2 * j + 1][c] = decimalContrib(a, row0 + 1);
                    }
                    else {
                        if (a.kind != K_COUNT) {
                            if (pair) { const ulonglong2 t = ldStream128(a.values + row0); r0 = t.x; r1 = t.y; }
                            else { if (v0) r0 = ldStream64(a.values + row0); if (v1) r1 = ldStream64(a.values + row0 + 1); }
                        }
                        contrib[2 * j][c] = contribOf(a.kind, r0); contrib[2 * j + 1][c] = contribOf(a.kind, r1);
                    }
                    if (add0) addMask[2 * j] |= 1u << c;
                    if (add1) addMask[2 * j + 1] |= 1u << c;
                }
            }
        }
        // ---- accumulate phase ----
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t rowIdx = rowIdxOf[r];
            const bool v = valid[r];
            uint64_t h = 0;
            int sslot = -1;
            bool done = false;
            if (v) {
                h = hashKeyWords<W>(kw[r]);
                if (!warpBypass && !spell[r]) sslot = special[r] ? (int) st.S + special[r] - 1 : st.find(kw[r], h);
                if (sslot >= 0) {
                    const uint32_t id = st.sid[sslot];
                    if (id < (uint32_t) K0) {
                        // tier 0: this thread's private accumulators of group `id` (bank-conflict free: consecutive threads, consecutive words)
                        uint64_t* mine = tier0 + (size_t) id * MAXV * kAggThreads + tid;
#pragma unroll
                        for (int c = 0; c < MAXV; c++) {
                            if (c < nAcc && ((addMask[r] >> c) & 1u)) {
                                uint64_t* p = mine + c * kAggThreads;
                                *p = combineAcc(*p, contrib[r][c], opOf(c));
                            }
                        }
                        if (!((seen0 >> id) & 1u)) { seen0 |= 1u << id; atomicMin(&st.first[sslot], rowIdx); }   // rows ascend per thread: first sight = this thread's minimum
                        done = true;
                    }
                }
            }
            const bool hit = v && sslot >= 0 && !done;
            const unsigned hits = __ballot_sync(0xffffffffu, hit);
            warpHits += __popc(__ballot_sync(0xffffffffu, v && sslot >= 0));
            warpRows += __popc(__ballot_sync(0xffffffffu, v));
            if (hits) {
                // tier 1: group the warp's lanes by shared slot; lanes that missed form singleton groups
                const int matchKey = hit ? sslot : -1 - lane;
                const unsigned peers = __match_any_sync(0xffffffffu, matchKey);
                const bool isLeader = hit && (lane == __ffs(peers) - 1);
                const unsigned leaders = __ballot_sync(0xffffffffu, isLeader);
                bool active = hit;
                uint32_t myAdd = addMask[r];
                if (__popc(leaders) <= kWarpAggMaxGroups && __popc(hits) > __popc(leaders)) {
                    // few distinct groups in this warp: reduce each group's lanes with shuffles, one shared atomic per group
                    uint64_t mine[MAXV];
#pragma unroll
                    for (int c = 0; c < MAXV; c++) mine[c] = 0;
                    uint32_t mineAdd = 0;
                    unsigned rem = leaders;
                    while (rem) {
                        const int gl = __ffs(rem) - 1;
                        rem &= rem - 1;
                        const unsigned gmask = __shfl_sync(0xffffffffu, peers, gl);
                        const bool in = hit && ((gmask >> lane) & 1u);
#pragma unroll
                        for (int c = 0; c < MAXV; c++) {
                            if (c < nAcc) {
                                const bool adds = in && ((addMask[r] >> c) & 1u);
                                const unsigned any = __ballot_sync(0xffffffffu, adds);
                                if (any) {
                                    uint64_t x = adds ? contrib[r][c] : 0;       // 0 is the identity of all three combine rules
                                    const int op = opOf(c);
#pragma unroll
                                    for (int m = 16; m >= 1; m >>= 1) x = combineAcc(x, shflXor64(x, m), op);
                                    if (lane == gl) { mine[c] = x; mineAdd |= 1u << c; }
                                }
                            }
                        }
                    }
                    active = isLeader;
                    myAdd = mineAdd;
#pragma unroll
                    for (int c = 0; c < MAXV; c++) contrib[r][c] = isLeader ? mine[c] : contrib[r][c];
                }
                if (active) {
#pragma unroll
                    for (int c = 0; c < MAXV; c++)
                        if (c < nAcc && ((myAdd >> c) & 1u)) accAddShared(&st.acc[c * st.S2 + sslot], contrib[r][c], opOf(c));
                    if (rowIdx < *(volatile uint32_t*) &st.first[sslot]) atomicMin(&st.first[sslot], rowIdx);
                }
            }
            if (v && sslot < 0) {
                // tier 2: HBM table (high cardinality, or the shared table is full)
                int gslot = special[r] ? (int) T.cap + special[r] - 1 : tableFind<W>(T, kw[r], h, true, false);
                if (gslot == -2) {
                    atomicOr(&P.bitmapOut[rowIdx >> 5], 1u << (rowIdx & 31));
                    atomicAdd(&T.counters[CNT_DEFERRED], 1u);
                }
                else {
                    const uint64_t g = firstRowWord(P.rowBase + rowIdx, spell[r]);
                    if (g < __ldcg(&T.firstRow[gslot])) atomicMin((unsigned long long*) &T.firstRow[gslot], (unsigned long long) g);
#pragma unroll
                    for (int c = 0; c < MAXV; c++)
                        if (c < nAcc && ((addMask[r] >> c) & 1u))
                            accAddGlobal(&T.acc[(size_t) gslot * T.nvPhys + P.acc[c].phys], contrib[r][c], opOf(c));
                }
            }
        }
        // a warp whose rows mostly miss a full shared table stops probing it (high-cardinality input)
        if (!warpBypass && warpRows >= 256) {
            if (*(volatile uint32_t*) st.count >= st.sLimit && warpHits * 2 < warpRows) warpBypass = true;
            warpHits = 0; warpRows = 0;
        }
    }

    // ---- fold tier 0 into the CTA's shared table, then flush that into the HBM table ----
    if (useShared) {
        __syncthreads();
        const uint32_t nT0 = min(st.count[1], (uint32_t) K0);
        const int warp = tid >> 5;
        for (uint32_t pairIdx = warp; pairIdx < nT0 * (uint32_t) MAXV; pairIdx += kAggThreads / 32) {
            const uint32_t id = pairIdx / MAXV, c = pairIdx % MAXV;
            if ((int) c >= nAcc) continue;
            const uint64_t* src = tier0 + (size_t) pairIdx * kAggThreads;
            const int op = opOf((int) c);
            uint64_t x = 0;
            for (int i = lane; i < kAggThreads; i += 32) x = combineAcc(x, src[i], op);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) x = combineAcc(x, shflXor64(x, m), op);
            if (lane == 0) {
                uint64_t* dst = &st.acc[c * st.S2 + st.count[2 + id]];
                *dst = combineAcc(*dst, x, op);
            }
        }
        __syncthreads();
        for (uint32_t s = tid; s < st.S2; s += kAggThreads) {
            const uint32_t f = st.first[s];
            if (f == kNoRow32) continue;
            int gslot;
            if (s >= st.S) gslot = (int) T.cap + (int) (s - st.S);
            else {
                uint64_t kw[W];
#pragma unroll
                for (int w = 0; w < W; w++) kw[w] = st.keys[w * st.S2 + s];
                gslot = tableFind<W>(T, kw, hashKeyWords<W>(kw), true, true);
            }
            if (gslot < 0) { atomicOr(&T.counters[CNT_ERROR], 0x80000000u); continue; }   // cannot happen: flush reserve (see HashAggregationOp)
            const uint64_t g = firstRowWord(P.rowBase + f);
            if (g < __ldcg(&T.firstRow[gslot])) atomicMin((unsigned long long*) &T.firstRow[gslot], (unsigned long long) g);
            for (int c = 0; c < nAcc; c++) {
                const uint64_t x = st.acc[c * st.S2 + s];
                if (x != 0) accAddGlobal(&T.acc[(size_t) gslot * T.nvPhys + P.acc[c].phys], x, opOf(c));
            }
        }
    }
}

// ---- specialised kernel: one non-null BIGINT key, NV non-null DOUBLE sums (+ optional row count), no masks ------------------
// Same three tiers and the same tables as hashAggKernel, but every runtime flag of the generic kernel is a compile-time
// constant here, and the tile loads are software pipelined (tile t+1's 128-bit loads are in flight while tile t is consumed).
template <int NV, bool CNT>
struct FastLaunch {
    GroupTable t;
    const uint64_t* key;
    const uint64_t* vals[NV];
    int32_t phys[NV + 1];        // physical accumulator column of each value column, then of the row count
    int64_t n;
    uint64_t rowBase;
    uint32_t* bitmapOut;
    uint32_t S, sLimit;
    int32_t nAccAlloc;
    uint32_t subsetCount, subsetIndex;   // mid mode: this launch aggregates only keys with hashSubset(key) == subsetIndex
    uint32_t capBits;                    // direct mode: log2(t.cap); subsets are contiguous slot ranges of the HBM table
    // packed-key mode (keyMode 1): two 32-bit key columns without NULLs on this page (Q1: returnflag, linestatus).  Inside the kernel
    // the key is (col1 << 32 | col0); at the HBM table it is re-packed into the table's W = 2 layout (KeyLayout: word / shift per column).
    int32_t warpGroups;                  // > 0: warp-private accumulator sets for the first warpGroups group ids instead of tier 0 (11 .. ~250 groups)
    int32_t keyMode;
    const uint32_t* key32[2];
    int32_t kWord[2], kShift[2];
};

// slot of a fast-kernel key in the HBM table (insert)
template <int NV, bool CNT>
__device__ __forceinline__ int fastGlobalFind(const FastLaunch<NV, CNT>& P, uint64_t key, bool fromFlush)
{
    if (P.keyMode == 0) {
        if (key == kEmptyKey) return (int) P.t.cap + 1;
        uint64_t kw[1] = {key};
        return tableFind<1>(P.t, kw, murmur64(key), true, fromFlush);
    }
    uint64_t kw[2] = {0, 0};
    const uint64_t v0 = key & 0xffffffffULL, v1 = key >> 32;
    if (P.kWord[0] == 0) kw[0] |= v0 << P.kShift[0]; else kw[1] |= v0 << P.kShift[0];
    if (P.kWord[1] == 0) kw[0] |= v1 << P.kShift[1]; else kw[1] |= v1 << P.kShift[1];
    return tableFind<2>(P.t, kw, hashKeyWords<2>(kw), true, fromFlush);
}

// Rows that miss the CTA's shared table: the HBM table (tier 2).  Kept out of line so the hot loop stays small.
template <int NV, bool CNT>
__device__ __noinline__ void fastGlobalRow(const FastLaunch<NV, CNT>& P, uint64_t key, uint64_t h, const uint64_t (&vals)[NV], uint32_t rowIdx)
{
    const GroupTable& T = P.t;
    uint64_t kw[1] = {key};
    int gslot = key == kEmptyKey ? (int) T.cap + 1 : tableFind<1>(T, kw, h, true, false);
    if (gslot == -2) {
        atomicOr(&P.bitmapOut[rowIdx >> 5], 1u << (rowIdx & 31));
        atomicAdd(&T.counters[CNT_DEFERRED], 1u);
        return;
    }
    const uint64_t g = firstRowWord(P.rowBase + rowIdx);
    if (g < __ldcg(&T.firstRow[gslot])) atomicMin((unsigned long long*) &T.firstRow[gslot], (unsigned long long) g);
    uint64_t* acc = T.acc + (size_t) gslot * T.nvPhys;
#pragma unroll
    for (int c = 0; c < NV; c++) atomicAdd((double*) &acc[P.phys[c]], __longlong_as_double((long long) vals[c]));
    if (CNT) atomicAdd((unsigned long long*) &acc[P.phys[NV]], 1ULL);
}

// ---- direct mode: every row goes straight to the HBM table (10^4 .. 10^7 groups) ---------------------------------------------
// Past a few thousand groups a shared-memory table stops catching rows, and each row costs its L2 atomics anyway: one key probe, NV
// native f64 REDs (REDG.ADD.F64), one firstRow min.  What is left to engineer is WHERE those atomics land: the launch aggregates
// only the keys whose home slot lies in slot range `subsetIndex` of `subsetCount` (contiguous ranges of the table), the host running
// as many passes over the chunk as it takes for one range's keys + accumulators to stay L2 resident — the streamed input is loaded
// with an evict-first policy so that it does not wash the table out of L2.  Per pass the kernel reads every key (8 B/row) and the
// value sectors of the rows it owns.
__device__ __forceinline__ ulonglong2 ldStreamEvictFirst128(const uint64_t* p, uint64_t policy)
{
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint64_t ldStreamEvictFirst64(const uint64_t* p, uint64_t policy)
{
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
    return v;
}

// Each warp scans keys (two per lane per step, software pipelined) and queues the rows it owns in shared memory; whenever 32 are
// queued the whole warp aggregates them, one row per lane — so with 1/8 of the rows owned the dependent chain (value loads, key
// probe, firstRow, REDs) runs once per ~256 scanned rows instead of once per 64, with every lane busy.
constexpr int kDirectQueue = 96;      // per warp: up to 31 left over + 64 new entries
template <int NV, bool CNT>
__global__ void __launch_bounds__(256, 4) hashAggDirectKernel(const __grid_constant__ FastLaunch<NV, CNT> P)
{
    constexpr int THREADS = 256, TILE = THREADS * 2;
    __shared__ uint64_t qKey[THREADS / 32][kDirectQueue];
    __shared__ uint32_t qRow[THREADS / 32][kDirectQueue];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1;
    const GroupTable& T = P.t;
    uint64_t streamPolicy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(streamPolicy));
    const int64_t nTiles = (P.n + TILE - 1) / TILE;
    ulonglong2 kA, kB;
    auto loadKeys = [&](int64_t tile, ulonglong2& k) {
        const int64_t row0 = tile * TILE + 2 * tid;
        k.x = 0; k.y = 0;
        if (row0 + 1 < P.n) k = ldStreamEvictFirst128(P.key + row0, streamPolicy);
        else if (row0 < P.n) k.x = ldStreamEvictFirst64(P.key + row0, streamPolicy);
    };
    auto owns = [&](uint64_t h) {
        return P.subsetCount <= 1 || (uint32_t) (((uint64_t) ((uint32_t) h & T.mask) * P.subsetCount) >> P.capBits) == P.subsetIndex;
    };
    auto aggregateRow = [&](uint64_t key, uint32_t rowIdx) {
        uint64_t vals[NV];
#pragma unroll
        for (int c = 0; c < NV; c++) vals[c] = ldStreamEvictFirst64(P.vals[c] + rowIdx, streamPolicy);
        uint64_t kw[1] = {key};
        const int gslot = key == kEmptyKey ? (int) T.cap + 1 : tableFind<1>(T, kw, murmur64(key), true, false);
        if (gslot == -2) {
            atomicOr(&P.bitmapOut[rowIdx >> 5], 1u << (rowIdx & 31));
            atomicAdd(&T.counters[CNT_DEFERRED], 1u);
            return;
        }
        const uint64_t g = firstRowWord(P.rowBase + rowIdx);
        if (g < __ldcg(&T.firstRow[gslot])) atomicMin((unsigned long long*) &T.firstRow[gslot], (unsigned long long) g);
        uint64_t* acc = T.acc + (size_t) gslot * T.nvPhys;
#pragma unroll
        for (int c = 0; c < NV; c++) atomicAdd((double*) &acc[P.phys[c]], __longlong_as_double((long long) vals[c]));
        if (CNT) atomicAdd((unsigned long long*) &acc[P.phys[NV]], 1ULL);
    };
    uint32_t qCount = 0;       // warp-uniform
    auto push = [&](bool mine, uint64_t key, uint32_t rowIdx) {
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        if (mine) { const uint32_t at = qCount + __popc(m & lt); qKey[warp][at] = key; qRow[warp][at] = rowIdx; }
        qCount += __popc(m);
    };
    auto drain = [&](uint32_t keep) {       // aggregate queued rows, 32 at a time, until at most `keep` are left
        __syncwarp();
        while (qCount > keep) {
            const uint32_t take = min(32u, qCount), at = qCount - take + lane;
            if (lane < take) aggregateRow(qKey[warp][at], qRow[warp][at]);
            qCount -= take;
            __syncwarp();
        }
    };
    auto consumeTile = [&](int64_t tile, const ulonglong2& k) {
        const int64_t row0 = tile * TILE + 2 * tid;
        push(row0 < P.n && owns(murmur64(k.x)), k.x, (uint32_t) row0);
        push(row0 + 1 < P.n && owns(murmur64(k.y)), k.y, (uint32_t) (row0 + 1));
        if (qCount >= 32) drain(31);
    };
    int64_t tile = blockIdx.x;
    if (tile < nTiles) loadKeys(tile, kA);
    while (tile < nTiles) {
        const int64_t t1 = tile + gridDim.x;
        if (t1 < nTiles) loadKeys(t1, kB);
        consumeTile(tile, kA);
        if (t1 >= nTiles) break;
        const int64_t t2 = t1 + gridDim.x;
        if (t2 < nTiles) loadKeys(t2, kA);
        consumeTile(t1, kB);
        tile = t2;
    }
    drain(0);
}

// MID = false: tiers 0/1/2, 256 threads, two CTAs per SM (few groups: TPC-H Q1).
// MID = true : no tier 0; one 512-thread CTA per SM owns a 4096-slot shared table and the launch only aggregates the keys of hash
//              subset P.subsetIndex of P.subsetCount — the host runs subsetCount passes over the chunk so that every pass's groups
//              fit in shared memory instead of paying 4 L2 atomics per row (10^3 .. 10^5 groups).
template <int NV, bool CNT, bool MID>
__global__ void __launch_bounds__(MID ? 512 : kAggThreads, MID ? 1 : 2) hashAggFastKernel(const __grid_constant__ FastLaunch<NV, CNT> P)
{
    constexpr int THREADS = MID ? 512 : kAggThreads;
    constexpr int NC = FastGeom<NV, CNT>::NC;
    constexpr int K0 = MID ? 0 : FastGeom<NV, CNT>::K0;
    constexpr int TILE = THREADS * 2;            // one 128-bit pair per thread per column per tile
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const GroupTable& T = P.t;

    SharedTable<1> st;
    // warp mode (host decides: more groups than tier 0 holds, few enough for per-warp sets): the tier-0 region holds, per warp,
    // [NC accumulator columns + firstRow][KW] instead; a warp updates its own set with plain LDS / DADD / STS, resolving lanes that
    // hit the same group in rounds (rank among the __match_any_sync peers) — no CAS loops, no contention across warps
    const int KW = MID ? 0 : P.warpGroups;
    st.S = P.S; st.S2 = P.S + 2; st.sLimit = P.sLimit; st.tier0Groups = KW > 0 ? KW : K0;
    const SharedLayout lay(P.S, 1, P.nAccAlloc, !MID);
    st.keys = (volatile uint64_t*) (smemRaw + lay.keys);
    st.acc = (uint64_t*) (smemRaw + lay.acc);
    st.first = (uint32_t*) (smemRaw + lay.first);
    st.state = nullptr;
    st.sid = (volatile uint8_t*) (smemRaw + lay.sid);
    st.count = (uint32_t*) (smemRaw + lay.misc);
    uint64_t* tier0 = (uint64_t*) (smemRaw + lay.tier0);     // [K0][NC][256]
    for (uint32_t i = tid; i < st.S2; i += THREADS) { st.keys[i] = kEmptyKey; st.first[i] = kNoRow32; st.sid[i] = 0xFF; }
    for (uint32_t i = tid; i < (uint32_t) P.nAccAlloc * st.S2; i += THREADS) st.acc[i] = 0;
    for (int i = tid; i < 320; i += THREADS) st.count[i] = 0;
    uint64_t* wtab = tier0 + (size_t) (tid >> 5) * (NC + 1) * (KW > 0 ? KW : 1);    // this warp's [NC + 1][KW]
    if (!MID) {
        if (KW > 0) {
            for (int i = tid; i < (THREADS / 32) * (NC + 1) * KW; i += THREADS) tier0[i] = ((i / KW) % (NC + 1)) == NC ? ~0ULL : 0ULL;
        }
        else {
#pragma unroll
            for (int i = 0; i < K0 * NC; i++) tier0[i * THREADS + tid] = 0;
        }
    }
    __syncthreads();

    uint32_t seen0 = 0;
    uint32_t probeHits = 0, probeMisses = 0;
    bool bypass = false;
    const int64_t nTiles = (P.n + TILE - 1) / TILE;
    const uint64_t* keyBase = P.key;

    // double-buffered tile registers
    ulonglong2 kA, kB, vA[NV], vB[NV];
    auto loadTile = [&](int64_t tile, ulonglong2& k, ulonglong2 (&v)[NV]) {
        const int64_t row0 = tile * TILE + 2 * tid;
        if (row0 + 1 < P.n) {
            if (P.keyMode == 0) k = ldStream128(keyBase + row0);
            else {      // two rows of each 32-bit key column in one 64-bit load; key = col1 << 32 | col0
                const uint64_t a = ldStream64((const uint64_t*) (P.key32[0] + row0)), b = ldStream64((const uint64_t*) (P.key32[1] + row0));
                k.x = (a & 0xffffffffULL) | (b << 32);
                k.y = (a >> 32) | (b & 0xffffffff00000000ULL);
            }
#pragma unroll
            for (int c = 0; c < NV; c++) v[c] = ldStream128(P.vals[c] + row0);
        }
        else if (row0 < P.n) {
            if (P.keyMode == 0) k.x = ldStream64(keyBase + row0);
            else k.x = (uint64_t) ldStream32(P.key32[0] + row0) | ((uint64_t) ldStream32(P.key32[1] + row0) << 32);
            k.y = 0;
#pragma unroll
            for (int c = 0; c < NV; c++) { v[c].x = ldStream64(P.vals[c] + row0); v[c].y = 0; }
        }
    };
    // shared-table slot and group id of a key (sslot < 0: not in the CTA's table)
    auto lookup = [&](uint64_t key, uint64_t h, int& sslot, uint32_t& id) {
        sslot = -1;
        if (key == kEmptyKey) sslot = (int) st.S + 1;      // the sentinel-valued key has its own slot (an empty slot must never match it)
        else if (!bypass) {
            const uint32_t s = (uint32_t) (h >> 32) & (st.S - 1);
            if (st.keys[s] == key) sslot = (int) s;
            else { uint64_t kw[1] = {key}; sslot = st.find(kw, h); }
            // a thread whose rows keep missing a full shared table stops probing it (high-cardinality input)
            if (sslot >= 0) probeHits++; else probeMisses++;
            if (((probeHits + probeMisses) & 63u) == 0) {
                if (probeMisses > 3 * probeHits && *(volatile uint32_t*) st.count >= st.sLimit) bypass = true;
                probeHits = 0; probeMisses = 0;
            }
        }
        id = 0xFF;
        if (sslot >= 0) id = st.sid[sslot];
    };
    // tiers 1 and 2
    auto applyShared = [&](uint64_t key, uint64_t h, const uint64_t (&vals)[NV], uint32_t rowIdx, int sslot) {
        if (sslot >= 0) {
            // tier 1: the CTA's shared table, shared-space reductions (no return value needed)
#pragma unroll
            for (int c = 0; c < NV; c++) sharedAddF64(&st.acc[c * st.S2 + sslot], vals[c]);
            if (CNT) sharedAddU64(&st.acc[NV * st.S2 + sslot], 1);
            if (rowIdx < *(volatile uint32_t*) &st.first[sslot]) sharedMinU32(&st.first[sslot], rowIdx);
        }
        else fastGlobalRow<NV, CNT>(P, key, h, vals, rowIdx);
    };
    auto consumeRow = [&](uint64_t key, const uint64_t (&vals)[NV], uint32_t rowIdx) {
        const uint64_t h = murmur64(key);
        // subset from the LOW hash word (the shared slot index uses bits 32..43: the two must be independent)
        if (MID && P.subsetCount > 1 && (uint32_t) (((h & 0xffffffffULL) * (uint64_t) P.subsetCount) >> 32) != P.subsetIndex) return;   // another pass owns this key (the caller reconverges the warp)
        int sslot; uint32_t id;
        lookup(key, h, sslot, id);
        if (!MID && id < (uint32_t) K0) {
            uint64_t* mine = tier0 + (size_t) id * NC * THREADS + tid;
#pragma unroll
            for (int c = 0; c < NV; c++) {
                double* p = (double*) (mine + c * THREADS);
                *p = *p + __longlong_as_double((long long) vals[c]);
            }
            if (CNT) mine[NV * THREADS] += 1;
            if (!((seen0 >> id) & 1u)) { seen0 |= 1u << id; atomicMin(&st.first[sslot], rowIdx); }
        }
        else applyShared(key, h, vals, rowIdx, sslot);
    };
    // warp mode: the whole warp takes one row per lane together (live = the lane has a row)
    auto consumeRowWarp = [&](bool live, uint64_t key, const uint64_t (&vals)[NV], uint32_t rowIdx) {
        const uint64_t h = murmur64(key);
        int sslot = -1; uint32_t id = 0xFF;
        if (live) lookup(key, h, sslot, id);
        __syncwarp();
        const bool mine = live && id < (uint32_t) KW;
        const int wid = mine ? (int) id : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, wid);
        const int rank = __popc(peers & ((1u << lane) - 1));
        const int rounds = (int) __reduce_max_sync(0xffffffffu, mine ? (unsigned) __popc(peers) : 0u);
        for (int r = 0; r < rounds; r++) {
            if (mine && rank == r) {
                volatile uint64_t* w = wtab + wid;
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    const double cur = __longlong_as_double((long long) w[c * KW]);
                    w[c * KW] = (uint64_t) __double_as_longlong(cur + __longlong_as_double((long long) vals[c]));
                }
                if (CNT) w[NV * KW] = w[NV * KW] + 1;
                if ((uint64_t) rowIdx < w[NC * KW]) w[NC * KW] = rowIdx;
            }
            __syncwarp();
        }
        if (live && !mine) applyShared(key, h, vals, rowIdx, sslot);
        __syncwarp();
    };
    auto consumeTile = [&](int64_t tile, const ulonglong2& k, const ulonglong2 (&v)[NV]) {
        const int64_t row0 = tile * TILE + 2 * tid;
        if (KW > 0) {
            uint64_t vals[NV];
#pragma unroll
            for (int c = 0; c < NV; c++) vals[c] = v[c].x;
            consumeRowWarp(row0 < P.n, k.x, vals, (uint32_t) row0);
#pragma unroll
            for (int c = 0; c < NV; c++) vals[c] = v[c].y;
            consumeRowWarp(row0 + 1 < P.n, k.y, vals, (uint32_t) (row0 + 1));
            return;
        }
        if (row0 < P.n) {
            uint64_t vals[NV];
#pragma unroll
            for (int c = 0; c < NV; c++) vals[c] = v[c].x;
            consumeRow(k.x, vals, (uint32_t) row0);
        }
        if (row0 + 1 < P.n) {
            uint64_t vals[NV];
#pragma unroll
            for (int c = 0; c < NV; c++) vals[c] = v[c].y;
            consumeRow(k.y, vals, (uint32_t) (row0 + 1));
        }
    };

    int64_t tile = blockIdx.x;
    if (tile < nTiles) loadTile(tile, kA, vA);
    while (tile < nTiles) {
        const int64_t t1 = tile + gridDim.x;
        if (t1 < nTiles) loadTile(t1, kB, vB);
        consumeTile(tile, kA, vA);
        __syncwarp();      // rows take different tiers (or skip: hash subsets); the next tile's 128-bit loads must be issued by whole warps
        if (t1 >= nTiles) break;
        const int64_t t2 = t1 + gridDim.x;
        if (t2 < nTiles) loadTile(t2, kA, vA);
        consumeTile(t1, kB, vB);
        __syncwarp();
        tile = t2;
    }

    // ---- fold tier 0 (or the warp sets) into the shared table, then flush the shared table into the HBM table ----
    __syncthreads();
    if (KW > 0) {
        const uint32_t nIds = min(st.count[1], (uint32_t) KW);
        for (uint32_t i = tid; i < nIds * (uint32_t) (NC + 1); i += THREADS) {
            const uint32_t id = i / (NC + 1), c = i % (NC + 1);
            const uint32_t slot = st.count[2 + id];
            if (c == (uint32_t) NC) {
                uint64_t f = ~0ULL;
                for (int w = 0; w < THREADS / 32; w++) f = min(f, tier0[((size_t) w * (NC + 1) + c) * KW + id]);
                if (f != ~0ULL) st.first[slot] = min(st.first[slot], (uint32_t) f);
            }
            else if (c < (uint32_t) NV) {
                double d = 0;
                for (int w = 0; w < THREADS / 32; w++) d += __longlong_as_double((long long) tier0[((size_t) w * (NC + 1) + c) * KW + id]);
                double* dst = (double*) &st.acc[c * st.S2 + slot];
                *dst = *dst + d;
            }
            else {
                uint64_t x = 0;
                for (int w = 0; w < THREADS / 32; w++) x += tier0[((size_t) w * (NC + 1) + c) * KW + id];
                st.acc[c * st.S2 + slot] += x;
            }
        }
        __syncthreads();
    }
    const uint32_t nT0 = (MID || KW > 0) ? 0u : min(st.count[1], (uint32_t) K0);
    const int warp = tid >> 5;
    for (uint32_t pairIdx = warp; pairIdx < nT0 * (uint32_t) NC; pairIdx += THREADS / 32) {
        const uint32_t id = pairIdx / NC, c = pairIdx % NC;
        const uint64_t* src = tier0 + (size_t) pairIdx * THREADS;
        const bool isF = c < (uint32_t) NV;
        uint64_t x = 0;
        if (isF) {
            double d = 0;
            for (int i = lane; i < THREADS; i += 32) d += __longlong_as_double((long long) src[i]);
            x = (uint64_t) __double_as_longlong(d);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) x = (uint64_t) __double_as_longlong(__longlong_as_double((long long) x) + __longlong_as_double((long long) shflXor64(x, m)));
        }
        else {
            for (int i = lane; i < THREADS; i += 32) x += src[i];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) x += shflXor64(x, m);
        }
        if (lane == 0) {
            uint64_t* dst = &st.acc[c * st.S2 + st.count[2 + id]];
            if (isF) *(double*) dst = *(double*) dst + __longlong_as_double((long long) x); else *dst = *dst + x;
        }
    }
    __syncthreads();
    for (uint32_t s = tid; s < st.S2; s += THREADS) {
        const uint32_t f = st.first[s];
        if (f == kNoRow32) continue;
        // (slot S + 1 holds the key whose value equals the empty sentinel; slot S, the NULL key, is never used by this kernel)
        const int gslot = fastGlobalFind<NV, CNT>(P, s >= st.S ? kEmptyKey : (uint64_t) st.keys[s], true);
        if (gslot < 0) { atomicOr(&T.counters[CNT_ERROR], 0x80000000u); continue; }
        const uint64_t g = firstRowWord(P.rowBase + f);
        if (g < __ldcg(&T.firstRow[gslot])) atomicMin((unsigned long long*) &T.firstRow[gslot], (unsigned long long) g);
        uint64_t* acc = T.acc + (size_t) gslot * T.nvPhys;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint64_t x = st.acc[c * st.S2 + s];
            if (x != 0) accAddGlobal(&acc[P.phys[c]], x, c < NV);
        }
    }
}

// ---- table maintenance kernels --------------------------------------------------------------------------------------------
__global__ void fillU64Kernel(uint64_t* p, uint64_t v, size_t n)
{
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}

template <int W>
__global__ void rehashKernel(GroupTable oldT, GroupTable newT)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < oldT.capP; s += gridDim.x * blockDim.x) {
        int ns;
        if (s >= oldT.cap) {
            if (oldT.firstRow[s] == kNoRow) continue;
            ns = (int) newT.cap + (int) (s - oldT.cap);
        }
        else {
            uint64_t kw[W];
            if (W == 1) { kw[0] = oldT.keys[s]; if (kw[0] == kEmptyKey) continue; }
            else {
                if (oldT.state[s] == kStEmpty) continue;
#pragma unroll
                for (int w = 0; w < W; w++) kw[w] = oldT.keys[(size_t) w * oldT.capP + s];
            }
            ns = tableFind<W>(newT, kw, hashKeyWords<W>(kw), true, true);
        }
        newT.firstRow[ns] = oldT.firstRow[s];
        for (uint32_t c = 0; c < oldT.nvPhys; c++) newT.acc[(size_t) ns * newT.nvPhys + c] = oldT.acc[(size_t) s * oldT.nvPhys + c];
        if (oldT.gid) newT.gid[ns] = oldT.gid[s];
    }
}

// acc[slot][dst] = src >= 0 ? acc[slot][src] : 1   for every touched slot (lazy activation of a per-channel non-null count)
__global__ void backfillCountKernel(GroupTable t, int dst, int src)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < t.capP; s += gridDim.x * blockDim.x)
        if (t.firstRow[s] != kNoRow) t.acc[(size_t) s * t.nvPhys + dst] = src >= 0 ? t.acc[(size_t) s * t.nvPhys + src] : 1;
}

// ---- aggregation state hand-off (PARTIAL → FINAL between GPUs without pages) ----------------------------------------------------
// The reference ships a PARTIAL HashAggregationOperator's groups as intermediate pages through an exchange to the FINAL operator
// (HashAggregationOperator.java:451-501 getOutput → PartitionedOutput/ExchangeClient → addInput of the FINAL step, which runs
// GroupedAccumulator.addIntermediate).  Between GPUs of one NVSwitch domain the same hand-off is: every source table compacted into a
// fixed-capacity wire buffer in HBM (one kernel, no host round trip), ONE all-gather of those buffers, one merge kernel per source on
// the destination.  The wire carries the packed key words and the physical accumulator columns, i.e. exactly the intermediate state.
struct AggStateHeader {
    uint32_t magic, W, nvPhys, maxGroups;
    uint32_t nGroups, overflow;          // written by the export kernel
    uint64_t configHash;                 // key layout + accumulator kinds: a state only merges into an operator planned identically
    uint64_t pad[4];
};
static_assert(sizeof(AggStateHeader) == 64, "wire header is 64 bytes");
constexpr uint32_t kAggStateMagic = 0x50474153;   // "PGAS"
struct AggStateView {
    AggStateHeader* header;
    uint64_t* keys;       // [W][maxGroups]
    uint64_t* firstRow;   // [maxGroups]
    uint64_t* special;    // [maxGroups] 0 ordinary key, 1 the Bigint path's NULL key, 2 the key equal to the W == 1 sentinel
    uint64_t* acc;        // [maxGroups][nvPhys]
    uint32_t maxGroups;
};
static size_t aggStateBytes(int W, uint32_t nvPhys, uint64_t maxGroups) { return sizeof(AggStateHeader) + (size_t) maxGroups * 8 * ((size_t) W + 2 + nvPhys); }
static AggStateView aggStateView(void* buffer, int W, uint32_t nvPhys, uint32_t maxGroups)
{
    AggStateView v;
    char* p = (char*) buffer;
    v.header = (AggStateHeader*) p; p += sizeof(AggStateHeader);
    v.keys = (uint64_t*) p; p += (size_t) maxGroups * 8 * W;
    v.firstRow = (uint64_t*) p; p += (size_t) maxGroups * 8;
    v.special = (uint64_t*) p; p += (size_t) maxGroups * 8;
    v.acc = (uint64_t*) p;
    v.maxGroups = maxGroups;
    return v;
}

// inactiveMask: lazily activated non-null counts that never saw a NULL hold nothing yet; their value is the group's row count
// (column countSrc), or 1 when no row count is kept (then only "> 0" is ever asked of them) — backfillCountKernel's rule
template <int W>
__global__ void exportStateKernel(GroupTable T, AggStateView out, uint64_t configHash, uint32_t inactiveMask, int countSrc)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out.header->magic = kAggStateMagic; out.header->W = W; out.header->nvPhys = T.nvPhys; out.header->maxGroups = out.maxGroups;
        out.header->configHash = configHash;
    }
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < T.capP; s += gridDim.x * blockDim.x) {
        const uint64_t f = T.firstRow[s];
        if (f == kNoRow) continue;
        const uint32_t i = atomicAdd(&out.header->nGroups, 1u);
        if (i >= out.maxGroups) { out.header->overflow = 1; continue; }
#pragma unroll
        for (int w = 0; w < W; w++) out.keys[(size_t) w * out.maxGroups + i] = s < T.cap ? T.keys[(size_t) w * T.capP + s] : 0;
        out.firstRow[i] = f;
        out.special[i] = s < T.cap ? 0 : (uint64_t) (s - T.cap + 1);
        for (uint32_t c = 0; c < T.nvPhys; c++) {
            uint64_t v = T.acc[(size_t) s * T.nvPhys + c];
            if ((inactiveMask >> c) & 1u) v = countSrc >= 0 ? T.acc[(size_t) s * T.nvPhys + countSrc] : 1;
            out.acc[(size_t) i * T.nvPhys + c] = v;
        }
    }
}

// One source state into the table: every group occurs once per state and the states are merged by consecutive launches, so each
// accumulator receives its contributions in a fixed order (source 0, 1, 2 ...: DOUBLE sums are reproducible).  `orderTag` goes on top
// of the first-row words so that first-appearance order across sources is (source, row) — the order the FINAL operator would see the
// sources' pages in.
template <int W>
__global__ void mergeStateKernel(GroupTable T, AggStateView in, uint32_t nGroups, uint64_t orderTag, uint32_t floatMask, uint32_t maxMask)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nGroups; i += gridDim.x * blockDim.x) {
        const uint64_t special = in.special[i];
        int slot;
        if (special) slot = (int) T.cap + (int) special - 1;
        else {
            uint64_t kw[W];
#pragma unroll
            for (int w = 0; w < W; w++) kw[w] = in.keys[(size_t) w * in.maxGroups + i];
            slot = tableFind<W>(T, kw, hashKeyWords<W>(kw), true, true);
            if (slot < 0) { atomicOr(&T.counters[CNT_ERROR], 0x80000000u); continue; }
        }
        const uint64_t f = in.firstRow[i] | orderTag;
        if (f < T.firstRow[slot]) T.firstRow[slot] = f;
        for (uint32_t c = 0; c < T.nvPhys; c++) {
            uint64_t* p = &T.acc[(size_t) slot * T.nvPhys + c];
            const int op = ((maxMask >> c) & 1u) ? OP_MAX_U64 : (((floatMask >> c) & 1u) ? OP_ADD_F64 : OP_ADD_I64);
            *p = combineAcc(*p, in.acc[(size_t) i * T.nvPhys + c], op);
        }
    }
}

__global__ void __launch_bounds__(256) collectSlotsKernel(GroupTable t, uint64_t* sortKeys, uint32_t* sortVals, unsigned int* count /* [0] = slots, [1..2] = OR of the first-row words */)
{
    __shared__ unsigned int warpCount[8], blockBase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long seen = 0;
    const uint32_t rounds = (t.capP + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (uint32_t k = 0; k < rounds; k++) {
        const uint32_t s = k * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
        const uint64_t f = s < t.capP ? t.firstRow[s] : kNoRow;
        const bool live = f != kNoRow;
        // one counter increment per CTA and round (a million groups would otherwise queue on this single address)
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (lane == 0) warpCount[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int total = 0;
            for (int i = 0; i < 8; i++) { const unsigned int c = warpCount[i]; warpCount[i] = total; total += c; }
            blockBase = total ? atomicAdd(count, total) : 0;
        }
        __syncthreads();
        if (live) {
            const unsigned int i = blockBase + warpCount[warp] + __popc(m & ((1u << lane) - 1));
            sortKeys[i] = f;
            sortVals[i] = s;
            seen |= f;
        }
        __syncthreads();
    }
    // which bits the sort has to look at
    for (int o = 16; o >= 1; o >>= 1) seen |= __shfl_xor_sync(0xffffffffu, seen, o);
    if (lane == 0 && seen) { atomicOr(&count[1], (unsigned int) seen); atomicOr(&count[2], (unsigned int) (seen >> 32)); }
}

// ---- output ------------------------------------------------------------------------------------------------------------
struct KeyOut {          // one output key column
    void* values; uint8_t* nulls; int32_t* lens;   // VARCHAR: lens[g] first, bytes in a second pass
    int32_t type, word, shift, bits;
    const uint64_t* internMeta; const uint64_t* internHash;     // interned VARCHAR key (varchar_interner.h): the field is the value's id
};
struct AggOut {
    int32_t function;      // pg_aggregate_function
    int32_t val0, val1;    // physical value columns (sum; lo/hi for bigint sum)
    int32_t val2, val3;    // sum(decimal) over 128-bit inputs: limbs 2 and 3 (-1: a short decimal, lo/hi only)
    int32_t countCol;      // physical non-null-count column, or -1 = every touched group has >= 1 non-null input
    int32_t partial;       // PARTIAL step: emit intermediate columns
    void* out0; uint8_t* null0;   // first output column
    void* out1;                   // second output column (avg partial: count; sum(decimal) partial: the state block's offsets)
};
struct EmitParams {
    GroupTable t;
    const uint32_t* order;   // slots in group-id order
    int32_t nGroups;
    int32_t W;
    KeyOut keys[kMaxKeyColumns]; int32_t nKeys; int32_t bigintPath; int32_t nullWord, nullShift;
    int64_t* hashOut;        // $hashvalue output or nullptr
    AggOut aggs[32]; int32_t nAggs;
};

__device__ inline uint64_t fieldOf(const uint64_t* kw, int word, int shift, int bits)
{
    uint64_t v = kw[word] >> shift;
    return bits >= 64 ? v : (v & ((1ULL << bits) - 1));
}

__global__ void emitKernel(const __grid_constant__ EmitParams P)
{
    const GroupTable& T = P.t;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < P.nGroups; g += gridDim.x * blockDim.x) {
        const uint32_t slot = P.order[g];
        uint64_t kw[kMaxKeyWords] = {0, 0, 0, 0};
        bool bigintNull = false;
        if (slot >= T.cap) { if (slot == T.cap) bigintNull = true; else kw[0] = kEmptyKey; }
        else for (int w = 0; w < P.W; w++) kw[w] = T.keys[(size_t) w * T.capP + slot];
        const uint32_t nullMask = P.bigintPath ? (bigintNull ? 1u : 0u) : (uint32_t) fieldOf(kw, P.nullWord, P.nullShift, 8);
        const uint32_t spelling = (uint32_t) T.firstRow[slot] & ((1u << kFirstRowShift) - 1);   // -0.0 spelling of the group's first row
        int64_t rawHash = 0;
        for (int k = 0; k < P.nKeys; k++) {
            const KeyOut& ko = P.keys[k];
            const bool isNull = (nullMask >> k) & 1u;
            uint64_t v = isNull ? 0 : fieldOf(kw, ko.word, ko.shift, ko.bits);
            int64_t th = 0;   // TypeUtils.hashPosition: NULL → 0
            switch (ko.type) {
                case PG_TYPE_BIGINT: ((int64_t*) ko.values)[g] = (int64_t) v; if (!isNull) th = hashLong((int64_t) v); break;
                case PG_TYPE_DOUBLE:
                    if (!isNull) th = hashLong(doubleHashBits((int64_t) v));
                    if (!isNull && v == 0 && ((spelling >> k) & 1u)) v = 0x8000000000000000ULL;     // the first row of the group spelled it -0.0
                    ((int64_t*) ko.values)[g] = (int64_t) v;
                    break;
                case PG_TYPE_INTEGER: case PG_TYPE_DATE: ((int32_t*) ko.values)[g] = (int32_t) v; if (!isNull) th = hashLong((int64_t) (int32_t) v); break;
                case PG_TYPE_BOOLEAN: ((uint8_t*) ko.values)[g] = (uint8_t) v; if (!isNull) th = v ? 1231 : 1237; break;
                case PG_TYPE_VARCHAR: {
                    if (ko.internMeta) {         // id → length now, bytes gathered from the interner's pool after the scan of the lengths
                        ko.lens[g] = isNull ? 0 : (int32_t) internLengthOf(ko.internMeta[(uint32_t) v]);
                        ((int32_t*) ko.values)[g] = (int32_t) v;
                        if (!isNull) th = (int64_t) ko.internHash[(uint32_t) v];
                        break;
                    }
                    int len = isNull ? 0 : (int) (v & 0xff);
                    ko.lens[g] = len;
                    ((uint64_t*) ko.values)[g] = v >> 8;     // packed bytes, expanded by expandVarcharKernel
                    if (!isNull) th = xxHash64Short(v >> 8, len);
                    break;
                }
            }
            ko.nulls[g] = isNull ? 1 : 0;
            rawHash = (int64_t) (31ULL * (uint64_t) rawHash + (uint64_t) th);
        }
        if (P.hashOut) P.hashOut[g] = P.bigintPath ? (bigintNull ? 0 : hashLong((int64_t) kw[0])) : rawHash;

        const uint64_t* acc = T.acc + (size_t) slot * T.nvPhys;
        for (int a = 0; a < P.nAggs; a++) {
            const AggOut& ao = P.aggs[a];
            const int64_t count = ao.countCol >= 0 ? (int64_t) acc[ao.countCol] : 1;
            switch (ao.function) {
                case PG_AGG_COUNT_STAR: case PG_AGG_COUNT_COLUMN:
                    ((int64_t*) ao.out0)[g] = (int64_t) acc[ao.val0]; ao.null0[g] = 0; break;
                case PG_AGG_SUM_DOUBLE:
                    ((uint64_t*) ao.out0)[g] = count > 0 ? acc[ao.val0] : 0; ao.null0[g] = count > 0 ? 0 : 1; break;
                case PG_AGG_SUM_BIGINT: {
                    __int128 total = ((__int128) (int64_t) acc[ao.val1] << 32) + (__int128) acc[ao.val0];
                    if (total > (__int128) INT64_MAX || total < (__int128) INT64_MIN) { atomicOr(&T.counters[CNT_ERROR], (unsigned int) ERR_BIGINT_OVERFLOW); total = 0; }
                    ((int64_t*) ao.out0)[g] = count > 0 ? (int64_t) total : 0; ao.null0[g] = count > 0 ? 0 : 1; break;
                }
                case PG_AGG_SUM_SHORT_DECIMAL: case PG_AGG_SUM_LONG_DECIMAL: {
                    // exact total of the limb sums in 192 bits (t2 : t1 : t0, t2 signed); DecimalSumAggregation keeps 127 bits of
                    // magnitude + an overflow count, i.e. the same number as overflow * 2^127 + value
                    unsigned __int128 lowPart = (unsigned __int128) acc[ao.val0];
                    __int128 top = 0;                 // multiples of 2^64
                    auto addShifted = [&](__int128 v, int shift) {        // += v << shift   (shift = 32 or 96 handled by splitting)
                        // v fits 65 bits (a signed / unsigned 64-bit sum)
                        const unsigned __int128 lo = (unsigned __int128) v << shift;          // low 128 bits of v << shift
                        const __int128 hi = v >> (128 - shift);                                // what fell off the top (arithmetic)
                        const unsigned __int128 before = lowPart;
                        lowPart += lo;
                        top += hi + (lowPart < before ? 1 : 0);
                    };
                    // `top` counts multiples of 2^128 here
                    if (ao.val2 < 0) addShifted((__int128) (int64_t) acc[ao.val1], 32);
                    else {
                        addShifted((__int128) (unsigned __int128) acc[ao.val1], 32);
                        addShifted((__int128) (unsigned __int128) acc[ao.val2], 64);
                        addShifted((__int128) (int64_t) acc[ao.val3], 96);
                    }
                    // value = top * 2^128 + lowPart (two's complement over 192+ bits): magnitude and sign
                    bool negative = top < 0;
                    unsigned __int128 magLow = lowPart; __int128 magTop = top;
                    if (negative) { magLow = ~lowPart + 1; magTop = ~top + (magLow == 0 ? 1 : 0); }
                    const uint64_t mLow = (uint64_t) magLow, mHigh = (uint64_t) (magLow >> 64);
                    const bool isNull = count <= 0;
                    if (ao.partial) {
                        // {overflow, low, high}: magnitude split at 2^127 (any split with |value| < 2^127 is the same number)
                        const __int128 over = (magTop << 1) | (__int128) (mHigh >> 63);
                        int64_t* st = (int64_t*) ao.out0 + 3 * (size_t) g;
                        st[0] = isNull ? 0 : (int64_t) (negative ? -over : over);
                        st[1] = isNull ? 0 : (int64_t) mLow;
                        st[2] = isNull ? 0 : (int64_t) ((mHigh & 0x7fffffffffffffffULL) | (negative && (mLow | (mHigh & 0x7fffffffffffffffULL)) ? 0x8000000000000000ULL : 0));
                        int32_t* offs = (int32_t*) ao.out1;
                        offs[g + 1] = 24 * (g + 1);
                        if (g == 0) offs[0] = 0;
                        ao.null0[g] = isNull ? 1 : 0;
                    }
                    else {
                        // outputLongDecimal: overflow != 0 or |value| >= 10^38 -> "Decimal overflow"
                        const bool tooBig = magTop != 0 || mHigh > 0x4b3b4ca85a86c47aULL || (mHigh == 0x4b3b4ca85a86c47aULL && mLow >= 0x098a224000000000ULL);
                        if (!isNull && tooBig) atomicOr(&T.counters[CNT_ERROR], (unsigned int) ERR_DECIMAL_OVERFLOW);
                        uint64_t* o = (uint64_t*) ao.out0 + 2 * (size_t) g;
                        o[0] = isNull || tooBig ? 0 : mLow;
                        o[1] = isNull || tooBig ? 0 : (mHigh | (negative && (mLow | mHigh) ? 0x8000000000000000ULL : 0));
                        ao.null0[g] = isNull ? 1 : 0;
                    }
                    break;
                }
                case PG_AGG_MIN_DOUBLE: case PG_AGG_MAX_DOUBLE: case PG_AGG_MIN_BIGINT: case PG_AGG_MAX_BIGINT: {
                    const bool isMin = ao.function == PG_AGG_MIN_DOUBLE || ao.function == PG_AGG_MIN_BIGINT;
                    const uint64_t o = isMin ? ~acc[ao.val0] : acc[ao.val0];
                    const uint64_t v = (ao.function == PG_AGG_MIN_DOUBLE || ao.function == PG_AGG_MAX_DOUBLE) ? f64OfOrdered(o) : i64OfOrdered(o);
                    ((uint64_t*) ao.out0)[g] = count > 0 ? v : 0; ao.null0[g] = count > 0 ? 0 : 1; break;
                }
                case PG_AGG_AVG_DOUBLE: case PG_AGG_AVG_BIGINT:
                    if (ao.partial) { ((uint64_t*) ao.out0)[g] = acc[ao.val0]; ao.null0[g] = 0; ((int64_t*) ao.out1)[g] = count; }
                    else if (count == 0) { ((uint64_t*) ao.out0)[g] = 0; ao.null0[g] = 1; }
                    else { ((double*) ao.out0)[g] = __longlong_as_double((long long) acc[ao.val0]) / (double) count; ao.null0[g] = 0; }
                    break;
            }
        }
    }
}

__global__ void expandVarcharKernel(const uint64_t* packed, const int32_t* offsets, int32_t n, uint8_t* bytes)
{
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        int32_t o = offsets[g], len = offsets[g + 1] - o;
        uint64_t v = packed[g];
        for (int i = 0; i < len; i++) bytes[o + i] = (uint8_t) (v >> (8 * i));
    }
}

// ======================================================================================================================
// Host side
// ======================================================================================================================
struct PhysCol { int kind; int channel; int mask; bool lazy; bool active; };

struct AggPlan {
    pg_aggregate spec;
    int val0 = -1, val1 = -1, val2 = -1, val3 = -1;
    int countCol = -1;        // physical column with the non-null count of (channel, mask); -1 never needed
    bool countExact = false;  // count value is an output (avg / count) rather than a null flag (sum)
};

class GroupTableStorage {
public:
    DeviceBuffer keys, state, firstRow, acc, gid, counters;
    uint32_t cap = 0; int W = 1; uint32_t nvPhys = 0; bool withGid = false;
    uint32_t flushReserve = 0;
    Operator* op = nullptr;

    GroupTable view() const
    {
        GroupTable t{};
        t.keys = keys.as<uint64_t>(); t.state = W > 1 ? state.as<uint32_t>() : nullptr; t.firstRow = firstRow.as<uint64_t>();
        t.acc = acc.as<uint64_t>(); t.gid = withGid ? gid.as<int32_t>() : nullptr;
        t.cap = cap; t.mask = cap - 1; t.capP = cap + 2; t.nvPhys = nvPhys;
        uint32_t lim = cap / 4 * 3;
        t.rowClaimLimit = lim;
        t.counters = counters.as<unsigned int>();
        return t;
    }
    void allocate(uint32_t capacity)
    {
        cap = capacity;
        size_t capP = (size_t) cap + 2;
        for (DeviceBuffer* b : {&keys, &state, &firstRow, &acc, &gid}) b->accounting = &op->stats.device_bytes;
        keys.reserve(capP * 8 * W);
        if (W > 1) state.reserve(capP * 4);
        firstRow.reserve(capP * 8);
        acc.reserve(std::max<size_t>(capP * 8 * nvPhys, 8));
        if (withGid) gid.reserve(capP * 4);
        cudaStream_t s = op->stream;
        if (W == 1) { fillU64Kernel<<<1024, 256, 0, s>>>(keys.as<uint64_t>(), kEmptyKey, capP); op->stats.kernel_launches++; }
        else { PG_CUDA(cudaMemsetAsync(keys.ptr, 0, capP * 8 * W, s)); PG_CUDA(cudaMemsetAsync(state.ptr, 0, capP * 4, s)); }
        PG_CUDA(cudaMemsetAsync(firstRow.ptr, 0xFF, capP * 8, s));
        if (nvPhys) PG_CUDA(cudaMemsetAsync(acc.ptr, 0, capP * 8 * nvPhys, s));
        if (withGid) PG_CUDA(cudaMemsetAsync(gid.ptr, 0xFF, capP * 4, s));
        PG_CUDA(cudaGetLastError());
    }
};

static uint32_t nextPow2(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return (uint32_t) std::min<uint64_t>(p, 1ULL << 31); }


// ---- result materialisation shared by HashAggregationOperator and GroupByHash ---------------------------------------------
struct AggEmit { int function, val0, val1, countCol; bool partial; int val2 = -1, val3 = -1; };

// Slots that were touched, ordered by first-touch row = ascending group id.  Returns the group count.
// `scratch` keeps the sort's temporaries alive until the caller's next synchronisation (blocks go back to the caching allocator only
// after the stream that used them was synchronised): no sync is needed here beyond reading the group count.
static int64_t sortSlotsByFirstRow(Operator* op, const GroupTable& T, DeviceBuffer& order, std::vector<DeviceBuffer>& scratch)
{
    const size_t capP = T.capP;
    scratch.clear(); scratch.resize(5);
    DeviceBuffer& sortKeys = scratch[0]; DeviceBuffer& sortVals = scratch[1]; DeviceBuffer& sortKeysOut = scratch[2]; DeviceBuffer& count = scratch[3]; DeviceBuffer& temp = scratch[4];
    sortKeys.reserve(capP * 8); sortVals.reserve(capP * 4); count.reserve(16);
    PG_CUDA(cudaMemsetAsync(count.ptr, 0, 16, op->stream));
    PG_LAUNCH(op, collectSlotsKernel, 1024, 256, 0, T, sortKeys.as<uint64_t>(), sortVals.as<uint32_t>(), count.as<unsigned int>());
    unsigned int h[3] = {0, 0, 0};
    PG_CUDA(cudaMemcpyAsync(h, count.ptr, 12, cudaMemcpyDeviceToHost, op->stream));
    PG_CUDA(cudaStreamSynchronize(op->stream));
    const unsigned int g = h[0];
    if (g == 0) return 0;
    // radix passes only over the bits some first-row word has set (600 M rows: 38 of the 64)
    const uint64_t seen = ((uint64_t) h[2] << 32) | h[1];
    int endBit = 1;
    while (endBit < 64 && (seen >> endBit) != 0) endBit++;
    sortKeysOut.reserve((size_t) g * 8); order.reserve((size_t) g * 4);
    size_t tempBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, sortKeys.as<uint64_t>(), sortKeysOut.as<uint64_t>(), sortVals.as<uint32_t>(), order.as<uint32_t>(), (int) g, 0, endBit, op->stream);
    temp.reserve(tempBytes);
    cub::DeviceRadixSort::SortPairs(temp.ptr, tempBytes, sortKeys.as<uint64_t>(), sortKeysOut.as<uint64_t>(), sortVals.as<uint32_t>(), order.as<uint32_t>(), (int) g, 0, endBit, op->stream);
    op->stats.kernel_launches += 4;
    return g;
}

struct ResultBuilder {
    struct Buf {
        void* ptr = nullptr;
        template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
    };
    std::vector<Buf> outBufs;
    std::vector<DeviceBuffer> owned;
    std::vector<PinnedBuffer> outHost;
    // Small results bound for the host (a partial aggregation's few groups, a Q1 answer) are built inside ONE device arena and cross
    // PCIe in one copy with the error counters: one synchronisation instead of two, one copy instead of two or three per column.
    static constexpr size_t kArenaBytes = 1 << 20;
    DeviceBuffer arena;
    PinnedBuffer arenaHost;
    std::vector<pg_block> outBlocks, outSlice;
    std::vector<int> outTypes;

    // Output schema: group keys [, $hashvalue], aggregates (HashAggregationOperator.java:357 toTypes), `g` rows in `order`.
    void build(Operator* op, const GroupTable& T, const KeyLayout& layout, const uint32_t* order, int64_t g, bool withHash, const std::vector<AggEmit>& aggs,
               const std::vector<std::unique_ptr<VarcharInterner>>* interners = nullptr)
    {
        cudaStream_t stream = op->stream;
        EmitParams E{};
        E.t = T; E.order = order; E.nGroups = (int) g; E.W = layout.W;
        E.nKeys = layout.nKeys; E.bigintPath = layout.bigintPath; E.nullWord = layout.nullWord < 0 ? 0 : layout.nullWord; E.nullShift = layout.nullShift;
        outTypes.clear(); outBufs.clear(); outBufs.reserve(64); owned.clear(); owned.reserve(64);
        struct ColBufs { int values, nulls; };
        std::vector<ColBufs> colBufs;
        const bool toHost = op->options.output_memory == PG_MEM_HOST;
        auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
        size_t scanTemp = 0, bound = 0, arenaUsed = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scanTemp, (const int32_t*) nullptr, (int32_t*) nullptr, (int) g + 1, stream);
        {
            size_t nCols = (size_t) layout.nKeys + (withHash ? 1 : 0), nVarchar = 0;
            for (const AggEmit& a : aggs) nCols += (a.partial && (a.function == PG_AGG_AVG_DOUBLE || a.function == PG_AGG_AVG_BIGINT)) ? 2 : 1;
            for (int k = 0; k < layout.nKeys; k++) nVarchar += layout.types[k] == PG_TYPE_VARCHAR;
            size_t nDecimal = 0;
            for (const AggEmit& a : aggs) nDecimal += a.function == PG_AGG_SUM_SHORT_DECIMAL || a.function == PG_AGG_SUM_LONG_DECIMAL;
            bound = nCols * (align((size_t) g * 8 + 16) + align((size_t) g + 16)) + nVarchar * (2 * align((size_t) (g + 1) * 4) + align(scanTemp + 16) + align((size_t) g * 8 + 16))
                  + nDecimal * (align((size_t) g * 24 + 16) + align((size_t) (g + 1) * 4 + 16));
        }
        const bool arenaMode = toHost && g > 0 && bound <= kArenaBytes && !layout.anyInterned();   // interned keys: byte totals are only known after the scan
        if (arenaMode) {
            arena.accounting = &op->stats.device_bytes;
            arena.reserve(bound);
            PG_CUDA(cudaMemsetAsync(arena.ptr, 0, bound, stream));      // every null vector / length column starts at zero
        }
        auto newBuf = [&](size_t bytes) {
            bytes = std::max<size_t>(bytes, 16);
            Buf b;
            if (arenaMode) {
                PG_CHECK_STATE(arenaUsed + align(bytes) <= bound, "result arena bound too small");
                b.ptr = (char*) arena.ptr + arenaUsed; arenaUsed += align(bytes);
            }
            else { owned.emplace_back(); owned.back().accounting = &op->stats.device_bytes; owned.back().reserve(bytes); b.ptr = owned.back().ptr; }
            outBufs.push_back(b);
            return (int) outBufs.size() - 1;
        };
        auto zero = [&](void* p, size_t bytes) { if (!arenaMode) PG_CUDA(cudaMemsetAsync(p, 0, bytes, stream)); };
        auto newColumn = [&](int type) {
            ColBufs cb{-1, -1};
            int w = typeWidth(type);
            cb.values = newBuf((size_t) g * (w ? w : 8));
            cb.nulls = newBuf(g);
            outTypes.push_back(type); colBufs.push_back(cb);
            return (int) colBufs.size() - 1;
        };
        std::vector<int> varcharLens(layout.nKeys, -1);
        std::map<int, int> stateOffsets;          // output column -> buffer with the offsets of its VARBINARY states
        for (int k = 0; k < layout.nKeys; k++) {
            int c = newColumn(layout.types[k]);
            KeyOut& ko = E.keys[k];
            ko.values = outBufs[colBufs[c].values].ptr; ko.nulls = outBufs[colBufs[c].nulls].as<uint8_t>(); ko.lens = nullptr;
            ko.type = layout.types[k]; ko.word = layout.word[k]; ko.shift = layout.shift[k]; ko.bits = layout.bits[k];
            ko.internMeta = nullptr; ko.internHash = nullptr;
            if (layout.interned[k]) {
                PG_CHECK_STATE(interners && (*interners)[k], "internal: interned key without its interner");
                const InternView iv = (*interners)[k]->view();
                ko.internMeta = iv.entryMeta; ko.internHash = iv.entryHash;
            }
            if (layout.types[k] == PG_TYPE_VARCHAR) {
                varcharLens[k] = newBuf((size_t) (g + 1) * 4);
                zero(outBufs[varcharLens[k]].ptr, (size_t) (g + 1) * 4);
                ko.lens = outBufs[varcharLens[k]].as<int32_t>();
            }
        }
        if (withHash) { int c = newColumn(PG_TYPE_BIGINT); E.hashOut = outBufs[colBufs[c].values].as<int64_t>(); zero(outBufs[colBufs[c].nulls].ptr, g); }
        PG_CHECK_ARG(aggs.size() <= 32, "at most 32 aggregates");
        E.nAggs = (int) aggs.size();
        for (size_t i = 0; i < aggs.size(); i++) {
            const AggEmit& a = aggs[i];
            AggOut& ao = E.aggs[i];
            ao.function = a.function; ao.val0 = a.val0; ao.val1 = a.val1; ao.val2 = a.val2; ao.val3 = a.val3; ao.partial = a.partial ? 1 : 0; ao.countCol = a.countCol;
            if (a.function == PG_AGG_SUM_SHORT_DECIMAL || a.function == PG_AGG_SUM_LONG_DECIMAL) {
                if (!a.partial) {
                    int c = newColumn(PG_TYPE_LONG_DECIMAL);
                    ao.out0 = outBufs[colBufs[c].values].ptr; ao.null0 = outBufs[colBufs[c].nulls].as<uint8_t>(); ao.out1 = nullptr;
                }
                else {      // one 24-byte state per group in a VARBINARY column (null groups keep their 24 zero bytes: offsets stay 24 * g)
                    ColBufs cb{newBuf((size_t) g * 24), newBuf(g)};
                    outTypes.push_back(PG_TYPE_VARCHAR); colBufs.push_back(cb);
                    const int offs = newBuf((size_t) (g + 1) * 4);
                    stateOffsets[(int) colBufs.size() - 1] = offs;
                    ao.out0 = outBufs[cb.values].ptr; ao.null0 = outBufs[cb.nulls].as<uint8_t>(); ao.out1 = outBufs[offs].ptr;
                }
                continue;
            }
            int type = (a.function == PG_AGG_SUM_DOUBLE || a.function == PG_AGG_AVG_DOUBLE || a.function == PG_AGG_AVG_BIGINT || a.function == PG_AGG_MIN_DOUBLE
                        || a.function == PG_AGG_MAX_DOUBLE) ? PG_TYPE_DOUBLE : PG_TYPE_BIGINT;
            int c = newColumn(type);
            ao.out0 = outBufs[colBufs[c].values].ptr; ao.null0 = outBufs[colBufs[c].nulls].as<uint8_t>(); ao.out1 = nullptr;
            if (a.partial && (a.function == PG_AGG_AVG_DOUBLE || a.function == PG_AGG_AVG_BIGINT)) {
                int c2 = newColumn(PG_TYPE_BIGINT);
                ao.out1 = outBufs[colBufs[c2].values].ptr;
                zero(outBufs[colBufs[c2].nulls].ptr, g);
            }
        }
        PG_LAUNCH(op, emitKernel, (int) std::min<int64_t>(divUp(g, 256), 2048), 256, 0, E);
        // VARCHAR keys: lens → offsets (exclusive scan) → bytes
        std::vector<int> varcharBytes(layout.nKeys, -1), varcharOffsets(layout.nKeys, -1);
        for (int k = 0; k < layout.nKeys; k++) {
            if (layout.types[k] != PG_TYPE_VARCHAR) continue;
            int offs = newBuf((size_t) (g + 1) * 4);
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, outBufs[varcharLens[k]].as<int32_t>(), outBufs[offs].as<int32_t>(), (int) g + 1, stream);
            const int t2 = newBuf(tb);       // (lives in outBufs until the next build: no synchronisation needed to release it)
            cub::DeviceScan::ExclusiveSum(outBufs[t2].ptr, tb, outBufs[varcharLens[k]].as<int32_t>(), outBufs[offs].as<int32_t>(), (int) g + 1, stream);
            op->stats.kernel_launches += 2;
            if (layout.interned[k]) {
                int32_t total = 0;
                PG_CUDA(cudaMemcpyAsync(&total, outBufs[offs].as<int32_t>() + g, 4, cudaMemcpyDeviceToHost, stream));
                PG_CUDA(cudaStreamSynchronize(stream));
                if (total < 0) throw Error(PG_ERR_INSUFFICIENT_RESOURCES, "VARCHAR key column of the result exceeds 2 GB");
                const int bytesBuf = newBuf((size_t) total);
                (*interners)[k]->gatherBytes(op, (const int32_t*) outBufs[colBufs[k].values].ptr, outBufs[colBufs[k].nulls].as<uint8_t>(), outBufs[offs].as<int32_t>(), g,
                                             outBufs[bytesBuf].as<uint8_t>());
                varcharBytes[k] = bytesBuf; varcharOffsets[k] = offs;
                continue;
            }
            int bytes = newBuf((size_t) g * 8);
            PG_LAUNCH(op, expandVarcharKernel, (int) std::min<int64_t>(divUp(g, 256), 2048), 256, 0, (const uint64_t*) outBufs[colBufs[k].values].ptr, outBufs[offs].as<int32_t>(), (int32_t) g, outBufs[bytes].as<uint8_t>());
            varcharBytes[k] = bytes; varcharOffsets[k] = offs;
        }
        unsigned int cnt[CNT_COUNT];
        PG_CUDA(cudaMemcpyAsync(cnt, T.counters, sizeof(cnt), cudaMemcpyDeviceToHost, stream));

        // describe the full result; slice() hands out row ranges of it
        outBlocks.assign(colBufs.size(), pg_block{});
        outHost.clear();
        outHost.reserve(colBufs.size() * 3 + 4);
        if (arenaMode) {
            arenaHost.reserve(arenaUsed + 16);
            PG_CUDA(cudaMemcpyAsync(arenaHost.ptr, arena.ptr, arenaUsed, cudaMemcpyDeviceToHost, stream));
            op->stats.d2h_bytes += (int64_t) arenaUsed;
        }
        auto place = [&](const void* dptr, size_t bytes) -> const void* {
            if (!toHost) return dptr;
            if (arenaMode) return (const char*) arenaHost.ptr + ((const char*) dptr - (const char*) arena.ptr);
            outHost.emplace_back();
            outHost.back().reserve(std::max<size_t>(bytes, 16));
            if (bytes) PG_CUDA(cudaMemcpyAsync(outHost.back().ptr, dptr, bytes, cudaMemcpyDeviceToHost, stream));
            op->stats.d2h_bytes += (int64_t) bytes;
            return outHost.back().ptr;
        };
        std::vector<int32_t> totalBytes(layout.nKeys, 0);
        for (int k = 0; k < layout.nKeys; k++)
            if (varcharOffsets[k] >= 0 && !arenaMode) PG_CUDA(cudaMemcpyAsync(&totalBytes[k], outBufs[varcharOffsets[k]].as<int32_t>() + g, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));      // the error counters and the VARCHAR byte totals (arena: the whole result) arrive together
        if (cnt[CNT_ERROR] & ERR_BIGINT_OVERFLOW) throw Error(PG_ERR_NUMERIC_OVERFLOW, "bigint addition overflow");
        if (cnt[CNT_ERROR] & ERR_DECIMAL_OVERFLOW) throw Error(PG_ERR_NUMERIC_OVERFLOW, "Decimal overflow");
        for (size_t c = 0; c < colBufs.size(); c++) {
            pg_block& b = outBlocks[c];
            b.type = outTypes[c]; b.encoding = PG_ENC_FLAT; b.position_count = g; b.dictionary_size = 0; b.ids = nullptr; b.offsets = nullptr;
            if (outTypes[c] == PG_TYPE_VARCHAR && stateOffsets.count((int) c)) {
                b.values = place(outBufs[colBufs[c].values].ptr, (size_t) g * 24);
                b.offsets = (const int32_t*) place(outBufs[stateOffsets[(int) c]].ptr, (size_t) (g + 1) * 4);
            }
            else if (outTypes[c] == PG_TYPE_VARCHAR) {
                int k = (int) c;   // key columns come first
                b.values = place(outBufs[varcharBytes[k]].ptr, (size_t) totalBytes[k]);
                b.offsets = (const int32_t*) place(outBufs[varcharOffsets[k]].ptr, (size_t) (g + 1) * 4);
            }
            else b.values = place(outBufs[colBufs[c].values].ptr, (size_t) g * typeWidth(outTypes[c]));
            b.is_null = (const uint8_t*) place(outBufs[colBufs[c].nulls].ptr, g);
        }
        if (!arenaMode) PG_CUDA(cudaStreamSynchronize(stream));
    }

    void slice(int64_t begin, int64_t m, int memory, pg_page* out)
    {
        outSlice = outBlocks;
        for (auto& b : outSlice) {
            b.position_count = m;
            if (b.type == PG_TYPE_VARCHAR) b.offsets += begin;   // offsets stay absolute into the shared byte slice
            else b.values = (const char*) b.values + begin * typeWidth(b.type);
            b.is_null += begin;
        }
        out->position_count = m; out->channel_count = (int32_t) outSlice.size(); out->memory = memory; out->blocks = outSlice.data();
    }
};

// GroupByHash.createGroupByHash + the accumulator factories of one HashAggregationOperator, planned on the host: the packed key layout
// and the physical accumulator columns (one per distinct (kind, channel, mask)) every aggregate reads.  Needs no device.
static void planAggregation(const pg_hash_aggregation_config* c, KeyLayout& layout, std::vector<AggPlan>& aggs, std::vector<PhysCol>& phys, int& rowCountCol)
{
    auto physColumn = [&](int kind, int channel, int mask, bool lazy) {
        for (size_t i = 0; i < phys.size(); i++)
            if (phys[i].kind == kind && phys[i].channel == channel && phys[i].mask == mask) return (int) i;
        phys.push_back({kind, channel, mask, lazy, !lazy});
        return (int) phys.size() - 1;
    };
    PG_CHECK_ARG(c->expected_groups > 0, "expectedSize must be greater than zero");
    PG_CHECK_ARG(c->n_aggregates >= 0 && c->n_aggregates <= 32, "at most 32 aggregates");
    layout.plan(c->n_group_channels, c->group_channels, c->group_types, c->group_max_lengths);
    const int step = c->step;
    PG_CHECK_ARG(step == PG_STEP_SINGLE || step == PG_STEP_PARTIAL || step == PG_STEP_FINAL, "bad aggregation step");
    const bool final = step == PG_STEP_FINAL;
    for (int i = 0; i < c->n_aggregates; i++) {
        AggPlan a; a.spec = c->aggregates[i];
        const int ch = a.spec.input_channel, mk = a.spec.mask_channel;
        PG_CHECK_ARG(!(final && mk >= 0), "FINAL step takes no mask");
        switch (a.spec.function) {
            case PG_AGG_COUNT_STAR:
                if (final) a.val0 = physColumn(K_I64_ADD, ch, -1, false);
                else a.val0 = physColumn(K_COUNT, -1, mk, false);
                break;
            case PG_AGG_COUNT_COLUMN:
                if (final) a.val0 = physColumn(K_I64_ADD, ch, -1, false);
                else { PG_CHECK_ARG(ch >= 0, "count(column) needs an input channel"); a.val0 = physColumn(K_COUNT, ch, mk, false); }
                break;
            case PG_AGG_SUM_DOUBLE:
                a.val0 = physColumn(K_F64, ch, mk, false);
                a.countCol = physColumn(K_COUNT, ch, mk, mk < 0);   // lazily activated: only once a page carries NULLs in `ch`
                break;
            case PG_AGG_SUM_BIGINT:
                a.val0 = physColumn(K_LO32, ch, mk, false); a.val1 = physColumn(K_HI32, ch, mk, false);
                a.countCol = physColumn(K_COUNT, ch, mk, mk < 0);
                break;
            case PG_AGG_SUM_SHORT_DECIMAL: case PG_AGG_SUM_LONG_DECIMAL:
                // DecimalSumAggregation: exact 128-bit sums as 32-bit limbs (see K_DEC_*); FINAL reads the serialized states
                if (final) { a.val0 = physColumn(K_DST_L0, ch, -1, false); a.val1 = physColumn(K_DST_L1, ch, -1, false); a.val2 = physColumn(K_DST_L2, ch, -1, false); a.val3 = physColumn(K_DST_L3, ch, -1, false); }
                else if (a.spec.function == PG_AGG_SUM_SHORT_DECIMAL) { a.val0 = physColumn(K_LO32, ch, mk, false); a.val1 = physColumn(K_HI32, ch, mk, false); }
                else { a.val0 = physColumn(K_DEC_L0, ch, mk, false); a.val1 = physColumn(K_DEC_L1, ch, mk, false); a.val2 = physColumn(K_DEC_L2, ch, mk, false); a.val3 = physColumn(K_DEC_L3, ch, mk, false); }
                a.countCol = physColumn(K_COUNT, ch, final ? -1 : mk, (final ? -1 : mk) < 0);
                break;
            case PG_AGG_AVG_DOUBLE: case PG_AGG_AVG_BIGINT:
                a.countExact = true;
                if (final) { a.val0 = physColumn(K_F64, ch, -1, false); a.countCol = physColumn(K_I64_ADD, ch + 1, -1, false); }
                else {
                    a.val0 = physColumn(a.spec.function == PG_AGG_AVG_DOUBLE ? K_F64 : K_F64_FROM_I64, ch, mk, false);
                    a.countCol = physColumn(K_COUNT, ch, mk, mk < 0);
                }
                break;
            case PG_AGG_MIN_DOUBLE: case PG_AGG_MAX_DOUBLE: case PG_AGG_MIN_BIGINT: case PG_AGG_MAX_BIGINT: {
                // AbstractMinMaxAggregationFunction: NULL inputs are skipped, a group without a non-null input yields NULL; the intermediate
                // state (PARTIAL -> FINAL) is the value itself, so the FINAL step is the same function over the intermediate channel
                const int kind = a.spec.function == PG_AGG_MIN_DOUBLE ? K_MIN_F64 : a.spec.function == PG_AGG_MAX_DOUBLE ? K_MAX_F64
                               : a.spec.function == PG_AGG_MIN_BIGINT ? K_MIN_I64 : K_MAX_I64;
                a.val0 = physColumn(kind, ch, final ? -1 : mk, false);
                a.countCol = physColumn(K_COUNT, ch, final ? -1 : mk, (final ? -1 : mk) < 0);
                break;
            }
            default: throw Error(PG_ERR_UNSUPPORTED, "aggregate function is not on the GPU path (supported: count, sum, avg, min, max over BIGINT/DOUBLE, sum over decimals)");
        }
        if (a.spec.function != PG_AGG_COUNT_STAR) PG_CHECK_ARG(ch >= 0, "aggregate needs an input channel");
        aggs.push_back(a);
    }
    // exact counts of lazily counted channels fall back on the shared row count until the channel shows a NULL
    bool needRowCount = false;
    for (auto& a : aggs) if (a.countExact && a.countCol >= 0 && phys[a.countCol].lazy) needRowCount = true;
    if (needRowCount) rowCountCol = physColumn(K_COUNT, -1, -1, false);
    PG_CHECK_ARG(phys.size() <= 16, "too many distinct accumulator columns (max 16)");

}

class HashAggregationOp : public Operator {
protected:
    KeyLayout layout;
    int hashChannel, step;
    std::vector<AggPlan> aggs;
    std::vector<PhysCol> phys;
    int rowCountCol = -1;
    GroupTableStorage table;
    PageStager stager;
    std::vector<int> stagedChannels;          // input channels the kernels read
    PinnedBuffer hostCounters;
    DeviceBuffer bitmapA, bitmapB;
    uint64_t rowsSeen = 0;
    int64_t knownGroups = 0;
    bool finishing = false, finished = false, outputBuilt = false;
    int gridCTAs = 0; uint32_t sharedSlots = 1024, sharedLimit = 768;
    int MAXV = 4;
    // output
    ResultBuilder result;
    DeviceBuffer order;
    std::vector<DeviceBuffer> sortScratch;
    int64_t outGroups = 0, outNext = 0;
    // fused pipeline input (FusedAggregationOp below): while set, addInput stages the RAW page's channels and every chunk runs through
    // `fusedLaunch` (filter / project / probe / aggregate in one kernel) instead of the aggregation kernels
    const std::vector<int>* inputChannels = nullptr;
    std::function<void(const std::vector<DeviceColumn>&, int64_t, const uint32_t*, uint32_t*)> fusedLaunch;
    bool lastFused = false;
    // interned VARCHAR keys (KeyLayout::interned): one interner per such key, the id column of the chunk in flight
    std::vector<std::unique_ptr<VarcharInterner>> interners;
    std::vector<DeviceBuffer> internIds, internNulls;
    std::vector<DeviceColumn> internCols;

public:
    explicit HashAggregationOp(const pg_hash_aggregation_config* c) : Operator(c->options), stager(this)
    {
        planAggregation(c, layout, aggs, phys, rowCountCol);
        hashChannel = c->hash_channel;
        step = c->step;
        MAXV = phys.size() <= 4 ? 4 : (phys.size() <= 8 ? 8 : 16);

        // channels to stage: keys, hash channel is not needed (slot function is recomputed from the key), agg inputs, masks
        auto need = [&](int ch) { if (ch >= 0 && std::find(stagedChannels.begin(), stagedChannels.end(), ch) == stagedChannels.end()) stagedChannels.push_back(ch); };
        for (int k = 0; k < layout.nKeys; k++) need(layout.channels[k]);
        for (auto& p : phys) { need(p.channel); need(p.mask); }

        // launch geometry: persistent CTAs, grid = SMs x resident CTAs
        table.op = this; table.W = layout.W; table.nvPhys = (uint32_t) phys.size();
        configureGeometry();
        uint32_t minCap = nextPow2((uint64_t) 8 * gridCTAs * sharedLimit);
        uint32_t cap = std::max(minCap, nextPow2((uint64_t) c->expected_groups * 2));
        if (cap > (1u << 30)) throw Error(PG_ERR_INSUFFICIENT_RESOURCES, "Size of hash table cannot exceed 1 billion entries");
        table.counters.accounting = &stats.device_bytes;
        table.counters.reserve(CNT_COUNT * sizeof(unsigned int));
        PG_CUDA(cudaMemsetAsync(table.counters.ptr, 0, CNT_COUNT * sizeof(unsigned int), stream));
        table.allocate(cap);
        hostCounters.reserve(CNT_COUNT * sizeof(unsigned int));
        bitmapA.accounting = bitmapB.accounting = &stats.device_bytes;
        interners.resize(layout.nKeys); internIds.resize(layout.nKeys); internNulls.resize(layout.nKeys); internCols.resize(layout.nKeys);
        for (int k = 0; k < layout.nKeys; k++) {
            if (!layout.interned[k]) continue;
            interners[k].reset(new VarcharInterner(this, &stats.device_bytes));
            internIds[k].accounting = internNulls[k].accounting = &stats.device_bytes;
            // a key channel no accumulator reads stays DictionaryBlock / RLE on the device: its dictionary is interned, not its rows
            bool keyOnly = true;
            for (auto& p : phys) if (p.channel == layout.channels[k] || p.mask == layout.channels[k]) keyOnly = false;
            for (int j = 0; j < layout.nKeys; j++) if (j != k && layout.channels[j] == layout.channels[k]) keyOnly = false;
            if (keyOnly) stager.keepEncoded.push_back(layout.channels[k]);
        }
        PG_CUDA(cudaStreamSynchronize(stream));
    }

    // Interned VARCHAR keys of this chunk → id columns (MultiChannelGroupByHash compares the slices themselves, :293-358; here equal
    // slices get equal ids first and the packed key carries the id).  Called after settle(): the previous chunk no longer needs its ids.
    void internKeys(const std::vector<DeviceColumn>& cols, int64_t n)
    {
        for (int k = 0; k < layout.nKeys; k++) {
            if (!layout.interned[k]) continue;
            const size_t i = std::find(stagedChannels.begin(), stagedChannels.end(), layout.channels[k]) - stagedChannels.begin();
            const DeviceColumn& src = cols[i];
            PG_CHECK_ARG(src.type == PG_TYPE_VARCHAR, "group-by channel type differs from the operator's configuration");
            internIds[k].reserve((size_t) n * 4 + 16);
            DeviceColumn ids; ids.type = PG_TYPE_INTEGER; ids.n = n; ids.values = internIds[k].ptr; ids.nulls = src.nulls; ids.offsets = nullptr;
            if (src.encoding != PG_ENC_FLAT) {
                uint8_t* rowNulls = nullptr;
                if (src.nulls) { internNulls[k].reserve((size_t) n + 16); rowNulls = internNulls[k].as<uint8_t>(); }
                interners[k]->internEncoded(this, src, n, internIds[k].as<int32_t>(), rowNulls, true);
                ids.nulls = rowNulls;
                stats.kernel_paths |= PG_PATH_DICTIONARY_KEYS;
            }
            else interners[k]->intern(this, src, n, internIds[k].as<int32_t>(), true);
            internCols[k] = ids;
            stats.kernel_paths |= PG_PATH_INTERNED_KEYS;
        }
    }

    template <int W, int MV> void configureFor()
    {
        size_t smem = sharedBytes(sharedSlots);
        auto kern = hashAggKernel<W, MV>;
        PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        int perSM = 0;
        PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kAggThreads, smem));
        (void) perSM;
        gridCTAs = numSMs * 2;   // the shared-memory budget (tier 0 + table) is sized for two resident CTAs per SM
    }
    size_t sharedBytes(uint32_t S) const { return SharedLayout(S, layout.W, (int) phys.size(), true).total; }
    void configureGeometry()
    {
        sharedSlots = 1024;
        while (sharedBytes(sharedSlots) > 112 * 1024 && sharedSlots > 128) sharedSlots >>= 1;   // two CTAs per SM
        sharedLimit = sharedSlots / 4 * 3;
        dispatch([&](auto w, auto mv) { configureFor<decltype(w)::value, decltype(mv)::value>(); });
    }
    template <typename F> void dispatch(F&& f) { dispatchFor(MAXV, f); }
    // kernels are instantiated for 4 / 8 / 16 accumulator columns; a launch picks the smallest that holds its ACTIVE columns
    template <typename F> void dispatchActive(F&& f) { int a = activeCount(); dispatchFor(a <= 4 ? 4 : (a <= 8 ? 8 : 16), f); }
    template <typename F> void dispatchFor(int maxv, F&& f)
    {
        auto withW = [&](auto w) {
            if (maxv == 4) f(w, std::integral_constant<int, 4>());
            else if (maxv == 8) f(w, std::integral_constant<int, 8>());
            else f(w, std::integral_constant<int, 16>());
        };
        switch (layout.W) {
            case 1: withW(std::integral_constant<int, 1>()); break;
            case 2: withW(std::integral_constant<int, 2>()); break;
            case 3: withW(std::integral_constant<int, 3>()); break;
            default: withW(std::integral_constant<int, 4>()); break;
        }
    }

    bool needsInput() override { return !finishing && !outputBuilt; }

    void readCounters(unsigned int* out)
    {
        PG_CUDA(cudaMemcpyAsync(hostCounters.ptr, table.counters.ptr, CNT_COUNT * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        memcpy(out, hostCounters.ptr, CNT_COUNT * sizeof(unsigned int));
        stats.d2h_bytes += CNT_COUNT * sizeof(unsigned int);
    }
    void checkErrors(const unsigned int* cnt)
    {
        if (cnt[CNT_ERROR] & ERR_VARCHAR_TOO_LONG) throw Error(PG_ERR_UNSUPPORTED, "VARCHAR group-by key longer than its declared max length (GPU path packs keys of <= 7 bytes)");
        if (cnt[CNT_ERROR] & 0x80000000u) throw Error(PG_ERR_INVALID_STATE, "internal: shared-table flush could not claim a slot");
        // the PageProcessor half of a fused fragment
        if (cnt[CNT_FPERR] & 4u) throw Error(PG_ERR_UNSUPPORTED, "VARCHAR group-by key longer than its declared max length (GPU path packs keys of <= 7 bytes)");
        if (cnt[CNT_FPERR] & 1u) throw Error(PG_ERR_NUMERIC_OVERFLOW, "bigint arithmetic overflow");
        if (cnt[CNT_FPERR] & 2u) throw Error(PG_ERR_INVALID_ARGUMENT, "Division by zero");
    }

    void grow(uint64_t wantGroups)
    {
        uint64_t want = std::max<uint64_t>((uint64_t) table.cap * 2, wantGroups * 4);
        if (want > (1ULL << 30)) throw Error(PG_ERR_INSUFFICIENT_RESOURCES, "Size of hash table cannot exceed 1 billion entries");
        GroupTableStorage nt;
        nt.op = this; nt.W = table.W; nt.nvPhys = table.nvPhys; nt.withGid = false;
        nt.counters = std::move(table.counters);
        unsigned int zero = 0;
        PG_CUDA(cudaMemcpyAsync(nt.counters.as<unsigned int>() + CNT_GROUPS, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
        nt.allocate(nextPow2(want));
        GroupTable oldV = table.view(); oldV.counters = nt.counters.as<unsigned int>();
        GroupTable newV = nt.view();
        switch (table.W) {
            case 1: PG_LAUNCH(this, rehashKernel<1>, 1024, 256, 0, oldV, newV); break;
            case 2: PG_LAUNCH(this, rehashKernel<2>, 1024, 256, 0, oldV, newV); break;
            case 3: PG_LAUNCH(this, rehashKernel<3>, 1024, 256, 0, oldV, newV); break;
            default: PG_LAUNCH(this, rehashKernel<4>, 1024, 256, 0, oldV, newV); break;
        }
        PG_CUDA(cudaStreamSynchronize(stream));
        table.keys = std::move(nt.keys); table.state = std::move(nt.state); table.firstRow = std::move(nt.firstRow);
        table.acc = std::move(nt.acc); table.counters = std::move(nt.counters); table.cap = nt.cap;
    }

    int activeCount() const { int n = 0; for (auto& p : phys) n += p.active ? 1 : 0; return n; }

    // BIGINT key without NULLs, only non-null DOUBLE sums (+ at most one plain row count), 16-byte aligned columns → fast kernel
    template <int NV, bool CNT>
    bool launchFastT(const std::vector<DeviceColumn>& cols, int64_t n, uint32_t* bitmapOut, const std::vector<int>& f64Cols, int countCol)
    {
        FastLaunch<NV, CNT> L{};
        L.t = table.view();
        auto colOf = [&](int ch) -> const DeviceColumn& { return cols[std::find(stagedChannels.begin(), stagedChannels.end(), ch) - stagedChannels.begin()]; };
        L.key = (const uint64_t*) colOf(layout.channels[0]).values;
        L.keyMode = layout.bigintPath ? 0 : 1;
        if (L.keyMode == 1)
            for (int k = 0; k < 2; k++) { L.key32[k] = (const uint32_t*) colOf(layout.channels[k]).values; L.kWord[k] = layout.word[k]; L.kShift[k] = layout.shift[k]; }
        for (int c = 0; c < NV; c++) { L.vals[c] = (const uint64_t*) colOf(phys[f64Cols[c]].channel).values; L.phys[c] = f64Cols[c]; }
        L.phys[NV] = CNT ? countCol : 0;
        L.n = n; L.rowBase = rowsSeen; L.bitmapOut = bitmapOut; L.S = sharedSlots; L.sLimit = sharedLimit; L.nAccAlloc = (int) phys.size();
        L.subsetCount = 1; L.subsetIndex = 0;
        // ---- mid mode: more groups than the two-CTA shared table holds, few enough for a handful of hash-subset passes ----
        constexpr int NC = NV + (CNT ? 1 : 0);
        uint32_t midSlots = 4096;
        while (SharedLayout(midSlots, 1, NC, false).total > 200 * 1024 && midSlots > 512) midSlots >>= 1;
        const uint32_t midCapacity = midSlots / 4 * 3;
        const int64_t passes = (int64_t) ((double) knownGroups * 1.25 / midCapacity) + 1;
        const char* modeEnv0 = getenv("PG_AGG_MODE");
        const bool forceDirect = modeEnv0 && !strcmp(modeEnv0, "direct");
        const bool forceMid = modeEnv0 && !strcmp(modeEnv0, "mid");
        // warp-private accumulator sets: more groups than tier 0 holds (K0), at most what the per-warp sets hold (KW)
        constexpr int K0host = FastGeom<NV, CNT>::K0;
        const int KWcap = std::min<int>(254, kTier0Bytes / ((kAggThreads / 32) * (NC + 1) * 8));
        const bool forceWarp = modeEnv0 && !strcmp(modeEnv0, "warp");
        const bool warpMode = !forceMid && !forceDirect && !(modeEnv0 && !strcmp(modeEnv0, "tiered")) && knownGroups > K0host && (knownGroups <= kWarpMaxGroups || forceWarp) && knownGroups <= KWcap;
        L.warpGroups = warpMode ? KWcap : 0;
        // measured (256 M rows, 4 sums, G rows/s): tiers 0/1 — 109 up to 10 groups (all tier 0), 44 / 39 / 34 / 33 / 18 at 11 / 16 / 32 / 64 / 128;
        // one 4096-slot table per SM — 22 / 27 / 36 / 44 / 47 / 49 at 11 / 16 / 32 / 64 / 128 / 256..1000 (fewer threads per hot address)
        if (!forceDirect && !warpMode && (knownGroups > kMidMinGroups || forceMid) && passes <= kMaxSubsetPasses) {   // one pass; beyond that direct mode is faster
            auto kern = hashAggFastKernel<NV, CNT, true>;
            L.S = midSlots; L.sLimit = midCapacity; L.nAccAlloc = NC; L.subsetCount = (uint32_t) passes;
            const size_t smem = SharedLayout(midSlots, 1, NC, false).total;
            PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            const int64_t tiles = divUp(n, 512 * 2);
            const int grid = (int) std::min<int64_t>((int64_t) numSMs, std::max<int64_t>(tiles, 1));
            for (uint32_t j = 0; j < L.subsetCount; j++) {
                L.subsetIndex = j;
                kern<<<grid, 512, smem, stream>>>(L);
                stats.kernel_launches++;
            }
            PG_CUDA(cudaGetLastError());
            return true;
        }
        // ---- direct mode: too many groups for shared-memory tables: straight to the HBM table, one L2-resident slot range per pass ----
        const char* modeEnv = getenv("PG_AGG_MODE");
        if ((knownGroups > (int64_t) sharedLimit * 9 / 10 && !(modeEnv && !strcmp(modeEnv, "tiered"))) || (modeEnv && !strcmp(modeEnv, "direct"))) {
            auto kern = hashAggDirectKernel<NV, CNT>;
            int64_t budgetMB = 96;
            if (const char* e = getenv("PG_AGG_L2_BUDGET_MB")) budgetMB = std::max<int64_t>(1, atoll(e));
            // bytes the probes and atomics of one full pass touch: every key slot (probing reads empty slots too), firstRow + accumulators of live groups
            const double live = std::min<double>((double) table.cap, (double) std::max<int64_t>(knownGroups, 1) * 1.25);
            const double footprint = (double) table.cap * 8.0 + live * (32.0 + 8.0 * table.nvPhys);
            L.subsetCount = (uint32_t) std::min<double>(64.0, std::max<double>(1.0, std::ceil(footprint / ((double) budgetMB * 1048576.0))));
            L.capBits = 0; while ((1u << L.capBits) < table.cap) L.capBits++;
            const int64_t tiles = divUp(n, 512);
            const int grid = (int) std::min<int64_t>((int64_t) numSMs * 4, std::max<int64_t>(tiles, 1));
            for (uint32_t j = 0; j < L.subsetCount; j++) {
                L.subsetIndex = j;
                kern<<<grid, 256, 0, stream>>>(L);
                stats.kernel_launches++;
            }
            PG_CUDA(cudaGetLastError());
            return true;
        }
        auto kern = hashAggFastKernel<NV, CNT, false>;
        const size_t smem = sharedBytes(sharedSlots);
        PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int perSM = 0;
        PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kAggThreads, smem));
        perSM = std::max(1, std::min(perSM, 2));
        const int64_t tiles = divUp(n, kAggThreads * 2);
        const int grid = (int) std::min<int64_t>((int64_t) numSMs * perSM, std::max<int64_t>(tiles, 1));
        kern<<<grid, kAggThreads, smem, stream>>>(L);
        stats.kernel_launches++;
        PG_CUDA(cudaGetLastError());
        return true;
    }

    bool tryLaunchFast(const std::vector<DeviceColumn>& cols, int64_t n, uint32_t* bitmapOut)
    {
        auto colOf = [&](int ch) -> const DeviceColumn& { return cols[std::find(stagedChannels.begin(), stagedChannels.end(), ch) - stagedChannels.begin()]; };
        // key: one BIGINT column, or two 32-bit columns (INTEGER / DATE: Q1's returnflag, linestatus codes); no NULL keys on this page
        const bool packed32 = !layout.bigintPath && layout.nKeys == 2 && layout.W == 2 && layout.bits[0] == 32 && layout.bits[1] == 32
                              && layout.types[0] != PG_TYPE_VARCHAR && layout.types[1] != PG_TYPE_VARCHAR;
        if (!layout.bigintPath && !packed32) return false;
        for (int k = 0; k < layout.nKeys; k++) {
            const DeviceColumn& kc = colOf(layout.channels[k]);
            PG_CHECK_ARG(kc.type == layout.types[k], "group-by channel type differs from the operator's configuration");
            if (kc.nulls || ((uintptr_t) kc.values % 16) != 0) return false;
        }
        if (packed32 && knownGroups > (int64_t) sharedLimit * 9 / 10) return false;    // mid / direct modes address the W = 1 table
        std::vector<int> f64Cols; int countCol = -1;
        for (size_t p = 0; p < phys.size(); p++) {
            if (!phys[p].active) continue;
            if (phys[p].kind == K_F64 && phys[p].mask < 0) {
                const DeviceColumn& vc = colOf(phys[p].channel);
                PG_CHECK_ARG(typeWidth(vc.type) == 8, "sum / avg inputs must be BIGINT or DOUBLE columns");
                if (vc.nulls || ((uintptr_t) vc.values % 16) != 0) return false;
                f64Cols.push_back((int) p);
            }
            else if (phys[p].kind == K_COUNT && phys[p].channel < 0 && phys[p].mask < 0 && countCol < 0) countCol = (int) p;
            else return false;
        }
        const int nv = (int) f64Cols.size();
        if (nv < 1 || nv > 8) return false;
        // the shared-memory budget assumes 2 CTAs per SM; flushReserve was sized from gridCTAs >= this grid
#define PG_FAST_CASE(N) case N: return countCol >= 0 ? launchFastT<N, true>(cols, n, bitmapOut, f64Cols, countCol) : launchFastT<N, false>(cols, n, bitmapOut, f64Cols, countCol);
        switch (nv) { PG_FAST_CASE(1) PG_FAST_CASE(2) PG_FAST_CASE(3) PG_FAST_CASE(4) PG_FAST_CASE(5) PG_FAST_CASE(6) PG_FAST_CASE(7) PG_FAST_CASE(8) }
#undef PG_FAST_CASE
        return false;
    }

    template <int W, int MV>
    void launchRange(const std::vector<DeviceColumn>& cols, int64_t begin, int64_t n, const uint32_t* bitmapIn, uint32_t* bitmapOut)
    {
        AggLaunch<MV> L{};
        L.t = table.view();
        auto colOf = [&](int ch) -> const DeviceColumn& {
            size_t i = std::find(stagedChannels.begin(), stagedChannels.end(), ch) - stagedChannels.begin();
            return cols[i];
        };
        std::vector<DeviceColumn> keyCols;
        for (int k = 0; k < layout.nKeys; k++) {
            DeviceColumn kc = layout.interned[k] ? internCols[k] : colOf(layout.channels[k]);
            PG_CHECK_ARG(layout.interned[k] || kc.type == layout.types[k], "group-by channel type differs from the operator's configuration");
            int w = typeWidth(kc.type);
            if (w) kc.values = (const char*) kc.values + begin * w; else kc.offsets = kc.offsets + begin;
            if (kc.nulls) kc.nulls += begin;
            keyCols.push_back(kc);
        }
        layout.fill(L.kp, keyCols, table.counters.as<unsigned int>() + CNT_ERROR);
        int na = 0;
        for (size_t p = 0; p < phys.size(); p++) {
            if (!phys[p].active) continue;
            AccInput& a = L.acc[na];
            a.kind = phys[p].kind; a.phys = (int) p; a.values = nullptr; a.nulls = nullptr; a.mask = nullptr; a.maskNulls = nullptr; a.offsets = nullptr;
            if (phys[p].channel >= 0) {
                const DeviceColumn& vc = colOf(phys[p].channel);
                if (kindIsDecimal(a.kind)) {
                    PG_CHECK_ARG(vc.type == PG_TYPE_LONG_DECIMAL && (uintptr_t) vc.values % 16 == 0, "sum(long decimal) needs a 16-byte aligned PG_TYPE_LONG_DECIMAL column");
                    a.values = (const uint64_t*) vc.values + 2 * begin;
                }
                else if (kindIsDecimalState(a.kind)) {
                    PG_CHECK_ARG(vc.type == PG_TYPE_VARCHAR && vc.offsets, "the FINAL step of sum(decimal) reads VARBINARY states");
                    a.values = (const uint64_t*) vc.values; a.offsets = vc.offsets + begin;
                }
                else if (phys[p].kind != K_COUNT) {
                    PG_CHECK_ARG(typeWidth(vc.type) == 8, "sum / avg inputs must be BIGINT or DOUBLE columns");
                    a.values = (const uint64_t*) vc.values + begin;
                }
                a.nulls = vc.nulls ? vc.nulls + begin : nullptr;
            }
            if (phys[p].mask >= 0) {
                const DeviceColumn& mc = colOf(phys[p].mask);
                PG_CHECK_ARG(mc.type == PG_TYPE_BOOLEAN, "mask channel must be BOOLEAN");
                a.mask = (const uint8_t*) mc.values + begin; a.maskNulls = mc.nulls ? mc.nulls + begin : nullptr;
            }
            if (kindIsFloat(a.kind)) L.floatMask |= 1u << na;
            if (kindIsMax(a.kind)) L.maxMask |= 1u << na;
            na++;
        }
        L.nAcc = na;
        L.n = n; L.rowBase = rowsSeen + (uint64_t) begin;
        L.bitmapIn = bitmapIn; L.bitmapOut = bitmapOut;
        L.S = bitmapIn ? 0 : sharedSlots; L.sLimit = sharedLimit;
        L.nAccAlloc = (int) phys.size();
        bool aligned = true;
        if (layout.bigintPath) aligned = aligned && ((uintptr_t) L.kp.f[0].values % 16 == 0);
        for (int c = 0; c < na; c++) if (L.acc[c].values && L.acc[c].kind < K_DEC_L0) aligned = aligned && ((uintptr_t) L.acc[c].values % 16 == 0);
        L.vec = aligned ? 1 : 0;
        size_t smem = L.S ? sharedBytes(sharedSlots) : 64;
        int64_t tiles = divUp(n, kAggThreads * RowsPerThread<MV>::value);
        int grid = (int) std::min<int64_t>(gridCTAs, std::max<int64_t>(tiles, 1));
        auto kern = hashAggKernel<W, MV>;
        PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        kern<<<grid, kAggThreads, smem, stream>>>(L);
        stats.kernel_launches++;
        PG_CUDA(cudaGetLastError());
    }

    void activateLazyCounts(const std::vector<DeviceColumn>& cols)
    {
        for (size_t p = 0; p < phys.size(); p++) {
            if (!phys[p].lazy || phys[p].active) continue;
            size_t i = std::find(stagedChannels.begin(), stagedChannels.end(), phys[p].channel) - stagedChannels.begin();
            if (cols[i].nulls == nullptr) continue;
            // first page with NULLs in this channel: every group seen so far has had only non-null inputs
            PG_LAUNCH(this, backfillCountKernel, 1024, 256, 0, table.view(), (int) p, rowCountCol);
            phys[p].active = true;
        }
    }

    void addInput(const pg_page* page) override
    {
        PG_CHECK_STATE(needsInput(), "Operator is already finishing");   // HashAggregationOperator.java:415
        const int64_t n = page->position_count;
        stats.input_positions += n;
        if (n == 0) return;
        // host pages are staged in sub-chunks so the PCIe copy of chunk k+1 overlaps the kernel of chunk k
        const int64_t chunk = page->memory == PG_MEM_HOST ? (int64_t(1) << 23) : (int64_t(1) << 30);
        // a large first page starts with a small probe chunk: its group count picks the kernel mode for the rest
        const int64_t probe = (rowsSeen == 0 && n > (int64_t(1) << 24)) ? (int64_t(1) << 22) : 0;
        for (int64_t begin = 0; begin < n;) {
            const int64_t m = std::min(begin == 0 && probe ? probe : chunk, n - begin);
            pg_page sub = *page;
            std::vector<pg_block> blocks(page->blocks, page->blocks + page->channel_count);
            sub.blocks = blocks.data(); sub.position_count = m;
            for (auto& b : blocks) {
                if (b.encoding == PG_ENC_FLAT) {
                    int w = typeWidth(b.type);
                    if (w) b.values = (const char*) b.values + begin * w; else b.offsets = b.offsets + begin;
                    if (b.is_null) b.is_null += begin;
                }
                else if (b.encoding == PG_ENC_DICTIONARY) b.ids += begin;
                b.position_count = m;
            }
            std::vector<DeviceColumn> cols;
            int slot = stager.stage(&sub, inputChannels ? *inputChannels : stagedChannels, cols);   // PCIe copy of this chunk overlaps the previous chunk's kernel
            settle();                                              // previous chunk: counters, deferred rows, growth
            if (!fusedLaunch) { activateLazyCounts(cols); internKeys(cols, m); }
            processChunk(cols, m);
            stager.markCompute(slot);
            rowsSeen += (uint64_t) m;
            begin += m;
        }
        // the borrowed (device) page must not be referenced after return, and deferred rows need its columns: settle now
        settle();
    }

    std::vector<DeviceColumn> lastCols; int64_t lastN = 0; uint64_t lastBase = 0; bool lastPending = false;

    void processChunk(const std::vector<DeviceColumn>& cols, int64_t n)
    {
        size_t bmBytes = (size_t) divUp(n, 32) * 4;
        bitmapA.reserve(bmBytes);
        PG_CUDA(cudaMemsetAsync(bitmapA.ptr, 0, bmBytes, stream));
        if (fusedLaunch) fusedLaunch(cols, n, nullptr, bitmapA.as<uint32_t>());
        else if (!tryLaunchFast(cols, n, bitmapA.as<uint32_t>()))
            dispatchActive([&](auto w, auto mv) { launchRange<decltype(w)::value, decltype(mv)::value>(cols, 0, n, nullptr, bitmapA.as<uint32_t>()); });
        lastCols = cols; lastN = n; lastBase = rowsSeen; lastPending = true; lastFused = (bool) fusedLaunch;
    }

    // Wait for the in-flight chunk, re-run rows that found the table at its claim limit after growing it.
    void settle()
    {
        if (!lastPending) return;
        unsigned int cnt[CNT_COUNT];
        readCounters(cnt);
        checkErrors(cnt);
        uint32_t* in = bitmapA.as<uint32_t>();
        uint64_t savedRows = rowsSeen;
        rowsSeen = lastBase;
        int rounds = 0;
        while (cnt[CNT_DEFERRED] > 0) {
            PG_CHECK_STATE(++rounds < 40, "internal: deferred rows did not converge");
            grow((uint64_t) cnt[CNT_GROUPS] + cnt[CNT_DEFERRED]);
            size_t bmBytes = (size_t) divUp(lastN, 32) * 4;
            DeviceBuffer& outB = (in == bitmapA.as<uint32_t>()) ? bitmapB : bitmapA;
            outB.reserve(bmBytes);
            PG_CUDA(cudaMemsetAsync(outB.ptr, 0, bmBytes, stream));
            unsigned int zero = 0;
            PG_CUDA(cudaMemcpyAsync(table.counters.as<unsigned int>() + CNT_DEFERRED, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
            if (lastFused) fusedLaunch(lastCols, lastN, in, outB.as<uint32_t>());     // the fragment's row program again, restricted to the deferred rows
            else dispatchActive([&](auto w, auto mv) { launchRange<decltype(w)::value, decltype(mv)::value>(lastCols, 0, lastN, in, outB.as<uint32_t>()); });
            in = outB.as<uint32_t>();
            readCounters(cnt);
            checkErrors(cnt);
        }
        rowsSeen = savedRows;
        knownGroups = cnt[CNT_GROUPS];
        if ((uint64_t) knownGroups * 2 > table.cap) grow((uint64_t) knownGroups);
        lastPending = false;
    }

    // ---- state hand-off (see exportStateKernel) ---------------------------------------------------------------------------
    uint64_t stateConfigHash() const
    {
        uint64_t h = 0xcbf29ce484222325ULL;
        auto mix = [&](uint64_t v) { h = (h ^ v) * 0x100000001b3ULL; };
        mix(layout.W); mix(layout.nKeys); mix(layout.bigintPath);
        for (int k = 0; k < layout.nKeys; k++) { mix(layout.types[k]); mix(layout.word[k]); mix(layout.shift[k]); mix(layout.bits[k]); }
        mix(phys.size());
        for (auto& p : phys) mix(p.kind);
        return h;
    }
    size_t stateBytes(int64_t maxGroups) const
    {
        PG_CHECK_ARG(maxGroups >= 1 && maxGroups <= (int64_t(1) << 30), "max_groups out of range");
        return aggStateBytes(layout.W, (uint32_t) phys.size(), (uint64_t) maxGroups);
    }
    // Everything this operator has aggregated so far, compacted into `buffer` (device memory of this operator's GPU) on its stream;
    // asynchronous — a consumer on another stream orders itself after the operator's stream.  The operator itself is unchanged.
    void exportState(void* buffer, size_t capacity, int64_t maxGroups)
    {
        PG_CHECK_ARG(buffer != nullptr && capacity >= stateBytes(maxGroups), "state buffer smaller than pg_hash_aggregation_state_bytes(max_groups)");
        PG_CHECK_STATE(!outputBuilt, "the operator has already produced output");
        if (layout.anyInterned()) throw Error(PG_ERR_UNSUPPORTED, "aggregation states with interned VARCHAR keys (ids are private to one operator): use the page-based PARTIAL / FINAL steps");
        settle();
        uint32_t inactive = 0;
        for (size_t p = 0; p < phys.size(); p++) if (phys[p].lazy && !phys[p].active) inactive |= 1u << p;
        const AggStateView v = aggStateView(buffer, layout.W, (uint32_t) phys.size(), (uint32_t) maxGroups);
        PG_CUDA(cudaMemsetAsync(buffer, 0, sizeof(AggStateHeader), stream));
        const GroupTable T = table.view();
        const int grid = (int) std::min<int64_t>(divUp(T.capP, 256), 8 * numSMs);
        const uint64_t hash = stateConfigHash();
        switch (layout.W) {
            case 1: PG_LAUNCH(this, exportStateKernel<1>, grid, 256, 0, T, v, hash, inactive, rowCountCol); break;
            case 2: PG_LAUNCH(this, exportStateKernel<2>, grid, 256, 0, T, v, hash, inactive, rowCountCol); break;
            case 3: PG_LAUNCH(this, exportStateKernel<3>, grid, 256, 0, T, v, hash, inactive, rowCountCol); break;
            default: PG_LAUNCH(this, exportStateKernel<4>, grid, 256, 0, T, v, hash, inactive, rowCountCol); break;
        }
    }
    // `count` states of identically planned operators, `stride` bytes apart in this GPU's memory (an all-gather's receive buffer),
    // written before this call was made or by work this operator's stream has been ordered after.  Source j is merged as if its rows
    // followed those of source j-1 (and this operator's own input when orderBase > 0).
    void mergeStates(const void* buffers, int32_t count, size_t stride, int64_t maxGroups, int32_t orderBase)
    {
        PG_CHECK_ARG(buffers != nullptr && count >= 1 && count <= 200 && orderBase >= 0 && orderBase + count <= 255, "between 1 and 200 states, order tags below 255");
        PG_CHECK_ARG(stride >= stateBytes(maxGroups), "stride smaller than one state");
        PG_CHECK_STATE(needsInput(), "Operator is already finishing");
        if (layout.anyInterned()) throw Error(PG_ERR_UNSUPPORTED, "aggregation states with interned VARCHAR keys (ids are private to one operator): use the page-based PARTIAL / FINAL steps");
        settle();
        // one strided copy brings every header; the group counts size the table, the hashes guard against a mismatched plan
        hostCounters.reserve(std::max<size_t>((size_t) count * sizeof(AggStateHeader), CNT_COUNT * sizeof(unsigned int)));
        PG_CUDA(cudaMemcpy2DAsync(hostCounters.ptr, sizeof(AggStateHeader), buffers, stride, sizeof(AggStateHeader), (size_t) count, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        stats.d2h_bytes += (int64_t) count * (int64_t) sizeof(AggStateHeader);
        std::vector<AggStateHeader> headers((const AggStateHeader*) hostCounters.ptr, (const AggStateHeader*) hostCounters.ptr + count);
        uint64_t incoming = 0;
        for (const AggStateHeader& h : headers) {
            if (h.magic == 0 && h.nGroups == 0) continue;       // an unused slot of the gather
            PG_CHECK_ARG(h.magic == kAggStateMagic && h.W == (uint32_t) layout.W && h.nvPhys == phys.size() && h.maxGroups == (uint32_t) maxGroups
                         && h.configHash == stateConfigHash(), "aggregation state of a differently planned operator");
            if (h.overflow || h.nGroups > h.maxGroups)
                throw Error(PG_ERR_INSUFFICIENT_RESOURCES, "aggregation state holds " + std::to_string(h.nGroups) + " groups, its buffer only " + std::to_string(h.maxGroups));
            incoming += h.nGroups;
        }
        if (((uint64_t) knownGroups + incoming) * 2 > table.cap) grow((uint64_t) knownGroups + incoming);
        uint32_t floatMask = 0, maxMask = 0;
        for (size_t p = 0; p < phys.size(); p++) {
            if (kindIsFloat(phys[p].kind)) floatMask |= 1u << p;
            if (kindIsMax(phys[p].kind)) maxMask |= 1u << p;
            if (phys[p].lazy && !phys[p].active) {      // the states carry every count explicitly
                PG_LAUNCH(this, backfillCountKernel, 1024, 256, 0, table.view(), (int) p, rowCountCol);
                phys[p].active = true;
            }
        }
        const GroupTable T = table.view();
        for (int32_t j = 0; j < count; j++) {
            const AggStateHeader& h = headers[j];
            if (h.nGroups == 0) continue;
            const AggStateView v = aggStateView((char*) buffers + (size_t) j * stride, layout.W, (uint32_t) phys.size(), (uint32_t) maxGroups);
            const uint64_t tag = (uint64_t) (orderBase + j) << 56;
            const int grid = (int) std::min<int64_t>(divUp(h.nGroups, 256), 8 * numSMs);
            switch (layout.W) {
                case 1: PG_LAUNCH(this, mergeStateKernel<1>, grid, 256, 0, T, v, h.nGroups, tag, floatMask, maxMask); break;
                case 2: PG_LAUNCH(this, mergeStateKernel<2>, grid, 256, 0, T, v, h.nGroups, tag, floatMask, maxMask); break;
                case 3: PG_LAUNCH(this, mergeStateKernel<3>, grid, 256, 0, T, v, h.nGroups, tag, floatMask, maxMask); break;
                default: PG_LAUNCH(this, mergeStateKernel<4>, grid, 256, 0, T, v, h.nGroups, tag, floatMask, maxMask); break;
            }
        }
        knownGroups += (int64_t) incoming;       // an upper bound until the next settle() reads the counter
        stats.input_positions += (int64_t) incoming;
    }

    void finish() override { finishing = true; }
    bool isFinished() override { return finished; }
    void syncStats() override { if (lastPending) settle(); stats.group_count = knownGroups; }

    // ---- output (buildResult) -------------------------------------------------------------------------------------------
    void buildOutput()
    {
        settle();
        outputBuilt = true;
        GroupTable T = table.view();
        int64_t g = sortSlotsByFirstRow(this, T, order, sortScratch);
        outGroups = g; outNext = 0;
        stats.group_count = g; knownGroups = g;
        if (g == 0) return;
        std::vector<AggEmit> emits;
        for (auto& a : aggs) {
            AggEmit e; e.function = a.spec.function; e.val0 = a.val0; e.val1 = a.val1; e.val2 = a.val2; e.val3 = a.val3; e.partial = step == PG_STEP_PARTIAL; e.countCol = -1;
            if (a.countCol >= 0) e.countCol = phys[a.countCol].active ? a.countCol : (a.countExact ? rowCountCol : -1);
            emits.push_back(e);
        }
        result.build(this, T, layout, order.as<uint32_t>(), g, hashChannel >= 0, emits, &interners);
    }

    bool getOutput(pg_page* out) override   // HashAggregationOperator.java:451-501
    {
        if (finished || !finishing) return false;
        if (!outputBuilt) buildOutput();
        if (outNext >= outGroups) { finished = true; return false; }
        const int64_t m = std::min<int64_t>(maxOutputRows(), outGroups - outNext);
        result.slice(outNext, m, options.output_memory, out);
        outNext += m;
        stats.output_positions += m;
        if (outNext >= outGroups) finished = true;
        return true;
    }
};


// ======================================================================================================================
// GroupByHash (standalone) — MB/operator/GroupByHash.java:61-89.  Exact incremental group ids per page:
//   insert kernel (find-or-claim, atomicMin of the first-touch row of every not-yet-numbered slot, claimers append the slot to
//   a "new" list) → sort the new slots by first-touch row → gid = nextGroupId + rank → lookup kernel writes the ids.
// ======================================================================================================================
template <int W>
__global__ void gbhInsertKernel(GroupTable T, KeyPlan kp, int64_t n, uint64_t rowBase, uint32_t* newSlots)
{
    for (int64_t row = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; row < n; row += (int64_t) gridDim.x * blockDim.x) {
        uint64_t kw[W];
        uint32_t spelling = 0;
        int special = packKey<W>(kp, row, kw, &spelling);
        int slot;
        if (special) slot = (int) T.cap + special - 1;
        else {
            bool isNew = false;
            slot = tableFind<W>(T, kw, hashKeyWords<W>(kw), true, true, &isNew);
            if (isNew) newSlots[atomicAdd(&T.counters[CNT_NEW], 1u)] = (uint32_t) slot;
        }
        if (T.gid[slot] < 0) {
            const uint64_t g = firstRowWord(rowBase + (uint64_t) row, spelling);
            if (g < __ldcg(&T.firstRow[slot])) atomicMin((unsigned long long*) &T.firstRow[slot], (unsigned long long) g);
        }
    }
}
__global__ void gbhSpecialsKernel(GroupTable T, uint32_t* newSlots)
{
    if (threadIdx.x < 2) {
        uint32_t slot = T.cap + threadIdx.x;
        if (T.gid[slot] < 0 && T.firstRow[slot] != kNoRow) newSlots[atomicAdd(&T.counters[CNT_NEW], 1u)] = slot;
    }
}
__global__ void gbhGatherFirstRowsKernel(GroupTable T, const uint32_t* newSlots, uint32_t nNew, uint64_t* keysOut)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nNew; i += gridDim.x * blockDim.x) keysOut[i] = T.firstRow[newSlots[i]];
}
__global__ void gbhAssignKernel(GroupTable T, const uint32_t* sortedSlots, uint32_t nNew, int32_t nextGroupId, uint32_t* slotOfGid)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nNew; i += gridDim.x * blockDim.x) {
        uint32_t slot = sortedSlots[i];
        T.gid[slot] = nextGroupId + (int32_t) i;
        slotOfGid[nextGroupId + i] = slot;
    }
}
template <int W>
__global__ void gbhLookupKernel(GroupTable T, KeyPlan kp, int64_t n, int64_t* idsOut, uint8_t* containsOut)
{
    for (int64_t row = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; row < n; row += (int64_t) gridDim.x * blockDim.x) {
        uint64_t kw[W];
        int special = packKey<W>(kp, row, kw);
        int slot = special ? (int) T.cap + special - 1 : tableFind<W>(T, kw, hashKeyWords<W>(kw), false, true);
        int32_t gid = slot >= 0 ? T.gid[slot] : -1;
        if (idsOut) idsOut[row] = gid;
        if (containsOut) containsOut[row] = gid >= 0 ? 1 : 0;
    }
}
__global__ void gbhRebuildSlotOfGidKernel(GroupTable T, uint32_t* slotOfGid)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < T.capP; s += gridDim.x * blockDim.x)
        if (T.gid[s] >= 0) slotOfGid[T.gid[s]] = s;
}

class GroupByHashImpl : public Operator {
    KeyLayout layout;
    int hashChannel;
    GroupTableStorage table;
    PageStager stager;
    std::vector<int> keyChannels;
    DeviceBuffer newSlots, slotOfGid, scratchIds, scratchFlags;
    int64_t nextGroupId = 0;
    uint64_t rowsSeen = 0;
    ResultBuilder result;
    std::vector<std::unique_ptr<VarcharInterner>> interners;     // interned VARCHAR keys (KeyLayout::interned)
    std::vector<DeviceBuffer> internIds, internNulls;

    // cols[k] (one per key) with interned VARCHAR keys replaced by their id columns; insert = false only looks values up (contains)
    std::vector<DeviceColumn> internKeys(const std::vector<DeviceColumn>& cols, int64_t n, bool insert)
    {
        std::vector<DeviceColumn> out = cols;
        for (int k = 0; k < layout.nKeys; k++) {
            if (!layout.interned[k]) continue;
            PG_CHECK_ARG(cols[k].type == PG_TYPE_VARCHAR, "group-by channel type differs from the configuration");
            internIds[k].reserve((size_t) n * 4 + 16);
            if (cols[k].encoding != PG_ENC_FLAT) {
                uint8_t* rowNulls = nullptr;
                if (cols[k].nulls) { internNulls[k].accounting = &stats.device_bytes; internNulls[k].reserve((size_t) n + 16); rowNulls = internNulls[k].as<uint8_t>(); }
                interners[k]->internEncoded(this, cols[k], n, internIds[k].as<int32_t>(), rowNulls, insert);
                out[k].nulls = rowNulls; out[k].encoding = PG_ENC_FLAT; out[k].ids = nullptr; out[k].entries = 0;
            }
            else interners[k]->intern(this, cols[k], n, internIds[k].as<int32_t>(), insert);
            out[k].type = PG_TYPE_INTEGER; out[k].values = internIds[k].ptr; out[k].offsets = nullptr;
        }
        return out;
    }

public:
    explicit GroupByHashImpl(const pg_group_by_hash_config* c) : Operator(c->options), stager(this)
    {
        PG_CHECK_ARG(c->expected_size > 0, "expectedSize must be greater than zero");
        layout.plan(c->n_channels, c->channels, c->types, c->max_lengths);
        hashChannel = c->hash_channel;
        for (int k = 0; k < layout.nKeys; k++) keyChannels.push_back(layout.channels[k]);
        table.op = this; table.W = layout.W; table.nvPhys = 0; table.withGid = true;
        table.counters.reserve(CNT_COUNT * sizeof(unsigned int));
        PG_CUDA(cudaMemsetAsync(table.counters.ptr, 0, CNT_COUNT * sizeof(unsigned int), stream));
        table.allocate(std::max<uint32_t>(1024, nextPow2((uint64_t) c->expected_size * 2)));
        interners.resize(layout.nKeys); internIds.resize(layout.nKeys); internNulls.resize(layout.nKeys);
        for (int k = 0; k < layout.nKeys; k++) {
            if (!layout.interned[k]) continue;
            interners[k].reset(new VarcharInterner(this, &stats.device_bytes)); internIds[k].accounting = &stats.device_bytes;
            bool once = true;
            for (int j = 0; j < layout.nKeys; j++) if (j != k && layout.channels[j] == layout.channels[k]) once = false;
            if (once) stager.keepEncoded.push_back(layout.channels[k]);       // dictionary path: intern the dictionary, not the rows
        }
        PG_CUDA(cudaStreamSynchronize(stream));
    }
    bool needsInput() override { return true; }
    void addInput(const pg_page*) override {}
    bool getOutput(pg_page*) override { return false; }
    void finish() override {}
    bool isFinished() override { return false; }

    void ensureCapacity(int64_t extraRows)
    {
        uint64_t need = ((uint64_t) nextGroupId + (uint64_t) extraRows) * 2;
        if (need <= table.cap) return;
        if (need > (1ULL << 30)) throw Error(PG_ERR_INSUFFICIENT_RESOURCES, "Size of hash table cannot exceed 1 billion entries");
        GroupTableStorage nt;
        nt.op = this; nt.W = table.W; nt.nvPhys = 0; nt.withGid = true;
        nt.counters = std::move(table.counters);
        PG_CUDA(cudaMemsetAsync(nt.counters.ptr, 0, CNT_COUNT * sizeof(unsigned int), stream));
        nt.allocate(nextPow2(need));
        GroupTable oldV = table.view(); oldV.counters = nt.counters.as<unsigned int>();
        GroupTable newV = nt.view();
        switch (table.W) {
            case 1: PG_LAUNCH(this, rehashKernel<1>, 1024, 256, 0, oldV, newV); break;
            case 2: PG_LAUNCH(this, rehashKernel<2>, 1024, 256, 0, oldV, newV); break;
            case 3: PG_LAUNCH(this, rehashKernel<3>, 1024, 256, 0, oldV, newV); break;
            default: PG_LAUNCH(this, rehashKernel<4>, 1024, 256, 0, oldV, newV); break;
        }
        if (nextGroupId > 0) PG_LAUNCH(this, gbhRebuildSlotOfGidKernel, 1024, 256, 0, newV, slotOfGid.as<uint32_t>());
        PG_CUDA(cudaStreamSynchronize(stream));
        table.keys = std::move(nt.keys); table.state = std::move(nt.state); table.firstRow = std::move(nt.firstRow);
        table.acc = std::move(nt.acc); table.gid = std::move(nt.gid); table.counters = std::move(nt.counters); table.cap = nt.cap;
    }

    template <typename F> void withW(F&& f)
    {
        switch (layout.W) {
            case 1: f(std::integral_constant<int, 1>()); break;
            case 2: f(std::integral_constant<int, 2>()); break;
            case 3: f(std::integral_constant<int, 3>()); break;
            default: f(std::integral_constant<int, 4>()); break;
        }
    }
    void checkErrors()
    {
        unsigned int cnt[CNT_COUNT];
        PG_CUDA(cudaMemcpyAsync(cnt, table.counters.ptr, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        if (cnt[CNT_ERROR] & ERR_VARCHAR_TOO_LONG) throw Error(PG_ERR_UNSUPPORTED, "VARCHAR group-by key longer than its declared max length (GPU path packs keys of <= 7 bytes)");
    }

    // addPage (idsOut == nullptr) / getGroupIds.  idsOut lives where the page lives.
    void insertPage(const pg_page* page, int64_t* idsOut)
    {
        const int64_t chunk = int64_t(1) << 24;
        for (int64_t begin = 0; begin < page->position_count; begin += chunk) {
            const int64_t n = std::min(chunk, page->position_count - begin);
            pg_page sub = *page;
            std::vector<pg_block> blocks(page->blocks, page->blocks + page->channel_count);
            sub.blocks = blocks.data(); sub.position_count = n;
            for (auto& b : blocks) {
                PG_CHECK_ARG(b.encoding == PG_ENC_FLAT || begin == 0, "encoded blocks larger than 2^24 rows");
                if (b.encoding == PG_ENC_FLAT) {
                    int w = typeWidth(b.type);
                    if (w) b.values = (const char*) b.values + begin * w; else b.offsets += begin;
                    if (b.is_null) b.is_null += begin;
                }
                b.position_count = n;
            }
            stats.input_positions += n;
            std::vector<DeviceColumn> cols;
            stager.stage(&sub, keyChannels, cols);
            if (layout.anyInterned()) cols = internKeys(cols, n, true);
            ensureCapacity(n);
            slotOfGid.accounting = newSlots.accounting = &stats.device_bytes;
            {   // slotOfGid must hold nextGroupId + n entries, preserving what is there
                size_t needBytes = ((size_t) nextGroupId + (size_t) n + 2) * 4;
                if (needBytes > slotOfGid.capacity) {
                    DeviceBuffer bigger; bigger.accounting = &stats.device_bytes; bigger.reserve(needBytes * 2);
                    if (nextGroupId) PG_CUDA(cudaMemcpyAsync(bigger.ptr, slotOfGid.ptr, (size_t) nextGroupId * 4, cudaMemcpyDeviceToDevice, stream));
                    PG_CUDA(cudaStreamSynchronize(stream));
                    slotOfGid = std::move(bigger);
                }
            }
            newSlots.reserve(((size_t) n + 2) * 4);
            KeyPlan kp{};
            layout.fill(kp, cols, table.counters.as<unsigned int>() + CNT_ERROR);
            GroupTable T = table.view();
            unsigned int zero = 0;
            PG_CUDA(cudaMemcpyAsync(T.counters + CNT_NEW, &zero, 4, cudaMemcpyHostToDevice, stream));
            const int grid = (int) std::min<int64_t>(divUp(n, 256), numSMs * 8);
            withW([&](auto w) { PG_LAUNCH(this, gbhInsertKernel<decltype(w)::value>, grid, 256, 0, T, kp, n, rowsSeen, newSlots.as<uint32_t>()); });
            PG_LAUNCH(this, gbhSpecialsKernel, 1, 32, 0, T, newSlots.as<uint32_t>());
            unsigned int nNew = 0;
            PG_CUDA(cudaMemcpyAsync(&nNew, T.counters + CNT_NEW, 4, cudaMemcpyDeviceToHost, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
            if (nNew) {
                DeviceBuffer fk, fkOut, sorted, temp;
                fk.reserve((size_t) nNew * 8); fkOut.reserve((size_t) nNew * 8); sorted.reserve((size_t) nNew * 4);
                PG_LAUNCH(this, gbhGatherFirstRowsKernel, (int) std::min<int64_t>(divUp(nNew, 256), 1024), 256, 0, T, newSlots.as<uint32_t>(), nNew, fk.as<uint64_t>());
                size_t tb = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, tb, fk.as<uint64_t>(), fkOut.as<uint64_t>(), newSlots.as<uint32_t>(), sorted.as<uint32_t>(), (int) nNew, 0, 64, stream);
                temp.reserve(tb);
                cub::DeviceRadixSort::SortPairs(temp.ptr, tb, fk.as<uint64_t>(), fkOut.as<uint64_t>(), newSlots.as<uint32_t>(), sorted.as<uint32_t>(), (int) nNew, 0, 64, stream);
                stats.kernel_launches += 4;
                PG_LAUNCH(this, gbhAssignKernel, (int) std::min<int64_t>(divUp(nNew, 256), 1024), 256, 0, T, sorted.as<uint32_t>(), nNew, (int32_t) nextGroupId, slotOfGid.as<uint32_t>());
                PG_CUDA(cudaStreamSynchronize(stream));
                nextGroupId += nNew;
            }
            if (idsOut) lookupInto(cols, n, idsOut + begin, nullptr, page->memory);
            rowsSeen += (uint64_t) n;
            checkErrors();
        }
        stats.group_count = nextGroupId;
    }

    void lookupInto(const std::vector<DeviceColumn>& cols, int64_t n, int64_t* idsOut, uint8_t* flagsOut, int memory)
    {
        KeyPlan kp{};
        layout.fill(kp, cols, table.counters.as<unsigned int>() + CNT_ERROR);
        GroupTable T = table.view();
        int64_t* dIds = idsOut; uint8_t* dFlags = flagsOut;
        if (memory == PG_MEM_HOST) {
            if (idsOut) { scratchIds.reserve((size_t) n * 8); dIds = scratchIds.as<int64_t>(); }
            if (flagsOut) { scratchFlags.reserve((size_t) n); dFlags = scratchFlags.as<uint8_t>(); }
        }
        const int grid = (int) std::min<int64_t>(divUp(n, 256), numSMs * 8);
        withW([&](auto w) { PG_LAUNCH(this, gbhLookupKernel<decltype(w)::value>, grid, 256, 0, T, kp, n, dIds, dFlags); });
        if (memory == PG_MEM_HOST) {
            if (idsOut) { PG_CUDA(cudaMemcpyAsync(idsOut, dIds, (size_t) n * 8, cudaMemcpyDeviceToHost, stream)); stats.d2h_bytes += n * 8; }
            if (flagsOut) { PG_CUDA(cudaMemcpyAsync(flagsOut, dFlags, (size_t) n, cudaMemcpyDeviceToHost, stream)); stats.d2h_bytes += n; }
        }
        PG_CUDA(cudaStreamSynchronize(stream));
    }

    void contains(const pg_page* page, uint8_t* flagsOut)
    {
        if (page->position_count == 0) return;
        std::vector<DeviceColumn> cols;
        stager.stage(page, keyChannels, cols);
        if (layout.anyInterned()) cols = internKeys(cols, page->position_count, false);
        lookupInto(cols, page->position_count, nullptr, flagsOut, page->memory);
        checkErrors();
    }

    int64_t groupCount() const { return nextGroupId; }

    void getKeys(pg_page* out)
    {
        if (nextGroupId == 0) { out->position_count = 0; out->channel_count = 0; out->memory = options.output_memory; out->blocks = nullptr; return; }
        result.build(this, table.view(), layout, slotOfGid.as<uint32_t>(), nextGroupId, hashChannel >= 0, {}, &interners);
        result.slice(0, nextGroupId, options.output_memory, out);
    }
};


struct FusedJoinShape { int joinType; std::vector<int32_t> hashChannels, outputChannels; };

// Describes a FilterAndProject → LookupJoin* → HashAggregation fragment for the code generator; false + why when it has to run
// operator by operator.  Host only (also behind pg_fused_aggregation_explain, which needs no device).
static bool describeFragment(FusedProgramDesc& D, const std::vector<int32_t>& inputTypes, const std::vector<pg_expr_node>& nodes, int filterNode,
                             const std::vector<int32_t>& projections, const std::vector<FusedJoinShape>& joins, const std::vector<LookupSourceView>& views,
                             const KeyLayout& layout, const std::vector<PhysCol>& phys, const std::vector<bool>& inputHasNulls, std::string& why)
{
    D = FusedProgramDesc();
    D.fp.nodes = nodes; D.fp.inputTypes.assign(inputTypes.begin(), inputTypes.end()); D.fp.inputHasNulls = inputHasNulls;
    D.fp.filterNode = filterNode; D.fp.projections.assign(projections.begin(), projections.end());
    for (size_t j = 0; j < joins.size(); j++) {
        const FusedJoinShape& J = joins[j];
        const LookupSourceView& v = views[j];
        if (J.joinType != PG_JOIN_INNER) { why = "outer join"; return false; }
        if (J.hashChannels.size() != 1 || v.nKeys != 1) { why = "multi-column join key"; return false; }
        if (!v.dense) { why = "lookup source is a hash table (duplicate or sparse keys), not the dense index"; return false; }
        FusedJoinDesc fj;
        fj.keyChannel = J.hashChannels[0];
        fj.probeOutputChannels.assign(J.outputChannels.begin(), J.outputChannels.end());
        fj.buildOutputs = v.buildOutputs; fj.strideWords = v.strideWords;
        D.joins.push_back(fj);
    }
    if (layout.anyInterned()) { why = "VARCHAR group key without a declared length of <= 7 bytes (interned per chunk)"; return false; }
    for (int k = 0; k < layout.nKeys; k++) {
        FusedKeyField f; f.channel = layout.channels[k]; f.type = layout.types[k]; f.word = layout.word[k]; f.shift = layout.shift[k]; f.bits = layout.bits[k]; f.maxLen = layout.maxLens[k];
        D.keys.push_back(f);
    }
    D.bigintPath = layout.bigintPath; D.W = layout.W; D.nullWord = layout.nullWord < 0 ? 0 : layout.nullWord; D.nullShift = layout.nullShift;
    for (size_t p = 0; p < phys.size(); p++) {
        if (!phys[p].active) continue;
        if ((phys[p].kind == K_F64 || phys[p].kind == K_F64_FROM_I64) && phys[p].mask < 0) D.accs.push_back({phys[p].channel, phys[p].kind == K_F64_FROM_I64});
        else if (phys[p].kind == K_COUNT && phys[p].channel < 0 && phys[p].mask < 0 && !D.countRows) D.countRows = true;
        else { why = "aggregate outside the fused set (sums / averages of doubles and count(*))"; return false; }
    }
    return true;
}

// ======================================================================================================================
// FusedAggregationOp — FilterAndProject → LookupJoin* → HashAggregation as one operator (include/presto_gpu.h, fused_agg.cuh,
// fused_codegen.cpp).  It IS the HashAggregationOperator of the chain (same group table, same output path); only the way rows get
// into the table differs: one generated kernel over the raw page instead of three operators and two intermediate pages.
// ======================================================================================================================
class FusedAggregationOp : public HashAggregationOp {
    struct JoinCfg {
        JoinBridge* bridge; int joinType; std::vector<int32_t> types, hashChannels, outputChannels; int precomputedHash;
    };
    std::vector<int32_t> inputTypes, projections;
    std::vector<pg_expr_node> nodes;
    ExprLiterals literals;
    int filterNode;
    std::vector<JoinCfg> joins;
    std::vector<int> rawChannels;                       // raw input channels the fragment reads (staged for the fused kernel)
    struct Compiled { const JitKernel* kernel = nullptr; const JitKernel* kernelTma = nullptr; FusedKernelInfo info; std::string reason; };
    std::map<uint64_t, Compiled> compiled;              // by null signature of the raw page
    std::unique_ptr<Operator> fpOp;
    std::vector<std::unique_ptr<Operator>> joinOps;
    int64_t fusedPages = 0, unfusedPages = 0;
    int lastRegisters = 0;
    std::string lastReason;
    std::vector<LookupSourceView> views;

public:
    explicit FusedAggregationOp(const pg_fused_aggregation_config* c) : HashAggregationOp(&c->aggregation)
    {
        const pg_filter_project_config& f = c->page_processor;
        PG_CHECK_ARG(f.n_input_channels >= 0 && f.n_input_channels <= kFusedMaxInputs, "at most 32 input channels");
        PG_CHECK_ARG(f.n_nodes >= 0 && f.n_nodes <= 64 && f.n_projections >= 0 && f.n_projections <= 32, "page processor program too large");
        PG_CHECK_ARG(c->n_joins >= 0 && c->n_joins <= kFusedMaxJoins, "at most 2 joins in a fused fragment");
        inputTypes.assign(f.input_types, f.input_types + f.n_input_channels);
        nodes.assign(f.nodes, f.nodes + f.n_nodes);
        literals.adopt(nodes);       // the unfused chain builds its FilterAndProjectOperator from `nodes` later
        filterNode = f.filter_node;
        projections.assign(f.projection_nodes, f.projection_nodes + f.n_projections);
        PG_CHECK_ARG(filterNode < f.n_nodes, "filter node out of range");
        for (int p : projections) PG_CHECK_ARG(p >= 0 && p < f.n_nodes, "projection node out of range");
        for (int i = 0; i < f.n_nodes; i++) {
            const pg_expr_node& n = nodes[i];
            if (n.op == PG_EX_INPUT) PG_CHECK_ARG(n.a >= 0 && n.a < f.n_input_channels, "input channel out of range");
            else for (int o : {n.a, n.b, n.c}) PG_CHECK_ARG(o < i, "expression nodes must be topologically ordered");
        }
        for (int j = 0; j < c->n_joins; j++) {
            const pg_lookup_join_config& jc = c->joins[j];
            PG_CHECK_ARG(c->bridges && c->bridges[j], "join bridge is null");
            JoinCfg J;
            J.bridge = (JoinBridge*) c->bridges[j]; J.joinType = jc.join_type; J.precomputedHash = jc.precomputed_hash_channel;
            J.types.assign(jc.types, jc.types + jc.n_channels);
            J.hashChannels.assign(jc.hash_channels, jc.hash_channels + jc.n_hash_channels);
            J.outputChannels.assign(jc.output_channels, jc.output_channels + jc.n_output_channels);
            joins.push_back(J);
        }
        PG_CHECK_ARG(step != PG_STEP_FINAL, "a fused fragment aggregates raw rows: step must be SINGLE or PARTIAL");
        for (int ch = 0; ch < f.n_input_channels; ch++) {
            bool used = false;
            for (const auto& n : nodes) if (n.op == PG_EX_INPUT && n.a == ch) used = true;
            if (used) rawChannels.push_back(ch);
        }
        views.resize(joins.size());
    }

    bool isBlocked() override
    {
        for (size_t j = 0; j < joins.size(); j++) { if (!views[j].ready) describeLookupSource(joins[j].bridge, views[j]); if (!views[j].ready) return true; }
        return false;
    }
    // LookupJoinOperator.needsInput :198-210: blocked until every lookup source future completes
    bool needsInput() override
    {
        if (!HashAggregationOp::needsInput()) return false;
        for (size_t j = 0; j < joins.size(); j++) { if (!views[j].ready) describeLookupSource(joins[j].bridge, views[j]); if (!views[j].ready) return false; }
        return true;
    }

    // Describes the fragment for the code generator; false + why when it has to run operator by operator.
    bool describe(FusedProgramDesc& D, const std::vector<bool>& inputHasNulls, std::string& why) const
    {
        std::vector<FusedJoinShape> shapes;
        for (const auto& J : joins) shapes.push_back({J.joinType, J.hashChannels, J.outputChannels});
        return describeFragment(D, inputTypes, nodes, filterNode, projections, shapes, views, layout, phys, inputHasNulls, why);
    }
    void physMapping(int32_t (&out)[kFusedMaxSums + 1]) const
    {
        int nv = 0, cnt = -1;
        for (size_t p = 0; p < phys.size(); p++) {
            if (!phys[p].active) continue;
            if (phys[p].kind == K_F64 || phys[p].kind == K_F64_FROM_I64) out[nv++] = (int) p; else cnt = (int) p;
        }
        out[nv] = cnt < 0 ? 0 : cnt;
    }

    Compiled& kernelFor(const pg_page* page)
    {
        uint64_t sig = 0;
        std::vector<bool> hasNulls(inputTypes.size(), false);
        for (int ch : rawChannels) if (page->blocks[ch].is_null) { sig |= 1ULL << ch; hasNulls[ch] = true; }
        auto it = compiled.find(sig);
        if (it != compiled.end()) return it->second;
        Compiled& c = compiled[sig];
        const char* off = getenv("PG_FUSED");
        if (off && off[0] == '0') { c.reason = "PG_FUSED=0"; return c; }
        FusedProgramDesc D;
        if (!describe(D, hasNulls, c.reason)) return c;
        const std::string src = generateFusedAggSource(D, c.info);
        if (!c.info.ok) { c.reason = c.info.reason; return c; }
        std::string err;
        c.kernel = getOrCompileKernel(options.device, src, "fused_kernel", &err);
        if (!c.kernel) c.reason = err;
        const char* tmaOff = getenv("PG_FUSED_TMA");
        if (c.kernel && c.info.tma && !(tmaOff && tmaOff[0] == '0')) c.kernelTma = getOrCompileKernel(options.device, src, "fused_kernel_tma", &err);
        return c;
    }

    void addInput(const pg_page* page) override
    {
        PG_CHECK_STATE(needsInput(), "fused aggregation does not need input (lookup sources still building, or already finishing)");
        PG_CHECK_ARG(page->channel_count == (int) inputTypes.size(), "page channel count differs from the page processor's input types");
        if (page->position_count == 0) return;
        for (int ch : rawChannels) PG_CHECK_ARG(page->blocks[ch].type == inputTypes[ch], "input channel type differs from the operator's configuration");
        Compiled& c = kernelFor(page);
        bool fusable = c.kernel != nullptr;
        if (fusable && page->memory == PG_MEM_DEVICE) {
            // eagerly loaded columns come in as 128-bit pairs: device pages must be 16-byte aligned (staged host pages always are)
            for (int ch : c.info.eagerInputs) if (((uintptr_t) page->blocks[ch].values % 16) != 0 || page->blocks[ch].encoding != PG_ENC_FLAT) { fusable = false; lastReason = "misaligned or encoded device column"; }
        }
        if (!fusable) { if (!c.reason.empty()) lastReason = c.reason; addInputUnfused(page); return; }
        lastRegisters = c.kernel->registers;
        int32_t physMap[kFusedMaxSums + 1] = {0};
        physMapping(physMap);
        const CUfunction fn = c.kernel->function;
        DriverApi& api = DriverApi::get();
        int smemPerSM = 0;
        PG_CUDA(cudaDeviceGetAttribute(&smemPerSM, cudaDevAttrMaxSharedMemoryPerMultiprocessor, options.device));
        fusedLaunch = [&, fn, smemPerSM](const std::vector<DeviceColumn>& cols, int64_t n, const uint32_t* bitmapIn, uint32_t* bitmapOut) {
            FusedParams L{};
            L.t = table.view();
            L.n = n; L.rowBase = rowsSeen; L.bitmapIn = bitmapIn; L.bitmapOut = bitmapOut;
            L.S = sharedSlots; L.sLimit = sharedLimit; L.nAccAlloc = (int) phys.size();
            L.direct = (!c.info.tiers || bitmapIn != nullptr || knownGroups > (int64_t) sharedLimit * 9 / 10) ? 1 : 0;
            memcpy(L.phys, physMap, sizeof(physMap));
            for (size_t i = 0; i < rawChannels.size(); i++) {
                L.values[rawChannels[i]] = cols[i].values; L.nulls[rawChannels[i]] = cols[i].nulls; L.offsets[rawChannels[i]] = cols[i].offsets;
            }
            for (size_t j = 0; j < joins.size(); j++) { L.joins[j].D = views[j].D; L.joins[j].rows = views[j].rows; }
            L.fpError = table.counters.as<unsigned int>() + CNT_FPERR;
            const int NC = c.info.nv + (c.info.cnt ? 1 : 0);
            const int k0max = std::min(31, kTier0Columns / std::max(NC, 1));
            L.t0Groups = k0max;
            const int64_t tiles = divUp(n, 256 * c.info.rowsPerThread);
            // ---- TMA ring: once the page's group count is known (probe chunk), tier 0 is sized for exactly those groups and the rest of
            // the shared-memory budget of two resident CTAs becomes cp.async.bulk stages for the eager columns ----
            CUfunction launchFn = fn;
            unsigned block = 256;
            size_t smem = (L.direct ? 64 : sharedBytes(sharedSlots)) + c.info.queueBytes;
            // (measured on Q3's lineitem fragment: the ring is SLOWER than the register double buffer in direct mode — 3.46 vs 2.78 ms; the
            // consumers of a selective fragment stall on the probe's L2 loads, not on the tile loads — so it stays opt-in there)
            const char* tmaJoins = getenv("PG_FUSED_TMA_DIRECT");
            if (c.kernelTma && L.direct && !bitmapIn && tiles > 4 * (int64_t) numSMs && tmaJoins && tmaJoins[0] == '1') {
                // direct mode (every survivor goes to the HBM table): no tier layout, only the survivor queue sits in front of the ring
                const size_t lay = (c.info.queueBytes + 127) & ~size_t(127);
                const int stages = 4;
                L.tmaStages = stages; L.tmaTiles = n / (256 * c.info.rowsPerThread) - 1;
                launchFn = c.kernelTma->function; block = 256 + 32;
                smem = lay + (size_t) stages * ((size_t) c.info.tmaStageBytes + 16);
            }
            else if (c.kernelTma && !L.direct && knownGroups >= 1 && knownGroups <= k0max && tiles > 4 * (int64_t) numSMs) {
                const int t0 = (int) knownGroups;
                // a handful of groups does not need the full shared table: 4 slots per known group leave the ring more room
                uint32_t slots = 64;
                while (slots < 4 * (uint32_t) t0 && slots < sharedSlots) slots <<= 1;
                const size_t lay = (SharedLayout(slots, 1, (int) phys.size(), (uint32_t) (t0 * NC * 256 * 8), 0).total + 127) & ~size_t(127);
                const size_t budget = (size_t) smemPerSM / 2 - 1024 - 64;
                const int stages = lay < budget ? (int) std::min<size_t>(4, (budget - lay) / ((size_t) c.info.tmaStageBytes + 16)) : 0;
                if (stages >= 2) {
                    L.t0Groups = t0; L.tmaStages = stages; L.tmaTiles = n / (256 * c.info.rowsPerThread) - 1;
                    L.S = slots; L.sLimit = slots / 4 * 3;
                    launchFn = c.kernelTma->function; block = 256 + 32;
                    smem = lay + (size_t) stages * ((size_t) c.info.tmaStageBytes + 16);
                }
            }
            if (api.funcSetAttribute(launchFn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 200 * 1024) != CUDA_SUCCESS) throw Error(PG_ERR_CUDA, "cuFuncSetAttribute failed for a fused pipeline kernel");
            int perSM = 0;
            api.occupancyMaxActiveBlocks(&perSM, launchFn, (int) block, smem);
            perSM = std::max(1, std::min(perSM, L.direct ? 8 : 2));    // flushReserve of the table was sized for 2 resident CTAs per SM with a shared table
            const unsigned grid = (unsigned) std::min<int64_t>((int64_t) numSMs * perSM, std::max<int64_t>(tiles, 1));
            void* args[1] = {&L};
            const CUresult r = api.launchKernel(launchFn, grid, 1, 1, block, 1, 1, (unsigned) smem, (CUstream) stream, args, nullptr);
            if (r != CUDA_SUCCESS) throw Error(PG_ERR_CUDA, "cuLaunchKernel of a fused pipeline kernel failed (" + std::to_string((int) r) + ")");
            stats.kernel_launches++;
            stats.kernel_paths |= (L.direct ? PG_PATH_FUSED_DIRECT : PG_PATH_FUSED_TIERS) | (L.tmaStages ? PG_PATH_FUSED_TMA : 0);
            lastRegisters = L.tmaStages ? c.kernelTma->registers : c.kernel->registers;
        };
        inputChannels = &rawChannels;
        struct Reset { FusedAggregationOp* o; ~Reset() { o->fusedLaunch = nullptr; o->inputChannels = nullptr; } } reset{this};
        HashAggregationOp::addInput(page);      // chunking, staging, deferred rows and growth are the aggregation's own
        fusedPages++;
    }

    // The same fragment operator by operator (still on the GPU): FilterAndProjectOperator → LookupJoinOperator* → this aggregation.
    void addInputUnfused(const pg_page* page)
    {
        pg_operator_options o = options;
        o.output_memory = PG_MEM_DEVICE; o.cuda_stream = (void*) stream;
        if (!fpOp) {
            pg_filter_project_config f{};
            f.n_input_channels = (int) inputTypes.size(); f.input_types = inputTypes.data(); f.n_nodes = (int) nodes.size(); f.nodes = nodes.data();
            f.filter_node = filterNode; f.n_projections = (int) projections.size(); f.projection_nodes = projections.data(); f.options = o;
            fpOp.reset(createFilterProject(&f));
            for (auto& J : joins) {
                pg_lookup_join_config jc{};
                jc.join_type = J.joinType; jc.n_channels = (int) J.types.size(); jc.types = J.types.data();
                jc.n_hash_channels = (int) J.hashChannels.size(); jc.hash_channels = J.hashChannels.data();
                jc.n_output_channels = (int) J.outputChannels.size(); jc.output_channels = J.outputChannels.data();
                jc.precomputed_hash_channel = J.precomputedHash; jc.options = o;
                joinOps.emplace_back(createLookupJoin(&jc, J.bridge));
            }
        }
        unfusedPages++;
        stats.kernel_paths |= PG_PATH_UNFUSED_CHAIN;
        const int64_t before = stats.input_positions;
        fpOp->addInput(page);
        pg_page cur{};
        bool has = fpOp->getOutput(&cur);
        for (auto& j : joinOps) {
            if (!has || cur.position_count == 0) break;
            j->addInput(&cur);
            has = j->getOutput(&cur);
        }
        if (has && cur.position_count > 0) HashAggregationOp::addInput(&cur);
        stats.input_positions = before + page->position_count;     // rows of the RAW page, like the fused path
    }

    void syncStats() override
    {
        HashAggregationOp::syncStats();
    }
    int64_t chainLaunches() const
    {
        int64_t n = fpOp ? fpOp->stats.kernel_launches : 0;
        for (auto& j : joinOps) n += j->stats.kernel_launches;
        return n;
    }
    void info(int64_t* fused, int64_t* unfused, int32_t* regs, char* reason, size_t cap) const
    {
        if (fused) *fused = fusedPages;
        if (unfused) *unfused = unfusedPages;
        if (regs) *regs = lastRegisters;
        if (reason && cap) snprintf(reason, cap, "%s", lastReason.c_str());
    }
};

Operator* createHashAggregation(const pg_hash_aggregation_config* c) { return new HashAggregationOp(c); }
Operator* createFusedAggregation(const pg_fused_aggregation_config* c) { return new FusedAggregationOp(c); }

}  // namespace pg

extern "C" int32_t pg_hash_aggregation_create(const pg_hash_aggregation_config* config, pg_operator** out)
{
    try {
        if (!config || !out) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "null argument");
        *out = (pg_operator*) pg::createHashAggregation(config);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}

extern "C" int32_t pg_hash_aggregation_state_bytes(pg_operator* op, int64_t max_groups, size_t* bytes)
{
    try {
        auto* p = dynamic_cast<pg::HashAggregationOp*>((pg::Operator*) op);
        if (!p || !bytes) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a hash aggregation operator");
        *bytes = p->stateBytes(max_groups);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    return PG_OK;
}
extern "C" int32_t pg_hash_aggregation_export_state(pg_operator* op, void* device_buffer, size_t capacity, int64_t max_groups)
{
    try {
        auto* p = dynamic_cast<pg::HashAggregationOp*>((pg::Operator*) op);
        if (!p) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a hash aggregation operator");
        p->bind();
        p->exportState(device_buffer, capacity, max_groups);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
extern "C" int32_t pg_hash_aggregation_merge_states(pg_operator* op, const void* device_buffers, int32_t count, size_t stride, int64_t max_groups, int32_t order_base)
{
    try {
        auto* p = dynamic_cast<pg::HashAggregationOp*>((pg::Operator*) op);
        if (!p) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a hash aggregation operator");
        p->bind();
        p->mergeStates(device_buffers, count, stride, max_groups, order_base);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}

extern "C" int32_t pg_fused_aggregation_create(const pg_fused_aggregation_config* config, pg_operator** out)
{
    try {
        if (!config || !out) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "null argument");
        *out = (pg_operator*) pg::createFusedAggregation(config);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}
extern "C" int32_t pg_fused_aggregation_get_info(pg_operator* op, int64_t* fused, int64_t* unfused, int32_t* registers, char* reason, size_t capacity)
{
    try {
        auto* p = dynamic_cast<pg::FusedAggregationOp*>((pg::Operator*) op);
        if (!p) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "not a fused aggregation operator");
        p->info(fused, unfused, registers, reason, capacity);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    return PG_OK;
}

// Payload row layout HashBuilderOperator gives a dense index (hash_join.cu tryBuildDense): widest columns first, no NULL bytes here.
extern "C" int32_t pg_fused_aggregation_explain(const pg_fused_aggregation_config* c, uint64_t nullableInputs, const int32_t* const* buildTypes,
                                                const int32_t* nBuildTypes, char* buffer, size_t capacity, size_t* needed)
{
    try {
        using namespace pg;
        PG_CHECK_ARG(c && needed, "null argument");
        const pg_filter_project_config& f = c->page_processor;
        std::vector<int32_t> inputTypes(f.input_types, f.input_types + f.n_input_channels), projections(f.projection_nodes, f.projection_nodes + f.n_projections);
        std::vector<pg_expr_node> nodes(f.nodes, f.nodes + f.n_nodes);
        std::vector<FusedJoinShape> shapes; std::vector<LookupSourceView> views;
        for (int j = 0; j < c->n_joins; j++) {
            const pg_lookup_join_config& jc = c->joins[j];
            shapes.push_back({jc.join_type, std::vector<int32_t>(jc.hash_channels, jc.hash_channels + jc.n_hash_channels), std::vector<int32_t>(jc.output_channels, jc.output_channels + jc.n_output_channels)});
            LookupSourceView v; v.ready = true; v.dense = true; v.nKeys = 1; v.keyType = PG_TYPE_BIGINT;
            int off = 0;
            const int nb = nBuildTypes ? nBuildTypes[j] : 0;
            v.buildOutputs.resize(nb);
            for (int width : {8, 4, 1})
                for (int i = 0; i < nb; i++)
                    if (typeWidth(buildTypes[j][i]) == width) { FusedBuildColumn& b = v.buildOutputs[i]; b.type = buildTypes[j][i]; b.width = width; b.word = off >> 3; b.shift = 8 * (off & 7); off += width; }
            v.strideWords = nb ? std::max(1, (off + 7) / 8) : 0;
            if (v.strideWords == 3) v.strideWords = 4;
            views.push_back(v);
        }
        KeyLayout layout; std::vector<AggPlan> aggs; std::vector<PhysCol> phys; int rowCountCol = -1;
        planAggregation(&c->aggregation, layout, aggs, phys, rowCountCol);
        std::vector<bool> hasNulls(inputTypes.size(), false);
        for (size_t i = 0; i < inputTypes.size(); i++) hasNulls[i] = (nullableInputs >> i) & 1ULL;
        FusedProgramDesc D; std::string why;
        if (!describeFragment(D, inputTypes, nodes, f.filter_node, projections, shapes, views, layout, phys, hasNulls, why)) throw Error(PG_ERR_UNSUPPORTED, "fragment is not fusable: " + why);
        FusedKernelInfo info;
        const std::string src = generateFusedAggSource(D, info);
        if (!info.ok) throw Error(PG_ERR_UNSUPPORTED, "fragment is not fusable: " + info.reason);
        *needed = src.size() + 1;
        if (buffer && capacity >= src.size() + 1) memcpy(buffer, src.c_str(), src.size() + 1);
    }
    catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; }
    catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; }
    return PG_OK;
}

extern "C" {
#define GBH_TRY try {
#define GBH_CATCH } catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; } catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; } return PG_OK;

int32_t pg_group_by_hash_create(const pg_group_by_hash_config* config, pg_group_by_hash** out)
{
    GBH_TRY
    if (!config || !out) throw pg::Error(PG_ERR_INVALID_ARGUMENT, "null argument");
    *out = (pg_group_by_hash*) new pg::GroupByHashImpl(config);
    GBH_CATCH
}
int32_t pg_group_by_hash_add_page(pg_group_by_hash* h, const pg_page* page) { GBH_TRY auto* g = (pg::GroupByHashImpl*) h; g->bind(); g->insertPage(page, nullptr); GBH_CATCH }
int32_t pg_group_by_hash_get_group_ids(pg_group_by_hash* h, const pg_page* page, int64_t* out) { GBH_TRY auto* g = (pg::GroupByHashImpl*) h; g->bind(); g->insertPage(page, out); GBH_CATCH }
int32_t pg_group_by_hash_get_group_count(pg_group_by_hash* h, int64_t* out) { GBH_TRY *out = ((pg::GroupByHashImpl*) h)->groupCount(); GBH_CATCH }
int32_t pg_group_by_hash_contains(pg_group_by_hash* h, const pg_page* page, uint8_t* out) { GBH_TRY auto* g = (pg::GroupByHashImpl*) h; g->bind(); g->contains(page, out); GBH_CATCH }
int32_t pg_group_by_hash_get_keys(pg_group_by_hash* h, pg_page* out) { GBH_TRY auto* g = (pg::GroupByHashImpl*) h; g->bind(); g->getKeys(out); GBH_CATCH }
int32_t pg_group_by_hash_destroy(pg_group_by_hash* h) { GBH_TRY auto* g = (pg::GroupByHashImpl*) h; if (g) { cudaSetDevice(g->options.device); if (g->stream) cudaStreamSynchronize(g->stream); delete g; } GBH_CATCH }
}
