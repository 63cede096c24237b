// presto_b200 — the kernel skeleton of a fused pipeline fragment: FilterAndProject → LookupJoin* → HashAggregation in ONE pass over
// the raw input page (reference: the operators of one Driver, MB/operator/Driver.java:402-536 chained by
// LocalExecutionPlanner.java:1633-1666 / 2650-2686 / 3544-3640; InMemoryHashAggregationBuilder.processPage :204-220 as the sink).
//
// This header is compiled twice: by nvcc into libprestogpu.so (FusedParams is the launch ABI the host fills) and by NVRTC at run time
// together with a generated `struct Prog` (fused_codegen.cpp) — the row program of the fragment: tile loads, filter, dense-index
// probes, key packing.  What is fixed lives here: the persistent tile loop with a register double buffer, and the aggregation sink
// with the same three tiers as hashAggFastKernel (hash_agg.cu): per-thread private accumulators in shared memory for the first groups
// a CTA sees (TPC-H Q1: every row ends there), the CTA's shared table, the HBM table.
#pragma once
#include "agg_tiers.cuh"
#include "dense_index.cuh"

namespace pg {

constexpr int kFusedMaxInputs = 32;
constexpr int kFusedMaxJoins = 2;
constexpr int kFusedMaxSums = 8;

struct FusedJoinView {
    DenseIndex D;
    const uint64_t* rows;            // payload rows in key order (strideWords words each); nullptr when the join has no build output
};

struct FusedParams {
    GroupTable t;
    int64_t n;                       // rows of this launch
    uint64_t rowBase;                // global index of row 0 (first-appearance bookkeeping)
    const uint32_t* bitmapIn;        // when set: only rows whose bit is set are processed (deferred-row pass)
    uint32_t* bitmapOut;             // rows that found the table at its claim limit
    uint32_t S, sLimit;              // shared table slots / fill limit (W == 1, !direct)
    int32_t nAccAlloc;               // accumulator columns the shared table was sized for
    int32_t direct;                  // 1: no on-chip tiers, every row goes to the HBM table (many groups, or W > 1)
    int32_t t0Groups;                // groups that get per-thread private accumulators (tier 0): <= kTier0Columns / accumulator columns
    int32_t tmaStages;               // > 0: eager columns arrive through a cp.async.bulk ring of this many stages (fused_kernel_tma)
    int64_t tmaTiles;                // tiles [0, tmaTiles) come through the ring; the tail (and partial tiles) through direct loads
    int32_t phys[kFusedMaxSums + 1]; // physical accumulator column of each sum, then of the row count
    const void* values[kFusedMaxInputs];     // raw input channels: fixed-width values or VARCHAR bytes
    const uint8_t* nulls[kFusedMaxInputs];
    const int32_t* offsets[kFusedMaxInputs]; // VARCHAR
    FusedJoinView joins[kFusedMaxJoins];
    unsigned int* fpError;           // bit 0 overflow, bit 1 division by zero (PageProcessor), bit 2 VARCHAR key longer than declared
};

// ---- page sink: FilterAndProject → LookupJoin* materialised as an output page, rows in input order --------------------------------
// Two launches of the same generated row program: the selection pass leaves, per 64-row unit (one warp, one row pair per lane), the
// survivors among its even and odd rows and their count; after an exclusive scan of the counts the emit pass re-runs the probes for
// the survivors only and stores every output channel at its final position — the compacted page is the only thing written to HBM.
struct FusedPageParams {
    int64_t n;
    const void* values[kFusedMaxInputs];
    const uint8_t* nulls[kFusedMaxInputs];
    const int32_t* offsets[kFusedMaxInputs];
    FusedJoinView joins[kFusedMaxJoins];
    unsigned int* fpError;
    uint32_t* masks;                 // [units][2]
    uint32_t* counts;                // [units + 1]
    const uint32_t* unitOffsets;     // emit pass: exclusive scan of counts
    void* outValues[kFusedMaxInputs];
    uint8_t* outNulls[kFusedMaxInputs];
};

#ifdef __CUDACC__
template <class Prog>
__device__ __forceinline__ void fusedSelectBody(const FusedPageParams& P)
{
    constexpr int TILE = Prog::THREADS * Prog::R;
    const long long nTiles = (P.n + TILE - 1) / TILE;
    typename Prog::Tile tA, tB;
    long long tile = blockIdx.x;
    if (tile < nTiles) Prog::load(P, tile * TILE, tA);
    while (tile < nTiles) {
        const long long t1 = tile + gridDim.x;
        if (t1 < nTiles) Prog::load(P, t1 * TILE, tB);
        Prog::select(P, tA, tile * TILE);
        __syncwarp();
        if (t1 >= nTiles) break;
        const long long t2 = t1 + gridDim.x;
        if (t2 < nTiles) Prog::load(P, t2 * TILE, tA);
        Prog::select(P, tB, t1 * TILE);
        __syncwarp();
        tile = t2;
    }
}
template <class Prog>
__device__ __forceinline__ void fusedEmitBody(const FusedPageParams& P)
{
    constexpr int TILE = Prog::THREADS * Prog::R;
    const long long nTiles = (P.n + TILE - 1) / TILE;
    for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) { Prog::emit(P, tile * TILE); __syncwarp(); }
}

// ---- the aggregation sink -----------------------------------------------------------------------------------------------------
// The on-chip tiers work on ONE 64-bit word per group: the packed key itself when W == 1, otherwise (Prog::TIERS) the key's fields
// squeezed together by the generated Prog::tierKey — e.g. Q1's two INTEGER codes, which the W-word layout spreads over two words to
// make room for the NULL flags no row of the page uses.  Prog::expandKey restores the W-word key when a group is flushed to the HBM table.
template <class Prog>
struct FusedAggSink {
    static constexpr int W = Prog::W, NV = Prog::NV, THREADS = Prog::THREADS;
    static constexpr bool CNT = Prog::CNT;
    static constexpr int NC = NV + (CNT ? 1 : 0);
    static constexpr int NVA = NV > 0 ? NV : 1;
    static constexpr int K0MAX = (kTier0Columns / NC) > 31 ? 31 : (kTier0Columns / NC);
    const FusedParams& P;
    int K0;                          // tier-0 groups of this launch (P.t0Groups)
    SharedTable<1> st;
    uint64_t* tier0;                 // [K0][NC][THREADS]
    uint32_t seen0, probeHits, probeMisses;
    bool bypass, tiers;
    int tid;

    __device__ __forceinline__ FusedAggSink(const FusedParams& p) : P(p) {}

    __device__ __forceinline__ void init(unsigned char* smemRaw)
    {
        tid = threadIdx.x;
        seen0 = 0; probeHits = 0; probeMisses = 0; bypass = false;
        tiers = Prog::TIERS && !P.direct;
        K0 = min(max(P.t0Groups, 0), K0MAX);
        const SharedLayout lay(P.S, 1, P.nAccAlloc, (uint32_t) (K0 * NC * THREADS * 8), 0);
        if (Prog::QUEUE) initQueue(smemRaw + (tiers ? lay.total : 0));
        if (!tiers) return;
        st.S = P.S; st.S2 = P.S + 2; st.sLimit = P.sLimit; st.tier0Groups = K0;
        st.keys = (volatile uint64_t*) (smemRaw + lay.keys);
        st.acc = (uint64_t*) (smemRaw + lay.acc);
        st.first = (uint32_t*) (smemRaw + lay.first);
        st.state = nullptr;
        st.sid = (volatile uint8_t*) (smemRaw + lay.sid);
        st.count = (uint32_t*) (smemRaw + lay.misc);
        tier0 = (uint64_t*) (smemRaw + lay.tier0);
        if (tid < THREADS) {
            for (uint32_t i = tid; i < st.S2; i += THREADS) { st.keys[i] = kEmptyKey; st.first[i] = kNoRow32; st.sid[i] = 0xFF; }
            for (uint32_t i = tid; i < (uint32_t) P.nAccAlloc * st.S2; i += THREADS) st.acc[i] = 0;
            for (int i = tid; i < 320; i += THREADS) st.count[i] = 0;
        }
        if (tid < THREADS) for (int i = 0; i < K0 * NC; i++) tier0[i * THREADS + tid] = 0;
    }
    // bytes of the tier layout of this launch (a TMA ring, when used, starts behind it)
    __device__ __forceinline__ uint32_t layoutBytes() const { return tiers ? SharedLayout(P.S, 1, P.nAccAlloc, (uint32_t) (K0 * NC * THREADS * 8), 0).total : 0; }

    // the HBM table (tier 2 / direct mode)
    __device__ __forceinline__ void addGlobal(const uint64_t (&kw)[W], int special, uint32_t spell, const uint64_t (&vals)[NVA], uint32_t rowIdx)
    {
        const GroupTable& T = P.t;
        const int gslot = special ? (int) T.cap + special - 1 : tableFind<W>(T, kw, hashKeyWords<W>(kw), true, false);
        if (gslot == -2) {
            atomicOr(&P.bitmapOut[rowIdx >> 5], 1u << (rowIdx & 31));
            atomicAdd(&T.counters[CNT_DEFERRED], 1u);
            return;
        }
        const uint64_t g = firstRowWord(P.rowBase + rowIdx, spell);
        if (g < __ldcg(&T.firstRow[gslot])) atomicMin((unsigned long long*) &T.firstRow[gslot], (unsigned long long) g);
        uint64_t* acc = T.acc + (size_t) gslot * T.nvPhys;
#pragma unroll
        for (int c = 0; c < NV; c++) atomicAdd((double*) &acc[P.phys[c]], __longlong_as_double((long long) vals[c]));
        if (CNT) atomicAdd((unsigned long long*) &acc[P.phys[NV]], 1ULL);
    }

    // Slot function of the CTA's shared table.  Nothing observable depends on it (group ids come from first-touch rows), so it is the
    // cheapest mix that separates small codes in either half of the key: two 32-bit multiplies (murmur64 costs ~25 instructions per row).
    __device__ __forceinline__ static uint64_t tierHash(uint64_t key)
    {
        const uint32_t m = ((uint32_t) key * 0x9E3779B1u) ^ ((uint32_t) (key >> 32) * 0x85EBCA6Bu);
        return (uint64_t) (m ^ (m >> 15)) << 32;        // SharedTable::find takes the slot from bits 32.. of the hash
    }

    // Everything that is not "the key sits in its home slot of the shared table and owns a tier-0 id": probing / claiming a shared
    // slot, tier 1 (shared-space reductions), the HBM table.  (Inlined: a noinline member would force the whole sink — table pointers,
    // counters, the row arrays — into local memory, which cost 2x on Q1.)
    __device__ __forceinline__ void addSlow(const uint64_t (&kw)[W], int special, uint32_t spell, const uint64_t (&vals)[NVA], uint32_t rowIdx)
    {
        if (!Prog::TIERS || !tiers || spell) { addGlobal(kw, special, spell, vals, rowIdx); return; }
        const uint64_t key = Prog::tierKey(kw);
        if (W > 1 && key == kEmptyKey) { addGlobal(kw, special, spell, vals, rowIdx); return; }     // a squeezed key that looks like an empty slot
        const uint64_t h = tierHash(key);
        int sslot = -1;
        if (special) sslot = (int) st.S + special - 1;        // NULL key of the Bigint path / the key equal to the empty sentinel: dedicated slots
        else if (!bypass) {
            uint64_t k1[1] = {key};
            sslot = st.find(k1, h);
            // a thread whose rows keep missing a full shared table stops probing it (high-cardinality input)
            if (sslot >= 0) probeHits++; else probeMisses++;
            if (((probeHits + probeMisses) & 63u) == 0) {
                if (probeMisses > 3 * probeHits && *(volatile uint32_t*) st.count >= st.sLimit) bypass = true;
                probeHits = 0; probeMisses = 0;
            }
        }
        if (sslot < 0) { addGlobal(kw, special, spell, vals, rowIdx); return; }
        const uint32_t id = st.sid[sslot];       // the two dedicated slots never get a tier-0 id: their rows take tier 1
        if (id < (uint32_t) K0) { addTier0(id, (uint32_t) sslot, vals, rowIdx); return; }
        // tier 1: the CTA's shared table, shared-space reductions
#pragma unroll
        for (int c = 0; c < NV; c++) sharedAddF64(&st.acc[c * st.S2 + sslot], vals[c]);
        if (CNT) sharedAddU64(&st.acc[NV * st.S2 + sslot], 1);
        if (rowIdx < *(volatile uint32_t*) &st.first[sslot]) sharedMinU32(&st.first[sslot], rowIdx);
    }

    // tier 0: this thread's private accumulators of group `id` (consecutive threads, consecutive words: no bank conflicts)
    __device__ __forceinline__ void addTier0(uint32_t id, uint32_t sslot, const uint64_t (&vals)[NVA], uint32_t rowIdx)
    {
        uint64_t* mine = tier0 + (size_t) id * NC * THREADS + tid;
#pragma unroll
        for (int c = 0; c < NV; c++) {
            double* p = (double*) (mine + c * THREADS);
            *p = *p + __longlong_as_double((long long) vals[c]);
        }
        if (CNT) mine[NV * THREADS] += 1;
        if (Prog::QUEUE) { if (rowIdx < *(volatile uint32_t*) &st.first[sslot]) atomicMin(&st.first[sslot], rowIdx); }   // rows reach a thread in queue order
        else if (!((seen0 >> id) & 1u)) { seen0 |= 1u << id; atomicMin(&st.first[sslot], rowIdx); }                      // rows ascend per thread: first sight = its minimum
    }

    __device__ __forceinline__ void add(const uint64_t (&kw)[W], int special, uint32_t spell, const uint64_t (&vals)[NVA], uint32_t rowIdx)
    {
        if (Prog::TIERS && tiers) {
            const uint64_t key = Prog::tierKey(kw);
            const uint32_t s = (uint32_t) (tierHash(key) >> 32) & (st.S - 1);
            if (!special && !spell && st.keys[s] == key && (!Prog::TIERKEY_CAN_BE_EMPTY || key != kEmptyKey)) {
                const uint32_t id = st.sid[s];
                if (id < (uint32_t) K0) { addTier0(id, s, vals, rowIdx); return; }
            }
        }
        addSlow(kw, special, spell, vals, rowIdx);
    }

    // ---- survivor queue (Prog::QUEUE: fragments with joins) ----------------------------------------------------------------------
    // After a selective probe only a few lanes of a warp still hold a row; running the aggregation (a table probe plus atomics) under
    // that divergence costs a full warp instruction stream per survivor.  Survivors are queued per warp in shared memory instead and
    // aggregated 32 at a time with every lane busy (the same idea as hashAggDirectKernel's owned-row queue).
    static constexpr int kQueueSlots = 64;             // up to 31 left over + 32 new per push
    struct QueueEntry { uint64_t kw[W]; uint64_t vals[NVA]; uint32_t rowIdx; uint32_t meta; };   // meta = special | spell << 8
    QueueEntry* queue;                                 // this warp's slots
    uint32_t qCount;                                   // warp-uniform

    __device__ __forceinline__ static size_t queueBytes() { return Prog::QUEUE ? sizeof(QueueEntry) * kQueueSlots * (THREADS / 32) : 0; }
    __device__ __forceinline__ void initQueue(unsigned char* base) { queue = (QueueEntry*) base + (size_t) (tid >> 5) * kQueueSlots; qCount = 0; }

    __device__ __forceinline__ void drain(uint32_t keep)
    {
        __syncwarp();
        const int lane = tid & 31;
        while (qCount > keep) {
            const uint32_t take = min(32u, qCount), at = qCount - take + lane;
            if (lane < (int) take) {
                const QueueEntry& e = queue[at];
                uint64_t kw[W], vals[NVA];
#pragma unroll
                for (int w = 0; w < W; w++) kw[w] = e.kw[w];
#pragma unroll
                for (int c = 0; c < NVA; c++) vals[c] = e.vals[c];
                add(kw, (int) (e.meta & 0xff), e.meta >> 8, vals, e.rowIdx);
            }
            qCount -= take;
            __syncwarp();
        }
    }
    // called by ALL lanes of the warp (live = this lane has a surviving row; m = the warp's ballot of `live`, not zero)
    __device__ __forceinline__ void push(unsigned m, bool live, const uint64_t (&kw)[W], int special, uint32_t spell, const uint64_t (&vals)[NVA], uint32_t rowIdx)
    {
        if (live) {
            QueueEntry& e = queue[qCount + __popc(m & ((1u << (tid & 31)) - 1))];
#pragma unroll
            for (int w = 0; w < W; w++) e.kw[w] = kw[w];
#pragma unroll
            for (int c = 0; c < NVA; c++) e.vals[c] = vals[c];
            e.rowIdx = rowIdx; e.meta = (uint32_t) special | (spell << 8);
        }
        qCount += __popc(m);
        if (qCount >= 32) drain(31);
    }

    // fold tier 0 into the CTA's shared table, then the shared table into the HBM table
    __device__ __forceinline__ void flush()
    {
        if (Prog::QUEUE) drain(0);
        if (!tiers) return;
        __syncthreads();
        const GroupTable& T = P.t;
        const int lane = tid & 31, warp = tid >> 5;
        const uint32_t nT0 = min(st.count[1], (uint32_t) K0);
        for (uint32_t pairIdx = warp; pairIdx < nT0 * (uint32_t) NC && tid < THREADS; pairIdx += THREADS / 32) {
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
        for (uint32_t s = tid; s < st.S2 && tid < THREADS; s += THREADS) {
            const uint32_t f = st.first[s];
            if (f == kNoRow32) continue;
            int gslot;
            if (s >= st.S) gslot = (int) T.cap + (int) (s - st.S);
            else { uint64_t kw[W]; Prog::expandKey((uint64_t) st.keys[s], kw); gslot = tableFind<W>(T, kw, hashKeyWords<W>(kw), true, true); }
            if (gslot < 0) { atomicOr(&T.counters[CNT_ERROR], 0x80000000u); continue; }
            const uint64_t g = firstRowWord(P.rowBase + f);
            if (g < __ldcg(&T.firstRow[gslot])) atomicMin((unsigned long long*) &T.firstRow[gslot], (unsigned long long) g);
            uint64_t* acc = T.acc + (size_t) gslot * T.nvPhys;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const uint64_t x = st.acc[c * st.S2 + s];
                if (x != 0) {
                    if (c < NV) atomicAdd((double*) &acc[P.phys[c]], __longlong_as_double((long long) x));
                    else atomicAdd((unsigned long long*) &acc[P.phys[c]], (unsigned long long) x);
                }
            }
        }
    }
};

// ---- TMA staging ring (north_star: "TMA bulk staging of page data into shared memory") ----------------------------------------------
// cp.async.bulk (the 1-D form of TMA: no tensor map) copies one tile of every eagerly loaded column from HBM into a shared-memory stage
// and signals an mbarrier with the byte count; a dedicated producer warp keeps `stages` tiles in flight while the 8 consumer warps
// move a landed tile into registers (Prog::loadSmem), release the stage and run the row program.  Compared with the register double
// buffer of fusedAggBody the prefetch depth no longer costs registers and the consumers issue no global loads for these columns.
__device__ __forceinline__ uint32_t smemU32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemU32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemU32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbarArrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smemU32(bar)) : "memory"); }
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" :: "r"(smemU32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulkCopyToShared(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smemU32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemU32(bar)) : "memory");
}

template <class Prog>
__device__ __forceinline__ void fusedAggBodyTma(const FusedParams& P)
{
    extern __shared__ __align__(16) unsigned char fusedSmem[];
    constexpr int THREADS = Prog::THREADS;            // consumer threads; the block has THREADS + 32 (the producer warp)
    constexpr int TILE = THREADS * Prog::R;
    FusedAggSink<Prog> sink(P);
    sink.init(fusedSmem);
    const int stages = P.tmaStages;
    unsigned char* ring = fusedSmem + ((sink.layoutBytes() + (uint32_t) FusedAggSink<Prog>::queueBytes() + 127u) & ~127u);
    uint64_t* full = (uint64_t*) (ring + (size_t) stages * Prog::STAGE_BYTES);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbarInit(&full[s], 1); mbarInit(&empty[s], THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long nTiles = (P.n + TILE - 1) / TILE;
    const long long tmaTiles = min((long long) P.tmaTiles, nTiles);
    if (threadIdx.x >= THREADS) {
        // ---- producer warp: one lane issues the bulk copies of tile k into stage k % stages once the consumers released it ----
        if (threadIdx.x == THREADS) {
            long long k = 0;
            for (long long tile = blockIdx.x; tile < tmaTiles; tile += gridDim.x, k++) {
                const int s = (int) (k % stages);
                if (k >= stages) mbarWait(&empty[s], (uint32_t) ((k / stages - 1) & 1));
                mbarExpectTx(&full[s], Prog::STAGE_TX_BYTES);
                Prog::issueStage(P, tile * TILE, ring + (size_t) s * Prog::STAGE_BYTES, &full[s]);
            }
        }
    }
    else {
        // ---- consumers ----
        typename Prog::Tile t;
        long long k = 0;
        for (long long tile = blockIdx.x; tile < tmaTiles; tile += gridDim.x, k++) {
            const int s = (int) (k % stages);
            mbarWait(&full[s], (uint32_t) ((k / stages) & 1));
            Prog::loadSmem(ring + (size_t) s * Prog::STAGE_BYTES, t);
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbarArrive(&empty[s]);      // the stage is free as soon as the tile sits in registers
            Prog::process(P, t, tile * TILE, sink);
            __syncwarp();
        }
        // the tail: the last full tile and the partial one come through direct loads (a bulk copy must not read past a column's end)
        for (long long tile = tmaTiles + blockIdx.x; tile < nTiles; tile += gridDim.x) {
            Prog::load(P, tile * TILE, t);
            Prog::process(P, t, tile * TILE, sink);
            __syncwarp();
        }
    }
    sink.flush();
}

// ---- the kernel body: persistent CTAs, tiles of THREADS * R rows, the next tile's loads in flight while this one is consumed ----
template <class Prog>
__device__ __forceinline__ void fusedAggBody(const FusedParams& P)
{
    extern __shared__ __align__(16) unsigned char fusedSmem[];
    FusedAggSink<Prog> sink(P);
    sink.init(fusedSmem);
    __syncthreads();
    constexpr int TILE = Prog::THREADS * Prog::R;
    const long long nTiles = (P.n + TILE - 1) / TILE;
    typename Prog::Tile tA, tB;
    long long tile = blockIdx.x;
    if (tile < nTiles) Prog::load(P, tile * TILE, tA);
    while (tile < nTiles) {
        const long long t1 = tile + gridDim.x;
        if (t1 < nTiles) Prog::load(P, t1 * TILE, tB);
        Prog::process(P, tA, tile * TILE, sink);
        __syncwarp();      // rows take different paths (filter, probe miss, tiers): the next tile's vector loads are issued by whole warps
        if (t1 >= nTiles) break;
        const long long t2 = t1 + gridDim.x;
        if (t2 < nTiles) Prog::load(P, t2 * TILE, tA);
        Prog::process(P, tB, t1 * TILE, sink);
        __syncwarp();
        tile = t2;
    }
    sink.flush();
}
#endif  // __CUDACC__

}  // namespace pg
