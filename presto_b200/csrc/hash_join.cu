// presto_b200 — HashBuilderOperator + LookupJoinOperator on sm_100a.
//
// Replaces (MB = presto-main-base/src/main/java/com/facebook/presto/):
//   MB/operator/HashBuilderOperator.java:333-365,519-537,642-651   addInput / finishInput / buildLookupSource
//   MB/operator/PagesIndex.java:202-226                             addPage (build columns kept, one address per row)
//   MB/operator/PagesHash.java:55-132                               open-address int[] key + byte[] positionToHashes build loop
//   MB/operator/ArrayPositionLinks.java:38-100                      duplicate-key chains
//   MB/operator/PagesHash.java:169-181, JoinHash.java:81-124        getAddressIndex / getJoinPosition / getNextJoinPosition
//   MB/operator/JoinProbe.java:74-119, LookupJoinOperator.java:481-505,567-616, LookupJoinPageBuilder.java:75-137
//
// Build: the join key of every build row is packed into W words (group_table.cuh) and found-or-claimed in an HBM table in
// parallel; the slot's `head` is exchanged to the row's position and the row's link takes the previous head.  The reference's
// sequential loop makes head = the LAST row of a key and links descend (PagesHash.java:108-123); a parallel exchange produces
// the same set in arbitrary order, so chains of keys with duplicates are re-threaded into descending position order afterwards
// (skipped entirely when the build side has no duplicate keys — the TPC-H orders case).
// Probe: pass 1 finds each probe row's chain head and match count, an exclusive scan turns counts into output offsets, pass 2
// writes probe and build columns straight to their final positions — output rows are in probe order with each probe row's
// matches in chain order, exactly the order LookupJoinOperator emits (pinned by TestHashJoinOperator).
#include "common.cuh"
#include "fp_interp.cuh"
#include "group_table.cuh"
#include "varchar_interner.h"
#include "expr_literals.h"
#include "key_layout.h"
#include "dense_index.cuh"
#include "operators_internal.h"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <mutex>

namespace pg {

constexpr int kMaxJoinColumns = 16;

struct JoinTable {
    GroupTable t;             // keys / state (firstRow, acc unused)
    int32_t* head;            // [capP] last build position with the slot's key, -1 empty
    int32_t* links;           // [n]    next (smaller) position with the same key, -1 end
};

__device__ __forceinline__ bool keyHasNull(const KeyPlan& kp, const uint64_t* kw, int special)
{
    if (kp.bigintPath) return special == 1;
    return ((kw[kp.nullWord] >> kp.nullShift) & 0xff) != 0;
}

template <int W>
__global__ void joinBuildKernel(JoinTable J, KeyPlan kp, int64_t n, unsigned long long* stats /* [0]=duplicates [1]=null keys */)
{
    for (int64_t pos = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; pos < n; pos += (int64_t) gridDim.x * blockDim.x) {
        uint64_t kw[W];
        int special = packKey<W>(kp, pos, kw);
        if (keyHasNull(kp, kw, special)) { J.links[pos] = -1; atomicAdd(&stats[1], 1ULL); continue; }   // PagesHash.java:99-102
        int slot = special == 2 ? (int) J.t.cap + 1 : tableFind<W>(J.t, kw, hashKeyWords<W>(kw), true, true);
        int old = atomicExch(&J.head[slot], (int) pos);
        J.links[pos] = old;
        if (old >= 0) atomicAdd(&stats[0], 1ULL);
    }
}

// Re-thread every chain longer than one row into descending position order (one thread per slot; chains arrive almost sorted
// because the exchange order follows the row order, so the tail-append fast path makes this close to linear).
__global__ void joinSortChainsKernel(JoinTable J)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < J.t.capP; s += gridDim.x * blockDim.x) {
        int node = J.head[s];
        if (node < 0 || J.links[node] < 0) continue;
        int newHead = -1, newTail = -1;
        while (node >= 0) {
            int next = J.links[node];
            if (newHead < 0) { newHead = newTail = node; J.links[node] = -1; }
            else if (node < newTail) { J.links[newTail] = node; J.links[node] = -1; newTail = node; }
            else if (node > newHead) { J.links[node] = newHead; newHead = node; }
            else {
                int p = newHead;
                while (J.links[p] > node) p = J.links[p];
                J.links[node] = J.links[p];
                J.links[p] = node;
            }
            node = next;
        }
        J.head[s] = newHead;
    }
}

// ---- JoinFilterFunction (MB/operator/JoinFilterFunction.java, MB/sql/gen/JoinFilterFunctionCompiler.java:97-250) --------------------
// The non-equi part of a join condition: a BOOLEAN RowExpression over the build row at `buildPosition` (inputs [0, nBuildInputs): the
// build pages' channels) and the probe row at `probePosition` (the remaining inputs: the probe page's channels) — the layout
// LocalExecutionPlanner.createJoinSourcesLayout gives it.  JoinHash.isJoinPositionEligible :113-118 asks it for every candidate of a
// chain; only TRUE keeps the match.  Evaluated by the RowExpression interpreter (fp_interp.cuh) inside the probe / emit kernels.
constexpr int kMaxJoinFilterNodes = 32;
struct JoinFilterProgram {
    FpNode nodes[kMaxJoinFilterNodes];
    FpInput inputs[kMaxInputs];
    int32_t nNodes, filterNode, nBuildInputs;
    unsigned int* error;
};
__device__ __forceinline__ bool joinPositionEligible(const JoinFilterProgram* F, int64_t buildPosition, int64_t probePosition)
{
    uint64_t regs[kMaxJoinFilterNodes];
    uint32_t nullBits = 0;
    for (int k = 0; k < F->nNodes; k++) {
        const FpNode& nd = F->nodes[k];
        Val x{0, false}, y{0, false}, z{0, false};
        int64_t row = 0;
        if (nd.op == PG_EX_INPUT) row = nd.a < F->nBuildInputs ? buildPosition : probePosition;
        else {
            if (nd.a >= 0) x = Val{regs[nd.a], (bool) ((nullBits >> nd.a) & 1u)};
            if (nd.b >= 0) y = Val{regs[nd.b], (bool) ((nullBits >> nd.b) & 1u)};
            if (nd.c >= 0) z = Val{regs[nd.c], (bool) ((nullBits >> nd.c) & 1u)};
        }
        const Val v = evalNodeAt(F->nodes, F->inputs, F->error, nd, row, x, y, z);
        regs[k] = v.bits;
        if (v.isNull) nullBits |= 1u << k;
    }
    return !((nullBits >> F->filterNode) & 1u) && regs[F->filterNode] != 0;
}

struct ProbeParams {
    JoinTable J;
    KeyPlan kp;
    int64_t n;
    int32_t outer;            // PROBE_OUTER: unmatched probe rows emit one row with NULL build columns
    int32_t hasDuplicates;
    int32_t* matchHead;       // [n] chain head or -1
    uint32_t* counts;         // [n+1] output rows per probe row
    unsigned long long* total;   // exact 64-bit sum of counts (a skewed join can emit more than 2^32 rows for one probe page)
    const JoinFilterProgram* filter;   // device memory; nullptr = no filter function
};

template <int W>
__global__ void joinProbeKernel(const __grid_constant__ ProbeParams P)
{
    unsigned long long mine = 0;
    for (int64_t row = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; row < P.n; row += (int64_t) gridDim.x * blockDim.x) {
        uint64_t kw[W];
        int special = packKey<W>(P.kp, row, kw);
        int head = -1;
        if (!keyHasNull(P.kp, kw, special)) {   // JoinProbe.java:81-84: a NULL probe key never matches
            int slot = special == 2 ? (int) P.J.t.cap + 1 : tableFind<W>(P.J.t, kw, hashKeyWords<W>(kw), false, true);
            if (slot >= 0) head = P.J.head[slot];
        }
        uint32_t c = 0;
        if (head >= 0 && P.filter) {        // LookupJoinOperator.joinCurrentPosition :567-586: only eligible positions produce rows
            for (int p = head; p >= 0; p = P.hasDuplicates ? P.J.links[p] : -1) c += joinPositionEligible(P.filter, p, row) ? 1u : 0u;
            if (c == 0 && P.outer) c = 1;
        }
        else if (head >= 0) {
            c = 1;
            if (P.hasDuplicates) for (int p = P.J.links[head]; p >= 0; p = P.J.links[p]) c++;
        }
        else if (P.outer) c = 1;
        P.matchHead[row] = head;
        P.counts[row] = c;
        mine += c;
    }
    // every thread of the warp gets here (no early exits): one 64-bit atomic per warp
    for (int s = 16; s >= 1; s >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, s);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(P.total, mine);
}

struct JoinColumn {
    const void* src; const uint8_t* srcNulls;
    void* dst; uint8_t* dstNulls;
    int32_t width;
};
struct EmitJoinParams {
    const int32_t* matchHead; const uint32_t* offsets; const int32_t* links;
    int64_t n;
    int32_t hasDuplicates;
    int32_t nProbe, nBuild;
    uint8_t* visited;          // OuterLookupSource.appendTo :100-105: positionVisited(position); nullptr when the join is not build-outer
    const JoinFilterProgram* filter;
    JoinColumn probe[kMaxJoinColumns];
    JoinColumn build[kMaxJoinColumns];
};

__device__ __forceinline__ void copyValue(const JoinColumn& c, int64_t from, int64_t to)
{
    if (c.width == 8) ((uint64_t*) c.dst)[to] = ((const uint64_t*) c.src)[from];
    else if (c.width == 4) ((uint32_t*) c.dst)[to] = ((const uint32_t*) c.src)[from];
    else ((uint8_t*) c.dst)[to] = ((const uint8_t*) c.src)[from];
    if (c.dstNulls) c.dstNulls[to] = c.srcNulls ? c.srcNulls[from] : 0;
}
__device__ __forceinline__ void writeNull(const JoinColumn& c, int64_t to)
{
    if (c.width == 8) ((uint64_t*) c.dst)[to] = 0;
    else if (c.width == 4) ((uint32_t*) c.dst)[to] = 0;
    else ((uint8_t*) c.dst)[to] = 0;
    c.dstNulls[to] = 1;
}

__global__ void joinEmitKernel(const __grid_constant__ EmitJoinParams P)
{
    for (int64_t row = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; row < P.n; row += (int64_t) gridDim.x * blockDim.x) {
        const uint32_t begin = P.offsets[row], end = P.offsets[row + 1];
        if (begin == end) continue;
        int pos = P.matchHead[row];
        if (P.filter) {
            uint32_t o = begin;
            for (; pos >= 0 && o < end; pos = P.hasDuplicates ? P.links[pos] : -1) {
                if (!joinPositionEligible(P.filter, pos, row)) continue;
                for (int c = 0; c < P.nProbe; c++) copyValue(P.probe[c], row, o);
                for (int c = 0; c < P.nBuild; c++) copyValue(P.build[c], pos, o);
                if (P.visited) P.visited[pos] = 1;
                o++;
            }
            if (o == begin) {      // PROBE_OUTER and no eligible position
                for (int c = 0; c < P.nProbe; c++) copyValue(P.probe[c], row, o);
                for (int c = 0; c < P.nBuild; c++) writeNull(P.build[c], o);
            }
            continue;
        }
        for (uint32_t o = begin; o < end; o++) {
            for (int c = 0; c < P.nProbe; c++) copyValue(P.probe[c], row, o);
            if (pos >= 0) {
                for (int c = 0; c < P.nBuild; c++) copyValue(P.build[c], pos, o);
                if (P.visited) P.visited[pos] = 1;
                pos = P.hasDuplicates ? P.links[pos] : -1;
            }
            else for (int c = 0; c < P.nBuild; c++) writeNull(P.build[c], o);   // LookupJoinPageBuilder.appendNullForBuild
        }
    }
}

// ---- VARCHAR output channels ------------------------------------------------------------------------------------------------------
// Fixed-width output columns are written in place by the emit kernels.  A VARCHAR column (LookupJoinPageBuilder.appendRow ->
// type.appendTo: the slice of the probe row / of the build position) follows them: the emit kernels also write, like one more 4-byte
// column, the SOURCE index of every output row (an iota column as its "values"), and the slices are gathered through that index —
// lengths, inclusive scan, bytes.
__global__ void iotaU32Kernel(uint32_t* p, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) p[i] = (uint32_t) i;
}
__global__ void varcharIndexLengthsKernel(const int32_t* __restrict__ srcOffsets, const uint8_t* __restrict__ srcNulls, const uint32_t* __restrict__ index,
                                          const uint8_t* __restrict__ indexNulls, int64_t m, int32_t* __restrict__ outOffsets, uint8_t* __restrict__ outNulls)
{
    for (int64_t o = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; o < m; o += (int64_t) gridDim.x * blockDim.x) {
        const bool noSource = indexNulls && indexNulls[o];
        const uint32_t i = noSource ? 0u : index[o];
        const bool isNull = noSource || (srcNulls && srcNulls[i]);
        outOffsets[o + 1] = isNull ? 0 : srcOffsets[i + 1] - srcOffsets[i];
        if (outNulls) outNulls[o] = isNull ? 1 : 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) outOffsets[0] = 0;
}
__global__ void varcharIndexBytesKernel(const uint8_t* __restrict__ srcBytes, const int32_t* __restrict__ srcOffsets, const uint32_t* __restrict__ index,
                                        const int32_t* __restrict__ outOffsets, int64_t m, uint8_t* __restrict__ outBytes)
{
    // one warp per output row: slices are short (names, comments), lanes copy consecutive bytes
    const int lane = threadIdx.x & 31;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t o = warpId; o < m; o += nWarps) {
        const int32_t dst = outOffsets[o], len = outOffsets[o + 1] - dst;
        if (len <= 0) continue;
        const int32_t src = srcOffsets[index[o]];
        for (int32_t k = lane; k < len; k += 32) outBytes[dst + k] = srcBytes[src + k];
    }
}
__global__ void rebaseOffsetsKernel(const int32_t* __restrict__ in, int64_t n, int32_t delta, int32_t* __restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) out[i + 1] = in[i + 1] + delta;
}

// Gathers `m` slices of a VARCHAR column through `index` (indexNulls[o] != 0: the output row has no source, e.g. the NULL build side of an
// outer row) into offsets / bytes / nulls buffers of the caller.  Returns the byte total (one synchronisation).
static int64_t gatherVarcharByIndex(Operator* op, const uint8_t* srcBytes, const int32_t* srcOffsets, const uint8_t* srcNulls, const uint32_t* index,
                                    const uint8_t* indexNulls, int64_t m, DeviceBuffer& outBytes, DeviceBuffer& outOffsets, uint8_t* outNulls, DeviceBuffer& scanTemp)
{
    outOffsets.reserve((size_t) (m + 1) * 4);
    const int grid = (int) std::max<int64_t>(1, std::min<int64_t>(divUp(m, 256), (int64_t) op->numSMs * 16));
    PG_LAUNCH(op, varcharIndexLengthsKernel, grid, 256, 0, srcOffsets, srcNulls, index, indexNulls, m, outOffsets.as<int32_t>(), outNulls);
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, outOffsets.as<int32_t>() + 1, outOffsets.as<int32_t>() + 1, (int) m, op->stream);
    scanTemp.reserve(tb);
    cub::DeviceScan::InclusiveSum(scanTemp.ptr, tb, outOffsets.as<int32_t>() + 1, outOffsets.as<int32_t>() + 1, (int) m, op->stream);
    op->stats.kernel_launches += 2;
    int32_t total = 0;
    PG_CUDA(cudaMemcpyAsync(&total, outOffsets.as<int32_t>() + m, 4, cudaMemcpyDeviceToHost, op->stream));
    PG_CUDA(cudaStreamSynchronize(op->stream));
    op->stats.d2h_bytes += 4;
    PG_CHECK_ARG(total >= 0, "a VARCHAR output column exceeds 2 GB: feed smaller probe pages");
    outBytes.reserve(std::max<size_t>((size_t) total, 16));
    if (total > 0) PG_LAUNCH(op, varcharIndexBytesKernel, (int) std::max<int64_t>(1, std::min<int64_t>(divUp(m, 8), (int64_t) op->numSMs * 16)), 256, 0, srcBytes, srcOffsets, index, outOffsets.as<int32_t>(), m, outBytes.as<uint8_t>());
    return total;
}

// ---- dense unique integer keys: bitmap + rank index ------------------------------------------------------------------------
// When the build keys are unique and span a range of at most kDenseRangeFactor x n (TPC-H order keys: 150 M keys over 600 M; Q3's
// filtered orders: ~15 M keys over the same 600 M), the
// LookupSource is a succinct index instead of a hash table: one bit per value of the key range (L2 resident: 75 MB at SF100) and a
// rank directory (exclusive popcount per 256-bit block), with the build OUTPUT columns re-laid in key order.  A probe is one 32-byte
// bitmap sector (L2) + 4 B of directory → rank → payload[rank]; sorted probe keys (lineitem is clustered by orderkey) make every
// access sequential.  Results are identical to the hash path (unique keys: at most one match, no chains).
constexpr int kDenseRangeFactor = 64;                // keys may cover as little as 1/64 of their range (bitmap: range / 8 bytes)
constexpr int kDenseSmallRangeBits = 31;             // ranges up to 2^31 (256 MB bitmap + 32 MB directory) use the index at any density

struct KeyColumn { const void* values; const uint8_t* nulls; int32_t type; };

__device__ __forceinline__ bool loadKey(const KeyColumn& k, int64_t row, int64_t& key)
{
    if (k.nulls && k.nulls[row]) return false;
    key = k.type == PG_TYPE_BIGINT ? ((const int64_t*) k.values)[row] : (int64_t) ((const int32_t*) k.values)[row];
    return true;
}

__global__ void __launch_bounds__(256) denseMinMaxKernel(KeyColumn k, int64_t n, long long* minMax /* [0]=min [1]=max [2]=non-null count */)
{
    __shared__ long long sLo[8], sHi[8]; __shared__ unsigned long long sCnt[8];
    long long lo = LLONG_MAX, hi = LLONG_MIN; unsigned long long cnt = 0;
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        int64_t key;
        if (loadKey(k, i, key)) { lo = min(lo, (long long) key); hi = max(hi, (long long) key); cnt++; }
    }
    for (int m = 16; m >= 1; m >>= 1) {
        lo = min(lo, (long long) shflXor64((uint64_t) lo, m)); hi = max(hi, (long long) shflXor64((uint64_t) hi, m)); cnt += shflXor64(cnt, m);
    }
    if ((threadIdx.x & 31) == 0) { sLo[threadIdx.x >> 5] = lo; sHi[threadIdx.x >> 5] = hi; sCnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { lo = min(lo, sLo[w]); hi = max(hi, sHi[w]); cnt += sCnt[w]; }
        atomicMin(&minMax[0], lo); atomicMax(&minMax[1], hi); atomicAdd((unsigned long long*) &minMax[2], cnt);
    }
}

__global__ void denseSetBitsKernel(KeyColumn k, int64_t n, int64_t minKey, unsigned long long* bitmap, unsigned int* duplicate)
{
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        int64_t key;
        if (!loadKey(k, i, key)) continue;
        const uint64_t idx = (uint64_t) (key - minKey);
        const unsigned long long bit = 1ULL << (idx & 63);
        if (atomicOr(&bitmap[idx >> 6], bit) & bit) *duplicate = 1;   // the same key twice: not unique → hash path
    }
}

__global__ void denseBlockCountKernel(const uint64_t* bitmap, int64_t blocks, uint32_t* counts)
{
    for (int64_t b = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; b < blocks; b += (int64_t) gridDim.x * blockDim.x) {
        const uint64_t* w = bitmap + b * kDenseBlockWords;
        counts[b] = __popcll(w[0]) + __popcll(w[1]) + __popcll(w[2]) + __popcll(w[3]);
    }
}

// Build OUTPUT columns are stored row-wise in key order ("payload rows"): one random 8..32-byte write per build row when the index
// is built, one 8..32-byte read per matching probe row (sequential when the probe keys are clustered).
constexpr int kMaxPayloadWords = 4;                 // payload rows of up to 32 bytes use the dense index
struct PayloadColumn { int32_t offset, width, nullOffset; };   // byte offsets inside the row; nullOffset < 0: column has no NULLs
struct PayloadLayout {
    int32_t nCols, strideWords;
    PayloadColumn cols[kMaxJoinColumns];
};

__device__ __forceinline__ uint64_t loadWidth(const void* base, int64_t i, int width)
{
    if (width == 8) return ((const uint64_t*) base)[i];
    if (width == 4) return ((const uint32_t*) base)[i];
    return ((const uint8_t*) base)[i];
}
__device__ __forceinline__ void storeWidth(void* base, int64_t i, int width, uint64_t v)
{
    if (width == 8) ((uint64_t*) base)[i] = v;
    else if (width == 4) ((uint32_t*) base)[i] = (uint32_t) v;
    else ((uint8_t*) base)[i] = (uint8_t) v;
}
__device__ __forceinline__ uint64_t widthMask(int width) { return width >= 8 ? ~0ULL : ((1ULL << (8 * width)) - 1); }

struct DenseGatherParams {
    KeyColumn key; DenseIndex D; int64_t n;
    PayloadLayout L;
    const void* src[kMaxJoinColumns]; const uint8_t* srcNulls[kMaxJoinColumns];
    uint64_t* rows;   // [nonNull][strideWords]
};
__global__ void denseGatherKernel(const __grid_constant__ DenseGatherParams P)
{
    for (int64_t pos = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; pos < P.n; pos += (int64_t) gridDim.x * blockDim.x) {
        int64_t key;
        if (!loadKey(P.key, pos, key)) continue;
        const int64_t r = denseRank(P.D, key);
        uint64_t w[kMaxPayloadWords] = {0, 0, 0, 0};
        for (int c = 0; c < P.L.nCols; c++) {
            const PayloadColumn& pc = P.L.cols[c];
            const uint64_t v = loadWidth(P.src[c], pos, pc.width);
#pragma unroll
            for (int k = 0; k < kMaxPayloadWords; k++) if (k == pc.offset >> 3) w[k] |= v << (8 * (pc.offset & 7));
            if (pc.nullOffset >= 0 && P.srcNulls[c][pos]) {
#pragma unroll
                for (int k = 0; k < kMaxPayloadWords; k++) if (k == pc.nullOffset >> 3) w[k] |= 1ULL << (8 * (pc.nullOffset & 7));
            }
        }
#pragma unroll
        for (int k = 0; k < kMaxPayloadWords; k++) if (k < P.L.strideWords) P.rows[r * P.L.strideWords + k] = w[k];
    }
}

// ---- sorted build of the dense index (sparse key sets in arrival order: TPC-H Q3's filtered orders) ------------------------------
// Scattering one payload row per build row straight from the input order costs three dependent random sectors per row (bitmap, rank
// directory, payload).  For one-word payloads it is cheaper to sort (key - min, payload word) pairs with the radix sort: after it the
// rank of a key IS its index, the payload rows are already in key order in the destination buffer, the bitmap is set front to back,
// and a duplicate key is two equal neighbours.
__global__ void __launch_bounds__(256) densePackPairsKernel(const __grid_constant__ DenseGatherParams P, int64_t minKey, uint32_t* keys32, uint64_t* pay, unsigned int* counter, int compact)
{
    __shared__ unsigned int warpCount[8], blockBase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t rounds = (P.n + (int64_t) gridDim.x * blockDim.x - 1) / ((int64_t) gridDim.x * blockDim.x);
    for (int64_t k = 0; k < rounds; k++) {
        const int64_t pos = k * gridDim.x * blockDim.x + blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
        int64_t key = 0;
        const bool live = pos < P.n && loadKey(P.key, pos, key);
        uint64_t w = 0;
        if (live) {
            for (int c = 0; c < P.L.nCols; c++) {
                const PayloadColumn& pc = P.L.cols[c];
                w |= loadWidth(P.src[c], pos, pc.width) << (8 * (pc.offset & 7));
                if (pc.nullOffset >= 0 && P.srcNulls[c][pos]) w |= 1ULL << (8 * (pc.nullOffset & 7));
            }
        }
        int64_t at = pos;                       // no NULL keys: the pair keeps its position, nothing to coordinate
        if (compact) {                          // NULL keys drop out: one counter increment per CTA and round
            const unsigned m = __ballot_sync(0xffffffffu, live);
            if (lane == 0) warpCount[warp] = __popc(m);
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned int total = 0;
                for (int i = 0; i < 8; i++) { const unsigned int c = warpCount[i]; warpCount[i] = total; total += c; }
                blockBase = total ? atomicAdd(counter, total) : 0;
            }
            __syncthreads();
            at = (int64_t) blockBase + warpCount[warp] + __popc(m & ((1u << lane) - 1));
            __syncthreads();
        }
        if (live) { keys32[at] = (uint32_t) (uint64_t) (key - minKey); pay[at] = w; }
    }
}
__global__ void __launch_bounds__(256) denseSortedBitsKernel(const uint32_t* __restrict__ sorted, int64_t n, unsigned long long* bitmap, unsigned int* duplicate)
{
    for (int64_t r = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; r < n; r += (int64_t) gridDim.x * blockDim.x) {
        const uint32_t idx = sorted[r];
        if (r > 0 && sorted[r - 1] == idx) { *duplicate = 1; continue; }
        atomicOr(&bitmap[idx >> 6], 1ULL << (idx & 63));
    }
}

// ---- range-partitioned build of the dense index (large builds) --------------------------------------------------------------
// Setting one bitmap bit and scattering one payload row per build row, straight from the input order, costs four uncoalesced
// accesses per row (bit, bitmap sector, directory word, payload row), each a separate L1TEX wavefront (~2 cycles per lane per SM)
// and a DRAM row miss: ~12 ms for 150 M rows.  Large builds are therefore split first — one tile-staged radix pass on the high
// bits of (key - min) into up to 1024 key-range buckets — after which ONE CTA owns one bucket: it builds the bucket's bitmap
// slice and rank directory in shared memory (no global atomics), finds duplicates there, and scatters the payload rows into the
// bucket's contiguous slice of the payload (a couple of MB: L2 hits, full lines to DRAM).  Records travel as a 32-bit offset
// inside the bucket plus the packed payload words.
constexpr int kPartBuckets = 1024;
constexpr int kPartTile = 4096;                      // rows per CTA tile of the partition pass
constexpr int kPartThreads = 512;
constexpr int kPartRowsPerThread = kPartTile / kPartThreads;
constexpr int kPartMaxShift = 20;                    // bucket = 2^20 key values: 128 KB of bitmap + 16 KB of directory in shared memory
constexpr int kPartMinShift = 8;                     // a bucket is at least one 256-bit directory block
constexpr int kFinalThreads = 1024;
constexpr int64_t kPartBuildMinRows = 1 << 22;       // smaller builds scatter directly (their payload rows fit in L2 anyway)

struct DensePartParams {
    KeyColumn key; int64_t n, tiles; int64_t minKey; int32_t shift;
    uint32_t* hist;                 // pass 1 out [kPartBuckets * tiles + 1], bucket-major
    const uint32_t* base;           // exclusive scan of hist
    PayloadLayout L;
    const void* src[kMaxJoinColumns]; const uint8_t* srcNulls[kMaxJoinColumns];
    uint32_t* recIdx;               // [nonNull] (key - min) & (2^shift - 1), bucket after bucket
    uint64_t* recPay;               // [strideWords][nonNull] payload word planes
    int64_t recCap;
    // final pass
    uint64_t* bitmap; uint32_t* prefix; int64_t blocks;   // global index: blocks x 256 bits, blocks directory entries
    uint64_t* rows; unsigned int* duplicate;
};

// exclusive scan of one value per thread across the CTA (blockDim.x <= 1024); `scratch` has 33 words
__device__ __forceinline__ uint32_t blockExclusiveScan(uint32_t v, uint32_t* scratch, uint32_t& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nWarps ? scratch[lane] : 0, winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += o; }
        scratch[lane] = winc - w;
        if (lane == 31) scratch[32] = winc;
    }
    __syncthreads();
    const uint32_t r = scratch[warp] + inc - v;
    total = scratch[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kPartThreads) densePartHistKernel(const __grid_constant__ DensePartParams P)
{
    __shared__ uint32_t hist[kPartBuckets];
    for (int64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        for (int b = threadIdx.x; b < kPartBuckets; b += kPartThreads) hist[b] = 0;
        __syncthreads();
        int64_t key[kPartRowsPerThread]; bool valid[kPartRowsPerThread];
#pragma unroll
        for (int r = 0; r < kPartRowsPerThread; r++) { const int64_t row = tile * kPartTile + r * kPartThreads + threadIdx.x; valid[r] = row < P.n && loadKey(P.key, row, key[r]); }
#pragma unroll
        for (int r = 0; r < kPartRowsPerThread; r++) if (valid[r]) atomicAdd(&hist[(uint64_t) (key[r] - P.minKey) >> P.shift], 1u);
        __syncthreads();
        for (int b = threadIdx.x; b < kPartBuckets; b += kPartThreads) P.hist[(size_t) b * P.tiles + tile] = hist[b];
        __syncthreads();
    }
}

template <int SW>
__global__ void __launch_bounds__(kPartThreads, 2) densePartScatterKernel(const __grid_constant__ DensePartParams P)
{
    extern __shared__ __align__(16) unsigned char partSmem[];
    uint32_t* hist = (uint32_t*) partSmem;                 // [kPartBuckets] counts, then tile-local starts
    uint32_t* gbase = hist + kPartBuckets;                  // [kPartBuckets] global position of the bucket's first record of this tile - local start
    uint32_t* scratch = gbase + kPartBuckets;               // [64]
    uint32_t* sIdx = scratch + 64;                          // [kPartTile]
    uint16_t* sBucket = (uint16_t*) (sIdx + kPartTile);     // [kPartTile]
    uint64_t* sPay = (uint64_t*) (sBucket + kPartTile);     // [SW][kPartTile]
    const uint32_t idxMask = (1u << P.shift) - 1;
    for (int64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        for (int b = threadIdx.x; b < kPartBuckets; b += kPartThreads) hist[b] = 0;
        __syncthreads();
        // all independent global loads of the tile first (keys, payload columns, this tile's bucket bases), then the shared-memory ranking
        int64_t key[kPartRowsPerThread]; bool valid[kPartRowsPerThread];
        uint64_t w[kPartRowsPerThread][SW];
#pragma unroll
        for (int r = 0; r < kPartRowsPerThread; r++) {
            const int64_t row = tile * kPartTile + r * kPartThreads + threadIdx.x;
            valid[r] = row < P.n && loadKey(P.key, row, key[r]);
        }
#pragma unroll
        for (int r = 0; r < kPartRowsPerThread; r++)
#pragma unroll
            for (int k = 0; k < SW; k++) w[r][k] = 0;
        for (int c = 0; c < P.L.nCols; c++) {          // column by column: the loads of all rows of the thread are in flight together
            const PayloadColumn pc = P.L.cols[c];
            uint64_t v[kPartRowsPerThread]; uint8_t isNull[kPartRowsPerThread];
#pragma unroll
            for (int r = 0; r < kPartRowsPerThread; r++) {
                const int64_t row = tile * kPartTile + r * kPartThreads + threadIdx.x;
                v[r] = valid[r] ? loadWidth(P.src[c], row, pc.width) : 0;
                isNull[r] = (valid[r] && pc.nullOffset >= 0) ? P.srcNulls[c][row] : 0;
            }
#pragma unroll
            for (int r = 0; r < kPartRowsPerThread; r++) {
#pragma unroll
                for (int k = 0; k < SW; k++) {
                    if (k == pc.offset >> 3) w[r][k] |= v[r] << (8 * (pc.offset & 7));
                    if (isNull[r] && k == pc.nullOffset >> 3) w[r][k] |= 1ULL << (8 * (pc.nullOffset & 7));
                }
            }
        }
        const uint32_t gb0 = P.base[(size_t) (2 * threadIdx.x) * P.tiles + tile], gb1 = P.base[(size_t) (2 * threadIdx.x + 1) * P.tiles + tile];
        uint32_t idx[kPartRowsPerThread]; uint32_t slot[kPartRowsPerThread];   // slot = bucket << 16 | rank inside (tile, bucket); 0xFFFFFFFF = no record
#pragma unroll
        for (int r = 0; r < kPartRowsPerThread; r++) {
            slot[r] = 0xFFFFFFFFu; idx[r] = 0;
            if (valid[r]) {
                const uint64_t full = (uint64_t) (key[r] - P.minKey);
                const uint32_t b = (uint32_t) (full >> P.shift);
                idx[r] = (uint32_t) full & idxMask;
                slot[r] = b << 16 | atomicAdd(&hist[b], 1u);
            }
        }
        __syncthreads();
        // tile-local starts (2 buckets per thread) and the global position of each bucket's run
        const uint32_t c0 = hist[2 * threadIdx.x], c1 = hist[2 * threadIdx.x + 1];
        uint32_t total;
        const uint32_t start = blockExclusiveScan(c0 + c1, scratch, total);
        hist[2 * threadIdx.x] = start; hist[2 * threadIdx.x + 1] = start + c0;
        gbase[2 * threadIdx.x] = gb0 - start;
        gbase[2 * threadIdx.x + 1] = gb1 - (start + c0);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kPartRowsPerThread; r++) {
            if (slot[r] == 0xFFFFFFFFu) continue;
            const uint32_t b = slot[r] >> 16, pos = hist[b] + (slot[r] & 0xFFFFu);
            sIdx[pos] = idx[r]; sBucket[pos] = (uint16_t) b;
#pragma unroll
            for (int k = 0; k < SW; k++) sPay[k * kPartTile + pos] = w[r][k];
        }
        __syncthreads();
        // write out bucket run after bucket run: neighbouring threads write neighbouring records
        for (uint32_t j = threadIdx.x; j < total; j += kPartThreads) {
            const size_t dst = (size_t) gbase[sBucket[j]] + j;
            P.recIdx[dst] = sIdx[j];
#pragma unroll
            for (int k = 0; k < SW; k++) P.recPay[(size_t) k * P.recCap + dst] = sPay[k * kPartTile + j];
        }
        __syncthreads();
    }
}

// One CTA per bucket: bitmap slice + rank directory built in shared memory from the bucket's records, published to the global index.
__global__ void __launch_bounds__(kFinalThreads) densePartIndexKernel(const __grid_constant__ DensePartParams P)
{
    extern __shared__ __align__(16) unsigned char finalSmem[];
    const int wordsPerBucket = 1 << (P.shift - 6), blocksPerBucket = 1 << (P.shift - 8);
    uint64_t* bm = (uint64_t*) finalSmem;                          // [wordsPerBucket]
    uint32_t* dir = (uint32_t*) (bm + wordsPerBucket);              // [blocksPerBucket] keys of the bucket before the block
    uint32_t* scratch = dir + blocksPerBucket;                      // [64]
    uint32_t* bm32 = (uint32_t*) bm;
    for (int b = blockIdx.x; b < kPartBuckets; b += gridDim.x) {
        const int64_t firstBlock = (int64_t) b << (P.shift - 8);
        if (firstBlock >= P.blocks) break;
        const uint32_t s = P.base[(size_t) b * P.tiles], e = P.base[(size_t) (b + 1) * P.tiles];
        for (int i = threadIdx.x; i < wordsPerBucket; i += kFinalThreads) bm[i] = 0;
        __syncthreads();
        for (uint32_t i0 = s; i0 < e; i0 += 4 * kFinalThreads) {
            uint32_t idx[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { const uint32_t i = i0 + q * kFinalThreads + threadIdx.x; idx[q] = i < e ? P.recIdx[i] : 0xFFFFFFFFu; }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (idx[q] == 0xFFFFFFFFu) continue;
                const uint32_t bit = 1u << (idx[q] & 31);
                if (atomicOr(&bm32[idx[q] >> 5], bit) & bit) *P.duplicate = 1;   // the same key twice: not unique → hash path
            }
        }
        __syncthreads();
        // rank directory: exclusive popcount scan over the bucket's 256-bit blocks
        const int per = (blocksPerBucket + kFinalThreads - 1) / kFinalThreads;
        uint32_t mine = 0;
        for (int q = 0; q < per; q++) {
            const int blk = threadIdx.x * per + q;
            if (blk < blocksPerBucket) mine += __popcll(bm[4 * blk]) + __popcll(bm[4 * blk + 1]) + __popcll(bm[4 * blk + 2]) + __popcll(bm[4 * blk + 3]);
        }
        uint32_t total;
        uint32_t run = blockExclusiveScan(mine, scratch, total);
        for (int q = 0; q < per; q++) {
            const int blk = threadIdx.x * per + q;
            if (blk < blocksPerBucket) { dir[blk] = run; run += __popcll(bm[4 * blk]) + __popcll(bm[4 * blk + 1]) + __popcll(bm[4 * blk + 2]) + __popcll(bm[4 * blk + 3]); }
        }
        __syncthreads();
        const int64_t liveBlocks = min((int64_t) blocksPerBucket, P.blocks - firstBlock);
        for (int64_t i = threadIdx.x; i < liveBlocks * 4; i += kFinalThreads) P.bitmap[firstBlock * 4 + i] = bm[i];
        for (int64_t i = threadIdx.x; i < liveBlocks; i += kFinalThreads) P.prefix[firstBlock + i] = s + dir[i];
        __syncthreads();
    }
}

// kPayloadParts consecutive CTAs per bucket (so the CTAs in flight cover ~20 buckets = ~40 MB of payload rows: L2 resident until the
// sectors are complete): each stages the bucket's slice of the finished index in shared memory and moves its share of the records'
// payload words to their ranks.
constexpr int kPayloadParts = 8;
template <int SW>
__global__ void __launch_bounds__(kFinalThreads) densePartPayloadKernel(const __grid_constant__ DensePartParams P)
{
    extern __shared__ __align__(16) unsigned char finalSmem[];
    const int wordsPerBucket = 1 << (P.shift - 6), blocksPerBucket = 1 << (P.shift - 8);
    uint64_t* bm = (uint64_t*) finalSmem;
    uint32_t* dir = (uint32_t*) (bm + wordsPerBucket);
    const int b = blockIdx.x / kPayloadParts, part = blockIdx.x % kPayloadParts;
    const int64_t firstBlock = (int64_t) b << (P.shift - 8);
    if (firstBlock >= P.blocks) return;
    const uint32_t s = P.base[(size_t) b * P.tiles], e = P.base[(size_t) (b + 1) * P.tiles];
    const uint32_t len = e - s, chunk = (len + kPayloadParts - 1) / kPayloadParts;
    const uint32_t lo = s + min(len, part * chunk), hi = s + min(len, (part + 1) * chunk);
    if (lo >= hi) return;
    const int64_t liveBlocks = min((int64_t) blocksPerBucket, P.blocks - firstBlock);
    for (int64_t i = threadIdx.x; i < liveBlocks * 4; i += kFinalThreads) bm[i] = P.bitmap[firstBlock * 4 + i];
    for (int64_t i = threadIdx.x; i < liveBlocks; i += kFinalThreads) dir[i] = P.prefix[firstBlock + i];     // already includes the keys before the bucket
    __syncthreads();
    for (uint32_t i0 = lo; i0 < hi; i0 += 4 * kFinalThreads) {
        uint32_t idx[4]; uint64_t w[4][SW];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = i0 + q * kFinalThreads + threadIdx.x;
            idx[q] = i < hi ? P.recIdx[i] : 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < SW; k++) w[q][k] = i < hi ? P.recPay[(size_t) k * P.recCap + i] : 0;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (idx[q] == 0xFFFFFFFFu) continue;
            const uint32_t blk = idx[q] >> 8, wi = (idx[q] >> 6) & 3u;
            uint32_t r = dir[blk] + __popcll(bm[idx[q] >> 6] & ((1ULL << (idx[q] & 63)) - 1));
            if (wi > 0) r += __popcll(bm[4 * blk]);
            if (wi > 1) r += __popcll(bm[4 * blk + 1]);
            if (wi > 2) r += __popcll(bm[4 * blk + 2]);
#pragma unroll
            for (int k = 0; k < SW; k++) P.rows[(size_t) r * SW + k] = w[q][k];
        }
    }
}

constexpr int kUnitRounds = 4;
constexpr int kUnitRows = 32 * kUnitRounds;     // rows owned by one warp per step
constexpr int kPreload8 = 4;                    // 8-byte probe output columns whose loads are issued before the index lookups

// Output columns grouped by width so the per-row code has no width switches.
struct CopyColumn { const void* src; void* dst; };
struct BuildField { void* dst; uint8_t* dstNulls; int32_t word, shift, nullWord, nullShift; };   // nullWord < 0: never NULL in the build
struct NullCopy { const uint8_t* src; uint8_t* dst; };

struct DenseProbeParams {
    KeyColumn key; DenseIndex D; int64_t n; int64_t units;
    int32_t outer;
    uint32_t* unitCounts;          // pass A out [units + 1]: matches per unit
    uint32_t* unitMask;            // pass A out / pass B in [units][kUnitRounds]: which lanes of each round matched
    const uint32_t* unitOffsets;   // pass B in: exclusive scan of the counts
    int32_t nProbe8, nProbe4, nProbe1, nProbeNulls;
    CopyColumn probe8[kMaxJoinColumns], probe4[kMaxJoinColumns], probe1[kMaxJoinColumns];
    NullCopy probeNulls[kMaxJoinColumns];
    int32_t nBuild8, nBuild4, nBuild1;
    BuildField build8[kMaxJoinColumns], build4[kMaxJoinColumns], build1[kMaxJoinColumns];
    const uint64_t* rows;          // payload rows in key order
};

// pass A of the probe: matches per unit (keys + index only; ~9 B of HBM traffic per probe row)
__global__ void __launch_bounds__(256) denseCountKernel(const __grid_constant__ DenseProbeParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t unit = warpId; unit < P.units; unit += nWarps) {
        int64_t key[kUnitRounds]; bool valid[kUnitRounds];
#pragma unroll
        for (int j = 0; j < kUnitRounds; j++) {
            const int64_t row = unit * kUnitRows + j * 32 + lane;
            valid[j] = row < P.n && loadKey(P.key, row, key[j]);
        }
        uint32_t c = 0, mine = 0;
#pragma unroll
        for (int j = 0; j < kUnitRounds; j++) {
            const unsigned m = __ballot_sync(0xffffffffu, valid[j] && denseContains(P.D, key[j]));
            c += __popc(m);
            if (lane == j) mine = m;
        }
        if (lane == 0) P.unitCounts[unit] = c;
        if (lane < kUnitRounds) P.unitMask[unit * kUnitRounds + lane] = mine;
    }
}

template <int SW> __device__ __forceinline__ uint64_t payloadWord(const uint64_t (&w)[SW], int k)
{
    if (SW == 1) return w[0];
    uint64_t r = w[0];
#pragma unroll
    for (int i = 1; i < SW; i++) if (k == i) r = w[i];
    return r;
}

// One warp per unit of 128 probe rows, no block-level synchronisation: all independent loads of the unit (keys, 8-byte probe output
// columns) are in flight before the dependent index lookups, payload reads follow the ranks, output positions come from ballots.
template <int SW>
__global__ void __launch_bounds__(256) denseEmitKernel(const __grid_constant__ DenseProbeParams P)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    const int64_t warpId = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t unit = warpId; unit < P.units; unit += nWarps) {
        // inner join: pass A left one match bit per row; units without a match are skipped before any load, and rows without a match
        // are never loaded (sparse joins touch only the sectors of matching rows)
        uint32_t mk[kUnitRounds];
#pragma unroll
        for (int j = 0; j < kUnitRounds; j++) mk[j] = 0xffffffffu;
        if (!P.outer) {
            const uint4 m4 = *reinterpret_cast<const uint4*>(P.unitMask + unit * kUnitRounds);
            mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
            if ((m4.x | m4.y | m4.z | m4.w) == 0) continue;
        }
        int64_t key[kUnitRounds]; bool valid[kUnitRounds]; bool emit[kUnitRounds];
        uint64_t pv[kUnitRounds][kPreload8];
#pragma unroll
        for (int j = 0; j < kUnitRounds; j++) {
            const int64_t row = unit * kUnitRows + j * 32 + lane;
            emit[j] = row < P.n && ((mk[j] >> lane) & 1u);
            valid[j] = emit[j] && loadKey(P.key, row, key[j]);
#pragma unroll
            for (int c = 0; c < kPreload8; c++) { pv[j][c] = 0; if (c < P.nProbe8 && emit[j]) pv[j][c] = ldStream64((const uint64_t*) P.probe8[c].src + row); }
        }
        int64_t rank[kUnitRounds];
#pragma unroll
        for (int j = 0; j < kUnitRounds; j++) rank[j] = valid[j] ? denseRank(P.D, key[j]) : -1;
        uint64_t pw[kUnitRounds][SW];
#pragma unroll
        for (int j = 0; j < kUnitRounds; j++)
#pragma unroll
            for (int k = 0; k < SW; k++) pw[j][k] = rank[j] >= 0 ? P.rows[rank[j] * SW + k] : 0;
        uint32_t cursor = P.outer ? (uint32_t) (unit * kUnitRows) : P.unitOffsets[unit];
        uint32_t o[kUnitRounds];
#pragma unroll
        for (int j = 0; j < kUnitRounds; j++) {
            const unsigned m = P.outer ? __ballot_sync(0xffffffffu, emit[j]) : mk[j];
            o[j] = cursor + __popc(m & lt);
            cursor += __popc(m);
        }
        const int64_t row0 = unit * kUnitRows + lane;
        // column by column, the rounds innermost: pointers and field descriptors are fetched once per unit, not once per round
#pragma unroll
        for (int c = 0; c < kPreload8; c++) if (c < P.nProbe8) {
            uint64_t* dst = (uint64_t*) P.probe8[c].dst;
#pragma unroll
            for (int j = 0; j < kUnitRounds; j++) if (emit[j]) dst[o[j]] = pv[j][c];
        }
        for (int c = kPreload8; c < P.nProbe8; c++) {
            const uint64_t* src = (const uint64_t*) P.probe8[c].src; uint64_t* dst = (uint64_t*) P.probe8[c].dst;
#pragma unroll
            for (int j = 0; j < kUnitRounds; j++) if (emit[j]) dst[o[j]] = ldStream64(src + row0 + j * 32);
        }
        for (int c = 0; c < P.nProbe4; c++) {
            const uint32_t* src = (const uint32_t*) P.probe4[c].src; uint32_t* dst = (uint32_t*) P.probe4[c].dst;
#pragma unroll
            for (int j = 0; j < kUnitRounds; j++) if (emit[j]) dst[o[j]] = ldStream32(src + row0 + j * 32);
        }
        for (int c = 0; c < P.nProbe1; c++) {
            const uint8_t* src = (const uint8_t*) P.probe1[c].src; uint8_t* dst = (uint8_t*) P.probe1[c].dst;
#pragma unroll
            for (int j = 0; j < kUnitRounds; j++) if (emit[j]) dst[o[j]] = src[row0 + j * 32];
        }
        for (int c = 0; c < P.nProbeNulls; c++) {
            const uint8_t* src = P.probeNulls[c].src; uint8_t* dst = P.probeNulls[c].dst;
#pragma unroll
            for (int j = 0; j < kUnitRounds; j++) if (emit[j]) dst[o[j]] = src ? src[row0 + j * 32] : 0;
        }
        for (int c = 0; c < P.nBuild8; c++) {
            const BuildField f = P.build8[c];
#pragma unroll
            for (int j = 0; j < kUnitRounds; j++) if (emit[j]) {
                const bool isNull = rank[j] < 0 || (f.nullWord >= 0 && ((payloadWord<SW>(pw[j], f.nullWord) >> f.nullShift) & 0xff));
                ((uint64_t*) f.dst)[o[j]] = isNull ? 0 : payloadWord<SW>(pw[j], f.word);
                if (f.dstNulls) f.dstNulls[o[j]] = isNull;
            }
        }
        for (int c = 0; c < P.nBuild4; c++) {
            const BuildField f = P.build4[c];
#pragma unroll
            for (int j = 0; j < kUnitRounds; j++) if (emit[j]) {
                const bool isNull = rank[j] < 0 || (f.nullWord >= 0 && ((payloadWord<SW>(pw[j], f.nullWord) >> f.nullShift) & 0xff));
                ((uint32_t*) f.dst)[o[j]] = isNull ? 0 : (uint32_t) (payloadWord<SW>(pw[j], f.word) >> f.shift);
                if (f.dstNulls) f.dstNulls[o[j]] = isNull;
            }
        }
        for (int c = 0; c < P.nBuild1; c++) {
            const BuildField f = P.build1[c];
#pragma unroll
            for (int j = 0; j < kUnitRounds; j++) if (emit[j]) {
                const bool isNull = rank[j] < 0 || (f.nullWord >= 0 && ((payloadWord<SW>(pw[j], f.nullWord) >> f.nullShift) & 0xff));
                ((uint8_t*) f.dst)[o[j]] = isNull ? 0 : (uint8_t) (payloadWord<SW>(pw[j], f.word) >> f.shift);
                if (f.dstNulls) f.dstNulls[o[j]] = isNull;
            }
        }
    }
}

// ======================================================================================================================
// Host side
// ======================================================================================================================
struct LookupSourceDevice {
    int device = 0;
    int64_t positionCount = 0;
    std::vector<int> types;
    std::vector<DeviceBuffer> values, nulls;       // per build channel (only channels in `kept`)
    std::vector<DeviceBuffer> offsets;             // VARCHAR channels: int32[positionCount + 1] into values (the bytes)
    std::vector<int64_t> varcharBytes;             // VARCHAR channels: bytes stored so far
    bool hasVarcharOutput = false;
    std::vector<bool> hasNulls;
    std::vector<int> outputChannels;
    KeyLayout layout;
    DeviceBuffer keys, state, head, links, counters;
    uint32_t cap = 0;
    bool hasDuplicates = false;
    int64_t nullKeyCount = 0, duplicateCount = 0;
    int64_t deviceBytes = 0;
    // dense unique-key index (bitmap + rank) — used instead of the hash table when eligible
    bool dense = false;
    int64_t minKey = 0; uint64_t range = 0;
    DeviceBuffer bitmap, prefix;
    DeviceBuffer payloadRows;      // build OUTPUT columns, row-wise, in key order
    // JoinFilterFunction handed to the build side (HashBuilderOperator's filterFunctionFactory → JoinHash.filterFunction)
    std::vector<pg_expr_node> filterNodes; int filterNode = -1; std::vector<int> filterInputTypes; int filterBuildInputs = 0;
    DeviceBuffer filterLiterals; std::vector<int64_t> filterLiteralOffset;     // bytes of the filter's VARCHAR constants, per node offset (-1: none)
    bool hasFilter() const { return filterNode >= 0; }
    // VARCHAR join keys (PagesHashStrategy compares the slices of the key channels; here they are interned to value ids when the build
    // finishes, the packed key carries the id, and every probe page looks its slices up — an absent value cannot match)
    std::vector<std::unique_ptr<VarcharInterner>> interners;    // per key column, null unless KeyLayout::interned
    std::vector<DeviceBuffer> keyIds;                           // per key column: int32 id per build position
    DeviceBuffer visited;          // outer bridges: one byte per build position, set by the probes' emit kernels (OuterPositionTracker)
    PayloadLayout payload{};
    DenseIndex denseView() const { DenseIndex D{}; D.bitmap = bitmap.as<uint64_t>(); D.prefix = prefix.as<uint32_t>(); D.minKey = minKey; D.range = range; return D; }

    JoinTable view() const
    {
        JoinTable J{};
        J.t.keys = keys.as<uint64_t>(); J.t.state = layout.W > 1 ? state.as<uint32_t>() : nullptr;
        J.t.firstRow = nullptr; J.t.acc = nullptr; J.t.gid = nullptr;
        J.t.cap = cap; J.t.mask = cap - 1; J.t.capP = cap + 2; J.t.nvPhys = 0; J.t.rowClaimLimit = cap;
        J.t.counters = counters.as<unsigned int>();
        J.head = head.as<int32_t>(); J.links = links.as<int32_t>();
        return J;
    }
};

struct JoinBridge {
    int device;
    std::mutex mu;
    std::shared_ptr<LookupSourceDevice> lookupSource;   // set when the build finishes (the LookupSourceFactory future completes)
    // build-outer joins (OuterLookupSource.java:36-203): the lookup source remembers which build positions a probe emitted; the
    // LookupOuterOperator may read them once every probe operator is done (PartitionedLookupSourceFactory's outerPositionsFuture)
    bool outer = false, noMoreProbes = false;
    int liveProbes = 0;
    std::shared_ptr<LookupSourceDevice> get() { std::lock_guard<std::mutex> l(mu); return lookupSource; }
    void set(std::shared_ptr<LookupSourceDevice> ls) { std::lock_guard<std::mutex> l(mu); lookupSource = std::move(ls); }
    bool isOuter() { std::lock_guard<std::mutex> l(mu); return outer; }
    void probeOpened() { std::lock_guard<std::mutex> l(mu); liveProbes++; }
    void probeClosed() { std::lock_guard<std::mutex> l(mu); liveProbes--; }
    bool outerPositionsReady() { std::lock_guard<std::mutex> l(mu); return lookupSource != nullptr && noMoreProbes && liveProbes == 0; }
};

static uint32_t nextPow2u(uint64_t v) { uint64_t p = 2; while (p < v) p <<= 1; return (uint32_t) std::min<uint64_t>(p, 1ULL << 31); }
__global__ void fillU64KernelJ(uint64_t* p, uint64_t v, size_t n)
{
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}

static void checkJoinKeyTypes(const KeyLayout& l)
{
    for (int k = 0; k < l.nKeys; k++)
        if (l.types[k] != PG_TYPE_BIGINT && l.types[k] != PG_TYPE_INTEGER && l.types[k] != PG_TYPE_DATE && l.types[k] != PG_TYPE_VARCHAR)
            throw Error(PG_ERR_UNSUPPORTED, "join keys on the GPU path are BIGINT / INTEGER / DATE / VARCHAR channels");
}

class HashBuilderOp : public Operator {
    JoinBridge* bridge;
    std::shared_ptr<LookupSourceDevice> ls;
    std::vector<int> kept;          // build channels kept on the device (keys + output channels)
    std::vector<int64_t> capacityRows;
    PageStager stager;
    bool finishing = false;
    int64_t expected;

public:
    HashBuilderOp(const pg_hash_builder_config* c, JoinBridge* b) : Operator(c->options), bridge(b), stager(this)
    {
        PG_CHECK_ARG(b != nullptr, "join bridge is null");
        PG_CHECK_ARG(c->n_hash_channels >= 1, "hash channels are empty");
        ls = std::make_shared<LookupSourceDevice>();
        ls->device = options.device;
        ls->types.assign(c->types, c->types + c->n_channels);
        ls->outputChannels.assign(c->output_channels, c->output_channels + c->n_output_channels);
        PG_CHECK_ARG(c->n_output_channels <= kMaxJoinColumns, "too many build output channels");
        std::vector<int32_t> keyTypes;
        for (int i = 0; i < c->n_hash_channels; i++) {
            PG_CHECK_ARG(c->hash_channels[i] >= 0 && c->hash_channels[i] < c->n_channels, "hash channel out of range");
            keyTypes.push_back(ls->types[c->hash_channels[i]]);
        }
        ls->layout.plan(c->n_hash_channels, c->hash_channels, keyTypes.data(), nullptr);
        checkJoinKeyTypes(ls->layout);
        auto keep = [&](int ch) { if (std::find(kept.begin(), kept.end(), ch) == kept.end()) kept.push_back(ch); };
        for (int k = 0; k < ls->layout.nKeys; k++) keep(ls->layout.channels[k]);
        for (int oc : ls->outputChannels) {
            PG_CHECK_ARG(oc >= 0 && oc < c->n_channels, "output channel out of range");
            if (ls->types[oc] == PG_TYPE_VARCHAR) ls->hasVarcharOutput = true;      // (position-ordered columns + hash table: no payload rows)
            keep(oc);
        }
        if (c->filter && c->filter->filter_node >= 0) {
            const pg_filter_project_config& f = *c->filter;
            PG_CHECK_ARG(f.n_input_channels >= c->n_channels && f.n_input_channels <= kMaxInputs, "the join filter's inputs are the build channels followed by the probe channels (at most 32)");
            PG_CHECK_ARG(f.n_nodes >= 1 && f.n_nodes <= kMaxJoinFilterNodes && f.filter_node < f.n_nodes, "join filter: between 1 and 32 expression nodes");
            ls->filterNodes.assign(f.nodes, f.nodes + f.n_nodes);
            ls->filterInputTypes.assign(f.input_types, f.input_types + f.n_input_channels);
            ls->filterNode = f.filter_node; ls->filterBuildInputs = c->n_channels;
            PG_CHECK_ARG(ls->filterNodes[ls->filterNode].type == PG_TYPE_BOOLEAN, "filter must be BOOLEAN");
            if (const char* problem = varcharUseProblem(ls->filterNodes.data(), f.n_nodes)) throw Error(PG_ERR_UNSUPPORTED, problem);
            {   // VARCHAR constants of the filter function go to the device once; the caller's memory is not referenced after create
                ls->filterLiteralOffset.assign(f.n_nodes, -1);
                std::vector<uint8_t> bytes;
                for (int i = 0; i < f.n_nodes; i++) {
                    pg_expr_node& n = ls->filterNodes[i];
                    if (n.op != PG_EX_CONST_VARCHAR) continue;
                    ls->filterLiteralOffset[i] = (int64_t) bytes.size();
                    const uint8_t* src = (const uint8_t*) (uintptr_t) n.imm_i;
                    bytes.insert(bytes.end(), src, src + n.reserved);
                    n.imm_i = 0; n.a = n.b = n.c = -1;
                }
                if (!bytes.empty()) {
                    ls->filterLiterals.accounting = &ls->deviceBytes;
                    ls->filterLiterals.reserve(bytes.size());
                    PG_CUDA(cudaMemcpyAsync(ls->filterLiterals.ptr, bytes.data(), bytes.size(), cudaMemcpyHostToDevice, stream));
                    PG_CUDA(cudaStreamSynchronize(stream));
                }
            }
            for (int i = 0; i < f.n_nodes; i++) {
                const pg_expr_node& n = ls->filterNodes[i];
                if (n.op == PG_EX_INPUT) {
                    PG_CHECK_ARG(n.a >= 0 && n.a < f.n_input_channels && ls->filterInputTypes[n.a] == n.type, "join filter: input node does not match its channel");
                    if (n.type == PG_TYPE_LONG_DECIMAL) throw Error(PG_ERR_UNSUPPORTED, "join filter inputs are channels of at most 8 bytes, or VARCHAR, on the GPU path");
                    if (n.a < c->n_channels) { PG_CHECK_ARG(ls->types[n.a] == n.type, "join filter: build input type differs from the build channel's"); keep(n.a); }
                }
                else for (int o : {n.a, n.b, n.c}) PG_CHECK_ARG(o < i, "expression nodes must be topologically ordered");
            }
        }
        ls->values.resize(c->n_channels); ls->nulls.resize(c->n_channels); ls->hasNulls.assign(c->n_channels, false);
        ls->offsets.resize(c->n_channels); ls->varcharBytes.assign(c->n_channels, 0);
        for (auto& v : ls->offsets) v.accounting = &ls->deviceBytes;
        for (auto& v : ls->values) v.accounting = &ls->deviceBytes;
        for (auto& v : ls->nulls) v.accounting = &ls->deviceBytes;
        capacityRows.assign(c->n_channels, 0);
        expected = std::max<int64_t>(c->expected_positions, 1024);
    }

    bool needsInput() override { return !finishing; }

    void ensureRows(int ch, int64_t rows, bool withNulls)
    {
        int w = typeWidth(ls->types[ch]);
        if (rows > capacityRows[ch]) {
            int64_t newCap = std::max<int64_t>({rows, capacityRows[ch] * 2, expected});
            DeviceBuffer nv; nv.accounting = &ls->deviceBytes; nv.reserve((size_t) newCap * w);
            if (ls->positionCount) PG_CUDA(cudaMemcpyAsync(nv.ptr, ls->values[ch].ptr, (size_t) ls->positionCount * w, cudaMemcpyDeviceToDevice, stream));
            DeviceBuffer nn; nn.accounting = &ls->deviceBytes;
            if (ls->hasNulls[ch] || withNulls) {
                nn.reserve((size_t) newCap);
                if (ls->hasNulls[ch] && ls->positionCount) PG_CUDA(cudaMemcpyAsync(nn.ptr, ls->nulls[ch].ptr, (size_t) ls->positionCount, cudaMemcpyDeviceToDevice, stream));
                else if (ls->positionCount) PG_CUDA(cudaMemsetAsync(nn.ptr, 0, (size_t) ls->positionCount, stream));
            }
            PG_CUDA(cudaStreamSynchronize(stream));
            ls->values[ch] = std::move(nv);
            ls->nulls[ch] = std::move(nn);
            capacityRows[ch] = newCap;
            if (withNulls) ls->hasNulls[ch] = true;
        }
        else if (withNulls && !ls->hasNulls[ch]) {
            ls->nulls[ch].reserve((size_t) capacityRows[ch]);
            if (ls->positionCount) PG_CUDA(cudaMemsetAsync(ls->nulls[ch].ptr, 0, (size_t) ls->positionCount, stream));
            ls->hasNulls[ch] = true;
        }
    }

    // PagesIndex.addPage for a VariableWidthBlock channel: the page's slices are appended to the channel's byte store, its offsets rebased
    std::vector<int64_t> capacityVarBytes, capacityOffsets;
    void appendVarchar(int ch, const DeviceColumn& col, int64_t n)
    {
        if (capacityVarBytes.empty()) { capacityVarBytes.assign(ls->types.size(), 0); capacityOffsets.assign(ls->types.size(), 0); }
        PG_CHECK_ARG(col.offsets != nullptr, "VARCHAR block without offsets");
        int32_t ends[2] = {0, 0};
        PG_CUDA(cudaMemcpyAsync(&ends[0], col.offsets, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaMemcpyAsync(&ends[1], col.offsets + n, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        const int64_t pageBytes = (int64_t) ends[1] - ends[0], have = ls->varcharBytes[ch], pos = ls->positionCount;
        PG_CHECK_ARG(have + pageBytes < (int64_t(1) << 31), "a VARCHAR build channel exceeds 2 GB");
        if (have + pageBytes > capacityVarBytes[ch]) {
            const int64_t cap = std::max<int64_t>({have + pageBytes, capacityVarBytes[ch] * 2, 1 << 16});
            DeviceBuffer nb; nb.accounting = &ls->deviceBytes; nb.reserve((size_t) cap);
            if (have) PG_CUDA(cudaMemcpyAsync(nb.ptr, ls->values[ch].ptr, (size_t) have, cudaMemcpyDeviceToDevice, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
            ls->values[ch] = std::move(nb); capacityVarBytes[ch] = cap;
        }
        if (pos + n + 1 > capacityOffsets[ch]) {
            const int64_t cap = std::max<int64_t>({pos + n + 1, capacityOffsets[ch] * 2, expected + 1});
            DeviceBuffer no; no.accounting = &ls->deviceBytes; no.reserve((size_t) cap * 4);
            if (capacityOffsets[ch]) PG_CUDA(cudaMemcpyAsync(no.ptr, ls->offsets[ch].ptr, (size_t) (pos + 1) * 4, cudaMemcpyDeviceToDevice, stream));
            else PG_CUDA(cudaMemsetAsync(no.ptr, 0, 4, stream));
            DeviceBuffer nn; nn.accounting = &ls->deviceBytes; nn.reserve((size_t) cap);
            if (pos) { if (ls->hasNulls[ch]) PG_CUDA(cudaMemcpyAsync(nn.ptr, ls->nulls[ch].ptr, (size_t) pos, cudaMemcpyDeviceToDevice, stream)); else PG_CUDA(cudaMemsetAsync(nn.ptr, 0, (size_t) pos, stream)); }
            PG_CUDA(cudaStreamSynchronize(stream));
            ls->offsets[ch] = std::move(no); ls->nulls[ch] = std::move(nn); capacityOffsets[ch] = cap;
            ls->hasNulls[ch] = true;       // (a VARCHAR channel always carries its null map: one byte per position)
        }
        if (pageBytes) PG_CUDA(cudaMemcpyAsync(ls->values[ch].as<uint8_t>() + have, (const uint8_t*) col.values + ends[0], (size_t) pageBytes, cudaMemcpyDeviceToDevice, stream));
        PG_LAUNCH(this, rebaseOffsetsKernel, (int) std::max<int64_t>(1, std::min<int64_t>(divUp(n, 256), (int64_t) numSMs * 8)), 256, 0, col.offsets, n, (int32_t) (have - ends[0]), ls->offsets[ch].as<int32_t>() + pos);
        if (col.nulls) PG_CUDA(cudaMemcpyAsync(ls->nulls[ch].as<uint8_t>() + pos, col.nulls, (size_t) n, cudaMemcpyDeviceToDevice, stream));
        else PG_CUDA(cudaMemsetAsync(ls->nulls[ch].as<uint8_t>() + pos, 0, (size_t) n, stream));
        ls->varcharBytes[ch] = have + pageBytes;
    }

    void addInput(const pg_page* page) override   // HashBuilderOperator.addInput :333-349 → PagesIndex.addPage :202-226
    {
        PG_CHECK_STATE(!finishing, "Operator is already finishing");
        const int64_t n = page->position_count;
        if (n == 0) return;
        PG_CHECK_ARG(ls->positionCount + n < (int64_t(1) << 31), "build side exceeds 2^31 positions");
        PG_CHECK_ARG(page->channel_count == (int) ls->types.size(), "build page channel count differs from the operator's configuration");
        std::vector<DeviceColumn> cols;
        int slot = stager.stage(page, kept, cols);
        for (size_t i = 0; i < kept.size(); i++) {
            int ch = kept[i];
            PG_CHECK_ARG(cols[i].type == ls->types[ch], "build channel type differs from the operator's configuration");
            int w = typeWidth(ls->types[ch]);
            if (ls->types[ch] == PG_TYPE_VARCHAR) { appendVarchar(ch, cols[i], n); continue; }
            ensureRows(ch, ls->positionCount + n, cols[i].nulls != nullptr);
            PG_CUDA(cudaMemcpyAsync((char*) ls->values[ch].ptr + (size_t) ls->positionCount * w, cols[i].values, (size_t) n * w, cudaMemcpyDeviceToDevice, stream));
            if (ls->hasNulls[ch]) {
                if (cols[i].nulls) PG_CUDA(cudaMemcpyAsync(ls->nulls[ch].as<uint8_t>() + ls->positionCount, cols[i].nulls, (size_t) n, cudaMemcpyDeviceToDevice, stream));
                else PG_CUDA(cudaMemsetAsync(ls->nulls[ch].as<uint8_t>() + ls->positionCount, 0, (size_t) n, stream));
            }
        }
        stager.markCompute(slot);
        PG_CUDA(cudaStreamSynchronize(stream));   // the borrowed page is not referenced after return
        ls->positionCount += n;
        stats.input_positions += n;
        stats.device_bytes = ls->deviceBytes;
    }

    void finish() override   // finishInput :519-537 → buildLookupSource :642-651 → new PagesHash(...)
    {
        if (finishing) return;
        finishing = true;
        const int64_t n = ls->positionCount;
        if (bridge->isOuter()) {       // the outer operator walks build POSITIONS: keep the position-ordered PagesIndex columns + hash table
            ls->visited.accounting = &ls->deviceBytes;
            ls->visited.reserve((size_t) std::max<int64_t>(n, 1));
            PG_CUDA(cudaMemsetAsync(ls->visited.ptr, 0, (size_t) std::max<int64_t>(n, 1), stream));
        }
        if (n > 0 && !bridge->isOuter() && !ls->hasFilter() && !ls->hasVarcharOutput && tryBuildDense()) {
            stats.group_count = n;
            stats.device_bytes = ls->deviceBytes;
            bridge->set(ls);
            return;
        }
        // HashCommon.arraySize(positionCount, 0.75f): power of two >= n / 0.75 (PagesHash.java:67)
        uint64_t want = (uint64_t) ((double) std::max<int64_t>(n, 1) / 0.75) + 1;
        if (want > (1ULL << 30)) throw Error(PG_ERR_INSUFFICIENT_RESOURCES, "Too large (hash table cannot exceed 2^30 slots)");
        ls->cap = nextPow2u(want);
        const size_t capP = (size_t) ls->cap + 2;
        const int W = ls->layout.W;
        for (DeviceBuffer* b : {&ls->keys, &ls->state, &ls->head, &ls->links, &ls->counters}) b->accounting = &ls->deviceBytes;
        ls->keys.reserve(capP * 8 * W);
        if (W > 1) { ls->state.reserve(capP * 4); PG_CUDA(cudaMemsetAsync(ls->state.ptr, 0, capP * 4, stream)); PG_CUDA(cudaMemsetAsync(ls->keys.ptr, 0, capP * 8 * W, stream)); }
        else { fillU64KernelJ<<<1024, 256, 0, stream>>>(ls->keys.as<uint64_t>(), kEmptyKey, capP); stats.kernel_launches++; }
        ls->head.reserve(capP * 4);
        PG_CUDA(cudaMemsetAsync(ls->head.ptr, 0xFF, capP * 4, stream));
        ls->links.reserve(std::max<size_t>((size_t) n * 4, 16));
        ls->counters.reserve(64);
        PG_CUDA(cudaMemsetAsync(ls->counters.ptr, 0, 64, stream));
        if (n > 0) {
            std::vector<DeviceColumn> keyCols;
            ls->interners.resize(ls->layout.nKeys); ls->keyIds.resize(ls->layout.nKeys);
            for (int k = 0; k < ls->layout.nKeys; k++) {
                int ch = ls->layout.channels[k];
                DeviceColumn c; c.type = ls->types[ch]; c.n = n; c.values = ls->values[ch].ptr; c.nulls = ls->hasNulls[ch] ? ls->nulls[ch].as<uint8_t>() : nullptr;
                if (ls->layout.interned[k]) {      // the channel's slices (PagesIndex order) → one id per build position
                    c.offsets = ls->offsets[ch].as<int32_t>();
                    ls->interners[k].reset(new VarcharInterner(this, &ls->deviceBytes));
                    ls->keyIds[k].accounting = &ls->deviceBytes;
                    ls->keyIds[k].reserve((size_t) n * 4 + 16);
                    ls->interners[k]->intern(this, c, n, ls->keyIds[k].as<int32_t>(), true);
                    c.type = PG_TYPE_INTEGER; c.values = ls->keyIds[k].ptr; c.offsets = nullptr;
                }
                keyCols.push_back(c);
            }
            KeyPlan kp{};
            ls->layout.fill(kp, keyCols, ls->counters.as<unsigned int>() + CNT_ERROR);
            JoinTable J = ls->view();
            unsigned long long* dstats = (unsigned long long*) (ls->counters.as<unsigned int>() + CNT_COUNT);
            const int grid = (int) std::min<int64_t>(divUp(n, 256), numSMs * 16);
            switch (W) {
                case 1: PG_LAUNCH(this, joinBuildKernel<1>, grid, 256, 0, J, kp, n, dstats); break;
                case 2: PG_LAUNCH(this, joinBuildKernel<2>, grid, 256, 0, J, kp, n, dstats); break;
                case 3: PG_LAUNCH(this, joinBuildKernel<3>, grid, 256, 0, J, kp, n, dstats); break;
                default: PG_LAUNCH(this, joinBuildKernel<4>, grid, 256, 0, J, kp, n, dstats); break;
            }
            unsigned long long h[2] = {0, 0};
            PG_CUDA(cudaMemcpyAsync(h, dstats, sizeof(h), cudaMemcpyDeviceToHost, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
            ls->duplicateCount = (int64_t) h[0]; ls->nullKeyCount = (int64_t) h[1];
            ls->hasDuplicates = h[0] > 0;
            if (ls->hasDuplicates) {
                PG_LAUNCH(this, joinSortChainsKernel, (int) std::min<int64_t>(divUp((int64_t) capP, 256), numSMs * 16), 256, 0, J);
                PG_CUDA(cudaStreamSynchronize(stream));
            }
        }
        else PG_CUDA(cudaStreamSynchronize(stream));
        stats.group_count = n;
        stats.device_bytes = ls->deviceBytes;
        bridge->set(ls);
    }
    // Unique keys over a small range → bitmap + rank index (see DenseIndex).  Returns false to fall back to the hash table.
    bool tryBuildDense()
    {
        if (ls->layout.nKeys != 1 || ls->layout.anyInterned()) return false;
        const int64_t n = ls->positionCount;
        // payload row layout: widest columns first (natural alignment), then one null byte per nullable column
        PayloadLayout payloadLayout{};
        {
            int off = 0;
            for (int width : {8, 4, 1})
                for (size_t i = 0; i < ls->outputChannels.size(); i++)
                    if (typeWidth(ls->types[ls->outputChannels[i]]) == width) { payloadLayout.cols[i].offset = off; payloadLayout.cols[i].width = width; payloadLayout.cols[i].nullOffset = -1; off += width; }
            for (size_t i = 0; i < ls->outputChannels.size(); i++)
                if (ls->hasNulls[ls->outputChannels[i]]) payloadLayout.cols[i].nullOffset = off++;
            payloadLayout.strideWords = std::max(1, (off + 7) / 8);
            if (payloadLayout.strideWords == 3) payloadLayout.strideWords = 4;
            if (payloadLayout.strideWords > kMaxPayloadWords) return false;   // wide payloads keep the hash-table path
        }
        const int kch = ls->layout.channels[0];
        KeyColumn kc{ls->values[kch].ptr, ls->hasNulls[kch] ? ls->nulls[kch].as<uint8_t>() : nullptr, ls->types[kch]};
        DeviceBuffer mm; mm.reserve(64);
        long long init[3] = {LLONG_MAX, LLONG_MIN, 0};
        PG_CUDA(cudaMemcpyAsync(mm.ptr, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        const int grid = (int) std::min<int64_t>(divUp(n, 256), numSMs * 16);
        PG_LAUNCH(this, denseMinMaxKernel, grid, 256, 0, kc, n, mm.as<long long>());
        long long h[3];
        PG_CUDA(cudaMemcpyAsync(h, mm.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        const int64_t nonNull = h[2];
        if (nonNull == 0) return false;
        const unsigned __int128 span = (unsigned __int128) ((__int128) h[1] - (__int128) h[0]) + 1;
        // dense enough (bitmap <= 8 x the key column), or small in absolute terms: a hash-partitioned build side keeps the full key range
        // on every rank with 1/P of the keys — the bitmap (range / 8 bytes: 75 MB for TPC-H SF100 order keys) still fits L2
        const bool denseEnough = span <= (unsigned __int128) kDenseRangeFactor * (unsigned __int128) nonNull + 4096;
        const bool smallEnough = span <= ((unsigned __int128) 1 << kDenseSmallRangeBits) && nonNull >= 65536 && span <= (unsigned __int128) 4096 * (unsigned __int128) nonNull;   // (not for a handful of far-apart keys)
        if ((!denseEnough && !smallEnough) || span > ((unsigned __int128) 1 << 36)) return false;
        const uint64_t range = (uint64_t) span;
        const int64_t blocks = (int64_t) ((range + 255) >> 8);
        ls->bitmap.accounting = ls->prefix.accounting = &ls->deviceBytes;
        ls->bitmap.reserve((size_t) blocks * 32);
        PG_CUDA(cudaMemsetAsync(ls->bitmap.ptr, 0, (size_t) blocks * 32, stream));
        DeviceBuffer dup; dup.reserve(16);
        PG_CUDA(cudaMemsetAsync(dup.ptr, 0, 16, stream));
        // large builds with a payload over a range of at most 2^30: range-partition first (see DensePartParams)
        int64_t partMinRows = kPartBuildMinRows;
        if (const char* e = getenv("PG_JOIN_BUCKET_MIN_ROWS")) partMinRows = atoll(e);     // tests force the partitioned build on small inputs
        int shift = kPartMinShift;
        while (((range - 1) >> shift) >= (uint64_t) kPartBuckets) shift++;
        // measured on TPC-H Q3's second build (14.6 M filtered orders over a range of 600 M, density 1/41, arrival order random): scatter
        // 1.01 ms, sorted pairs 0.83 ms, range-partitioned 0.71 ms — the partition pass costs per row, the per-bucket bitmap slices per key value
        int64_t partDensity = kDenseRangeFactor;
        if (const char* e = getenv("PG_JOIN_BUCKET_DENSITY")) partDensity = std::max<int64_t>(1, atoll(e));
        const bool partitioned = n >= partMinRows && !ls->outputChannels.empty() && shift <= kPartMaxShift && range <= (uint64_t) partDensity * (uint64_t) nonNull;
        // sparse key sets in arrival order with a one-word payload: sort (key - min, payload) pairs instead of scattering (densePackPairsKernel)
        int64_t sortMinRows = 1 << 20;
        if (const char* e = getenv("PG_JOIN_SORTED_MIN_ROWS")) sortMinRows = atoll(e);      // tests force it on small inputs
        const bool sortedBuild = !partitioned && !ls->outputChannels.empty() && payloadLayout.strideWords == 1 && range <= 0xFFFFFFFFULL && nonNull >= sortMinRows;
        if (!partitioned && !sortedBuild) PG_LAUNCH(this, denseSetBitsKernel, grid, 256, 0, kc, n, (int64_t) h[0], ls->bitmap.as<unsigned long long>(), dup.as<unsigned int>());
        ls->prefix.reserve((size_t) (blocks + 1) * 4);
        ls->minKey = h[0]; ls->range = range;
        // build output columns as payload rows in key order
        DenseGatherParams G{};
        G.key = kc; G.n = n;
        G.L = payloadLayout; G.L.nCols = (int) ls->outputChannels.size();
        for (size_t i = 0; i < ls->outputChannels.size(); i++) {
            const int ch = ls->outputChannels[i];
            G.src[i] = ls->values[ch].ptr; G.srcNulls[i] = ls->hasNulls[ch] ? ls->nulls[ch].as<uint8_t>() : nullptr;
        }
        ls->payload = G.L;
        ls->payloadRows.accounting = &ls->deviceBytes;
        ls->payloadRows.reserve(std::max<size_t>((size_t) nonNull * G.L.strideWords * 8, 16));
        G.rows = ls->payloadRows.as<uint64_t>();
        auto failOnDuplicate = [&]() {
            unsigned int isDup = 0;
            PG_CUDA(cudaMemcpyAsync(&isDup, dup.ptr, 4, cudaMemcpyDeviceToHost, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
            if (isDup) { ls->bitmap.release(); ls->prefix.release(); ls->payloadRows.release(); }
            return isDup != 0;
        };
        if (partitioned) {
            DensePartParams B{};
            DeviceBuffer hist, base, temp, recIdx, recPay;
            B.key = kc; B.n = n; B.tiles = divUp(n, kPartTile); B.minKey = (int64_t) h[0]; B.shift = shift;
            const size_t histN = (size_t) kPartBuckets * B.tiles + 1;
            hist.reserve(histN * 4); base.reserve(histN * 4);
            B.hist = hist.as<uint32_t>(); B.base = base.as<uint32_t>();
            PG_CUDA(cudaMemsetAsync(B.hist + histN - 1, 0, 4, stream));
            const int SW = G.L.strideWords;
            // tile staging: counts/starts + global bases + scan scratch + idx + bucket + payload planes
            const size_t partSmem = (size_t) kPartBuckets * 8 + 256 + (size_t) kPartTile * (4 + 2 + 8 * SW);
            const size_t finalSmem = ((size_t) 1 << (shift - 3)) + ((size_t) 4 << (shift - 8)) + 256;
            const int pgrid = (int) std::min<int64_t>(B.tiles, (int64_t) numSMs * (partSmem > 100 * 1024 ? 1 : partSmem > 70 * 1024 ? 2 : 3));
            PG_LAUNCH(this, densePartHistKernel, (int) std::min<int64_t>(B.tiles, numSMs * 4), kPartThreads, 0, B);
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, B.hist, base.as<uint32_t>(), (int) histN, stream);
            temp.reserve(tb);
            cub::DeviceScan::ExclusiveSum(temp.ptr, tb, B.hist, base.as<uint32_t>(), (int) histN, stream);
            stats.kernel_launches += 2;
            recIdx.reserve((size_t) nonNull * 4); recPay.reserve((size_t) nonNull * SW * 8);
            B.recIdx = recIdx.as<uint32_t>(); B.recPay = recPay.as<uint64_t>(); B.recCap = nonNull;
            B.L = G.L;
            for (int i = 0; i < G.L.nCols; i++) { B.src[i] = G.src[i]; B.srcNulls[i] = G.srcNulls[i]; }
            B.bitmap = ls->bitmap.as<uint64_t>(); B.prefix = ls->prefix.as<uint32_t>(); B.blocks = blocks;
            B.rows = G.rows; B.duplicate = dup.as<unsigned int>();
            const int usedBuckets = (int) (((range - 1) >> shift) + 1);
            const int fgrid = usedBuckets;
#define PG_PART_BUILD(SWV) do { \
                PG_CUDA(cudaFuncSetAttribute(densePartScatterKernel<SWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) partSmem)); \
                PG_CUDA(cudaFuncSetAttribute(densePartIndexKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) finalSmem)); \
                PG_CUDA(cudaFuncSetAttribute(densePartPayloadKernel<SWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) finalSmem)); \
                PG_LAUNCH(this, densePartScatterKernel<SWV>, pgrid, kPartThreads, partSmem, B); \
                PG_LAUNCH(this, densePartIndexKernel, fgrid, kFinalThreads, finalSmem, B); \
                PG_LAUNCH(this, densePartPayloadKernel<SWV>, fgrid * kPayloadParts, kFinalThreads, finalSmem, B); } while (0)
            if (SW == 1) PG_PART_BUILD(1); else if (SW == 2) PG_PART_BUILD(2); else PG_PART_BUILD(4);
#undef PG_PART_BUILD
            const uint32_t total = (uint32_t) nonNull;
            PG_CUDA(cudaMemcpyAsync(ls->prefix.as<uint32_t>() + blocks, &total, 4, cudaMemcpyHostToDevice, stream));
            if (failOnDuplicate()) return false;     // (also waits for the kernels: the temporaries above are released on return)
        }
        else {
            DeviceBuffer keys32, keysSorted, payIn, sortTemp;
            if (sortedBuild) {
                keys32.reserve((size_t) nonNull * 4); keysSorted.reserve((size_t) nonNull * 4); payIn.reserve((size_t) nonNull * 8);
                PG_LAUNCH(this, densePackPairsKernel, grid, 256, 0, G, (int64_t) h[0], keys32.as<uint32_t>(), payIn.as<uint64_t>(), dup.as<unsigned int>() + 1, nonNull != n ? 1 : 0);
                int endBit = 1;
                while (endBit < 32 && ((range - 1) >> endBit) != 0) endBit++;
                size_t sb = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, sb, keys32.as<uint32_t>(), keysSorted.as<uint32_t>(), payIn.as<uint64_t>(), G.rows, (int) nonNull, 0, endBit, stream);
                sortTemp.reserve(sb);
                cub::DeviceRadixSort::SortPairs(sortTemp.ptr, sb, keys32.as<uint32_t>(), keysSorted.as<uint32_t>(), payIn.as<uint64_t>(), G.rows, (int) nonNull, 0, endBit, stream);
                stats.kernel_launches += 1 + (endBit + 7) / 8 + 1;
                PG_LAUNCH(this, denseSortedBitsKernel, grid, 256, 0, keysSorted.as<uint32_t>(), nonNull, ls->bitmap.as<unsigned long long>(), dup.as<unsigned int>());
            }
            if (failOnDuplicate()) return false;      // (waits for the kernels: the sort's temporaries may go)
            // rank directory
            DeviceBuffer counts, temp; counts.reserve((size_t) (blocks + 1) * 4);
            PG_CUDA(cudaMemsetAsync(counts.as<uint32_t>() + blocks, 0, 4, stream));
            PG_LAUNCH(this, denseBlockCountKernel, (int) std::min<int64_t>(divUp(blocks, 256), numSMs * 16), 256, 0, ls->bitmap.as<uint64_t>(), blocks, counts.as<uint32_t>());
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.as<uint32_t>(), ls->prefix.as<uint32_t>(), (int) (blocks + 1), stream);
            temp.reserve(tb);
            cub::DeviceScan::ExclusiveSum(temp.ptr, tb, counts.as<uint32_t>(), ls->prefix.as<uint32_t>(), (int) (blocks + 1), stream);
            stats.kernel_launches += 2;
            G.D = ls->denseView();
            if (G.L.nCols && !sortedBuild) PG_LAUNCH(this, denseGatherKernel, grid, 256, 0, G);
            PG_CUDA(cudaStreamSynchronize(stream));
        }
        ls->nullKeyCount = n - nonNull;
        ls->dense = true;
        // the unsorted copies of the build columns are no longer needed
        for (auto& v : ls->values) v.release();
        for (auto& v : ls->nulls) v.release();
        return true;
    }

    bool isFinished() override { return finishing; }
    bool getOutput(pg_page*) override { return false; }
};

class LookupJoinOp : public Operator {
    JoinBridge* bridge;
    int joinType;
    std::vector<int> types, outputChannels, staged;
    KeyLayout layout;
    PageStager stager;
    bool finishing = false, hasPending = false;
    DeviceBuffer matchHead, counts, offsets, scanTemp, unitMask, exactTotal;
    std::vector<DeviceBuffer> outValues, outNulls;
    std::vector<PinnedBuffer> hostValues, hostNulls;
    std::vector<pg_block> outBlocks;
    int64_t pendingRows = 0;
    bool trackedByBridge = false;     // counted in bridge->liveProbes until finish() / destruction
    DeviceBuffer filterDev;           // JoinFilterProgram + its error word, rebuilt per page (the probe columns move)
    // VARCHAR output channels (gatherVarcharByIndex): source index columns written by the emit kernels, then offsets / bytes per slot
    std::vector<DeviceBuffer> outOffsets; std::vector<PinnedBuffer> hostOffsets; std::vector<int64_t> varBytes;
    DeviceBuffer iotaProbe, iotaBuild, srcRow, srcPos, srcPosNull, varScan;
    std::vector<DeviceBuffer> probeKeyIds;      // VARCHAR key channels of the page in flight as value ids of the LookupSource's interners
    struct VarcharOut { size_t slot; bool fromBuild; int channel; };
    bool filterChannelsStaged = false;

public:
    ~LookupJoinOp() override { if (trackedByBridge) bridge->probeClosed(); }
    LookupJoinOp(const pg_lookup_join_config* c, JoinBridge* b) : Operator(c->options), bridge(b), stager(this)
    {
        PG_CHECK_ARG(b != nullptr, "join bridge is null");
        PG_CHECK_ARG(c->join_type >= PG_JOIN_INNER && c->join_type <= PG_JOIN_FULL_OUTER, "unknown join type");
        if ((c->join_type == PG_JOIN_LOOKUP_OUTER || c->join_type == PG_JOIN_FULL_OUTER) && !b->isOuter())
            throw Error(PG_ERR_INVALID_ARGUMENT, "LOOKUP_OUTER / FULL_OUTER joins need a bridge created for outer tracking (pg_join_bridge_set_outer)");
        // the probe side of LOOKUP_OUTER behaves like INNER, of FULL_OUTER like PROBE_OUTER (LookupJoinOperators.java: probeOuterJoin / fullOuterJoin)
        joinType = (c->join_type == PG_JOIN_PROBE_OUTER || c->join_type == PG_JOIN_FULL_OUTER) ? PG_JOIN_PROBE_OUTER : PG_JOIN_INNER;
        trackedByBridge = b->isOuter();
        if (trackedByBridge) b->probeOpened();
        types.assign(c->types, c->types + c->n_channels);
        outputChannels.assign(c->output_channels, c->output_channels + c->n_output_channels);
        PG_CHECK_ARG(c->n_output_channels <= kMaxJoinColumns, "too many probe output channels");
        std::vector<int32_t> keyTypes;
        for (int i = 0; i < c->n_hash_channels; i++) {
            PG_CHECK_ARG(c->hash_channels[i] >= 0 && c->hash_channels[i] < c->n_channels, "hash channel out of range");
            keyTypes.push_back(types[c->hash_channels[i]]);
        }
        layout.plan(c->n_hash_channels, c->hash_channels, keyTypes.data(), nullptr);
        checkJoinKeyTypes(layout);
        auto need = [&](int ch) { if (std::find(staged.begin(), staged.end(), ch) == staged.end()) staged.push_back(ch); };
        for (int k = 0; k < layout.nKeys; k++) need(layout.channels[k]);
        for (int oc : outputChannels) {
            PG_CHECK_ARG(oc >= 0 && oc < c->n_channels, "output channel out of range");
            need(oc);
        }
        for (DeviceBuffer* d : {&matchHead, &counts, &offsets, &scanTemp, &unitMask}) d->accounting = &stats.device_bytes;
    }

    // LookupJoinOperator.needsInput :198-210: blocked until the lookup source future completes
    bool needsInput() override { return !finishing && !hasPending && bridge->get() != nullptr; }
    bool isBlocked() override { return bridge->get() == nullptr; }

    void addInput(const pg_page* page) override
    {
        PG_CHECK_STATE(needsInput(), "lookup join does not need input (still building, finishing, or output pending)");
        auto ls = bridge->get();
        PG_CHECK_ARG(ls->layout.nKeys == layout.nKeys, "probe and build hash channel counts differ");
        for (int k = 0; k < layout.nKeys; k++) PG_CHECK_ARG(ls->layout.types[k] == layout.types[k], "probe and build join key types differ");
        const int64_t n = page->position_count;
        stats.input_positions += n;
        if (n == 0) return;
        PG_CHECK_ARG(n < (int64_t(1) << 31), "probe page exceeds 2^31 positions");
        if (ls->hasFilter() && !filterChannelsStaged) {      // the filter function reads probe channels of its own
            PG_CHECK_ARG((int) ls->filterInputTypes.size() == ls->filterBuildInputs + (int) types.size(), "the join filter's inputs are the build channels followed by the probe channels");
            for (const pg_expr_node& nd : ls->filterNodes)
                if (nd.op == PG_EX_INPUT && nd.a >= ls->filterBuildInputs) {
                    const int ch = nd.a - ls->filterBuildInputs;
                    PG_CHECK_ARG(types[ch] == nd.type, "join filter: probe input type differs from the probe channel's");
                    if (std::find(staged.begin(), staged.end(), ch) == staged.end()) staged.push_back(ch);
                }
            filterChannelsStaged = true;
        }
        std::vector<DeviceColumn> cols;
        int slot = stager.stage(page, staged, cols);
        auto colOf = [&](int ch) -> const DeviceColumn& { return cols[std::find(staged.begin(), staged.end(), ch) - staged.begin()]; };

        if (ls->dense) { addInputDense(*ls, cols, slot, n); return; }
        const JoinFilterProgram* filterProgram = nullptr;
        if (ls->hasFilter()) {
            JoinFilterProgram F{};
            F.nNodes = (int) ls->filterNodes.size(); F.filterNode = ls->filterNode; F.nBuildInputs = ls->filterBuildInputs;
            for (int i = 0; i < F.nNodes; i++) {
                FpNode& d = F.nodes[i]; const pg_expr_node& sn = ls->filterNodes[i];
                d.op = sn.op; d.type = sn.type; d.a = sn.a; d.b = sn.b; d.c = sn.c; d.isDouble = sn.type == PG_TYPE_DOUBLE;
                if (sn.op == PG_EX_CONST_DOUBLE) memcpy(&d.immI, &sn.imm_d, 8); else d.immI = sn.imm_i;
                if (sn.op == PG_EX_CONST_VARCHAR) { d.immI = (int64_t) (uintptr_t) (ls->filterLiterals.as<uint8_t>() + ls->filterLiteralOffset[i]); d.len = sn.reserved; d.a = d.b = d.c = -1; }
                if (sn.op != PG_EX_INPUT) continue;
                FpInput& in = F.inputs[sn.a];
                in.type = sn.type;
                if (sn.a < F.nBuildInputs) {
                    in.values = ls->values[sn.a].ptr; in.nulls = ls->hasNulls[sn.a] ? ls->nulls[sn.a].as<uint8_t>() : nullptr;
                    in.offsets = sn.type == PG_TYPE_VARCHAR ? ls->offsets[sn.a].as<int32_t>() : nullptr;
                }
                else { const DeviceColumn& c = colOf(sn.a - F.nBuildInputs); in.values = c.values; in.nulls = c.nulls; in.offsets = c.offsets; }
            }
            filterDev.accounting = &stats.device_bytes;
            filterDev.reserve(sizeof(JoinFilterProgram) + 16);
            F.error = (unsigned int*) ((char*) filterDev.ptr + sizeof(JoinFilterProgram));
            PG_CUDA(cudaMemsetAsync(F.error, 0, 4, stream));
            PG_CUDA(cudaMemcpyAsync(filterDev.ptr, &F, sizeof(F), cudaMemcpyHostToDevice, stream));     // (pageable source: staged before the call returns)
            filterProgram = (const JoinFilterProgram*) filterDev.ptr;
        }
        // pass 1: probe
        matchHead.reserve((size_t) n * 4); counts.reserve((size_t) (n + 1) * 4); offsets.reserve((size_t) (n + 1) * 4);
        ProbeParams P{};
        P.J = ls->view();
        std::vector<DeviceColumn> keyCols;
        for (int k = 0; k < layout.nKeys; k++) {
            keyCols.push_back(colOf(layout.channels[k]));
            PG_CHECK_ARG(keyCols.back().type == layout.types[k], "probe key channel type differs from configuration");
            if (!layout.interned[k]) continue;
            // VARCHAR key: the id the build gave this slice, -1 when the build never saw it (read-only lookup on this operator's stream)
            if (probeKeyIds.size() < (size_t) layout.nKeys) { probeKeyIds.resize(layout.nKeys); for (auto& b : probeKeyIds) b.accounting = &stats.device_bytes; }
            probeKeyIds[k].reserve((size_t) n * 4 + 16);
            DeviceColumn& c = keyCols.back();
            if (k < (int) ls->interners.size() && ls->interners[k]) ls->interners[k]->intern(this, c, n, probeKeyIds[k].as<int32_t>(), false);
            else PG_CUDA(cudaMemsetAsync(probeKeyIds[k].ptr, 0xFF, (size_t) n * 4, stream));       // empty build: nothing matches
            c.type = PG_TYPE_INTEGER; c.values = probeKeyIds[k].ptr; c.offsets = nullptr;
        }
        layout.fill(P.kp, keyCols, ls->counters.as<unsigned int>() + CNT_ERROR);
        P.n = n; P.outer = joinType == PG_JOIN_PROBE_OUTER; P.hasDuplicates = ls->hasDuplicates; P.matchHead = matchHead.as<int32_t>(); P.counts = counts.as<uint32_t>();
        PG_CUDA(cudaMemsetAsync(counts.as<uint32_t>() + n, 0, 4, stream));
        exactTotal.reserve(16);
        PG_CUDA(cudaMemsetAsync(exactTotal.ptr, 0, 8, stream));
        P.total = exactTotal.as<unsigned long long>();
        P.filter = filterProgram;
        const int grid = (int) std::min<int64_t>(divUp(n, 256), numSMs * 16);
        switch (layout.W) {
            case 1: PG_LAUNCH(this, joinProbeKernel<1>, grid, 256, 0, P); break;
            case 2: PG_LAUNCH(this, joinProbeKernel<2>, grid, 256, 0, P); break;
            case 3: PG_LAUNCH(this, joinProbeKernel<3>, grid, 256, 0, P); break;
            default: PG_LAUNCH(this, joinProbeKernel<4>, grid, 256, 0, P); break;
        }
        // counts → offsets
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.as<uint32_t>(), offsets.as<uint32_t>(), (int) (n + 1), stream);
        scanTemp.reserve(tb);
        cub::DeviceScan::ExclusiveSum(scanTemp.ptr, tb, counts.as<uint32_t>(), offsets.as<uint32_t>(), (int) (n + 1), stream);
        stats.kernel_launches += 2;
        uint32_t total = 0;
        unsigned long long total64 = 0;
        PG_CUDA(cudaMemcpyAsync(&total, offsets.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaMemcpyAsync(&total64, exactTotal.ptr, 8, cudaMemcpyDeviceToHost, stream));
        unsigned int filterError = 0;
        if (filterProgram) PG_CUDA(cudaMemcpyAsync(&filterError, (const char*) filterDev.ptr + sizeof(JoinFilterProgram), 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        stats.d2h_bytes += 12;
        if (filterError) {
            stager.markCompute(slot);
            if (filterError & FP_ERR_OVERFLOW) throw Error(PG_ERR_NUMERIC_OVERFLOW, "bigint arithmetic overflow in the join filter");
            throw Error(PG_ERR_INVALID_ARGUMENT, "Division by zero");
        }
        // the 32-bit scan wraps silently beyond 2^32 output rows; a Page holds < 2^31 positions (the reference bounds every output page
        // through LookupJoinPageBuilder; here the caller has to feed smaller probe pages)
        if (total64 >= (1ULL << 31)) {
            stager.markCompute(slot);
            throw Error(PG_ERR_INSUFFICIENT_RESOURCES, "join output of one probe page exceeds 2^31 positions (" + std::to_string(total64) + "): feed smaller probe pages");
        }
        pendingRows = total;
        if (total == 0) { stager.markCompute(slot); return; }

        // pass 2: write the output columns in place
        const size_t nOut = outputChannels.size() + ls->outputChannels.size();
        ensureOutputSlots(nOut);
        outBlocks.assign(nOut, pg_block{});
        std::vector<bool> hasNulls(nOut, false);
        std::vector<VarcharOut> vouts;
        EmitJoinParams E{};
        E.matchHead = matchHead.as<int32_t>(); E.offsets = offsets.as<uint32_t>(); E.links = ls->links.as<int32_t>();
        E.n = n; E.hasDuplicates = ls->hasDuplicates;
        E.visited = ls->visited.as<uint8_t>();
        E.filter = filterProgram;
        int np = 0, nb = 0;
        bool probeVarchar = false, buildVarchar = false;
        for (size_t i = 0; i < outputChannels.size(); i++) {
            const DeviceColumn& c = colOf(outputChannels[i]);
            outBlocks[i].type = c.type;
            if (c.type == PG_TYPE_VARCHAR) { vouts.push_back({i, false, outputChannels[i]}); probeVarchar = true; hasNulls[i] = c.nulls != nullptr; continue; }
            JoinColumn& jc = E.probe[np++];
            jc.width = typeWidth(c.type); jc.src = c.values; jc.srcNulls = c.nulls;
            outValues[i].reserve((size_t) total * jc.width);
            jc.dst = outValues[i].ptr; jc.dstNulls = nullptr;
            if (c.nulls) { outNulls[i].reserve(total); jc.dstNulls = outNulls[i].as<uint8_t>(); hasNulls[i] = true; }
        }
        for (size_t i = 0; i < ls->outputChannels.size(); i++) {
            const int ch = ls->outputChannels[i];
            const size_t o = outputChannels.size() + i;
            outBlocks[o].type = ls->types[ch];
            const bool nullable = ls->hasNulls[ch] || joinType == PG_JOIN_PROBE_OUTER;
            if (ls->types[ch] == PG_TYPE_VARCHAR) { vouts.push_back({o, true, ch}); buildVarchar = true; hasNulls[o] = nullable; continue; }
            JoinColumn& jc = E.build[nb++];
            jc.width = typeWidth(ls->types[ch]); jc.src = ls->values[ch].ptr; jc.srcNulls = ls->hasNulls[ch] ? ls->nulls[ch].as<uint8_t>() : nullptr;
            outValues[o].reserve((size_t) total * jc.width);
            jc.dst = outValues[o].ptr; jc.dstNulls = nullptr;
            if (nullable) { outNulls[o].reserve(total); jc.dstNulls = outNulls[o].as<uint8_t>(); hasNulls[o] = true; }
        }
        // VARCHAR outputs: one more 4-byte column per side carries the source index of every output row
        PG_CHECK_ARG(np + (probeVarchar ? 1 : 0) <= kMaxJoinColumns && nb + (buildVarchar ? 1 : 0) <= kMaxJoinColumns, "too many output channels");
        if (probeVarchar) {
            iotaProbe.reserve((size_t) n * 4); srcRow.reserve((size_t) total * 4);
            PG_LAUNCH(this, iotaU32Kernel, grid, 256, 0, iotaProbe.as<uint32_t>(), n);
            JoinColumn& jc = E.probe[np++];
            jc.width = 4; jc.src = iotaProbe.ptr; jc.srcNulls = nullptr; jc.dst = srcRow.ptr; jc.dstNulls = nullptr;
        }
        if (buildVarchar) {
            iotaBuild.reserve((size_t) std::max<int64_t>(ls->positionCount, 1) * 4); srcPos.reserve((size_t) total * 4);
            PG_LAUNCH(this, iotaU32Kernel, grid, 256, 0, iotaBuild.as<uint32_t>(), ls->positionCount);
            JoinColumn& jc = E.build[nb++];
            jc.width = 4; jc.src = iotaBuild.ptr; jc.srcNulls = nullptr; jc.dst = srcPos.ptr; jc.dstNulls = nullptr;
            if (joinType == PG_JOIN_PROBE_OUTER) { srcPosNull.reserve(total); jc.dstNulls = srcPosNull.as<uint8_t>(); }
        }
        E.nProbe = np; E.nBuild = nb;
        PG_LAUNCH(this, joinEmitKernel, grid, 256, 0, E);
        for (const VarcharOut& v : vouts) {
            outNulls[v.slot].reserve(total);
            uint8_t* dstNulls = hasNulls[v.slot] ? outNulls[v.slot].as<uint8_t>() : nullptr;
            if (v.fromBuild)
                varBytes[v.slot] = gatherVarcharByIndex(this, ls->values[v.channel].as<uint8_t>(), ls->offsets[v.channel].as<int32_t>(), ls->nulls[v.channel].as<uint8_t>(), srcPos.as<uint32_t>(),
                                                        joinType == PG_JOIN_PROBE_OUTER ? srcPosNull.as<uint8_t>() : nullptr, total, outValues[v.slot], outOffsets[v.slot], dstNulls, varScan);
            else {
                const DeviceColumn& c = colOf(v.channel);
                PG_CHECK_ARG(c.offsets != nullptr, "VARCHAR block without offsets");
                varBytes[v.slot] = gatherVarcharByIndex(this, (const uint8_t*) c.values, c.offsets, c.nulls, srcRow.as<uint32_t>(), nullptr, total, outValues[v.slot], outOffsets[v.slot], dstNulls, varScan);
            }
        }
        stager.markCompute(slot);
        publish(total, hasNulls);
    }

    void ensureOutputSlots(size_t nOut)
    {
        if (outValues.size() >= nOut) return;
        outValues.resize(nOut); outNulls.resize(nOut); outOffsets.resize(nOut); varBytes.resize(nOut, 0);
        for (auto* v : {&outValues, &outNulls, &outOffsets}) for (auto& d : *v) d.accounting = &stats.device_bytes;
    }

    // Probe of the bitmap + rank index: count per 128-row unit → scan of the unit counts → write the output columns in place.
    void addInputDense(const LookupSourceDevice& ls, const std::vector<DeviceColumn>& cols, int slot, int64_t n)
    {
        auto colOf = [&](int ch) -> const DeviceColumn& { return cols[std::find(staged.begin(), staged.end(), ch) - staged.begin()]; };
        const DeviceColumn& kcol = colOf(layout.channels[0]);
        DenseProbeParams P{};
        P.key = KeyColumn{kcol.values, kcol.nulls, kcol.type};
        P.D = ls.denseView(); P.n = n; P.units = divUp(n, kUnitRows); P.outer = joinType == PG_JOIN_PROBE_OUTER;
        P.rows = ls.payloadRows.as<uint64_t>();
        const int grid = (int) std::min<int64_t>(divUp(P.units, 8), (int64_t) numSMs * 8);
        int64_t total = n;
        if (!P.outer) {
            counts.reserve((size_t) (P.units + 1) * 4); offsets.reserve((size_t) (P.units + 1) * 4);
            PG_CUDA(cudaMemsetAsync(counts.as<uint32_t>() + P.units, 0, 4, stream));
            P.unitCounts = counts.as<uint32_t>();
            unitMask.reserve((size_t) P.units * kUnitRounds * 4);
            P.unitMask = unitMask.as<uint32_t>();
            PG_LAUNCH(this, denseCountKernel, grid, 256, 0, P);
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.as<uint32_t>(), offsets.as<uint32_t>(), (int) (P.units + 1), stream);
            scanTemp.reserve(tb);
            cub::DeviceScan::ExclusiveSum(scanTemp.ptr, tb, counts.as<uint32_t>(), offsets.as<uint32_t>(), (int) (P.units + 1), stream);
            stats.kernel_launches += 2;
            uint32_t t = 0;
            PG_CUDA(cudaMemcpyAsync(&t, offsets.as<uint32_t>() + P.units, 4, cudaMemcpyDeviceToHost, stream));
            PG_CUDA(cudaStreamSynchronize(stream));
            stats.d2h_bytes += 4;
            total = t;
            P.unitOffsets = offsets.as<uint32_t>();
        }
        pendingRows = total;
        if (total == 0) { stager.markCompute(slot); return; }
        const size_t nOut = outputChannels.size() + ls.outputChannels.size();
        ensureOutputSlots(nOut);
        outBlocks.assign(nOut, pg_block{});
        std::vector<bool> hasNulls(nOut, false);
        std::vector<VarcharOut> vouts;
        for (size_t i = 0; i < outputChannels.size(); i++) {
            const DeviceColumn& c = colOf(outputChannels[i]);
            if (c.type == PG_TYPE_VARCHAR) { vouts.push_back({i, false, outputChannels[i]}); outBlocks[i].type = c.type; hasNulls[i] = c.nulls != nullptr; continue; }
            const int w = typeWidth(c.type);
            outValues[i].reserve((size_t) total * w);
            CopyColumn cc{c.values, outValues[i].ptr};
            if (w == 8) P.probe8[P.nProbe8++] = cc; else if (w == 4) P.probe4[P.nProbe4++] = cc; else P.probe1[P.nProbe1++] = cc;
            if (c.nulls) { outNulls[i].reserve(total); P.probeNulls[P.nProbeNulls++] = NullCopy{c.nulls, outNulls[i].as<uint8_t>()}; hasNulls[i] = true; }
            outBlocks[i].type = c.type;
        }
        for (size_t i = 0; i < ls.outputChannels.size(); i++) {
            const int ch = ls.outputChannels[i];
            const size_t o = outputChannels.size() + i;
            const PayloadColumn& pc = ls.payload.cols[i];
            outValues[o].reserve((size_t) total * pc.width);
            BuildField f{};
            f.dst = outValues[o].ptr; f.dstNulls = nullptr;
            f.word = pc.offset >> 3; f.shift = 8 * (pc.offset & 7);
            f.nullWord = pc.nullOffset >= 0 ? pc.nullOffset >> 3 : -1; f.nullShift = pc.nullOffset >= 0 ? 8 * (pc.nullOffset & 7) : 0;
            if (pc.nullOffset >= 0 || P.outer) { outNulls[o].reserve(total); f.dstNulls = outNulls[o].as<uint8_t>(); hasNulls[o] = true; }
            if (pc.width == 8) P.build8[P.nBuild8++] = f; else if (pc.width == 4) P.build4[P.nBuild4++] = f; else P.build1[P.nBuild1++] = f;
            outBlocks[o].type = ls.types[ch];
        }
        if (!vouts.empty()) {      // the source row of every output row, as one more 4-byte probe column
            PG_CHECK_ARG(P.nProbe4 < kMaxJoinColumns, "too many output channels");
            iotaProbe.reserve((size_t) n * 4); srcRow.reserve((size_t) total * 4);
            PG_LAUNCH(this, iotaU32Kernel, grid, 256, 0, iotaProbe.as<uint32_t>(), n);
            P.probe4[P.nProbe4++] = CopyColumn{iotaProbe.ptr, srcRow.ptr};
        }
        switch (ls.payload.strideWords) {
            case 1: PG_LAUNCH(this, denseEmitKernel<1>, grid, 256, 0, P); break;
            case 2: PG_LAUNCH(this, denseEmitKernel<2>, grid, 256, 0, P); break;
            default: PG_LAUNCH(this, denseEmitKernel<4>, grid, 256, 0, P); break;
        }
        for (const VarcharOut& v : vouts) {
            const DeviceColumn& c = colOf(v.channel);
            PG_CHECK_ARG(c.offsets != nullptr, "VARCHAR block without offsets");
            outNulls[v.slot].reserve(total);
            varBytes[v.slot] = gatherVarcharByIndex(this, (const uint8_t*) c.values, c.offsets, c.nulls, srcRow.as<uint32_t>(), nullptr, total, outValues[v.slot], outOffsets[v.slot],
                                                    hasNulls[v.slot] ? outNulls[v.slot].as<uint8_t>() : nullptr, varScan);
        }
        stager.markCompute(slot);
        publish(total, hasNulls);
    }

    // describe (and, for host output, bring back) the `total` output rows sitting in outValues / outNulls
    void publish(int64_t total, const std::vector<bool>& hasNulls)
    {
        const size_t nOut = outBlocks.size();
        const bool toHost = options.output_memory == PG_MEM_HOST;
        if (toHost && hostValues.size() < nOut) { hostValues.resize(nOut); hostNulls.resize(nOut); hostOffsets.resize(nOut); }
        for (size_t i = 0; i < nOut; i++) {
            pg_block& b = outBlocks[i];
            const bool varchar = b.type == PG_TYPE_VARCHAR;
            const size_t vbytes = varchar ? (size_t) varBytes[i] : (size_t) total * typeWidth(b.type);
            b.encoding = PG_ENC_FLAT; b.position_count = total; b.offsets = varchar ? outOffsets[i].as<int32_t>() : nullptr; b.ids = nullptr; b.dictionary_size = 0;
            if (toHost) {
                hostValues[i].reserve(std::max<size_t>(vbytes, 16));
                if (vbytes) PG_CUDA(cudaMemcpyAsync(hostValues[i].ptr, outValues[i].ptr, vbytes, cudaMemcpyDeviceToHost, stream));
                if (varchar) {
                    hostOffsets[i].reserve((size_t) (total + 1) * 4);
                    PG_CUDA(cudaMemcpyAsync(hostOffsets[i].ptr, outOffsets[i].ptr, (size_t) (total + 1) * 4, cudaMemcpyDeviceToHost, stream));
                    b.offsets = hostOffsets[i].as<int32_t>();
                    stats.d2h_bytes += (total + 1) * 4;
                }
                b.values = hostValues[i].ptr; b.is_null = nullptr;
                stats.d2h_bytes += (int64_t) vbytes;
                if (hasNulls[i]) { hostNulls[i].reserve(total); PG_CUDA(cudaMemcpyAsync(hostNulls[i].ptr, outNulls[i].ptr, total, cudaMemcpyDeviceToHost, stream)); b.is_null = hostNulls[i].as<uint8_t>(); stats.d2h_bytes += total; }
            }
            else { b.values = outValues[i].ptr; b.is_null = hasNulls[i] ? outNulls[i].as<uint8_t>() : nullptr; }
            stats.output_bytes += (int64_t) vbytes;
        }
        PG_CUDA(cudaStreamSynchronize(stream));
        hasPending = true;
        stats.output_positions += total;
    }

    bool getOutput(pg_page* out) override
    {
        if (!hasPending) return false;
        out->position_count = pendingRows; out->channel_count = (int32_t) outBlocks.size(); out->memory = options.output_memory; out->blocks = outBlocks.data();
        hasPending = false;
        return true;
    }
    void finish() override
    {
        finishing = true;
        if (trackedByBridge) { trackedByBridge = false; bridge->probeClosed(); }     // this probe will visit no more positions
    }
    bool isFinished() override { return finishing && !hasPending; }
};

// ======================================================================================================================
// LookupOuterOperator (MB/operator/LookupOuterOperator.java:109-225): the source operator of RIGHT / FULL OUTER joins.  Once every probe
// operator is done it emits the build positions none of them visited — ascending position (SharedLookupOuterPositionIterator.appendToNext
// :139-151) — as rows of NULL probe channels + the build output channels.
// ======================================================================================================================
__global__ void outerUnvisitedKernel(const uint8_t* visited, int64_t n, uint32_t* flags)
{
    for (int64_t p = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; p <= n; p += (int64_t) gridDim.x * blockDim.x) flags[p] = (p < n && !visited[p]) ? 1u : 0u;
}
struct OuterEmitParams {
    const uint8_t* visited; const uint32_t* offsets; int64_t n;
    uint32_t* srcPos;          // when set: the build position of every output row (VARCHAR build channels are gathered through it)
    int32_t nProbe, nBuild;
    JoinColumn probe[kMaxJoinColumns];      // dst only: all NULL
    JoinColumn build[kMaxJoinColumns];
};
__global__ void outerEmitKernel(const __grid_constant__ OuterEmitParams P)
{
    for (int64_t p = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; p < P.n; p += (int64_t) gridDim.x * blockDim.x) {
        if (P.visited[p]) continue;
        const int64_t o = P.offsets[p];
        if (P.srcPos) P.srcPos[o] = (uint32_t) p;
        for (int c = 0; c < P.nProbe; c++) writeNull(P.probe[c], o);
        for (int c = 0; c < P.nBuild; c++) copyValue(P.build[c], p, o);
    }
}

class LookupOuterOp : public Operator {
    JoinBridge* bridge;
    std::vector<int> probeTypes;
    bool built = false, closed = false;
    int64_t total = 0, next = 0;
    std::vector<DeviceBuffer> outValues, outNulls;
    std::vector<PinnedBuffer> hostValues, hostNulls;
    std::vector<pg_block> outBlocks, slice;
    DeviceBuffer flags, offsets, scanTemp, srcPos, varScan;
    std::vector<DeviceBuffer> outOffsets; std::vector<PinnedBuffer> hostOffsets; std::vector<int64_t> varBytes;

public:
    LookupOuterOp(const pg_lookup_outer_config* c, JoinBridge* b) : Operator(c->options), bridge(b)
    {
        PG_CHECK_ARG(b != nullptr, "join bridge is null");
        PG_CHECK_ARG(b->isOuter(), "the bridge was not created for outer tracking (pg_join_bridge_set_outer)");
        PG_CHECK_ARG(c->n_probe_output_channels >= 0 && c->n_probe_output_channels <= kMaxJoinColumns, "too many probe output channels");
        probeTypes.assign(c->probe_output_types, c->probe_output_types + c->n_probe_output_channels);
    }
    bool needsInput() override { return false; }
    void addInput(const pg_page*) override { throw Error(PG_ERR_UNSUPPORTED, "LookupOuterOperator is a source operator"); }    // :171-175
    bool isBlocked() override { return !closed && !bridge->outerPositionsReady(); }     // the outerPositionsFuture :149-153
    void finish() override { closed = true; }                                            // :155-160: a source operator just terminates
    bool isFinished() override { return closed; }

    void build()
    {
        built = true;
        auto ls = bridge->get();
        const int64_t n = ls->positionCount;
        if (n == 0) return;
        PG_CHECK_ARG(n < (int64_t(1) << 31), "build side exceeds 2^31 positions");
        flags.reserve((size_t) (n + 1) * 4); offsets.reserve((size_t) (n + 1) * 4);
        const int grid = (int) std::min<int64_t>(divUp(n + 1, 256), numSMs * 16);
        PG_LAUNCH(this, outerUnvisitedKernel, grid, 256, 0, ls->visited.as<uint8_t>(), n, flags.as<uint32_t>());
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, flags.as<uint32_t>(), offsets.as<uint32_t>(), (int) (n + 1), stream);
        scanTemp.reserve(tb);
        cub::DeviceScan::ExclusiveSum(scanTemp.ptr, tb, flags.as<uint32_t>(), offsets.as<uint32_t>(), (int) (n + 1), stream);
        stats.kernel_launches += 2;
        uint32_t t = 0;
        PG_CUDA(cudaMemcpyAsync(&t, offsets.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, stream));
        PG_CUDA(cudaStreamSynchronize(stream));
        stats.d2h_bytes += 4;
        total = t;
        if (total == 0) return;
        const size_t nOut = probeTypes.size() + ls->outputChannels.size();
        outValues.resize(nOut); outNulls.resize(nOut); outOffsets.resize(nOut); varBytes.assign(nOut, 0);
        for (auto* v : {&outValues, &outNulls, &outOffsets}) for (auto& d : *v) d.accounting = &stats.device_bytes;
        outBlocks.assign(nOut, pg_block{});
        OuterEmitParams E{};
        E.visited = ls->visited.as<uint8_t>(); E.offsets = offsets.as<uint32_t>(); E.n = n;
        std::vector<bool> hasNulls(nOut, false);
        std::vector<std::pair<size_t, int>> varcharBuild;      // (slot, build channel)
        int np = 0, nb = 0;
        for (size_t i = 0; i < probeTypes.size(); i++) {
            outBlocks[i].type = probeTypes[i]; hasNulls[i] = true;
            outNulls[i].reserve(total);
            if (probeTypes[i] == PG_TYPE_VARCHAR) {     // an all-NULL VARCHAR column: every slice empty
                outOffsets[i].reserve((size_t) (total + 1) * 4); outValues[i].reserve(16);
                PG_CUDA(cudaMemsetAsync(outOffsets[i].ptr, 0, (size_t) (total + 1) * 4, stream));
                PG_CUDA(cudaMemsetAsync(outNulls[i].ptr, 1, (size_t) total, stream));
                continue;
            }
            JoinColumn& jc = E.probe[np++];
            jc.width = typeWidth(probeTypes[i]); jc.src = nullptr; jc.srcNulls = nullptr;
            outValues[i].reserve((size_t) total * jc.width);
            jc.dst = outValues[i].ptr; jc.dstNulls = outNulls[i].as<uint8_t>();
        }
        for (size_t i = 0; i < ls->outputChannels.size(); i++) {
            const int ch = ls->outputChannels[i];
            const size_t o = probeTypes.size() + i;
            outBlocks[o].type = ls->types[ch];
            if (ls->types[ch] == PG_TYPE_VARCHAR) { varcharBuild.push_back({o, ch}); hasNulls[o] = true; continue; }
            JoinColumn& jc = E.build[nb++];
            jc.width = typeWidth(ls->types[ch]); jc.src = ls->values[ch].ptr; jc.srcNulls = ls->hasNulls[ch] ? ls->nulls[ch].as<uint8_t>() : nullptr;
            outValues[o].reserve((size_t) total * jc.width);
            jc.dst = outValues[o].ptr; jc.dstNulls = nullptr;
            if (jc.srcNulls) { outNulls[o].reserve(total); jc.dstNulls = outNulls[o].as<uint8_t>(); hasNulls[o] = true; }
        }
        E.nProbe = np; E.nBuild = nb;
        if (!varcharBuild.empty()) { srcPos.reserve((size_t) total * 4); E.srcPos = srcPos.as<uint32_t>(); }
        PG_LAUNCH(this, outerEmitKernel, grid, 256, 0, E);
        for (auto& vb : varcharBuild) {
            outNulls[vb.first].reserve(total);
            varBytes[vb.first] = gatherVarcharByIndex(this, ls->values[vb.second].as<uint8_t>(), ls->offsets[vb.second].as<int32_t>(), ls->nulls[vb.second].as<uint8_t>(), srcPos.as<uint32_t>(), nullptr,
                                                      total, outValues[vb.first], outOffsets[vb.first], outNulls[vb.first].as<uint8_t>(), varScan);
        }
        const bool toHost = options.output_memory == PG_MEM_HOST;
        if (toHost) { hostValues.resize(nOut); hostNulls.resize(nOut); hostOffsets.resize(nOut); }
        for (size_t i = 0; i < nOut; i++) {
            pg_block& b = outBlocks[i];
            const bool varchar = b.type == PG_TYPE_VARCHAR;
            const size_t vbytes = varchar ? (size_t) varBytes[i] : (size_t) total * typeWidth(b.type);
            b.encoding = PG_ENC_FLAT; b.position_count = total; b.offsets = varchar ? outOffsets[i].as<int32_t>() : nullptr; b.ids = nullptr; b.dictionary_size = 0;
            if (toHost) {
                hostValues[i].reserve(std::max<size_t>(vbytes, 16));
                if (vbytes) PG_CUDA(cudaMemcpyAsync(hostValues[i].ptr, outValues[i].ptr, vbytes, cudaMemcpyDeviceToHost, stream));
                if (varchar) {
                    hostOffsets[i].reserve((size_t) (total + 1) * 4);
                    PG_CUDA(cudaMemcpyAsync(hostOffsets[i].ptr, outOffsets[i].ptr, (size_t) (total + 1) * 4, cudaMemcpyDeviceToHost, stream));
                    b.offsets = hostOffsets[i].as<int32_t>();
                }
                b.values = hostValues[i].ptr; b.is_null = nullptr;
                stats.d2h_bytes += (int64_t) vbytes;
                if (hasNulls[i]) { hostNulls[i].reserve(total); PG_CUDA(cudaMemcpyAsync(hostNulls[i].ptr, outNulls[i].ptr, total, cudaMemcpyDeviceToHost, stream)); b.is_null = hostNulls[i].as<uint8_t>(); stats.d2h_bytes += total; }
            }
            else { b.values = outValues[i].ptr; b.is_null = hasNulls[i] ? outNulls[i].as<uint8_t>() : nullptr; }
            stats.output_bytes += (int64_t) vbytes;
        }
        PG_CUDA(cudaStreamSynchronize(stream));
    }

    bool getOutput(pg_page* out) override      // :183-214
    {
        if (closed || !bridge->outerPositionsReady()) return false;       // tryGetFutureValue(outerPositionsFuture) == null -> return null
        if (!built) build();
        if (next >= total) { closed = true; return false; }
        const int64_t m = std::min<int64_t>(maxOutputRows(), total - next);
        slice = outBlocks;
        for (auto& b : slice) {
            b.position_count = m;
            if (b.type == PG_TYPE_VARCHAR) b.offsets += next;      // offsets stay absolute into the shared byte slice
            else b.values = (const char*) b.values + next * typeWidth(b.type);
            if (b.is_null) b.is_null += next;
        }
        out->position_count = m; out->channel_count = (int32_t) slice.size(); out->memory = options.output_memory; out->blocks = slice.data();
        next += m;
        stats.output_positions += m;
        if (next >= total) closed = true;
        return true;
    }
};

Operator* createLookupJoin(const pg_lookup_join_config* c, JoinBridge* bridge) { return new LookupJoinOp(c, bridge); }

void describeLookupSource(JoinBridge* bridge, LookupSourceView& out)
{
    out = LookupSourceView();
    auto ls = bridge->get();
    if (!ls) return;
    out.ready = true;
    out.nKeys = ls->layout.nKeys;
    out.keyType = ls->layout.types[0];
    out.dense = ls->dense;
    if (!ls->dense) return;
    out.D = ls->denseView();
    out.rows = ls->payloadRows.as<uint64_t>();
    out.strideWords = ls->outputChannels.empty() ? 0 : ls->payload.strideWords;
    for (size_t i = 0; i < ls->outputChannels.size(); i++) {
        const PayloadColumn& pc = ls->payload.cols[i];
        FusedBuildColumn b;
        b.type = ls->types[ls->outputChannels[i]]; b.width = pc.width;
        b.word = pc.offset >> 3; b.shift = 8 * (pc.offset & 7);
        b.nullWord = pc.nullOffset >= 0 ? pc.nullOffset >> 3 : -1; b.nullShift = pc.nullOffset >= 0 ? 8 * (pc.nullOffset & 7) : 0;
        out.buildOutputs.push_back(b);
    }
}

// ---- HashSemiJoinOperator (MB/operator/HashSemiJoinOperator.java:142-192) -------------------------------------------------------
// flag = contains(set, key) with SQL's three values: NULL key → NULL (false for a set without rows); absent → NULL if the set holds a NULL
__device__ __forceinline__ void semiResult(bool keyIsNull, bool contains, int setEmpty, int setHasNull, int64_t row, uint8_t* val, uint8_t* isNull)
{
    bool n, v;
    if (keyIsNull) { n = !setEmpty; v = false; }
    else if (!contains && setHasNull) { n = true; v = false; }
    else { n = false; v = contains; }
    val[row] = v ? 1 : 0; isNull[row] = n ? 1 : 0;
}
__global__ void semiJoinDenseKernel(KeyColumn key, DenseIndex D, int64_t n, int setEmpty, int setHasNull, uint8_t* val, uint8_t* isNull)
{
    for (int64_t row = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; row < n; row += (int64_t) gridDim.x * blockDim.x) {
        int64_t k;
        const bool have = loadKey(key, row, k);
        semiResult(!have, have && denseContains(D, k), setEmpty, setHasNull, row, val, isNull);
    }
}
template <int W>
__global__ void semiJoinHashKernel(const __grid_constant__ ProbeParams P, int setEmpty, int setHasNull, uint8_t* val, uint8_t* isNull)
{
    for (int64_t row = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; row < P.n; row += (int64_t) gridDim.x * blockDim.x) {
        uint64_t kw[W];
        const int special = packKey<W>(P.kp, row, kw);
        const bool keyIsNull = keyHasNull(P.kp, kw, special);
        bool contains = false;
        if (!keyIsNull && !setEmpty) {
            const int slot = special == 2 ? (int) P.J.t.cap + 1 : tableFind<W>(P.J.t, kw, hashKeyWords<W>(kw), false, true);
            contains = slot >= 0 && P.J.head[slot] >= 0;
        }
        semiResult(keyIsNull, contains, setEmpty, setHasNull, row, val, isNull);
    }
}

class SemiJoinOp : public Operator {
    JoinBridge* bridge;
    std::vector<int> types, staged;
    int probeChannel;
    KeyLayout layout;
    PageStager stager;
    bool finishing = false, hasPending = false;
    std::vector<DeviceBuffer> outValues, outNulls;
    std::vector<PinnedBuffer> hostValues, hostNulls;
    std::vector<pg_block> outBlocks;
    int64_t pendingRows = 0;
    // VARCHAR channels: the probe key as value ids of the set's interner; pass-through slices copied by an identity gather
    DeviceBuffer probeKeyIds, iota, varScan;
    std::vector<DeviceBuffer> outOffsets; std::vector<PinnedBuffer> hostOffsets; std::vector<int64_t> varBytes;

public:
    SemiJoinOp(const pg_semi_join_config* c, JoinBridge* b) : Operator(c->options), bridge(b), stager(this)
    {
        PG_CHECK_ARG(b != nullptr, "set bridge is null");
        PG_CHECK_ARG(c->n_channels >= 1 && c->n_channels < kMaxJoinColumns, "too many channels");
        types.assign(c->types, c->types + c->n_channels);
        probeChannel = c->probe_join_channel;
        PG_CHECK_ARG(probeChannel >= 0 && probeChannel < c->n_channels, "probe join channel out of range");
        for (int t : types) if (t == PG_TYPE_LONG_DECIMAL) throw Error(PG_ERR_UNSUPPORTED, "Int128 channels are not on the GPU semi-join path");
        int32_t kt = types[probeChannel];
        layout.plan(1, &probeChannel, &kt, nullptr);
        checkJoinKeyTypes(layout);
        for (int i = 0; i < c->n_channels; i++) staged.push_back(i);
        const size_t nOut = types.size() + 1;
        outValues.resize(nOut); outNulls.resize(nOut); outOffsets.resize(nOut); varBytes.assign(nOut, 0);
        for (auto& d : outValues) d.accounting = &stats.device_bytes;
        for (auto& d : outNulls) d.accounting = &stats.device_bytes;
        for (auto& d : outOffsets) d.accounting = &stats.device_bytes;
        for (DeviceBuffer* d : {&probeKeyIds, &iota, &varScan}) d->accounting = &stats.device_bytes;
    }
    bool needsInput() override { return !finishing && !hasPending && bridge->get() != nullptr; }   // "Set has not been built yet"
    bool isBlocked() override { return bridge->get() == nullptr; }

    void addInput(const pg_page* page) override
    {
        PG_CHECK_STATE(needsInput(), "semi join does not need input (set not built, finishing, or output pending)");
        auto ls = bridge->get();
        PG_CHECK_ARG(ls->layout.nKeys == 1 && ls->layout.types[0] == layout.types[0], "probe and set key types differ");
        const int64_t n = page->position_count;
        stats.input_positions += n;
        if (n == 0) return;
        PG_CHECK_ARG(n < (int64_t(1) << 31), "probe page exceeds 2^31 positions");
        PG_CHECK_ARG(page->channel_count == (int) types.size(), "page channel count differs from the operator's configuration");
        std::vector<DeviceColumn> cols;
        const int slot = stager.stage(page, staged, cols);
        const size_t flag = types.size();
        outValues[flag].reserve((size_t) n); outNulls[flag].reserve((size_t) n);
        const int setEmpty = ls->positionCount == 0, setHasNull = ls->nullKeyCount > 0;
        const int grid = (int) std::min<int64_t>(divUp(n, 256), (int64_t) numSMs * 16);
        const DeviceColumn& kc = cols[(size_t) probeChannel];
        PG_CHECK_ARG(kc.type == layout.types[0], "probe key channel type differs from configuration");
        if (ls->dense)
            PG_LAUNCH(this, semiJoinDenseKernel, grid, 256, 0, KeyColumn{kc.values, kc.nulls, kc.type}, ls->denseView(), n, setEmpty, setHasNull, outValues[flag].as<uint8_t>(), outNulls[flag].as<uint8_t>());
        else {
            ProbeParams P{};
            P.J = ls->view();
            std::vector<DeviceColumn> keyCols{kc};
            if (layout.interned[0]) {      // VARCHAR: the id the set gave this slice, -1 (no member) when the set never saw it
                probeKeyIds.reserve((size_t) n * 4 + 16);
                if (!ls->interners.empty() && ls->interners[0]) ls->interners[0]->intern(this, kc, n, probeKeyIds.as<int32_t>(), false);
                else PG_CUDA(cudaMemsetAsync(probeKeyIds.ptr, 0xFF, (size_t) n * 4, stream));
                keyCols[0].type = PG_TYPE_INTEGER; keyCols[0].values = probeKeyIds.ptr; keyCols[0].offsets = nullptr;
            }
            layout.fill(P.kp, keyCols, ls->counters.as<unsigned int>() + CNT_ERROR);
            P.n = n;
            switch (layout.W) {
                case 1: PG_LAUNCH(this, semiJoinHashKernel<1>, grid, 256, 0, P, setEmpty, setHasNull, outValues[flag].as<uint8_t>(), outNulls[flag].as<uint8_t>()); break;
                default: PG_LAUNCH(this, semiJoinHashKernel<2>, grid, 256, 0, P, setEmpty, setHasNull, outValues[flag].as<uint8_t>(), outNulls[flag].as<uint8_t>()); break;
            }
        }
        // the input channels travel with the page (Page.appendColumn in the reference; here: device copies owned by the operator)
        outBlocks.assign(flag + 1, pg_block{});
        std::vector<bool> hasNulls(flag + 1, false);
        for (size_t i = 0; i < flag; i++) {
            outBlocks[i].type = types[i];
            if (types[i] == PG_TYPE_VARCHAR) {      // the slices of every row, in order: an identity gather (rebased offsets, own bytes)
                PG_CHECK_ARG(cols[i].offsets != nullptr, "VARCHAR block without offsets");
                if (iota.capacity < (size_t) n * 4) { iota.reserve((size_t) n * 4); }
                PG_LAUNCH(this, iotaU32Kernel, grid, 256, 0, iota.as<uint32_t>(), n);
                outNulls[i].reserve((size_t) n);
                hasNulls[i] = cols[i].nulls != nullptr;
                varBytes[i] = gatherVarcharByIndex(this, (const uint8_t*) cols[i].values, cols[i].offsets, cols[i].nulls, iota.as<uint32_t>(), nullptr, n, outValues[i], outOffsets[i],
                                                   hasNulls[i] ? outNulls[i].as<uint8_t>() : nullptr, varScan);
                continue;
            }
            const int w = typeWidth(types[i]);
            outValues[i].reserve((size_t) n * w);
            PG_CUDA(cudaMemcpyAsync(outValues[i].ptr, cols[i].values, (size_t) n * w, cudaMemcpyDeviceToDevice, stream));
            if (cols[i].nulls) { outNulls[i].reserve((size_t) n); PG_CUDA(cudaMemcpyAsync(outNulls[i].ptr, cols[i].nulls, (size_t) n, cudaMemcpyDeviceToDevice, stream)); hasNulls[i] = true; }
        }
        outBlocks[flag].type = PG_TYPE_BOOLEAN; hasNulls[flag] = true;
        stager.markCompute(slot);
        const bool toHost = options.output_memory == PG_MEM_HOST;
        if (toHost && hostValues.size() < flag + 1) { hostValues.resize(flag + 1); hostNulls.resize(flag + 1); hostOffsets.resize(flag + 1); }
        for (size_t i = 0; i <= flag; i++) {
            pg_block& b = outBlocks[i];
            const bool varchar = b.type == PG_TYPE_VARCHAR;
            const size_t vbytes = varchar ? (size_t) varBytes[i] : (size_t) n * typeWidth(b.type);
            b.encoding = PG_ENC_FLAT; b.position_count = n; b.offsets = varchar ? outOffsets[i].as<int32_t>() : nullptr; b.ids = nullptr; b.dictionary_size = 0;
            if (toHost) {
                hostValues[i].reserve(std::max<size_t>(vbytes, 16));
                if (vbytes) PG_CUDA(cudaMemcpyAsync(hostValues[i].ptr, outValues[i].ptr, vbytes, cudaMemcpyDeviceToHost, stream));
                if (varchar) {
                    hostOffsets[i].reserve((size_t) (n + 1) * 4);
                    PG_CUDA(cudaMemcpyAsync(hostOffsets[i].ptr, outOffsets[i].ptr, (size_t) (n + 1) * 4, cudaMemcpyDeviceToHost, stream));
                    b.offsets = hostOffsets[i].as<int32_t>();
                    stats.d2h_bytes += (n + 1) * 4;
                }
                b.values = hostValues[i].ptr; b.is_null = nullptr;
                stats.d2h_bytes += (int64_t) vbytes;
                if (hasNulls[i]) { hostNulls[i].reserve((size_t) n); PG_CUDA(cudaMemcpyAsync(hostNulls[i].ptr, outNulls[i].ptr, (size_t) n, cudaMemcpyDeviceToHost, stream)); b.is_null = hostNulls[i].as<uint8_t>(); stats.d2h_bytes += n; }
            }
            else { b.values = outValues[i].ptr; b.is_null = hasNulls[i] ? outNulls[i].as<uint8_t>() : nullptr; }
            stats.output_bytes += (int64_t) vbytes;
        }
        PG_CUDA(cudaStreamSynchronize(stream));
        pendingRows = n; hasPending = true;
        stats.output_positions += n;
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
};

}  // namespace pg

extern "C" {
#define J_TRY try {
#define J_CATCH } catch (const pg::Error& e) { pg::setLastError(e.what()); return e.code; } catch (const std::exception& e) { pg::setLastError(e.what()); return PG_ERR_INVALID_STATE; } return PG_OK;

int32_t pg_join_bridge_create(int32_t device, pg_join_bridge** out)
{
    J_TRY
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) throw pg::Error(PG_ERR_NO_DEVICE, "no CUDA device visible: presto_b200 operators only run on the GPU (no CPU fallback)");
    auto* b = new pg::JoinBridge(); b->device = device; *out = (pg_join_bridge*) b;
    J_CATCH
}
int32_t pg_join_bridge_destroy(pg_join_bridge* b) { J_TRY auto* jb = (pg::JoinBridge*) b; if (jb) { cudaSetDevice(jb->device); delete jb; } J_CATCH }
int32_t pg_join_bridge_get_position_count(pg_join_bridge* b, int64_t* out) { J_TRY auto ls = ((pg::JoinBridge*) b)->get(); *out = ls ? ls->positionCount : -1; J_CATCH }
int32_t pg_hash_builder_create(const pg_hash_builder_config* c, pg_join_bridge* bridge, pg_operator** out) { J_TRY *out = (pg_operator*) new pg::HashBuilderOp(c, (pg::JoinBridge*) bridge); J_CATCH }
int32_t pg_lookup_join_create(const pg_lookup_join_config* c, pg_join_bridge* bridge, pg_operator** out) { J_TRY *out = (pg_operator*) new pg::LookupJoinOp(c, (pg::JoinBridge*) bridge); J_CATCH }
int32_t pg_join_bridge_set_outer(pg_join_bridge* b) { J_TRY PG_CHECK_ARG(b != nullptr, "null argument"); auto* jb = (pg::JoinBridge*) b; PG_CHECK_STATE(jb->get() == nullptr, "the build side has already finished"); std::lock_guard<std::mutex> l(jb->mu); jb->outer = true; J_CATCH }
int32_t pg_join_bridge_no_more_probe_operators(pg_join_bridge* b) { J_TRY PG_CHECK_ARG(b != nullptr, "null argument"); auto* jb = (pg::JoinBridge*) b; std::lock_guard<std::mutex> l(jb->mu); jb->noMoreProbes = true; J_CATCH }
int32_t pg_lookup_outer_create(const pg_lookup_outer_config* c, pg_join_bridge* bridge, pg_operator** out) { J_TRY PG_CHECK_ARG(c && out, "null argument"); *out = (pg_operator*) new pg::LookupOuterOp(c, (pg::JoinBridge*) bridge); J_CATCH }
int32_t pg_hash_semi_join_create(const pg_semi_join_config* c, pg_join_bridge* set, pg_operator** out) { J_TRY *out = (pg_operator*) new pg::SemiJoinOp(c, (pg::JoinBridge*) set); J_CATCH }
}
