// presto_b200 — device view of the dense unique-key join index (bitmap + rank directory + payload rows in key order) built by
// HashBuilderOperator (hash_join.cu) and probed by LookupJoinOperator and by the fused pipeline kernels (fused_agg.cuh).
#pragma once
#include "common.cuh"

namespace pg {

constexpr int kDenseBlockWords = 4;                  // 4 x 64 bit = one 32-byte sector per block

struct DenseIndex {
    const uint64_t* bitmap;    // [blocks * 4]
    const uint32_t* prefix;    // [blocks] set bits before the block
    int64_t minKey;
    uint64_t range;            // keys are in [minKey, minKey + range)
};

// rank of `key` among the build keys, or -1 when absent
__device__ __forceinline__ int64_t denseRank(const DenseIndex& D, int64_t key)
{
    const uint64_t idx = (uint64_t) (key - D.minKey);
    if (idx >= D.range) return -1;
    const uint64_t blk = idx >> 8;
    const ulonglong4 w = *reinterpret_cast<const ulonglong4*>(D.bitmap + blk * kDenseBlockWords);   // one 32-byte sector
    const uint32_t wi = (uint32_t) (idx >> 6) & 3u;
    const uint64_t word = wi == 0 ? w.x : wi == 1 ? w.y : wi == 2 ? w.z : w.w;
    const uint64_t bit = 1ULL << (idx & 63);
    if (!(word & bit)) return -1;
    uint32_t r = D.prefix[blk] + __popcll(word & (bit - 1));
    if (wi > 0) r += __popcll(w.x);
    if (wi > 1) r += __popcll(w.y);
    if (wi > 2) r += __popcll(w.z);
    return (int64_t) r;
}

// membership only (pass A of the probe): one 8-byte bitmap word, no directory
__device__ __forceinline__ bool denseContains(const DenseIndex& D, int64_t key)
{
    const uint64_t idx = (uint64_t) (key - D.minKey);
    if (idx >= D.range) return false;
    return (D.bitmap[idx >> 6] >> (idx & 63)) & 1ULL;
}

}  // namespace pg
