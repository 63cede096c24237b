// presto_b200 — LZ4 block codec for SerializedPage payloads (PagesSerde.java:118-143 compresses the raw page with the PageCompressor
// PagesSerdeFactory.java:75 creates: io.airlift.compress.lz4.Lz4Compressor, i.e. the LZ4 *block* format, no frame; :96 the decompressor).
//
// aircompressor 0.27 is a third-party dependency that is not under /root/reference.  The block FORMAT is fixed (token = literal length <<
// 4 | match length - 4, 255-extension bytes, literals, 2-byte little-endian offset, …; the last sequence is literals only; the last match
// starts at least 12 bytes before the end and the last 5 bytes are literals), so decoding is fully determined and any conforming encoder's
// output is read by Lz4Decompressor.  WHICH matches an encoder finds is its own business: this encoder's bytes differ from aircompressor's
// (parity unpinned for compressed bytes; pinned: what every stream decodes to, and against liblz4 through pyarrow in the tests).
//
// Parallel design: a payload is cut into segments of kSegmentBytes.  One thread finds the matches of one segment with a private hash table
// (matches never leave the segment, so segments are independent); a per-block pass turns the segments' sizes into output offsets — a
// segment's first literal run starts where the last match of the preceding segments ended, which is what makes the concatenation ONE valid
// block —; one thread per segment then writes its sequences.  Every function is __host__ __device__: the CPU tests run the very code the
// kernels run (pg_lz4_*_host in runtime.cu).
#pragma once
#include "common.cuh"

namespace pg {
namespace lz4 {

constexpr int kSegmentBytes = 8192;        // positions inside a segment fit the 16-bit table entries and match records
constexpr int kHashLog = 11;
constexpr int kHashEntries = 1 << kHashLog;
constexpr int kMinMatch = 4, kMfLimit = 12, kLastLiterals = 5;
constexpr int kMaxMatchesPerSegment = kSegmentBytes / kMinMatch;
constexpr uint16_t kNoEntry = 0xFFFF;

struct Match { uint16_t relPos, len, off; };      // start relative to the segment, length >= 4, distance >= 1 (inside the segment)
struct SegmentInfo {
    uint32_t nMatches;
    uint32_t firstPos;       // block-relative start of the first match
    uint32_t lastEnd;        // block-relative end of the last match
    uint32_t innerBytes;     // encoded bytes of the segment's sequences without the literal run (extension bytes + literals) of the first one
    // filled by planBlock:
    uint32_t anchor;         // where the first literal run starts (end of the last match before this segment)
    uint32_t outAt;          // offset of the segment's first byte in the compressed block
};
struct BlockInfo { uint32_t compressedBytes, tailAnchor, tailAt, pad; };

__host__ __device__ __forceinline__ uint32_t read32(const uint8_t* p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }
__host__ __device__ __forceinline__ uint32_t hash4(uint32_t v) { return (v * 2654435761u) >> (32 - kHashLog); }
__host__ __device__ __forceinline__ uint32_t lengthExtra(uint32_t v) { return v >= 15 ? 1 + (v - 15) / 255 : 0; }
__host__ __device__ __forceinline__ int64_t compressBound(int64_t n) { return n + n / 255 + 16; }
__host__ __device__ __forceinline__ int64_t segmentsOf(int64_t n) { return n <= 0 ? 1 : (n + kSegmentBytes - 1) / kSegmentBytes; }

// Pass 1: the matches of segment `seg` of src[0, n).  `table`: kHashEntries entries of scratch private to the caller.
__host__ __device__ inline void findMatches(const uint8_t* src, int64_t n, int64_t seg, uint16_t* table, Match* out, SegmentInfo& info)
{
    const int64_t segBegin = seg * kSegmentBytes, segEnd = segBegin + kSegmentBytes < n ? segBegin + kSegmentBytes : n;
    for (int i = 0; i < kHashEntries; i++) table[i] = kNoEntry;
    const int64_t matchEndLimit = segEnd < n - kLastLiterals ? segEnd : n - kLastLiterals;     // matches end here at the latest
    int64_t ip = segBegin, anchor = segBegin;
    uint32_t nM = 0, inner = 0, firstPos = 0, lastEnd = 0, misses = 0;
    while (ip + kMfLimit <= n && ip + kMinMatch <= matchEndLimit) {
        const uint32_t v = read32(src + ip), h = hash4(v);
        const uint16_t cand = table[h];
        table[h] = (uint16_t) (ip - segBegin);
        if (cand != kNoEntry) {
            int64_t c = segBegin + cand;
            if (read32(src + c) == v) {
                int64_t len = kMinMatch;
                while (ip + len < matchEndLimit && src[c + len] == src[ip + len]) len++;
                while (ip > anchor && c > segBegin && src[ip - 1] == src[c - 1]) { ip--; c--; len++; }
                const uint32_t lit = (uint32_t) (ip - anchor);
                out[nM].relPos = (uint16_t) (ip - segBegin); out[nM].len = (uint16_t) len; out[nM].off = (uint16_t) (ip - c);
                if (nM == 0) { firstPos = (uint32_t) ip; inner += 1 + 2 + lengthExtra((uint32_t) len - kMinMatch); }
                else inner += 1 + lengthExtra(lit) + lit + 2 + lengthExtra((uint32_t) len - kMinMatch);
                ip += len; anchor = ip; lastEnd = (uint32_t) ip; nM++; misses = 0;
                continue;
            }
        }
        ip += 1 + (misses++ >> 5);       // incompressible stretches are skipped faster and faster (as LZ4's acceleration does)
    }
    info.nMatches = nM; info.firstPos = firstPos; info.lastEnd = lastEnd; info.innerBytes = inner;
}

// Pass 2 (one caller per block): where every segment's bytes go, and the final literal run.
__host__ __device__ inline void planBlock(int64_t n, SegmentInfo* segs, int64_t nSegs, BlockInfo& block)
{
    uint32_t anchor = 0, at = 0;
    for (int64_t s = 0; s < nSegs; s++) {
        SegmentInfo& g = segs[s];
        g.anchor = anchor; g.outAt = at;
        if (g.nMatches == 0) continue;
        const uint32_t lit0 = g.firstPos - anchor;
        at += lengthExtra(lit0) + lit0 + g.innerBytes;
        anchor = g.lastEnd;
    }
    const uint32_t tail = (uint32_t) n - anchor;
    block.tailAnchor = anchor; block.tailAt = at;
    block.compressedBytes = at + 1 + lengthExtra(tail) + tail;
    block.pad = 0;
}

__host__ __device__ __forceinline__ uint8_t* putLength(uint8_t* op, uint32_t v)       // the 255-extension of a length whose token nibble is 15
{
    uint32_t rem = v - 15;
    while (rem >= 255) { *op++ = 255; rem -= 255; }
    *op++ = (uint8_t) rem;
    return op;
}

// Pass 3: the sequences of segment `seg`; the thread of the last segment also writes the block's final literals.
__host__ __device__ inline void emitSegment(const uint8_t* src, int64_t n, int64_t seg, int64_t nSegs, const Match* matches, const SegmentInfo& g, const BlockInfo& block, uint8_t* dst)
{
    const int64_t segBegin = seg * kSegmentBytes;
    uint8_t* op = dst + g.outAt;
    int64_t anchor = g.anchor;
    for (uint32_t k = 0; k < g.nMatches; k++) {
        const int64_t pos = segBegin + matches[k].relPos;
        const uint32_t lit = (uint32_t) (pos - anchor), ml = (uint32_t) matches[k].len - kMinMatch;
        *op++ = (uint8_t) (((lit < 15 ? lit : 15) << 4) | (ml < 15 ? ml : 15));
        if (lit >= 15) op = putLength(op, lit);
        for (uint32_t i = 0; i < lit; i++) op[i] = src[anchor + i];
        op += lit;
        *op++ = (uint8_t) (matches[k].off & 0xff); *op++ = (uint8_t) (matches[k].off >> 8);
        if (ml >= 15) op = putLength(op, ml);
        anchor = pos + matches[k].len;
    }
    if (seg == nSegs - 1) {
        op = dst + block.tailAt;
        const uint32_t lit = (uint32_t) n - block.tailAnchor;
        *op++ = (uint8_t) ((lit < 15 ? lit : 15) << 4);
        if (lit >= 15) op = putLength(op, lit);
        for (uint32_t i = 0; i < lit; i++) op[i] = src[block.tailAnchor + i];
    }
}

// Decoder.  Every participating lane (lane, lanes) runs the same parse — the token and length bytes are read by all of them — and
// copies its share of every literal run and match; between the copies the lanes synchronise (`WARP`, device only) so that a match can read
// what the previous copies wrote.  A match that overlaps its own output (offset < length) repeats a period of `offset` bytes, all of which
// were written before the match began.  Returns the decoded size, or -1 for a malformed or overlong stream (nothing is written out of bounds).
template <bool WARP>
__host__ __device__ inline int64_t decodeBlock(const uint8_t* src, int64_t srcLen, uint8_t* dst, int64_t dstCap, int lane, int lanes)
{
    int64_t ip = 0, op = 0;
    while (ip < srcLen) {
        const uint32_t token = src[ip++];
        int64_t lit = token >> 4;
        if (lit == 15) { uint32_t b; do { if (ip >= srcLen) return -1; b = src[ip++]; lit += b; } while (b == 255); }
        if (ip + lit > srcLen || op + lit > dstCap) return -1;
        for (int64_t i = lane; i < lit; i += lanes) dst[op + i] = src[ip + i];
        ip += lit; op += lit;
        if (ip >= srcLen) break;                         // the last sequence has no match
        if (ip + 2 > srcLen) return -1;
        const int64_t off = (int64_t) src[ip] | ((int64_t) src[ip + 1] << 8);
        ip += 2;
        if (off == 0 || off > op) return -1;
        int64_t ml = token & 15;
        if (ml == 15) { uint32_t b; do { if (ip >= srcLen) return -1; b = src[ip++]; ml += b; } while (b == 255); }
        ml += kMinMatch;
        if (op + ml > dstCap) return -1;
#ifdef __CUDA_ARCH__
        if (WARP) __syncwarp();
#endif
        const uint8_t* from = dst + op - off;
        if (off >= ml) for (int64_t i = lane; i < ml; i += lanes) dst[op + i] = from[i];
        else for (int64_t i = lane; i < ml; i += lanes) dst[op + i] = from[i % off];
        op += ml;      // (the next match synchronises before it reads; literal runs read the compressed stream only)
    }
    return op;
}

}  // namespace lz4
}  // namespace pg
