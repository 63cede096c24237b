// presto_b200 — host side of the chunked CRC-32 (java.util.zip.CRC32 = CRC-32/ISO-HDLC) used by the SerializedPage checksum
// (presto-spi/.../spi/page/PagesSerdeUtil.java:109-120): the device computes one CRC per 4 KB chunk, the host combines them with the
// operator "advance the register over len zero bytes" — a 32 x 32 matrix over GF(2), built by repeated squaring (the idea behind
// zlib's crc32_combine, restated).  Pure host code: tested on the CPU box against zlib (tests/test_abi.py).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>

namespace pg {

struct Crc32Host {
    uint32_t table[256];
    Crc32Host()
    {
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
    }
    uint32_t update(uint32_t crc, const uint8_t* p, size_t n) const { crc = ~crc; for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8); return ~crc; }
    static uint32_t times(const uint32_t* mat, uint32_t vec) { uint32_t s = 0; for (int i = 0; vec; vec >>= 1, i++) if (vec & 1) s ^= mat[i]; return s; }
    static void square(uint32_t* sq, const uint32_t* mat) { for (int i = 0; i < 32; i++) sq[i] = times(mat, mat[i]); }
    // matrix that advances a CRC register over `len` zero bytes (len > 0)
    static void zeroOperator(uint32_t* out, uint64_t len)
    {
        uint32_t even[32], odd[32];
        odd[0] = 0xEDB88320u;                       // one zero BIT
        for (int i = 1; i < 32; i++) odd[i] = 1u << (i - 1);
        square(even, odd);                          // two bits
        square(odd, even);                          // four bits
        bool have = false; uint32_t acc[32];
        // successive squarings give 8, 16, 32 ... bits = 1, 2, 4 ... bytes; multiply the ones selected by len together
        uint32_t* cur = even; uint32_t* nxt = odd;
        do {
            square(cur, nxt);                       // cur = operator for the next power-of-two BYTES
            if (len & 1) {
                if (!have) { memcpy(acc, cur, sizeof acc); have = true; }
                else { uint32_t t[32]; for (int i = 0; i < 32; i++) t[i] = times(cur, acc[i]); memcpy(acc, t, sizeof acc); }
            }
            len >>= 1;
            std::swap(cur, nxt);
        } while (len);
        memcpy(out, acc, sizeof acc);
    }
    // crc(A || B) from crc(A), crc(B) and the operator for |B| zero bytes
    static uint32_t combine(uint32_t crcA, uint32_t crcB, const uint32_t* opLenB) { return times(opLenB, crcA) ^ crcB; }
};


// CRC-32 of a host buffer computed the way the serde does it: per-chunk CRCs, then the combine chain
inline uint32_t crc32ChunkedHost(const uint8_t* data, uint64_t bytes, uint64_t chunk)
{
    Crc32Host H;
    if (bytes == 0) return 0;
    const uint64_t chunks = (bytes + chunk - 1) / chunk;
    uint32_t opFull[32], opLast[32];
    const uint64_t lastLen = bytes - (chunks - 1) * chunk;
    if (chunks > 1) Crc32Host::zeroOperator(opFull, chunk);
    Crc32Host::zeroOperator(opLast, lastLen);
    uint32_t crc = H.update(0, data, (size_t) std::min(chunk, bytes));
    for (uint64_t c = 1; c < chunks; c++) {
        const uint64_t len = c == chunks - 1 ? lastLen : chunk;
        crc = Crc32Host::combine(crc, H.update(0, data + c * chunk, (size_t) len), c == chunks - 1 ? opLast : opFull);
    }
    return crc;
}

}  // namespace pg
