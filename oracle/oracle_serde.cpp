/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 *
 * SerializedPage wire format, restated from
 *   SPI = presto-spi/src/main/java/com/facebook/presto/spi/page/, PC = presto-common/src/main/java/com/facebook/presto/common/block/
 *   SPI/PagesSerde.java:67-156 (serialize, deserialize, wrapSlice), SPI/PagesSerdeUtil.java:45-120 (raw page, page metadata, checksum),
 *   SPI/PageCodecMarker.java:25-45, PC/BlockEncodingManager.java:63-118 (length-prefixed encoding name),
 *   PC/LongArrayBlockEncoding.java:34-71 (same shape: IntArray, ByteArray), PC/VariableWidthBlockEncoding.java:35-75,
 *   PC/RunLengthBlockEncoding.java:30-55, PC/DictionaryBlockEncoding.java:30-80, PC/EncoderUtil.java:31-100 (null bits, MSB first).
 * Pinned by TestPagesSerde's serialized sizes (MB tests: execution/buffer/TestPagesSerde.java:83-127) and, for the checksum, by
 * java.util.zip.CRC32 == zlib crc32 (checked against python's zlib in tests/test_oracle_serde.py).
 */
#include "oracle_common.h"

namespace oracle {

// java.util.zip.CRC32 (CRC-32/ISO-HDLC, reflected polynomial 0xEDB88320)
static uint32_t crcTable[256];
static bool crcReady = false;
static void crcInit()
{
    if (crcReady) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crcTable[i] = c;
    }
    crcReady = true;
}
static uint32_t crcUpdate(uint32_t crc, const uint8_t* p, size_t n)
{
    crcInit();
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = crcTable[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

struct Writer {
    std::vector<uint8_t> b;
    void i32(int32_t v) { uint8_t t[4]; memcpy(t, &v, 4); b.insert(b.end(), t, t + 4); }
    void i64(int64_t v) { uint8_t t[8]; memcpy(t, &v, 8); b.insert(b.end(), t, t + 8); }
    void u8(uint8_t v) { b.push_back(v); }
    void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*) p; b.insert(b.end(), q, q + n); }
    void name(const char* s) { i32((int32_t) strlen(s)); bytes(s, strlen(s)); }   // BlockEncodingManager.writeLengthPrefixedString :113-118
};

// EncoderUtil.encodeNullsAsBits :31-63
static void writeNulls(Writer& w, const uint8_t* isNull, int64_t n)
{
    w.u8(isNull ? 1 : 0);            // block.mayHaveNull()
    if (!isNull) return;
    for (int64_t p = 0; p < (n & ~7LL); p += 8) {
        uint8_t v = 0;
        for (int k = 0; k < 8; k++) v |= isNull[p + k] ? (uint8_t) (0x80 >> k) : 0;
        w.u8(v);
    }
    if (n & 7) {
        uint8_t v = 0; int mask = 0x80;
        for (int64_t p = n & ~7LL; p < n; p++) { v |= isNull[p] ? mask : 0; mask >>= 1; }
        w.u8(v);
    }
}

static const char* flatEncodingName(int type)
{
    switch (type) {
        case PG_TYPE_BIGINT: case PG_TYPE_DOUBLE: return "LONG_ARRAY";
        case PG_TYPE_INTEGER: case PG_TYPE_DATE: return "INT_ARRAY";
        case PG_TYPE_BOOLEAN: return "BYTE_ARRAY";
        case PG_TYPE_VARCHAR: return "VARIABLE_WIDTH";
    }
    throw OracleError(PG_ERR_UNSUPPORTED, "type has no wire encoding here");
}

// flat storage of `n` positions → Long/Int/ByteArrayBlockEncoding.writeBlock or VariableWidthBlockEncoding.writeBlock
static void writeFlat(Writer& w, int type, int64_t n, const void* values, const uint8_t* isNull, const int32_t* offsets)
{
    w.name(flatEncodingName(type));
    w.i32((int32_t) n);
    if (type == PG_TYPE_VARCHAR) {
        // offsets relative to the first position (VariableWidthBlockEncoding :44-50), nulls, total length, bytes
        int32_t total = 0;
        for (int64_t p = 0; p < n; p++) { total += offsets[p + 1] - offsets[p]; w.i32(total); }
        writeNulls(w, isNull, n);
        w.i32(total);
        w.bytes((const uint8_t*) values + (n ? offsets[0] : 0), (size_t) total);
        return;
    }
    writeNulls(w, isNull, n);
    const int width = typeWidth(type);
    for (int64_t p = 0; p < n; p++)
        if (!isNull || !isNull[p]) w.bytes((const uint8_t*) values + (size_t) p * width, width);
}

static void writeBlock(Writer& w, const pg_block& b)
{
    if (b.encoding == PG_ENC_RLE) {          // RunLengthBlockEncoding.writeBlock :30-40: run length, then the single-position value block
        w.name("RLE");
        w.i32((int32_t) b.position_count);
        writeFlat(w, b.type, 1, b.values, b.is_null, b.offsets);
        return;
    }
    if (b.encoding == PG_ENC_DICTIONARY) {
        // DictionaryBlockEncoding.writeBlock :34-54: the dictionary is compacted first (DictionaryBlock.compact :584-628: referenced
        // entries only, renumbered in order of first reference; untouched when every entry is referenced), then position count, the
        // dictionary block, the ids, the dictionary instance id (DictionaryId.randomDictionaryId :29-32: this node's UUID + a sequence)
        const int64_t n = b.position_count, dictSize = b.dictionary_size;
        std::vector<int32_t> remap((size_t) dictSize, -1), keep, ids((size_t) n);
        for (int64_t p = 0; p < n; p++) {
            const int32_t id = b.ids[p];
            if (id < 0 || id >= dictSize) throw OracleError(PG_ERR_INVALID_ARGUMENT, "reference to a non-existent key");
            if (remap[id] == -1) { remap[id] = (int32_t) keep.size(); keep.push_back(id); }
        }
        w.name("DICTIONARY");
        w.i32((int32_t) n);
        if ((int64_t) keep.size() == dictSize) {
            writeFlat(w, b.type, dictSize, b.values, b.is_null, b.offsets);
            for (int64_t p = 0; p < n; p++) ids[p] = b.ids[p];
        }
        else {      // dictionary.copyPositions(dictionaryPositionsToCopy)
            const int64_t m = (int64_t) keep.size();
            std::vector<uint8_t> nulls;
            if (b.is_null) { nulls.resize((size_t) m + 1); for (int64_t k = 0; k < m; k++) nulls[k] = b.is_null[keep[k]]; }     // (+1: a null array exists even for 0 entries)
            if (b.type == PG_TYPE_VARCHAR) {
                std::vector<int32_t> off((size_t) m + 1, 0); std::vector<uint8_t> bytes;
                for (int64_t k = 0; k < m; k++) {
                    const int32_t o = b.offsets[keep[k]], len = b.offsets[keep[k] + 1] - o;
                    bytes.insert(bytes.end(), (const uint8_t*) b.values + o, (const uint8_t*) b.values + o + len);
                    off[k + 1] = off[k] + len;
                }
                writeFlat(w, b.type, m, bytes.data(), b.is_null ? nulls.data() : nullptr, off.data());
            }
            else {
                const int width = typeWidth(b.type);
                std::vector<uint8_t> vals((size_t) m * width);
                for (int64_t k = 0; k < m; k++) memcpy(&vals[(size_t) k * width], (const uint8_t*) b.values + (size_t) keep[k] * width, width);
                writeFlat(w, b.type, m, vals.data(), b.is_null ? nulls.data() : nullptr, nullptr);
            }
            for (int64_t p = 0; p < n; p++) ids[p] = remap[b.ids[p]];
        }
        w.bytes(ids.data(), (size_t) n * 4);
        static const uint64_t nodeId[2] = {0x6f7261636c652d6eULL, 0x6f64652d69640000ULL};     // "oracle-node-id": one per process, like the JVM's random UUID
        static int64_t sequence = 0;
        w.i64((int64_t) nodeId[0]); w.i64((int64_t) nodeId[1]); w.i64(sequence++);
        return;
    }
    writeFlat(w, b.type, b.position_count, b.values, b.is_null, b.offsets);
}

// PagesSerdeUtil.computeSerializedPageChecksum :109-120
static int64_t pageChecksum(const uint8_t* data, size_t n, uint8_t markers, int32_t positionCount, int32_t uncompressedSize)
{
    uint32_t c = crcUpdate(0, data, n);
    c = crcUpdate(c, &markers, 1);
    uint8_t t[4];
    memcpy(t, &positionCount, 4); c = crcUpdate(c, t, 4);       // updateCrc: the four bytes, low first
    memcpy(t, &uncompressedSize, 4); c = crcUpdate(c, t, 4);
    return (int64_t) (uint64_t) c;
}

struct Reader {
    const uint8_t* p; const uint8_t* end;
    void need(size_t n) const { if ((size_t) (end - p) < n) throw OracleError(PG_ERR_INVALID_ARGUMENT, "serialized page is truncated"); }
    int32_t i32() { need(4); int32_t v; memcpy(&v, p, 4); p += 4; return v; }
    int64_t i64() { need(8); int64_t v; memcpy(&v, p, 8); p += 8; return v; }
    uint8_t u8() { need(1); return *p++; }
    std::string name() { int32_t n = i32(); if (n < 0) throw OracleError(PG_ERR_INVALID_ARGUMENT, "bad encoding name"); need(n); std::string s((const char*) p, n); p += n; return s; }
};

// storage of one deserialized channel (kept alive by the serde object until the next call)
struct WireColumn {
    std::vector<uint8_t> values, nulls; std::vector<int32_t> offsets, ids;
    bool hasNulls = false;
};

// EncoderUtil.decodeNullBits :68-100
static bool readNulls(Reader& r, int64_t n, std::vector<uint8_t>& out)
{
    if (!r.u8()) return false;
    out.assign(n, 0);
    for (int64_t p = 0; p < (n & ~7LL); p += 8) {
        const uint8_t v = r.u8();
        for (int k = 0; k < 8; k++) out[p + k] = (v & (0x80 >> k)) != 0;
    }
    if (n & 7) {
        const uint8_t v = r.u8(); int mask = 0x80;
        for (int64_t p = n & ~7LL; p < n; p++) { out[p] = (v & mask) != 0; mask >>= 1; }
    }
    return true;
}

// reads a flat block of `type` whose encoding name has already been consumed
static void readFlat(Reader& r, const std::string& enc, int type, WireColumn& c, int64_t& n)
{
    if (enc != flatEncodingName(type)) throw OracleError(PG_ERR_INVALID_ARGUMENT, "block encoding " + enc + " does not fit the channel type");
    n = r.i32();
    if (n < 0) throw OracleError(PG_ERR_INVALID_ARGUMENT, "negative position count");
    if (type == PG_TYPE_VARCHAR) {
        c.offsets.assign(n + 1, 0);
        for (int64_t p = 0; p < n; p++) c.offsets[p + 1] = r.i32();
        c.hasNulls = readNulls(r, n, c.nulls);
        const int32_t total = r.i32();
        r.need(total);
        c.values.assign(r.p, r.p + total); r.p += total;
        return;
    }
    c.hasNulls = readNulls(r, n, c.nulls);
    const int width = typeWidth(type);
    c.values.assign((size_t) n * width, 0);
    for (int64_t p = 0; p < n; p++)
        if (!c.hasNulls || !c.nulls[p]) { r.need(width); memcpy(&c.values[(size_t) p * width], r.p, width); r.p += width; }
}

class PagesSerde {
public:
    bool checksumEnabled;
    int compression = PG_COMPRESSION_NONE;
    std::vector<uint8_t> wire, decoded;
    std::vector<WireColumn> cols; std::vector<pg_block> blocks;
    explicit PagesSerde(bool c, int z = PG_COMPRESSION_NONE) : checksumEnabled(c), compression(z) {}

    // ---- LZ4 block format (what io.airlift.compress.lz4.Lz4Compressor / Lz4Decompressor read and write, PagesSerdeFactory.java:75,96).
    // aircompressor is not under /root/reference: the FORMAT is restated from its published description; which matches the encoder
    // picks is implementation defined (parity unpinned for compressed bytes — the tests compare what streams decode to, across this
    // encoder, the GPU's and liblz4's).  A plain single-table greedy encoder: hash of 4 bytes → last position, extend, emit.
    static void lz4Length(std::vector<uint8_t>& out, size_t v) { v -= 15; while (v >= 255) { out.push_back(255); v -= 255; } out.push_back((uint8_t) v); }
    static std::vector<uint8_t> lz4Compress(const uint8_t* src, size_t n)
    {
        std::vector<uint8_t> out;
        std::vector<int64_t> table(1 << 14, -1);
        size_t anchor = 0, ip = 0;
        auto read32 = [&](size_t p) { uint32_t v; memcpy(&v, src + p, 4); return v; };
        auto sequence = [&](size_t litEnd, size_t offset, size_t matchLen) {          // matchLen == 0: the final literals
            const size_t lit = litEnd - anchor;
            const size_t ml = matchLen ? matchLen - 4 : 0;
            out.push_back((uint8_t) ((std::min<size_t>(lit, 15) << 4) | std::min<size_t>(ml, 15)));
            if (lit >= 15) lz4Length(out, lit);
            out.insert(out.end(), src + anchor, src + litEnd);
            if (!matchLen) return;
            out.push_back((uint8_t) (offset & 0xff)); out.push_back((uint8_t) (offset >> 8));
            if (ml >= 15) lz4Length(out, ml);
        };
        while (n >= 13 && ip + 12 <= n) {                       // the last match starts at least 12 bytes before the end
            const uint32_t v = read32(ip);
            const size_t h = (v * 2654435761u) >> 18;
            const int64_t cand = table[h];
            table[h] = (int64_t) ip;
            if (cand >= 0 && ip - (size_t) cand <= 65535 && read32((size_t) cand) == v) {
                size_t len = 4;
                while (ip + len < n - 5 && src[(size_t) cand + len] == src[ip + len]) len++;      // the last 5 bytes stay literals
                sequence(ip, ip - (size_t) cand, len);
                ip += len; anchor = ip;
            }
            else ip++;
        }
        sequence(n, 0, 0);
        return out;
    }
    static bool lz4Decompress(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t expect)
    {
        out.clear(); out.reserve(expect);
        size_t ip = 0;
        while (ip < n) {
            const uint8_t token = src[ip++];
            size_t lit = token >> 4;
            if (lit == 15) { uint8_t b; do { if (ip >= n) return false; b = src[ip++]; lit += b; } while (b == 255); }
            if (ip + lit > n || out.size() + lit > expect) return false;
            out.insert(out.end(), src + ip, src + ip + lit); ip += lit;
            if (ip >= n) break;
            if (ip + 2 > n) return false;
            const size_t off = src[ip] | ((size_t) src[ip + 1] << 8); ip += 2;
            if (off == 0 || off > out.size()) return false;
            size_t ml = token & 15;
            if (ml == 15) { uint8_t b; do { if (ip >= n) return false; b = src[ip++]; ml += b; } while (b == 255); }
            ml += 4;
            if (out.size() + ml > expect) return false;
            for (size_t i = 0; i < ml; i++) out.push_back(out[out.size() - off]);
        }
        return out.size() == expect;
    }

    void serialize(const pg_page* page, const uint8_t** bytes, int64_t* size)
    {
        Writer raw;
        raw.i32(page->channel_count);                               // PagesSerdeUtil.writeRawPage :45-51
        for (int ch = 0; ch < page->channel_count; ch++) writeBlock(raw, page->blocks[ch]);
        uint8_t markers = 0; int64_t checksum = 0;                  // PagesSerde.wrapSlice :118-156 (no cipher)
        const int32_t n = (int32_t) page->position_count, rawSize = (int32_t) raw.b.size();
        std::vector<uint8_t>* slice = &raw.b;
        std::vector<uint8_t> compressed;
        if (compression == PG_COMPRESSION_LZ4) {                    // :122-137: kept only when compressedSize / uncompressedSize <= 0.8
            compressed = lz4Compress(raw.b.data(), raw.b.size());
            if ((double) compressed.size() / (double) rawSize <= 0.8) { slice = &compressed; markers |= PG_PAGE_MARKER_COMPRESSED; }
        }
        if (checksumEnabled) { markers |= PG_PAGE_MARKER_CHECKSUMMED; checksum = pageChecksum(slice->data(), slice->size(), markers, n, rawSize); }
        Writer w;                                                   // PagesSerdeUtil.writeSerializedPageMetadata :70-77
        w.i32(n); w.u8(markers); w.i32(rawSize); w.i32((int32_t) slice->size()); w.i64(checksum);
        w.bytes(slice->data(), slice->size());
        wire.swap(w.b);
        *bytes = wire.data(); *size = (int64_t) wire.size();
    }

    void deserialize(const uint8_t* bytes, int64_t size, const int32_t* types, int nTypes, pg_page* out)
    {
        Reader r{bytes, bytes + size};
        const int32_t positionCount = r.i32(); const uint8_t markers = r.u8();      // readSerializedPage :79-88
        const int32_t uncompressed = r.i32(), sz = r.i32(); const int64_t checksum = r.i64();
        if (markers & PG_PAGE_MARKER_ENCRYPTED) throw OracleError(PG_ERR_UNSUPPORTED, "encrypted pages are not supported");
        r.need(sz);
        if ((markers & PG_PAGE_MARKER_CHECKSUMMED) && pageChecksum(r.p, sz, markers, positionCount, uncompressed) != checksum)
            throw OracleError(PG_ERR_INVALID_ARGUMENT, "serialized page checksum mismatch");
        r.end = r.p + sz;
        if (markers & PG_PAGE_MARKER_COMPRESSED) {                  // PagesSerde.deserialize :95-107
            if (uncompressed < 0 || !lz4Decompress(r.p, (size_t) sz, decoded, (size_t) uncompressed))
                throw OracleError(PG_ERR_INVALID_ARGUMENT, "page size changed after decompression (malformed LZ4 block)");
            r = Reader{decoded.data(), decoded.data() + decoded.size()};
        }
        const int32_t channels = r.i32();                                            // readRawPage :53-62
        if (channels != nTypes) throw OracleError(PG_ERR_INVALID_ARGUMENT, "channel count of the page differs from the types given");
        cols.assign(channels, WireColumn()); blocks.assign(channels, pg_block{});
        for (int ch = 0; ch < channels; ch++) {
            WireColumn& c = cols[ch]; pg_block& b = blocks[ch];
            const int type = types[ch];
            std::string enc = r.name();
            int64_t n = 0;
            b.type = type; b.encoding = PG_ENC_FLAT; b.dictionary_size = 0;
            if (enc == "RLE") {                                   // RunLengthBlockEncoding.readBlock :43-53
                const int32_t run = r.i32();
                std::string inner = r.name();
                int64_t one = 0;
                readFlat(r, inner, type, c, one);
                if (one != 1) throw OracleError(PG_ERR_INVALID_ARGUMENT, "RLE value block must hold one position");
                b.encoding = PG_ENC_RLE; b.position_count = run; b.dictionary_size = 1;
            }
            else if (enc == "DICTIONARY") {                       // DictionaryBlockEncoding.readBlock :57-80
                const int32_t count = r.i32();
                std::string inner = r.name();
                int64_t dictSize = 0;
                readFlat(r, inner, type, c, dictSize);
                c.ids.assign(count, 0);
                r.need((size_t) count * 4); memcpy(c.ids.data(), r.p, (size_t) count * 4); r.p += (size_t) count * 4;
                r.i64(); r.i64(); r.i64();                        // dictionary instance id
                b.encoding = PG_ENC_DICTIONARY; b.position_count = count; b.dictionary_size = dictSize; b.ids = c.ids.data();
            }
            else { readFlat(r, enc, type, c, n); b.position_count = n; }
            b.values = c.values.data(); b.is_null = c.hasNulls ? c.nulls.data() : nullptr;
            b.offsets = type == PG_TYPE_VARCHAR ? c.offsets.data() : nullptr;
        }
        out->position_count = positionCount; out->channel_count = channels; out->memory = PG_MEM_HOST; out->blocks = blocks.data();
    }
};

}  // namespace oracle

using oracle::OracleError;
extern "C" void or_set_error(const char* msg);

#define OR_TRY try {
#define OR_CATCH } catch (const OracleError& e) { or_set_error(e.what()); return e.code; } catch (const std::exception& e) { or_set_error(e.what()); return PG_ERR_INVALID_STATE; } return PG_OK;

extern "C" {
int32_t or_pages_serde_create(const pg_pages_serde_config* config, pg_pages_serde** out)
{
    OR_TRY *out = (pg_pages_serde*) new oracle::PagesSerde(config->checksum_enabled != 0, config->compression); OR_CATCH
}
int32_t or_pages_serde_serialize(pg_pages_serde* serde, const pg_page* page, const uint8_t** bytes, int64_t* size)
{
    OR_TRY ((oracle::PagesSerde*) serde)->serialize(page, bytes, size); OR_CATCH
}
// one SerializedPage per row range, back to back (OptimizedPartitionedOutputOperator.flush :774-797: one page per partition buffer)
int32_t or_pages_serde_serialize_ranges(pg_pages_serde* serde, const pg_page* page, const int64_t* rowOffsets, int32_t nRanges, const uint8_t** bytes, int64_t* byteOffsets)
{
    OR_TRY
    static thread_local std::vector<uint8_t> all;
    all.clear();
    byteOffsets[0] = 0;
    for (int32_t r = 0; r < nRanges; r++) {
        const int64_t lo = rowOffsets[r], hi = rowOffsets[r + 1];
        if (lo < 0 || lo > hi || hi > page->position_count) throw OracleError(PG_ERR_INVALID_ARGUMENT, "row offsets must be ascending and inside the page");
        std::vector<pg_block> blocks(page->blocks, page->blocks + page->channel_count);
        for (pg_block& b : blocks) {
            if (b.encoding == PG_ENC_FLAT) {
                const int w = oracle::typeWidth(b.type);
                if (w) b.values = (const char*) b.values + lo * w; else if (b.offsets) b.offsets += lo;
                if (b.is_null) b.is_null += lo;
            }
            else if (b.encoding == PG_ENC_DICTIONARY) b.ids += lo;
            b.position_count = hi - lo;
        }
        pg_page part = *page; part.blocks = blocks.data(); part.position_count = hi - lo;
        const uint8_t* one = nullptr; int64_t size = 0;
        ((oracle::PagesSerde*) serde)->serialize(&part, &one, &size);
        all.insert(all.end(), one, one + size);
        byteOffsets[r + 1] = (int64_t) all.size();
    }
    *bytes = all.data();
    OR_CATCH
}
int32_t or_pages_serde_deserialize(pg_pages_serde* serde, const uint8_t* bytes, int64_t size, int32_t memory, const int32_t* types, int32_t n_types, pg_page* out)
{
    OR_TRY
    if (memory != PG_MEM_HOST) throw OracleError(PG_ERR_INVALID_ARGUMENT, "the oracle reads host bytes");
    ((oracle::PagesSerde*) serde)->deserialize(bytes, size, types, n_types, out);
    OR_CATCH
}
int32_t or_pages_serde_destroy(pg_pages_serde* serde) { delete (oracle::PagesSerde*) serde; return PG_OK; }
}
