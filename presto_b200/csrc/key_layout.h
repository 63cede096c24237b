// presto_b200 — host-side planning of the packed group key (see group_table.cuh).
#pragma once
#include "group_table.cuh"

namespace pg {

struct KeyLayout {
    int nKeys = 0;
    int W = 1;
    bool bigintPath = false;
    bool anyInterned() const { for (int k = 0; k < nKeys; k++) if (interned[k]) return true; return false; }
    int types[kMaxKeyColumns] = {0};
    int channels[kMaxKeyColumns] = {0};
    int maxLens[kMaxKeyColumns] = {0};
    // VARCHAR keys that do not fit a packed field (no declared length, or more than 7 bytes) enter the key as the 32-bit id of their
    // value in a per-operator VarcharInterner (varchar_interner.h); the operator substitutes the id column before packing
    bool interned[kMaxKeyColumns] = {false};
    int word[kMaxKeyColumns] = {0}, shift[kMaxKeyColumns] = {0}, bits[kMaxKeyColumns] = {0};
    int nullWord = 0, nullShift = 0;

    // GroupByHash.createGroupByHash (MB/operator/GroupByHash.java:46-59): Bigint path iff exactly one BIGINT channel.
    void plan(int n, const int32_t* chans, const int32_t* tys, const int32_t* lens)
    {
        PG_CHECK_ARG(n >= 1 && n <= kMaxKeyColumns, "between 1 and 8 group-by channels are supported on the GPU path");
        nKeys = n;
        for (int k = 0; k < n; k++) { channels[k] = chans[k]; types[k] = tys[k]; maxLens[k] = lens ? lens[k] : 0; }
        bigintPath = (n == 1 && types[0] == PG_TYPE_BIGINT);
        if (bigintPath) { W = 1; word[0] = 0; shift[0] = 0; bits[0] = 64; nullWord = -1; return; }
        // field widths
        for (int k = 0; k < n; k++) {
            switch (types[k]) {
                case PG_TYPE_BIGINT: case PG_TYPE_DOUBLE: bits[k] = 64; break;
                case PG_TYPE_INTEGER: case PG_TYPE_DATE: bits[k] = 32; break;
                case PG_TYPE_BOOLEAN: bits[k] = 8; break;
                case PG_TYPE_VARCHAR:
                    if (maxLens[k] <= 0 || maxLens[k] > 7) { interned[k] = true; bits[k] = 32; }
                    else bits[k] = maxLens[k] <= 1 ? 16 : (maxLens[k] <= 3 ? 32 : 64);
                    break;
                default: throw Error(PG_ERR_UNSUPPORTED, "group-by key type");
            }
        }
        // greedy packing, no field straddles a word; the null-flag byte goes first so that a single-word key keeps bit 63 clear
        int used[kMaxKeyWords + 4] = {0};
        nullWord = 0; nullShift = 0; used[0] = 8;
        int nw = 1;
        for (int k = 0; k < n; k++) {
            int w = 0;
            while (used[w] + bits[k] > 64 || (w == 0 && used[0] + bits[k] > 63)) { w++; if (w >= kMaxKeyWords + 4) break; }
            // word 0 keeps its top bit clear so a W == 1 key can never equal the empty sentinel
            if (w >= kMaxKeyWords) throw Error(PG_ERR_UNSUPPORTED, "group-by key wider than 4 x 64 bits");
            word[k] = w; shift[k] = used[w]; used[w] += bits[k];
            if (w + 1 > nw) nw = w + 1;
        }
        W = nw;
    }

    void fill(KeyPlan& kp, const std::vector<DeviceColumn>& cols /* one per key */, unsigned int* errorFlag) const
    {
        kp.nKeys = nKeys; kp.bigintPath = bigintPath ? 1 : 0; kp.nullWord = nullWord; kp.nullShift = nullShift; kp.errorFlag = errorFlag;
        for (int k = 0; k < nKeys; k++) {
            kp.f[k].values = cols[k].values; kp.f[k].nulls = cols[k].nulls; kp.f[k].offsets = cols[k].offsets;
            kp.f[k].type = interned[k] ? PG_TYPE_INTEGER : types[k]; kp.f[k].word = word[k]; kp.f[k].shift = shift[k]; kp.f[k].bits = bits[k]; kp.f[k].maxLen = maxLens[k];
        }
    }
};

}  // namespace pg
