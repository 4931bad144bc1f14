// vint32-style codecs used by every `.data` file of a veloci index directory.
//
// Reference call sites: src/indices/indirect/indirect.rs:68,84 (VintArrayIterator),
// src/indices/persistence_score/token_to_anchor_score_vint.rs:37-48,131-159
// (VIntArrayEncodeMostCommon / VintArrayMostCommonIterator),
// src/indices/persistence_data_binary_search.rs:96,116,200.
//
// The codec itself lives in the third-party crate `vint32 = "0.3.0"` (feature
// `common-encoding`, Cargo.toml:44) whose source is NOT under /root/reference.
// What is restated here:
//   * plain vint: little-endian base-128, high bit = "another byte follows";
//   * VIntArray::serialize(): vint(byte_len) || vints   -- pinned by the
//     reference test persistence_data_binary_search.rs:252-253 ([5,6] stored at
//     offset 1 makes the next entry start at offset 4);
//   * "most common" array: vint(most_common_value) || vint(byte_len) || items,
//     where the first byte of an item is [more:1][is_most_common:1][payload:6]
//     and continuation bytes are [more:1][payload:7].
// The last layout is UNVERIFIED AGAINST UPSTREAM (no fixture in the reference
// pins it); it is isolated in this one header so it can be swapped.
#pragma once
#include <cstddef>
#include <cstdint>
#include <unordered_map>
#include <vector>

namespace vfmt {

inline void vint_encode(std::vector<uint8_t>& out, uint32_t v) {
    while (v >= 0x80) {
        out.push_back((uint8_t)((v & 0x7F) | 0x80));
        v >>= 7;
    }
    out.push_back((uint8_t)v);
}

// Returns bytes consumed (0 when the buffer ends inside a value).
inline size_t vint_decode(const uint8_t* p, const uint8_t* end, uint32_t& v) {
    uint32_t r = 0;
    int shift = 0;
    const uint8_t* s = p;
    while (p < end) {
        uint8_t b = *p++;
        r |= (uint32_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) {
            v = r;
            return (size_t)(p - s);
        }
        shift += 7;
        if (shift > 28 + 7) break;
    }
    v = r;
    return 0;
}

// VIntArray::serialize()
inline void vint_array_serialize(std::vector<uint8_t>& out, const uint32_t* vals, size_t n) {
    std::vector<uint8_t> body;
    body.reserve(n * 2);
    for (size_t i = 0; i < n; ++i) vint_encode(body, vals[i]);
    vint_encode(out, (uint32_t)body.size());
    out.insert(out.end(), body.begin(), body.end());
}

// VintArrayIterator::from_serialized_vint_array(): iterates the values of one
// serialized array that starts at `p` (the buffer may continue after it).
struct VintArrayIter {
    const uint8_t* p = nullptr;
    const uint8_t* end = nullptr;
    VintArrayIter() = default;
    VintArrayIter(const uint8_t* data, const uint8_t* data_end) {
        if (data == nullptr || data >= data_end) return;
        uint32_t len = 0;
        size_t used = vint_decode(data, data_end, len);
        if (used == 0) return;
        p = data + used;
        end = p + len;
        if (end > data_end) end = data_end;
    }
    bool next(uint32_t& v) {
        if (p == nullptr || p >= end) return false;
        size_t used = vint_decode(p, end, v);
        if (used == 0) {
            p = end;
            return false;
        }
        p += used;
        return true;
    }
    // upper bound of remaining items, like the crate's size_hint().1
    size_t size_hint() const { return p ? (size_t)(end - p) : 0; }
};

// ---- "most common value" variant (anchor/score postings) -------------------

inline uint32_t most_common_value(const uint32_t* vals, size_t n) {
    std::unordered_map<uint32_t, uint32_t> freq;
    freq.reserve(64);
    uint32_t best = 0, best_n = 0;
    for (size_t i = 0; i < n; ++i) {
        uint32_t c = ++freq[vals[i]];
        if (c > best_n || (c == best_n && vals[i] < best)) {
            best_n = c;
            best = vals[i];
        }
    }
    return best;
}

inline void vint_common_item(std::vector<uint8_t>& out, uint32_t v, uint32_t common) {
    if (v == common) {
        out.push_back(0x40);
        return;
    }
    if (v < 0x40) {
        out.push_back((uint8_t)v);
        return;
    }
    out.push_back((uint8_t)((v & 0x3F) | 0x80));
    v >>= 6;
    while (v >= 0x80) {
        out.push_back((uint8_t)((v & 0x7F) | 0x80));
        v >>= 7;
    }
    out.push_back((uint8_t)v);
}

// VIntArrayEncodeMostCommon::encode_vals + serialize
inline void vint_common_array_serialize(std::vector<uint8_t>& out, const uint32_t* vals, size_t n) {
    uint32_t common = most_common_value(vals, n);
    std::vector<uint8_t> body;
    body.reserve(n * 2);
    for (size_t i = 0; i < n; ++i) vint_common_item(body, vals[i], common);
    vint_encode(out, common);
    vint_encode(out, (uint32_t)body.size());
    out.insert(out.end(), body.begin(), body.end());
}

// VintArrayMostCommonIterator::from_slice
struct VintCommonIter {
    const uint8_t* p = nullptr;
    const uint8_t* end = nullptr;
    uint32_t common = 0;
    VintCommonIter() = default;
    VintCommonIter(const uint8_t* data, const uint8_t* data_end) {
        if (data == nullptr || data >= data_end) return;
        size_t used = vint_decode(data, data_end, common);
        if (used == 0) return;
        data += used;
        uint32_t len = 0;
        used = vint_decode(data, data_end, len);
        if (used == 0) return;
        p = data + used;
        end = p + len;
        if (end > data_end) end = data_end;
    }
    bool next(uint32_t& v) {
        if (p == nullptr || p >= end) return false;
        uint8_t b = *p++;
        if (b == 0x40) {
            v = common;
            return true;
        }
        uint32_t r = b & 0x3F;
        int shift = 6;
        while (b & 0x80) {
            if (p >= end) return false;
            b = *p++;
            r |= (uint32_t)(b & 0x7F) << shift;
            shift += 7;
        }
        v = r;
        return true;
    }
    size_t size_hint() const { return p ? (size_t)(end - p) : 0; }
};

}  // namespace vfmt
