// Readers and writers for the index files of a veloci directory (everything
// except the FST, which lives in fst.hpp).  All little-endian.
//
//   Indirect 1:n      src/indices/indirect/indirect.rs:10-89, indirect/mod.rs:12-20,
//                     writer create_indirect.rs:36-115
//   SingleArrayPacked src/indices/direct/single_array.rs:17-63,93-147,
//                     writer direct/create_direct.rs:40-84
//   AnchorScore       src/indices/persistence_score/token_to_anchor_score_vint.rs:37-48,128-204
//   PhrasePair        src/indices/persistence_data_binary_search.rs:51-92,126-203
//
// The structs here only *describe* byte buffers owned by someone else (the
// host loader keeps the file bytes alive); nothing is copied on the read side.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "vint.hpp"

namespace vfmt {

static const uint32_t kHighBit = 0x80000000u;

inline uint32_t load_u32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}
inline uint64_t load_u64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}

// IndexValuesMetadata (src/indices/metadata.rs:1-18)
struct IndexValuesMeta {
    uint32_t max_value_id = 0;
    float avg_join_size = 0.f;
    uint64_t num_values = 0;
    uint32_t num_ids = 0;
};

// ---------------------------------------------------------------- Indirect --
struct IndirectView {
    const uint8_t* start_pos = nullptr;  // `.indirect`
    size_t n_ids = 0;                    // start_pos bytes / 4
    const uint8_t* data = nullptr;       // `.data`
    size_t data_len = 0;

    // get_values(): false == None (id out of range or EMPTY_BUCKET)
    bool get_values(uint64_t id, std::vector<uint32_t>& out) const {
        out.clear();
        if (id >= n_ids) return false;
        uint32_t slot = load_u32(start_pos + id * 4);
        if (slot & kHighBit) {
            out.push_back(slot & ~kHighBit);
            return true;
        }
        if (slot == 0) return false;
        if (slot >= data_len) return true;  // Some(vec![]) on a dangling offset
        VintArrayIter it(data + slot, data + data_len);
        uint32_t v;
        while (it.next(v)) out.push_back(v);
        return true;
    }
    // appends instead of replacing (get_values_iter + extend)
    void append_values(uint64_t id, std::vector<uint32_t>& out) const {
        if (id >= n_ids) return;
        uint32_t slot = load_u32(start_pos + id * 4);
        if (slot & kHighBit) {
            out.push_back(slot & ~kHighBit);
            return;
        }
        if (slot == 0 || slot >= data_len) return;
        VintArrayIter it(data + slot, data + data_len);
        uint32_t v;
        while (it.next(v)) out.push_back(v);
    }
    // IndexIdToParent::get_value default: first element of get_values
    bool get_value(uint64_t id, uint32_t& v) const {
        if (id >= n_ids) return false;
        uint32_t slot = load_u32(start_pos + id * 4);
        if (slot & kHighBit) {
            v = slot & ~kHighBit;
            return true;
        }
        if (slot == 0 || slot >= data_len) return false;
        VintArrayIter it(data + slot, data + data_len);
        return it.next(v);
    }
};

// IndirectFlushingInOrderVint: ids must arrive in ascending order, once each.
struct IndirectWriter {
    std::vector<uint32_t> ids;
    std::vector<uint8_t> data;
    IndexValuesMeta meta;
    IndirectWriter() { data.push_back(0); }  // offset 0 is the EMPTY_BUCKET marker
    void add(uint32_t id, const uint32_t* vals, size_t n) {
        meta.num_values += 1;
        meta.num_ids += (uint32_t)n;
        if (ids.size() <= id) ids.resize((size_t)id + 1, 0);
        if (n == 1) {
            ids[id] = vals[0] | kHighBit;
        } else {
            ids[id] = (uint32_t)data.size();
            vint_array_serialize(data, vals, n);
        }
    }
    void add(uint32_t id, const std::vector<uint32_t>& vals) { add(id, vals.data(), vals.size()); }
    bool empty() const { return ids.empty(); }
    void finish() { meta.avg_join_size = (float)meta.num_values / (float)std::max<uint32_t>(1, meta.num_ids); }
};

// ------------------------------------------------------- SingleArrayPacked --
// get_bytes_required(): note the reference doubles the value (`val += val`).
inline int packed_bytes_required(uint32_t max_value_id) {
    uint32_t val = max_value_id + max_value_id;  // wraps like the u32 add in release builds
    if (val < (1u << 8)) return 1;
    if (val < (1u << 16)) return 2;
    if (val < (1u << 24)) return 3;
    return 4;
}

struct PackedView {
    const uint8_t* bytes = nullptr;
    size_t len = 0;
    int width = 4;
    bool get_value(uint64_t id, uint32_t& v) const {
        size_t pos = (size_t)id * (size_t)width;
        if (pos >= len) return false;
        uint32_t raw = 0;
        size_t n = std::min<size_t>((size_t)width, len - pos);
        memcpy(&raw, bytes + pos, n);
        if (raw == 0) return false;
        v = raw - 1;
        return true;
    }
};

struct PackedWriter {
    std::vector<uint32_t> cache;  // stored as val+1
    IndexValuesMeta meta;
    void add(uint32_t id, uint32_t val) {
        meta.num_values += 1;
        if (cache.size() <= id) cache.resize((size_t)id + 1, 0);
        cache[id] = val + 1;
        meta.max_value_id = std::max(meta.max_value_id, val);
    }
    std::vector<uint8_t> encode() {
        int w = packed_bytes_required(meta.max_value_id);
        std::vector<uint8_t> out(cache.size() * (size_t)w);
        for (size_t i = 0; i < cache.size(); ++i) memcpy(&out[i * w], &cache[i], (size_t)w);
        meta.avg_join_size = (float)meta.num_values / (float)std::max<size_t>(1, cache.size());
        return out;
    }
};

// ----------------------------------------------------------- AnchorScore ----
struct AnchorScoreView {
    const uint8_t* start_pos = nullptr;
    size_t start_len = 0;
    bool wide = false;  // data_type == U64
    const uint8_t* data = nullptr;
    size_t data_len = 0;

    size_t num_ids() const { return start_len / (wide ? 8 : 4); }
    // Calls f(anchor, raw_score_u32) for every posting of `id`, anchors ascending.
    template <class F>
    void for_each(uint32_t id, F&& f) const {
        if (id >= num_ids()) return;
        uint64_t pos = wide ? load_u64(start_pos + (size_t)id * 8) : load_u32(start_pos + (size_t)id * 4);
        if (pos == 0 || pos >= data_len) return;
        VintCommonIter it(data + pos, data + data_len);
        uint32_t cur = 0, d, s;
        while (it.next(d)) {
            if (!it.next(s)) break;
            cur += d;
            f(cur, s);
        }
    }
    // upper bound on the number of postings (size_hint().1 of the vint iterator: remaining bytes)
    size_t size_hint(uint32_t id) const {
        if (id >= num_ids()) return 0;
        uint64_t pos = wide ? load_u64(start_pos + (size_t)id * 8) : load_u32(start_pos + (size_t)id * 4);
        if (pos == 0 || pos >= data_len) return 0;
        VintCommonIter it(data + pos, data + data_len);
        return it.size_hint();
    }
};

// TokenToAnchorScoreVintFlushing: `pairs` = [anchor, score, anchor, score ...],
// anchors strictly ascending; delta coded here (delta_compress_data_block).
struct AnchorScoreWriter {
    std::vector<uint64_t> pos;
    std::vector<uint8_t> data;
    IndexValuesMeta meta;
    std::vector<uint32_t> scratch;
    AnchorScoreWriter() { data.push_back(0); }
    void set_scores(uint32_t id, const uint32_t* pairs, size_t n_u32) {
        if (pos.size() <= id) pos.resize((size_t)id + 1, 0);
        meta.num_values += n_u32 / 2;
        meta.num_ids += 1;
        pos[id] = data.size();
        scratch.assign(pairs, pairs + n_u32);
        uint32_t last = 0;
        for (size_t i = 0; i + 1 < n_u32; i += 2) {
            uint32_t a = scratch[i];
            scratch[i] = a - last;
            last = a;
        }
        vint_common_array_serialize(data, scratch.data(), scratch.size());
    }
    bool needs_u64() const { return data.size() >= (1ull << 32); }
    std::vector<uint8_t> encode_start_pos() const {
        bool wide = needs_u64();
        std::vector<uint8_t> out(pos.size() * (wide ? 8 : 4));
        for (size_t i = 0; i < pos.size(); ++i) {
            if (wide) memcpy(&out[i * 8], &pos[i], 8);
            else {
                uint32_t p = (uint32_t)pos[i];
                memcpy(&out[i * 4], &p, 4);
            }
        }
        return out;
    }
    void finish() { meta.avg_join_size = (float)meta.num_values / (float)std::max<uint32_t>(1, meta.num_ids); }
};

// ------------------------------------------------------------ PhrasePair ----
struct PhrasePairView {
    const uint8_t* recs = nullptr;  // 12-byte records (t1, t2, data offset), sorted by (t1, t2)
    size_t n = 0;
    const uint8_t* data = nullptr;
    size_t data_len = 0;

    // binary_search_slice(): same probing sequence as the reference.
    bool find(uint32_t t1, uint32_t t2, uint32_t& off) const {
        if (n == 0) return false;
        size_t size = n, base = 0;
        auto key_at = [&](size_t i, uint32_t& a, uint32_t& b) {
            a = load_u32(recs + i * 12);
            b = load_u32(recs + i * 12 + 4);
        };
        while (size > 1) {
            size_t half = size / 2, mid = base + half;
            uint32_t a, b;
            key_at(mid, a, b);
            bool greater = (a > t1) || (a == t1 && b > t2);
            base = greater ? base : mid;
            size -= half;
        }
        uint32_t a, b;
        key_at(base, a, b);
        if (a != t1 || b != t2) return false;
        off = load_u32(recs + base * 12 + 8);
        return true;
    }
    bool get_values(uint32_t t1, uint32_t t2, std::vector<uint32_t>& out) const {
        uint32_t off;
        if (!find(t1, t2, off)) return false;
        if (off >= data_len) return true;
        VintArrayIter it(data + off, data + data_len);
        uint32_t v;
        while (it.next(v)) out.push_back(v);
        return true;
    }
};

struct PhrasePairWriter {
    std::vector<uint8_t> recs;
    std::vector<uint8_t> data;
    IndexValuesMeta meta;
    PhrasePairWriter() { data.push_back(0); }
    // keys must arrive sorted; anchors sorted + deduped by the caller
    void add(uint32_t t1, uint32_t t2, const uint32_t* anchors, size_t n) {
        meta.num_values += 1;
        meta.num_ids += (uint32_t)n;
        uint32_t off = (uint32_t)data.size();
        size_t at = recs.size();
        recs.resize(at + 12);
        memcpy(&recs[at], &t1, 4);
        memcpy(&recs[at + 4], &t2, 4);
        memcpy(&recs[at + 8], &off, 4);
        vint_array_serialize(data, anchors, n);
    }
    bool empty() const { return recs.empty(); }
    void finish() { meta.avg_join_size = (float)meta.num_values / (float)std::max<uint32_t>(1, meta.num_ids); }
};

// -------------------------------------------------------------- paths -------
// util.rs:173-188 get_steps_to_anchor
inline std::vector<std::string> get_steps_to_anchor(const std::string& path) {
    std::vector<std::string> paths;
    std::string cur;
    size_t i = 0;
    while (i <= path.size()) {
        size_t j = path.find('.', i);
        if (j == std::string::npos) j = path.size();
        std::string part = path.substr(i, j - i);
        if (!cur.empty() || i > 0) cur += ".";
        cur += part;
        if (part.size() >= 2 && part.compare(part.size() - 2, 2, "[]") == 0) paths.push_back(cur);
        i = j + 1;
    }
    paths.push_back(path + ".textindex");
    return paths;
}

inline bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// util.rs:136-142 extract_field_name: drop the 10 trailing chars (".textindex")
inline std::string extract_field_name(const std::string& field) {
    // char-count based in the reference; ".textindex" is ASCII so bytes == chars for the tail
    if (field.size() < 10) return std::string();
    return field.substr(0, field.size() - 10);
}

}  // namespace vfmt
