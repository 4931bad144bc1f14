// Host-side view of a veloci index directory: `Persistence::load`.
//
// Mirrors src/persistence.rs:53-68 (the four path-keyed index maps + fst map),
// :206-305 (load_indices dispatch on IndexCategory x IndexCardinality,
// load_all_fst), :312-348 (get_* lookups and their error text) and
// src/metadata.rs:10-43 / src/indices/metadata.rs:1-51 (metaData.json schema).
// Files are read whole into memory (the reference mmaps them); the typed views
// of format/codecs.hpp point into those buffers.
#pragma once
#include <cstdio>
#include <map>
#include <optional>
#include <memory>
#include <stdexcept>
#include <string>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "doc_store.hpp"
#include "../format/codecs.hpp"
#include "../format/fst.hpp"
#include "../format/unicode.hpp"
#include "../vjson.hpp"

namespace vhost {

struct IoError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct PathNotFound : std::runtime_error {  // VelociError::StringError via path_not_found (persistence.rs:454-458)
    using std::runtime_error::runtime_error;
};
struct FstNotFound : std::runtime_error {  // VelociError::FstNotFound (error.rs)
    std::string path;
    explicit FstNotFound(const std::string& p) : std::runtime_error("field does not exist " + p + " (fst not found)"), path(p) {}
};

inline std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw IoError("could not open " + path);
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)std::max<long>(n, 0));
    if (n > 0 && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) {
        fclose(f);
        throw IoError("short read on " + path);
    }
    fclose(f);
    return buf;
}
inline void write_file(const std::string& path, const void* data, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw IoError("could not create " + path);
    if (n && fwrite(data, 1, n, f) != n) {
        fclose(f);
        throw IoError("short write on " + path);
    }
    fclose(f);
}

enum class IndexCategory { Boost, KeyValue, AnchorScore, Phrase };
enum class IndexCardinality { MultiValue, SingleValue };

struct IndexMetadata {
    std::string path;
    IndexCategory category = IndexCategory::KeyValue;
    IndexCardinality cardinality = IndexCardinality::MultiValue;
    bool is_empty = false;
    vfmt::IndexValuesMeta meta;
    bool data_type_u64 = false;
};

struct FieldInfo {
    std::string name;
    size_t num_text_ids = 0;
    size_t num_long_text_ids = 0;
    bool tokenize = true;
    std::optional<std::vector<std::string>> tokenize_on_chars;
    size_t do_not_store_text_longer_than = 64;
    std::vector<IndexMetadata> indices;
    bool is_anchor_identity_column = false;
    bool has_fst = false;
};

struct Metadata {
    uint64_t num_docs = 0;
    uint64_t bytes_indexed = 0;
    std::map<std::string, FieldInfo> columns;
};

inline vjson::Value metadata_to_json(const Metadata& m) {
    using vjson::Value;
    Value root = Value::make_object();
    root.set("num_docs", Value::make_u64(m.num_docs));
    root.set("bytes_indexed", Value::make_u64(m.bytes_indexed));
    Value cols = Value::make_object();
    for (auto& kv : m.columns) {
        const FieldInfo& fi = kv.second;
        Value c = Value::make_object();
        c.set("name", Value::make_string(fi.name));
        Value tm = Value::make_object();
        tm.set("num_text_ids", Value::make_u64(fi.num_text_ids));
        tm.set("num_long_text_ids", Value::make_u64(fi.num_long_text_ids));
        Value opt = Value::make_object();
        opt.set("tokenize", Value::make_bool(fi.tokenize));
        if (fi.tokenize_on_chars) {
            Value arr = Value::make_array();
            for (auto& s : *fi.tokenize_on_chars) arr.arr.push_back(Value::make_string(s));
            opt.set("tokenize_on_chars", arr);
        } else {
            opt.set("tokenize_on_chars", Value());
        }
        opt.set("do_not_store_text_longer_than", Value::make_u64(fi.do_not_store_text_longer_than));
        tm.set("options", opt);
        c.set("textindex_metadata", tm);
        Value idx = Value::make_array();
        for (auto& im : fi.indices) {
            Value e = Value::make_object();
            e.set("path", Value::make_string(im.path));
            const char* cat = im.category == IndexCategory::Boost ? "Boost" : im.category == IndexCategory::KeyValue ? "KeyValue" : im.category == IndexCategory::AnchorScore ? "AnchorScore" : "Phrase";
            e.set("index_category", Value::make_string(cat));
            e.set("index_cardinality", Value::make_string(im.cardinality == IndexCardinality::MultiValue ? "MultiValue" : "SingleValue"));
            e.set("is_empty", Value::make_bool(im.is_empty));
            Value md = Value::make_object();
            md.set("max_value_id", Value::make_u64(im.meta.max_value_id));
            Value avg;
            avg.kind = Value::Number;
            avg.num = im.meta.avg_join_size;
            md.set("avg_join_size", avg);
            md.set("num_values", Value::make_u64(im.meta.num_values));
            md.set("num_ids", Value::make_u64(im.meta.num_ids));
            e.set("metadata", md);
            e.set("data_type", Value::make_string(im.data_type_u64 ? "U64" : "U32"));
            idx.arr.push_back(e);
        }
        c.set("indices", idx);
        c.set("is_anchor_identity_column", Value::make_bool(fi.is_anchor_identity_column));
        c.set("has_fst", Value::make_bool(fi.has_fst));
        cols.set(kv.first, c);
    }
    root.set("columns", cols);
    return root;
}

inline Metadata metadata_from_json(const vjson::Value& root) {
    Metadata m;
    auto u64 = [](const vjson::Value* v) -> uint64_t { return (v && v->is_number()) ? (v->num_is_u64 ? v->u64 : (uint64_t)v->num) : 0; };
    m.num_docs = u64(root.get("num_docs"));
    m.bytes_indexed = u64(root.get("bytes_indexed"));
    const vjson::Value* cols = root.get("columns");
    if (!cols || !cols->is_object()) throw IoError("metaData.json: missing columns");
    for (auto& kv : cols->obj) {
        const vjson::Value& c = kv.second;
        FieldInfo fi;
        if (auto* n = c.get("name")) fi.name = n->str;
        if (auto* tm = c.get("textindex_metadata")) {
            fi.num_text_ids = (size_t)u64(tm->get("num_text_ids"));
            fi.num_long_text_ids = (size_t)u64(tm->get("num_long_text_ids"));
            if (auto* opt = tm->get("options")) {
                if (auto* t = opt->get("tokenize")) fi.tokenize = t->b;
                if (auto* t = opt->get("tokenize_on_chars"))
                    if (t->is_array()) {
                        std::vector<std::string> cs;
                        for (auto& e : t->arr) cs.push_back(e.str);
                        fi.tokenize_on_chars = cs;
                    }
                if (auto* t = opt->get("do_not_store_text_longer_than")) fi.do_not_store_text_longer_than = (size_t)u64(t);
            }
        }
        if (auto* idx = c.get("indices"))
            for (auto& e : idx->arr) {
                IndexMetadata im;
                if (auto* p = e.get("path")) im.path = p->str;
                if (auto* p = e.get("index_category")) {
                    const std::string& s = p->str;
                    im.category = s == "Boost" ? IndexCategory::Boost : s == "AnchorScore" ? IndexCategory::AnchorScore : s == "Phrase" ? IndexCategory::Phrase : IndexCategory::KeyValue;
                }
                if (auto* p = e.get("index_cardinality")) im.cardinality = p->str == "SingleValue" ? IndexCardinality::SingleValue : IndexCardinality::MultiValue;
                if (auto* p = e.get("is_empty")) im.is_empty = p->is_bool() && p->b;
                if (auto* md = e.get("metadata")) {
                    im.meta.max_value_id = (uint32_t)u64(md->get("max_value_id"));
                    if (auto* a = md->get("avg_join_size")) im.meta.avg_join_size = (float)a->num;
                    im.meta.num_values = u64(md->get("num_values"));
                    im.meta.num_ids = (uint32_t)u64(md->get("num_ids"));
                }
                if (auto* p = e.get("data_type")) im.data_type_u64 = p->str == "U64";
                fi.indices.push_back(std::move(im));
            }
        if (auto* p = c.get("is_anchor_identity_column")) fi.is_anchor_identity_column = p->is_bool() && p->b;
        if (auto* p = c.get("has_fst")) fi.has_fst = p->is_bool() && p->b;
        m.columns[kv.first] = std::move(fi);
    }
    return m;
}

// trait IndexIdToParent (persistence.rs:142-181) over the two concrete stores
struct KeyValueStore {
    enum Kind { Empty, Indirect, Packed } kind = Empty;
    vfmt::IndirectView ind;
    vfmt::PackedView packed;
    vfmt::IndexValuesMeta meta;

    bool get_values(uint64_t id, std::vector<uint32_t>& out) const {
        out.clear();
        if (kind == Indirect) return ind.get_values(id, out);
        if (kind == Packed) {
            uint32_t v;
            if (!packed.get_value(id, v)) return false;
            out.push_back(v);
            return true;
        }
        return false;
    }
    void append_values(uint64_t id, std::vector<uint32_t>& out) const {
        if (kind == Indirect) ind.append_values(id, out);
        else if (kind == Packed) {
            uint32_t v;
            if (packed.get_value(id, v)) out.push_back(v);
        }
    }
    bool get_value(uint64_t id, uint32_t& v) const {
        if (kind == Indirect) return ind.get_value(id, v);
        if (kind == Packed) return packed.get_value(id, v);
        return false;
    }
};

// Decoded term dictionary of one `.textindex.fst`: terms in byte order with their ids.
struct TermDict {
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> offsets;  // n + 1
    std::vector<uint32_t> ids;      // n, ascending
    size_t size() const { return ids.size(); }
    std::string term(size_t i) const { return std::string((const char*)&bytes[offsets[i]], offsets[i + 1] - offsets[i]); }
    // term id -> dictionary slot (ids are ascending, may have gaps)
    bool find_id(uint32_t id, size_t& slot) const {
        auto it = std::lower_bound(ids.begin(), ids.end(), id);
        if (it == ids.end() || *it != id) return false;
        slot = (size_t)(it - ids.begin());
        return true;
    }
};

struct Persistence {
    std::string dir;
    Metadata metadata;
    std::unordered_map<std::string, std::vector<uint8_t>> files;
    std::unordered_map<std::string, KeyValueStore> key_value_stores;
    std::unordered_map<std::string, vfmt::AnchorScoreView> token_to_anchor_score;
    std::unordered_map<std::string, vfmt::PhrasePairView> phrase_pair_to_anchor;
    std::unordered_map<std::string, KeyValueStore> boost_valueid_to_value;
    std::unordered_map<std::string, vfmt::FstReader> fst;
    std::unordered_map<std::string, TermDict> dict;  // same keys as `fst`

    // The document store (`data`, src/search.rs:78-79), opened on first use.
    std::mutex doc_mu;
    DocLoader docs;
    const DocLoader& doc_store() {
        std::lock_guard<std::mutex> lock(doc_mu);
        if (!docs.is_open()) {
            const std::vector<uint8_t>& f = file("data");
            try {
                docs = DocLoader(f.data(), f.size());
            } catch (const DocStoreError& e) {
                throw IoError(e.what());
            }
        }
        return docs;
    }
    std::string get_doc(uint32_t doc_id) {
        try {
            return doc_store().get_doc(doc_id);
        } catch (const DocStoreError& e) {
            throw IoError(e.what());
        }
    }

    const std::vector<uint8_t>& file(const std::string& name) {
        auto it = files.find(name);
        if (it != files.end()) return it->second;
        return files.emplace(name, read_file(dir + "/" + name)).first->second;
    }

    static std::unique_ptr<Persistence> load(const std::string& dir) {
        std::unique_ptr<Persistence> p(new Persistence());
        p->dir = dir;
        const std::vector<uint8_t>& mj = p->file("metaData.json");
        try {
            p->metadata = metadata_from_json(vjson::parse((const char*)mj.data(), mj.size()));
        } catch (const vjson::ParseError& e) {
            throw IoError(std::string("metaData.json: ") + e.what());
        }
        p->load_indices();
        return p;
    }

    void load_indices() {
        for (auto& col : metadata.columns)
            for (auto& el : col.second.indices) {
                switch (el.category) {
                    case IndexCategory::Phrase: {
                        vfmt::PhrasePairView v;
                        if (!el.is_empty) {
                            const auto& ind = file(el.path + ".indirect");
                            const auto& dat = file(el.path + ".data");
                            v.recs = ind.data();
                            v.n = ind.size() / 12;
                            v.data = dat.data();
                            v.data_len = dat.size();
                        }
                        phrase_pair_to_anchor[el.path] = v;
                        break;
                    }
                    case IndexCategory::AnchorScore: {
                        const auto& ind = file(el.path + ".indirect");
                        const auto& dat = file(el.path + ".data");
                        vfmt::AnchorScoreView v;
                        v.start_pos = ind.data();
                        v.start_len = ind.size();
                        v.wide = el.data_type_u64;
                        v.data = dat.data();
                        v.data_len = dat.size();
                        token_to_anchor_score[el.path] = v;
                        break;
                    }
                    case IndexCategory::Boost:
                    case IndexCategory::KeyValue: {
                        KeyValueStore s;
                        s.meta = el.meta;
                        if (el.category == IndexCategory::KeyValue && el.is_empty) {
                            s.kind = KeyValueStore::Empty;
                        } else if (el.cardinality == IndexCardinality::MultiValue) {
                            const auto& ind = file(el.path + ".indirect");
                            const auto& dat = file(el.path + ".data");
                            s.kind = KeyValueStore::Indirect;
                            s.ind.start_pos = ind.data();
                            s.ind.n_ids = ind.size() / 4;
                            s.ind.data = dat.data();
                            s.ind.data_len = dat.size();
                        } else {
                            const auto& dat = file(el.path);
                            s.kind = KeyValueStore::Packed;
                            s.packed.bytes = dat.data();
                            s.packed.len = dat.size();
                            s.packed.width = vfmt::packed_bytes_required(el.meta.max_value_id);
                        }
                        if (el.category == IndexCategory::Boost) boost_valueid_to_value[el.path] = s;
                        else key_value_stores[el.path] = s;
                        break;
                    }
                }
            }
        // load_all_fst
        for (auto& col : metadata.columns) {
            if (!col.second.has_fst) continue;
            std::string path = col.first + ".textindex";
            const auto& bytes = file(path + ".fst");
            vfmt::FstReader rd(bytes.data(), bytes.size());
            TermDict d;
            d.offsets.push_back(0);
            rd.for_each([&](const std::string& k, uint64_t v) {
                d.bytes.insert(d.bytes.end(), k.begin(), k.end());
                d.offsets.push_back((uint32_t)d.bytes.size());
                d.ids.push_back((uint32_t)v);
            });
            fst.emplace(path, rd);
            dict.emplace(path, std::move(d));
        }
    }

    [[noreturn]] void path_not_found(const std::string& path) const {
        std::string all;
        std::vector<std::string> keys;
        for (auto& kv : key_value_stores) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end());
        for (auto& k : keys) all += k + "\n";
        throw PathNotFound("Did not found path in indices " + path + "\nAll loaded indices: \n" + all);
    }

    const KeyValueStore& get_valueid_to_parent(const std::string& path) const {
        auto it = key_value_stores.find(path);
        if (it == key_value_stores.end()) path_not_found(path);
        return it->second;
    }
    bool has_index(const std::string& path) const { return key_value_stores.count(path) != 0; }
    const KeyValueStore& get_boost(const std::string& path) const {
        auto it = boost_valueid_to_value.find(path);
        if (it == boost_valueid_to_value.end()) path_not_found(path);
        return it->second;
    }
    const vfmt::AnchorScoreView& get_token_to_anchor(const std::string& path) const {
        std::string p = path + ".to_anchor_id_score";
        auto it = token_to_anchor_score.find(p);
        if (it == token_to_anchor_score.end()) path_not_found(p);
        return it->second;
    }
    const vfmt::PhrasePairView& get_phrase_pair_to_anchor(const std::string& path) const {
        auto it = phrase_pair_to_anchor.find(path);
        if (it == phrase_pair_to_anchor.end()) path_not_found(path);
        return it->second;
    }
    const TermDict& get_dict(const std::string& path) const {
        auto it = dict.find(path);
        if (it == dict.end()) throw FstNotFound(path);
        return it->second;
    }
    const vfmt::FstReader& get_fst(const std::string& path) const {
        auto it = fst.find(path);
        if (it == fst.end()) throw FstNotFound(path);
        return it->second;
    }
    bool is_anchor_identity_column(const std::string& textindex_path) const {
        auto it = metadata.columns.find(vfmt::extract_field_name(textindex_path));
        return it != metadata.columns.end() && it->second.is_anchor_identity_column;
    }
    bool is_tokenized(const std::string& textindex_path) const {
        auto it = metadata.columns.find(vfmt::extract_field_name(textindex_path));
        return it != metadata.columns.end() && it->second.tokenize;
    }
    // search_field.rs:520-526 get_text_for_id
    std::string get_text_for_id(const std::string& path, uint32_t id) const {
        std::string out;
        get_fst(path).ord_to_term(id, out);
        return out;
    }
};

}  // namespace vhost
