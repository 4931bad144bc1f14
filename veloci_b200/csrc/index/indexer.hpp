// JSON-lines -> veloci index directory.  NOT part of the accelerated path: it
// exists so that tests, the bench and the oracle have index directories in the
// reference's on-disk layout to load.
//
// Follows src/create.rs:187-283 (cb_text / callback_ids), :389-410 (posting
// merge rule), :580-720 (which file gets which writer), src/create/
// {create_fulltext.rs:27-114, calculate_score.rs:34-49, path_data.rs:57-139,
// features.rs:41-80, fields_config.rs:14-110}, json_converter/src/lib.rs:5-166
// and src/tokenizer/simple_tokenizer_group.rs:48-82.  Field configuration is
// accepted as JSON only (the reference also takes TOML).
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <optional>
#include <set>
#include <string>
#include <sys/stat.h>
#include <unordered_map>
#include <vector>

#include "../format/codecs.hpp"
#include "../format/fst.hpp"
#include "../format/unicode.hpp"
#include "../host/doc_store.hpp"
#include "../host/persistence.hpp"
#include "../vjson.hpp"

namespace vindex {

using vhost::FieldInfo;
using vhost::IndexCardinality;
using vhost::IndexCategory;
using vhost::IndexMetadata;

// tokenizer/mod.rs:17-19 DEFAULT_SEPERATORS
inline const std::vector<uint32_t>& default_separators() {
    static const std::vector<uint32_t> s = {' ', '\t', '\n', '\r', ':', '(', ')', ',', '.', 0x2026, ';', 0x30FB, 0x2019, 0x2014, '-', '\\',
                                            '[', ']', '{', '}', '<', '>', '\'', '"', 0x201C, 0x2122};
    return s;
}

// SimpleTokenizerGroupTokenIter: (token, is_separator) pieces covering the whole text
inline void tokenize(const std::string& text, const std::vector<uint32_t>& seps, std::vector<std::pair<std::string, bool>>& out) {
    out.clear();
    size_t last_returned = 0;
    bool last_was_token = false;  // the reference's name: true while inside a separator run
    size_t i = 0;
    while (i < text.size()) {
        size_t pos = i;
        uint32_t cp = vfmt::utf8_next((const uint8_t*)text.data(), text.size(), i);
        bool is_sep = std::find(seps.begin(), seps.end(), cp) != seps.end();
        if (is_sep) {
            if (pos == 0) {
                last_was_token = true;
            } else if (!last_was_token) {
                out.emplace_back(text.substr(last_returned, pos - last_returned), false);
                last_was_token = true;
                last_returned = pos;
            }
        } else if (last_was_token) {
            out.emplace_back(text.substr(last_returned, pos - last_returned), true);
            last_was_token = false;
            last_returned = pos;
        }
    }
    if (last_returned != text.size()) out.emplace_back(text.substr(last_returned), last_was_token);
}

// calculate_score.rs:34-49
inline uint32_t calculate_token_score_for_entry(uint32_t token_best_pos, uint32_t num_occurences, uint32_t num_tokens_in_text, bool is_exact) {
    float score = is_exact ? 400.f : 2000.f / (log2f((float)token_best_pos + 10.f) + 10.f);
    float m = log10f((float)num_occurences + 1000.f) - 2.f;
    m -= (m - 1.f) * 0.7f;
    score /= m;
    float t = log10f((float)(num_tokens_in_text + 10));
    t -= (t - 1.f) * 0.7f;
    score /= t;
    return (uint32_t)score;
}

enum IndexCreationType { TokensToTextID, TokenToAnchorIDScore, PhrasePairToAnchor, TextIDToTokenIds, TextIDToParent, ParentToTextID, ParentToValueID, ValueIDToParent, TextIDToAnchor, kNumIndexTypes };

struct FieldConfig {
    bool facet = false;
    bool tokenize = true;
    std::optional<std::vector<std::string>> tokenize_on_chars;
    size_t do_not_store_text_longer_than = 64;
    bool boost = false;
    bool disabled[kNumIndexTypes] = {false};
    bool enabled(IndexCreationType t) const { return !disabled[t]; }
    std::vector<uint32_t> separators() const {
        if (!tokenize_on_chars) return default_separators();
        std::vector<uint32_t> s;
        for (auto& c : *tokenize_on_chars) {
            size_t i = 0;
            if (!c.empty()) s.push_back(vfmt::utf8_next((const uint8_t*)c.data(), c.size(), i));
        }
        return s;
    }
};

// features.rs:41-80 features_to_disabled_indices
inline void apply_features(FieldConfig& c, const std::set<std::string>& f) {
    auto none_of = [&](std::initializer_list<const char*> names) {
        for (auto n : names)
            if (f.count(n)) return false;
        return true;
    };
    if (none_of({"All", "TokensToTextID", "BoostTextLocality", "Highlight", "BoostingFieldData"})) c.disabled[TokensToTextID] = true;
    if (none_of({"All", "Search"})) c.disabled[TokenToAnchorIDScore] = true;
    if (none_of({"All", "Select", "Facets"})) c.disabled[ParentToValueID] = true;
    if (none_of({"All", "BoostingFieldData"})) c.disabled[ValueIDToParent] = true;
    if (none_of({"All", "PhraseBoost"})) c.disabled[PhrasePairToAnchor] = true;
    if (none_of({"All", "Select", "WhyFound"})) c.disabled[TextIDToTokenIds] = true;
    if (none_of({"All", "BoostingFieldData"})) c.disabled[TextIDToParent] = true;
    if (none_of({"All", "Facets", "Select"})) c.disabled[ParentToTextID] = true;
    if (none_of({"All", "BoostTextLocality", "Select", "Filters"})) c.disabled[TextIDToAnchor] = true;
}

struct FieldsConfig {
    std::map<std::string, FieldConfig> fields;
    const FieldConfig& get(const std::string& path) const {
        std::string p = vfmt::ends_with(path, ".textindex") ? path.substr(0, path.size() - 10) : path;
        auto it = fields.find(p);
        if (it != fields.end()) return it->second;
        return fields.at("*GLOBAL*");
    }
    // fields_config.rs:95-110 config_from_string + :27-49 features_to_indices (JSON form)
    static FieldsConfig from_json(const std::string& json) {
        FieldsConfig fc;
        static const char* kTypeNames[kNumIndexTypes] = {"TokensToTextID", "TokenToAnchorIDScore", "PhrasePairToAnchor", "TextIDToTokenIds", "TextIDToParent", "ParentToTextID", "ParentToValueID", "ValueIDToParent", "TextIDToAnchor"};
        static const char* kAllFeatures[] = {"TokensToTextID", "BoostTextLocality", "BoostingFieldData", "Search", "Filters", "Facets", "Select", "WhyFound", "Highlight", "PhraseBoost"};
        std::string trimmed = json;
        vjson::Value root = trimmed.find('{') == std::string::npos ? vjson::Value::make_object() : vjson::parse(json);
        for (auto& kv : root.obj) {
            FieldConfig c;
            const vjson::Value& v = kv.second;
            if (auto* f = v.get("facet")) c.facet = f->is_bool() && f->b;
            if (auto* ft = v.get("fulltext"))
                if (ft->is_object()) {
                    if (auto* t = ft->get("tokenize")) c.tokenize = t->is_bool() ? t->b : true;
                    if (auto* t = ft->get("tokenize_on_chars"))
                        if (t->is_array()) {
                            std::vector<std::string> cs;
                            for (auto& e : t->arr) cs.push_back(e.str);
                            c.tokenize_on_chars = cs;
                        }
                    if (auto* t = ft->get("do_not_store_text_longer_than")) c.do_not_store_text_longer_than = (size_t)t->u64;
                }
            if (auto* b = v.get("boost")) c.boost = b->is_object();
            if (auto* d = v.get("disabled_indices"))
                if (d->is_array())
                    for (auto& e : d->arr)
                        for (int t = 0; t < kNumIndexTypes; ++t)
                            if (e.str == kTypeNames[t]) c.disabled[t] = true;
            const vjson::Value* feats = v.get("features");
            const vjson::Value* dis = v.get("disabled_features");
            if (feats && feats->is_array()) {
                std::set<std::string> f;
                for (auto& e : feats->arr) f.insert(e.str);
                apply_features(c, f);
            } else if (dis && dis->is_array()) {
                // Features::invert keeps the features that ARE listed (features.rs:23-38, sic)
                std::set<std::string> listed, f;
                for (auto& e : dis->arr) listed.insert(e.str);
                for (auto n : kAllFeatures)
                    if (listed.count(n)) f.insert(n);
                apply_features(c, f);
            }
            fc.fields[kv.first] = c;
        }
        if (!fc.fields.count("*GLOBAL*")) {
            FieldConfig c;  // FieldConfig::default(): features = {Search, TokensToTextID}
            apply_features(c, {"Search", "TokensToTextID"});
            fc.fields["*GLOBAL*"] = c;
        }
        return fc;
    }
};

struct TermInfo {
    uint32_t id = 0;
    uint32_t num_occurences = 0;
};

struct TermDataInPath {
    std::unordered_map<std::string, TermInfo> terms;
    size_t do_not_store_text_longer_than = 64;
    uint32_t id_counter_for_large_texts = 0;
};

typedef std::vector<std::pair<uint32_t, uint32_t>> KVList;

struct PathData {
    FieldConfig cfg;
    std::vector<uint32_t> seps;
    bool is_anchor_identity_column = false;
    TermDataInPath term_data;
    KVList tokens_to_text_id, text_id_to_parent, parent_to_text_id, value_id_to_anchor, text_id_to_anchor, anchor_to_text_id, boost;
    std::vector<std::pair<uint32_t, std::pair<uint32_t, uint32_t>>> token_to_anchor_id_score;
    std::vector<std::pair<std::pair<uint32_t, uint32_t>, uint32_t>> phrase_pair_to_anchor;
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> text_id_to_token_ids;
    std::set<uint32_t> text_ids_with_tokens;
    bool created = false;
};

struct PathDataIds {
    KVList value_to_parent, parent_to_value;
    bool has_v2p = false, has_p2v = false;
};

class Indexer {
  public:
    Indexer(const std::string& dir, const std::string& config_json) : dir_(dir), cfg_(FieldsConfig::from_json(config_json)) { mkdir(dir.c_str(), 0755); }

    void add_document_json(const std::string& line) { docs_.push_back(vjson::parse(line)); }
    void add_document(const vjson::Value& v) { docs_.push_back(v); }

    void finish() {
        // pass 1: collect terms (create_fulltext.rs:116-152)
        IdHolder h1;
        for (auto& d : docs_) {
            uint32_t root = h1.get_id("");
            std::string path;
            walk(d, root, h1, root, path, "", /*pass=*/1);
        }
        meta_.num_docs = docs_.size();
        {  // write_docs (src/create/write_docs.rs:11-34): the documents as JSON text, in the compressed store `data`
            vhost::DocStoreWriter store;
            std::vector<uint8_t> bytes;
            uint64_t indexed = 0;
            for (auto& d : docs_) {
                const std::string text = vjson::to_string(d);
                indexed += text.size();
                store.add_doc(text, bytes);
            }
            store.finish(bytes);
            vhost::write_file(dir_ + "/data", bytes.data(), bytes.size());
            meta_.bytes_indexed = indexed;
        }
        for (auto& kv : terms_in_path_) {
            const std::string& path = kv.first;
            TermDataInPath& td = kv.second;
            const FieldConfig& fc = cfg_.get(path);
            FieldInfo fi;
            fi.name = path;
            fi.has_fst = true;
            fi.tokenize = fc.tokenize;
            fi.tokenize_on_chars = fc.tokenize_on_chars;
            fi.do_not_store_text_longer_than = fc.do_not_store_text_longer_than;
            bool all_once = true;
            for (auto& t : td.terms) all_once = all_once && t.second.num_occurences == 1;
            fi.is_anchor_identity_column = path.find("[]") == std::string::npos && docs_.size() == td.terms.size() && all_once;
            fi.num_text_ids = td.terms.size();
            // set_ids + store_fst
            std::vector<std::pair<const std::string*, TermInfo*>> sorted;
            sorted.reserve(td.terms.size());
            for (auto& t : td.terms) sorted.emplace_back(&t.first, &t.second);
            std::sort(sorted.begin(), sorted.end(), [](auto& a, auto& b) { return *a.first < *b.first; });
            vfmt::FstWriter w;
            for (size_t i = 0; i < sorted.size(); ++i) {
                sorted[i].second->id = (uint32_t)i;
                if (sorted[i].first->size() <= fc.do_not_store_text_longer_than) w.insert(*sorted[i].first, i);
            }
            std::vector<uint8_t> bytes = w.finish();
            vhost::write_file(dir_ + "/" + path + ".textindex.fst", bytes.data(), bytes.size());
            meta_.columns[path] = fi;
        }
        // pass 2
        IdHolder h2;
        for (auto& d : docs_) {
            uint32_t root = h2.get_id("");
            std::string path;
            walk(d, root, h2, root, path, "", /*pass=*/2);
        }
        write_indices();
        std::string mj = vjson::to_string(vhost::metadata_to_json(meta_), 2);
        vhost::write_file(dir_ + "/metaData.json", mj.data(), mj.size());
    }

  private:
    struct IdHolder {
        std::unordered_map<std::string, uint32_t> ids;
        uint32_t get_id(const std::string& path) {
            auto it = ids.find(path);
            if (it != ids.end()) return ++it->second;
            ids.emplace(path, 0);
            return 0;
        }
    };

    std::string dir_;
    FieldsConfig cfg_;
    std::vector<vjson::Value> docs_;
    vhost::Metadata meta_;
    std::map<std::string, TermDataInPath> terms_in_path_;
    std::map<std::string, PathData> path_data_;
    std::map<std::string, PathDataIds> tuples_;

    static std::string convert_to_string(const vjson::Value& v) {
        switch (v.kind) {
            case vjson::Value::String: return v.str;
            case vjson::Value::Number:
                if (v.num_is_u64) return std::to_string(v.u64);
                if (!v.num_is_i64) return vjson::f64_to_string(v.num);
                return "";
            case vjson::Value::Bool: return v.b ? "true" : "false";
            default: return "";
        }
    }

    void walk(const vjson::Value& data, uint32_t anchor, IdHolder& h, uint32_t parent, std::string& path, const std::string& name, int pass) {
        if (data.is_array()) {
            path += name;
            path += "[]";
            size_t prev = path.size();
            for (auto& el : data.arr) {
                uint32_t id = h.get_id(path);
                if (pass == 2) cb_ids(path, id, parent);
                walk(el, anchor, h, id, path, "", pass);
                path.resize(prev);
            }
        } else if (data.is_object()) {
            path += name;
            if (!path.empty()) path += ".";
            size_t prev = path.size();
            std::vector<const std::pair<std::string, vjson::Value>*> sorted;
            for (auto& kv : data.obj) sorted.push_back(&kv);
            std::sort(sorted.begin(), sorted.end(), [](auto a, auto b) { return a->first < b->first; });
            for (auto kv : sorted) {
                walk(kv->second, anchor, h, parent, path, kv->first, pass);
                path.resize(prev);
            }
        } else if (!data.is_null()) {
            path += name;
            std::string value = convert_to_string(data);
            if (pass == 1) cb_text_pass1(value, path);
            else cb_text_pass2(anchor, value, path, parent);
        }
    }

    void cb_text_pass1(const std::string& text, const std::string& path) {
        const FieldConfig& fc = cfg_.get(path);
        auto it = terms_in_path_.find(path);
        if (it == terms_in_path_.end()) {
            TermDataInPath td;
            td.do_not_store_text_longer_than = fc.do_not_store_text_longer_than;
            it = terms_in_path_.emplace(path, std::move(td)).first;
        }
        TermDataInPath& td = it->second;
        auto count = [&](const std::string& t) {
            TermInfo& ti = td.terms[t];
            if (ti.num_occurences != UINT32_MAX) ti.num_occurences++;
        };
        if (td.do_not_store_text_longer_than < text.size()) td.id_counter_for_large_texts++;
        else count(text);
        if (fc.tokenize) {
            std::vector<std::pair<std::string, bool>> toks;
            tokenize(text, fc.separators(), toks);
            if (toks.size() >= 2)
                for (auto& t : toks) count(t.first);
        }
    }

    void cb_ids(const std::string& path, uint32_t value_id, uint32_t parent_val_id) {
        auto it = tuples_.find(path);
        if (it == tuples_.end()) {
            PathDataIds p;
            const FieldConfig& fc = cfg_.get(path);
            p.has_v2p = fc.enabled(ValueIDToParent);
            p.has_p2v = fc.enabled(ParentToValueID);
            it = tuples_.emplace(path, std::move(p)).first;
        }
        if (it->second.has_v2p) it->second.value_to_parent.emplace_back(value_id, parent_val_id);
        if (it->second.has_p2v) it->second.parent_to_value.emplace_back(parent_val_id, value_id);
    }

    void cb_text_pass2(uint32_t anchor, const std::string& value, const std::string& path, uint32_t parent_val_id) {
        PathData& data = path_data_[path];
        if (!data.created) {
            data.created = true;
            data.cfg = cfg_.get(path);
            data.seps = data.cfg.separators();
            data.term_data = std::move(terms_in_path_[path]);
            data.is_anchor_identity_column = meta_.columns[path].is_anchor_identity_column;
        }
        const FieldConfig& fc = data.cfg;
        // get_text_info (create.rs:143-161)
        TermInfo text_info;
        if (data.term_data.do_not_store_text_longer_than < value.size()) {
            data.term_data.id_counter_for_large_texts++;
            text_info.id = (uint32_t)data.term_data.terms.size() + 1 + data.term_data.id_counter_for_large_texts;
            text_info.num_occurences = 1;
        } else {
            text_info = data.term_data.terms.at(value);
        }
        if (fc.enabled(TextIDToParent)) data.text_id_to_parent.emplace_back(text_info.id, parent_val_id);
        if (fc.enabled(ParentToTextID)) data.parent_to_text_id.emplace_back(parent_val_id, text_info.id);
        if (fc.enabled(TextIDToAnchor) && !data.is_anchor_identity_column) data.text_id_to_anchor.emplace_back(text_info.id, anchor);
        if (fc.facet && path.find("[]") != std::string::npos) data.anchor_to_text_id.emplace_back(anchor, text_info.id);
        if (fc.boost) {
            // value.trim() != "" -> parse::<f32>()
            size_t b = value.find_first_not_of(" \t\r\n");
            if (b != std::string::npos) {
                float f = strtof(value.c_str(), nullptr);
                if (!std::isnan(f)) {
                    uint32_t bits;
                    memcpy(&bits, &f, 4);
                    data.boost.emplace_back(parent_val_id, bits);
                }
            }
            data.value_id_to_anchor.emplace_back(parent_val_id, anchor);
        }
        bool scores = fc.enabled(TokenToAnchorIDScore);
        if (scores) data.token_to_anchor_id_score.push_back({text_info.id, {anchor, calculate_token_score_for_entry(0, text_info.num_occurences, 1, true)}});
        if (fc.tokenize) {
            std::vector<std::pair<std::string, bool>> toks;
            tokenize(value, data.seps, toks);
            if (toks.size() >= 2) {
                uint32_t current_token_pos = 0;
                bool already = fc.enabled(TextIDToTokenIds) && data.text_ids_with_tokens.count(text_info.id);
                bool has_prev = false;
                uint32_t prev_token = 0;
                std::vector<uint32_t> tokens_ids;
                struct ValIdPairToken {
                    uint32_t id, num_occurences, pos;
                };
                std::vector<ValIdPairToken> toks_to_anchor;
                for (auto& t : toks) {
                    const TermInfo& ti = data.term_data.terms.at(t.first);
                    if (!already) tokens_ids.push_back(ti.id);
                    if (fc.enabled(TokensToTextID)) data.tokens_to_text_id.emplace_back(ti.id, text_info.id);
                    if (scores) {
                        toks_to_anchor.push_back({ti.id, ti.num_occurences, current_token_pos});
                        current_token_pos++;
                    }
                    if (!t.second && fc.enabled(PhrasePairToAnchor)) {
                        if (has_prev) data.phrase_pair_to_anchor.push_back({{prev_token, ti.id}, anchor});
                        prev_token = ti.id;
                        has_prev = true;
                    }
                }
                if (!already && fc.enabled(TextIDToTokenIds)) {
                    data.text_ids_with_tokens.insert(text_info.id);
                    data.text_id_to_token_ids.emplace_back(text_info.id, tokens_ids);
                }
                if (scores) {
                    // calculate_and_add_token_score_in_doc
                    std::sort(toks_to_anchor.begin(), toks_to_anchor.end(), [](auto& a, auto& b) { return a.id != b.id ? a.id < b.id : a.pos < b.pos; });
                    for (size_t i = 0; i < toks_to_anchor.size();) {
                        size_t j = i;
                        while (j < toks_to_anchor.size() && toks_to_anchor[j].id == toks_to_anchor[i].id) ++j;
                        uint32_t score = calculate_token_score_for_entry(toks_to_anchor[i].pos, toks_to_anchor[i].num_occurences, current_token_pos, false);
                        data.token_to_anchor_id_score.push_back({toks_to_anchor[i].id, {anchor, score}});
                        i = j;
                    }
                }
            }
        }
    }

    void push_index(const std::string& col, IndexMetadata im) {
        auto it = meta_.columns.find(col);
        if (it == meta_.columns.end()) {
            FieldInfo fi;
            fi.has_fst = false;
            it = meta_.columns.emplace(col, fi).first;
        }
        it->second.indices.push_back(std::move(im));
    }

    void write_multi(const std::string& col, const std::string& path, KVList& kv, bool sort_by_key, bool sort_and_dedup, IndexCategory cat = IndexCategory::KeyValue) {
        if (sort_by_key) std::stable_sort(kv.begin(), kv.end(), [](auto& a, auto& b) { return a.first < b.first; });
        vfmt::IndirectWriter w;
        uint32_t max_value = 0;
        std::vector<uint32_t> group;
        for (size_t i = 0; i < kv.size();) {
            size_t j = i;
            group.clear();
            while (j < kv.size() && kv[j].first == kv[i].first) {
                group.push_back(kv[j].second);
                max_value = std::max(max_value, kv[j].second);
                ++j;
            }
            if (sort_and_dedup) {
                std::sort(group.begin(), group.end());
                group.erase(std::unique(group.begin(), group.end()), group.end());
            }
            w.add(kv[i].first, group);
            i = j;
        }
        w.finish();
        w.meta.max_value_id = max_value;
        vhost::write_file(dir_ + "/" + path + ".indirect", w.ids.data(), w.ids.size() * 4);
        vhost::write_file(dir_ + "/" + path + ".data", w.data.data(), w.data.size());
        IndexMetadata im;
        im.path = path;
        im.category = cat;
        im.cardinality = IndexCardinality::MultiValue;
        im.is_empty = w.empty();
        im.meta = w.meta;
        push_index(col, im);
    }

    void write_single(const std::string& col, const std::string& path, KVList& kv) {
        vfmt::PackedWriter w;
        for (auto& e : kv) w.add(e.first, e.second);
        std::vector<uint8_t> bytes = w.encode();
        vhost::write_file(dir_ + "/" + path, bytes.data(), bytes.size());
        IndexMetadata im;
        im.path = path;
        im.category = IndexCategory::KeyValue;
        im.cardinality = IndexCardinality::SingleValue;
        im.is_empty = w.cache.empty();
        im.meta = w.meta;
        push_index(col, im);
    }

    void write_indices() {
        for (auto& kv : path_data_) {
            const std::string col = kv.first;
            const std::string path = col + ".textindex";
            PathData& d = kv.second;
            const FieldConfig& fc = d.cfg;
            if (fc.enabled(TokensToTextID)) write_multi(col, path + ".tokens_to_text_id", d.tokens_to_text_id, true, true);
            if (fc.enabled(TokenToAnchorIDScore)) {
                auto& v = d.token_to_anchor_id_score;
                std::sort(v.begin(), v.end(), [](auto& a, auto& b) { return a.first != b.first ? a.first < b.first : a.second.first < b.second.first; });
                vfmt::AnchorScoreWriter w;
                std::vector<uint32_t> pairs;
                for (size_t i = 0; i < v.size();) {
                    size_t j = i;
                    pairs.clear();
                    while (j < v.size() && v[j].first == v[i].first) {
                        // dedup_keep_best_score_by: same anchor -> max score + min(group len, 5)
                        size_t k = j;
                        uint32_t best = 0;
                        while (k < v.size() && v[k].first == v[i].first && v[k].second.first == v[j].second.first) {
                            best = std::max(best, v[k].second.second);
                            ++k;
                        }
                        if (k - j > 1) best += (uint32_t)std::min<size_t>(k - j, 5);
                        pairs.push_back(v[j].second.first);
                        pairs.push_back(best);
                        j = k;
                    }
                    w.set_scores(v[i].first, pairs.data(), pairs.size());
                    i = j;
                }
                w.finish();
                std::vector<uint8_t> sp = w.encode_start_pos();
                vhost::write_file(dir_ + "/" + path + ".to_anchor_id_score.indirect", sp.data(), sp.size());
                vhost::write_file(dir_ + "/" + path + ".to_anchor_id_score.data", w.data.data(), w.data.size());
                IndexMetadata im;
                im.path = path + ".to_anchor_id_score";
                im.category = IndexCategory::AnchorScore;
                im.meta = w.meta;
                im.data_type_u64 = w.needs_u64();
                push_index(col, im);
            }
            if (fc.enabled(PhrasePairToAnchor)) {
                auto& v = d.phrase_pair_to_anchor;
                std::sort(v.begin(), v.end());
                vfmt::PhrasePairWriter w;
                std::vector<uint32_t> group;
                for (size_t i = 0; i < v.size();) {
                    size_t j = i;
                    group.clear();
                    while (j < v.size() && v[j].first == v[i].first) group.push_back(v[j++].second);
                    group.erase(std::unique(group.begin(), group.end()), group.end());
                    w.add(v[i].first.first, v[i].first.second, group.data(), group.size());
                    i = j;
                }
                w.finish();
                vhost::write_file(dir_ + "/" + path + ".phrase_pair_to_anchor.indirect", w.recs.data(), w.recs.size());
                vhost::write_file(dir_ + "/" + path + ".phrase_pair_to_anchor.data", w.data.data(), w.data.size());
                IndexMetadata im;
                im.path = path + ".phrase_pair_to_anchor";
                im.category = IndexCategory::Phrase;
                im.is_empty = w.empty();
                im.meta = w.meta;
                push_index(col, im);
            }
            if (fc.enabled(TextIDToTokenIds)) {
                auto& v = d.text_id_to_token_ids;
                std::stable_sort(v.begin(), v.end(), [](auto& a, auto& b) { return a.first < b.first; });
                KVList flat;
                for (auto& e : v)
                    for (uint32_t t : e.second) flat.emplace_back(e.first, t);
                write_multi(col, path + ".text_id_to_token_ids", flat, false, false);
            }
            if (fc.enabled(TextIDToParent)) write_multi(col, path + ".value_id_to_parent", d.text_id_to_parent, true, false);
            if (fc.boost) write_multi(col, col + ".value_id_to_anchor", d.value_id_to_anchor, false, false);
            if (fc.enabled(ParentToTextID)) write_single(col, path + ".parent_to_value_id", d.parent_to_text_id);
            if (fc.enabled(TextIDToAnchor)) write_multi(col, path + ".text_id_to_anchor", d.text_id_to_anchor, true, true);
            if (fc.facet && col.find("[]") != std::string::npos) write_multi(col, path + ".anchor_to_text_id", d.anchor_to_text_id, false, false);
            if (fc.boost) write_multi(col, col + ".boost_valid_to_value", d.boost, false, false, IndexCategory::Boost);
        }
        for (auto& kv : tuples_) {
            const std::string& path = kv.first;
            if (kv.second.has_v2p) write_single(path, path + ".value_id_to_parent", kv.second.value_to_parent);
            if (kv.second.has_p2v) write_multi(path, path + ".parent_to_value_id", kv.second.parent_to_value, false, false);
        }
    }
};

}  // namespace vindex
