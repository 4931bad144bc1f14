// `select`: documents rebuilt from the indices instead of read from the document store (src/search/read_document.rs:8-59,
// src/search.rs:242-279), and the why_found that goes with it (src/search/why_found.rs:11-49, highlight_document
// src/highlight_field.rs:187-271): the text of a hit is its list of token ids (text_id_to_token_ids), the matched term ids
// mark the hit positions.  Host code over the index files (Persistence); no kernel.
#pragma once
#include <algorithm>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../format/codecs.hpp"
#include "../vjson.hpp"
#include "highlight.hpp"
#include "persistence.hpp"

namespace vhost {

struct MissingTextId : std::runtime_error {  // VelociError::MissingTextId (src/search.rs:256-259)
    using std::runtime_error::runtime_error;
};

// join_and_get_text_for_ids (src/search.rs:242-269): the text of `prop` under parent `id`; false = the parent has none
inline bool text_of(const Persistence& p, uint32_t id, const std::string& prop, std::string& text) {
    const std::string field = prop + ".textindex";
    uint32_t text_id;
    if (!p.get_valueid_to_parent(field + ".parent_to_value_id").get_value(id, text_id)) return false;
    auto col = p.metadata.columns.find(prop);
    if (col == p.metadata.columns.end()) throw IoError("no column " + prop + " in metaData.json");  // (an index panic in the reference)
    if (text_id >= (uint32_t)col->second.num_text_ids) {  // a long text is not in the dictionary: it is the concatenation of its tokens
        std::vector<uint32_t> tokens;
        if (!p.get_valueid_to_parent(field + ".text_id_to_token_ids").get_values(text_id, tokens))
            throw MissingTextId("missing text_id " + std::to_string(text_id) + " in index " + field + ".text_id_to_token_ids, therefore could not load text");
        text.clear();
        for (uint32_t t : tokens) text += p.get_text_for_id(field, t);
    } else {
        text = p.get_text_for_id(field, text_id);
    }
    return true;
}

// The tree of the selected fields (util.rs:175-229 get_all_steps_to_anchor + to_node_tree): key = the path up to a segment,
// a node without children is a leaf; where a selected field is a prefix of another one the shorter wins (is_leaf).
struct ReadTree {
    std::map<std::string, ReadTree> kids;
};
inline ReadTree read_tree_of(const Persistence& p, const std::vector<std::string>& fields) {
    ReadTree root;
    for (const std::string& field : fields) {
        if (!p.has_index(field + ".textindex.parent_to_value_id")) continue;  // get_read_tree_from_fields (search.rs:272-279)
        ReadTree* at = &root;
        for (size_t dot = field.find('.');; dot = field.find('.', dot + 1)) {
            at = &at->kids[field.substr(0, dot)];
            if (dot == std::string::npos) break;
        }
    }
    // a node that ends a selected field is a leaf, whatever hangs below it
    std::function<void(ReadTree&)> prune = [&](ReadTree& t) {
        for (auto& kv : t.kids) {
            if (std::find(fields.begin(), fields.end(), kv.first) != fields.end() && p.has_index(kv.first + ".textindex.parent_to_value_id")) kv.second.kids.clear();
            else prune(kv.second);
        }
    };
    prune(root);
    return root;
}

inline std::string prop_name(const std::string& path) {  // util.rs:138-144 extract_prop_name
    size_t from = path.rfind('.');
    std::string last = path.substr(from == std::string::npos ? 0 : from + 1);
    if (last.size() >= 2 && last.compare(last.size() - 2, 2, "[]") == 0) last.resize(last.size() - 2);
    return last;
}

// read_tree (src/search/read_document.rs:13-59)
inline vjson::Value read_tree(const Persistence& p, uint32_t id, const ReadTree& tree) {
    vjson::Value json = vjson::Value::make_object();
    for (auto& kv : tree.kids) {
        const std::string& prop = kv.first;
        const std::string current = prop + ".parent_to_value_id";
        const bool is_array = prop.size() >= 2 && prop.compare(prop.size() - 2, 2, "[]") == 0;
        std::vector<uint32_t> sub_ids;
        if (kv.second.kids.empty()) {
            std::string text;
            if (is_array) {
                if (p.get_valueid_to_parent(current).get_values(id, sub_ids)) {
                    vjson::Value arr = vjson::Value::make_array();
                    for (uint32_t sub : sub_ids)
                        if (text_of(p, sub, prop, text)) arr.arr.push_back(vjson::Value::make_string(text));
                    json.set(prop_name(prop), arr);
                }
            } else if (text_of(p, id, prop, text)) {
                json.set(prop_name(prop), vjson::Value::make_string(text));
            }
        } else if (!p.has_index(current)) {  // an object in an object: no 1:n information to follow
            json.set(prop_name(prop), read_tree(p, id, kv.second));
        } else if (p.get_valueid_to_parent(current).get_values(id, sub_ids)) {
            if (is_array) {
                vjson::Value arr = vjson::Value::make_array();
                for (uint32_t sub : sub_ids) arr.arr.push_back(read_tree(p, sub, kv.second));
                json.set(prop_name(prop), arr);
            } else if (!sub_ids.empty()) {
                json.set(prop_name(prop), read_tree(p, sub_ids[0], kv.second));
            }
        }
    }
    std::sort(json.obj.begin(), json.obj.end(), [](auto& a, auto& b) { return a.first < b.first; });  // serde_json's map is ordered by key
    return json;
}

// read_data (src/search/read_document.rs:8-11)
inline vjson::Value read_data(const Persistence& p, uint32_t id, const std::vector<std::string>& fields) { return read_tree(p, id, read_tree_of(p, fields)); }

// highlight_document (src/highlight_field.rs:187-271): `text_id` of `path` (= "<field>.textindex") with the tokens in `hit_ids`
// marked; false = nothing to show
inline bool highlight_by_token_ids(const Persistence& p, const std::string& path, uint32_t text_id, const std::set<uint32_t>& hit_ids, const SnippetInfo& opt, std::string& out) {
    std::vector<uint32_t> doc;
    if (!p.get_valueid_to_parent(path + ".text_id_to_token_ids").get_values(text_id, doc)) {
        if (!hit_ids.count(text_id)) return false;
        out = opt.start_tag + p.get_text_for_id(path, text_id) + opt.end_tag;  // the text is one token and it is a hit
        return true;
    }
    std::vector<int64_t> hit_pos;
    for (size_t i = 0; i < doc.size(); ++i)
        if (hit_ids.count(doc[i])) hit_pos.push_back((int64_t)i);
    if (hit_pos.empty()) return false;
    const int64_t around = opt.num_words_around_snippet * 2;
    out.clear();
    uint64_t n_windows = 0;
    for (size_t g = 0; g < hit_pos.size() && n_windows < opt.max_snippets;) {
        size_t last = g;
        while (last + 1 < hit_pos.size() && hit_pos[last + 1] - hit_pos[last] < around) ++last;
        const size_t from = (size_t)std::max<int64_t>(hit_pos[g] - around, 0), to = (size_t)std::min<int64_t>(hit_pos[last] + around + 1, (int64_t)doc.size());
        if (n_windows++) out += opt.connector;
        for (size_t i = from; i < to; ++i) {
            const bool hit = hit_ids.count(doc[i]) != 0;
            if (hit) out += opt.start_tag;
            out += p.get_text_for_id(path, doc[i]);
            if (hit) out += opt.end_tag;
        }
        g = last + 1;
    }
    if (hit_pos.front() > around) out.insert(0, opt.connector);
    if (hit_pos.back() < (int64_t)doc.size() - around) out += opt.connector;
    return true;
}

// get_why_found (src/search/why_found.rs:11-49) for one anchor: `term_ids_in_field`: "<field>.textindex" -> the term ids the
// request's parts matched there.  field -> highlighted texts.
inline std::map<std::string, std::vector<std::string>> why_found_by_ids(const Persistence& p, uint32_t anchor, const std::map<std::string, std::set<uint32_t>>& term_ids_in_field) {
    std::map<std::string, std::vector<std::string>> out;
    const SnippetInfo opt;
    for (auto& kv : term_ids_in_field) {
        if (kv.second.empty()) continue;
        const std::string field = vfmt::extract_field_name(kv.first);
        const std::vector<std::string> steps = vfmt::get_steps_to_anchor(field);  // every "[]" level, then "<field>.textindex"
        std::vector<uint32_t> ids{anchor}, next;
        for (auto& step : steps) {  // join_anchor_to_leaf (facet.rs:75-93)
            const KeyValueStore& kv_store = p.get_valueid_to_parent(step + ".parent_to_value_id");
            next.clear();
            for (uint32_t id : ids) kv_store.append_values(id, next);
            ids.swap(next);
        }
        for (uint32_t text_id : ids) {
            std::string snippet;
            if (highlight_by_token_ids(p, steps.back(), text_id, kv.second, opt, snippet)) out[field].push_back(std::move(snippet));
        }
    }
    return out;
}

}  // namespace vhost
