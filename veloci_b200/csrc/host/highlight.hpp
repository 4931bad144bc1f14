// why_found on the stored document (src/highlight_field.rs:98-186): the fast path of search::to_documents
// (src/search.rs:65-101) when the request has no `select` -- the document's text values are tokenized again and the
// tokens that equal a matched term of their field are wrapped in <b>..</b>; only windows of five words around the hits
// are kept, joined by " ... ".
//
// `why_found_terms`: field path with ".textindex" -> the texts of the terms the request matched in that field
// (SearchResult::why_found_terms = term_text_in_field, src/search.rs:186, search_field.rs:386-389).
#pragma once
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../format/unicode.hpp"
#include "../vjson.hpp"
#include "persistence.hpp"

namespace vhost {

using TermSets = std::map<std::string, std::set<std::string>>;

struct SnippetInfo {  // src/search/request/snippet_info.rs:1-39 (the defaults: DEFAULT_SNIPPETINFO)
    int64_t num_words_around_snippet = 5;
    std::string start_tag = "<b>", end_tag = "</b>", connector = " ... ";
    uint64_t max_snippets = 0xFFFFFFFFull;
};

// The pieces of `text`: maximal runs of separator scalars and of other scalars, alternating, covering the text
// (tokenizer/simple_tokenizer_group.rs:48-82 yields exactly these; a piece is a view into `text`).
inline std::vector<std::pair<size_t, size_t>> split_runs(const std::string& text, const std::vector<uint32_t>& separators) {
    std::vector<std::pair<size_t, size_t>> runs;  // (begin, end)
    size_t i = 0, run_begin = 0;
    int run_kind = -1;
    while (i < text.size()) {
        const size_t at = i;
        const uint32_t cp = vfmt::utf8_next((const uint8_t*)text.data(), text.size(), i);
        int kind = 0;
        for (uint32_t s : separators) kind |= s == cp;
        if (run_kind >= 0 && kind != run_kind) runs.emplace_back(run_begin, at), run_begin = at;
        run_kind = kind;
    }
    if (run_begin < text.size()) runs.emplace_back(run_begin, text.size());
    return runs;
}

inline const std::vector<uint32_t>& default_token_separators() {  // tokenizer/mod.rs:17-19
    static const std::vector<uint32_t> s = {' ', '\t', '\n', '\r', ':', '(', ')', ',', '.', 0x2026, ';', 0x30FB, 0x2019, 0x2014, '-', '\\',
                                            '[', ']', '{', '}', '<', '>', '\'', '"', 0x201C, 0x2122};
    return s;
}

// highlight_text (src/highlight_field.rs:98-141): nullopt-like = false
inline bool highlight_text(const std::string& text, const std::set<std::string>& terms, const SnippetInfo& opt, const std::vector<uint32_t>* separators, std::string& out) {
    if (terms.size() == 1 && terms.count(text)) {  // the one hit is the whole text
        out = opt.start_tag + text + opt.end_tag;
        return true;
    }
    if (!separators) return false;  // field without a tokenizer
    const auto runs = split_runs(text, *separators);
    std::vector<char> is_hit(runs.size(), 0);
    std::vector<int64_t> hit_pos;
    for (size_t p = 0; p < runs.size(); ++p)
        if (terms.count(text.substr(runs[p].first, runs[p].second - runs[p].first))) is_hit[p] = 1, hit_pos.push_back((int64_t)p);
    if (hit_pos.empty()) return false;
    const int64_t around = opt.num_words_around_snippet * 2;  // every second piece is a separator run
    out.clear();
    uint64_t n_windows = 0;
    for (size_t g = 0; g < hit_pos.size() && n_windows < opt.max_snippets;) {  // hits closer than `around` pieces share a window
        size_t last = g;
        while (last + 1 < hit_pos.size() && hit_pos[last + 1] - hit_pos[last] < around) ++last;
        const size_t from = (size_t)std::max<int64_t>(hit_pos[g] - around, 0), to = (size_t)std::min<int64_t>(hit_pos[last] + around + 1, (int64_t)runs.size());
        if (n_windows++) out += opt.connector;
        for (size_t p = from; p < to; ++p) {
            if (is_hit[p]) out += opt.start_tag;
            out.append(text, runs[p].first, runs[p].second - runs[p].first);
            if (is_hit[p]) out += opt.end_tag;
        }
        g = last + 1;
    }
    if (hit_pos.front() > around) out.insert(0, opt.connector);                       // ellipsis_snippet (:80-96)
    if (hit_pos.back() < (int64_t)runs.size() - around) out += opt.connector;
    return true;
}

// for_each_element's text callback (json_converter): every scalar value with its field path ("a.b[].c", "tags[]")
template <class F>
inline void for_each_text(const vjson::Value& v, std::string& path, const std::string& name, F&& cb) {
    const size_t mark = path.size();
    if (v.is_array()) {
        path += name, path += "[]";
        for (auto& el : v.arr) for_each_text(el, path, "", cb);
    } else if (v.is_object()) {
        path += name;
        if (!path.empty()) path += ".";
        for (auto& kv : v.obj) for_each_text(kv.second, path, kv.first, cb);
    } else if (!v.is_null()) {
        path += name;
        std::string text;
        if (v.is_string()) text = v.str;
        else if (v.is_bool()) text = v.b ? "true" : "false";
        else if (v.num_is_u64) text = std::to_string(v.u64);
        else if (v.num_is_i64) text = std::to_string(v.i64);
        else text = vjson::f64_to_string(v.num);
        cb(text, path);
    }
    path.resize(mark);
}

// highlight_on_original_document (src/highlight_field.rs:148-186): field -> highlighted texts, in document order
inline std::map<std::string, std::vector<std::string>> highlight_document(const Metadata& meta, const std::string& doc_json, const TermSets& why_found_terms) {
    std::map<std::string, std::vector<std::string>> out;
    if (why_found_terms.empty()) return out;
    const vjson::Value doc = vjson::parse(doc_json);
    const SnippetInfo opt;
    std::string path;
    for_each_text(doc, path, "", [&](const std::string& text, const std::string& field) {
        auto terms = why_found_terms.find(field + ".textindex");
        if (terms == why_found_terms.end()) return;
        auto col = meta.columns.find(field);
        if (col == meta.columns.end()) throw IoError("could not find metadata for \"" + field + "\"");  // a panic in the reference
        std::vector<uint32_t> seps;
        const std::vector<uint32_t>* use = nullptr;
        if (col->second.tokenize) {
            if (col->second.tokenize_on_chars) {
                for (auto& c : *col->second.tokenize_on_chars) {
                    size_t i = 0;
                    if (!c.empty()) seps.push_back(vfmt::utf8_next((const uint8_t*)c.data(), c.size(), i));
                }
                use = &seps;
            } else {
                use = &default_token_separators();
            }
        }
        std::string snippet;
        if (highlight_text(text, terms->second, opt, use, snippet)) out[field].push_back(std::move(snippet));
    });
    return out;
}

inline void write_highlights(std::string& out, const std::map<std::string, std::vector<std::string>>& h) {
    out += '{';
    bool first = true;
    for (auto& kv : h) {
        if (!first) out += ',';
        first = false;
        vjson::write_string(out, kv.first);
        out += ":[";
        for (size_t i = 0; i < kv.second.size(); ++i) {
            if (i) out += ',';
            vjson::write_string(out, kv.second[i]);
        }
        out += ']';
    }
    out += '}';
}

}  // namespace vhost
