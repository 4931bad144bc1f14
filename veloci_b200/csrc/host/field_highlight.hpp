// search_field::highlight (src/search/search_field.rs:232-245): the texts of a field that a search part hits, with the hit
// tokens marked -- get_term_ids_in_field, then resolve_token_hits_to_text_id (:549-636) with snippets, ordered by score.
// The device matches and scores the part's terms (as for suggest); what follows here is host work on its few term hits:
// token ids -> text ids (`.tokens_to_text_id`), one hit per text with the largest |score| of its tokens, the text
// rebuilt from its token ids with the hits tagged (highlight_document, src/highlight_field.rs:187-271).
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "part_hits.hpp"
#include "read_document.hpp"
#include "request.hpp"

namespace vhost {

inline bool is_white_space(uint32_t c) {  // Unicode White_Space: what `\s` and str::trim mean in Rust
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F ||
           c == 0x3000;
}

// util::normalize_text (src/util.rs:11-30): five replacements one after the other, then to_lowercase and trim.
// (`\d` is matched for ASCII digits only; the regex crate also takes the other decimal digits of Unicode.)
inline std::string normalize_text(const std::string& text) {
    std::vector<uint32_t> s, t;
    vfmt::utf8_decode(text, s);
    auto is_digit = [](uint32_t c) { return c >= '0' && c <= '9'; };
    // \([fmn\d]\) -> " "
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '(' && i + 2 < s.size() && s[i + 2] == ')' && (s[i + 1] == 'f' || s[i + 1] == 'm' || s[i + 1] == 'n' || is_digit(s[i + 1]))) {
            t.push_back(' ');
            i += 2;
        } else {
            t.push_back(s[i]);
        }
    }
    s.swap(t), t.clear();
    for (uint32_t c : s) t.push_back((c == '(' || c == ')') ? (uint32_t)' ' : c);  // [\(\)] -> " "
    s.swap(t), t.clear();
    for (uint32_t c : s)  // [{}'"“] -> ""
        if (!(c == '{' || c == '}' || c == '\'' || c == '"' || c == 0x201C)) t.push_back(c);
    s.swap(t), t.clear();
    for (size_t i = 0; i < s.size();) {  // \s\s+ -> " "
        size_t j = i;
        while (j < s.size() && is_white_space(s[j])) ++j;
        if (j - i >= 2) {
            t.push_back(' ');
            i = j;
        } else {
            t.push_back(s[i]);
            ++i;
        }
    }
    s.swap(t), t.clear();
    for (uint32_t c : s)  // [,.…;・’-] -> ""
        if (!(c == ',' || c == '.' || c == 0x2026 || c == ';' || c == 0x30FB || c == 0x2019 || c == '-')) t.push_back(c);
    std::string out;
    for (uint32_t c : t) vfmt::utf8_append(out, c);
    out = vfmt::to_lowercase(out);
    std::vector<uint32_t> low;
    vfmt::utf8_decode(out, low);
    size_t a = 0, b = low.size();
    while (a < b && is_white_space(low[a])) ++a;
    while (b > a && is_white_space(low[b - 1])) --b;
    std::string trimmed;
    for (size_t i = a; i < b; ++i) vfmt::utf8_append(trimmed, low[i]);
    return trimmed;
}

struct HighlightRequest {
    SearchPart part;       // terms normalized (search_field.rs:234)
    bool snippet = false;  // RequestSearchPart::snippet
    SnippetInfo info;      // RequestSearchPart::snippet_info, DEFAULT_SNIPPETINFO without it
};

inline HighlightRequest parse_highlight_request(const vjson::Value& v) {
    HighlightRequest r;
    r.part = parse_search_part(v);
    for (std::string& t : r.part.terms) t = normalize_text(t);
    if (const vjson::Value* s = v.get("snippet"))
        if (!s->is_null()) {
            if (!s->is_bool()) throw RequestError("invalid type for snippet: expected a boolean");
            r.snippet = s->b;
        }
    if (const vjson::Value* s = v.get("snippet_info"))
        if (!s->is_null()) {
            if (!s->is_object()) throw RequestError("snippet_info must be an object");
            if (const vjson::Value* f = s->get("num_words_around_snippet")) r.info.num_words_around_snippet = (int64_t)f->num;
            if (const vjson::Value* f = s->get("snippet_start_tag")) r.info.start_tag = f->str;
            if (const vjson::Value* f = s->get("snippet_end_tag")) r.info.end_tag = f->str;
            if (const vjson::Value* f = s->get("snippet_connector")) r.info.connector = f->str;
            if (const vjson::Value* f = s->get("max_snippets")) r.info.max_snippets = (uint64_t)f->num;
        }
    return r;
}

struct FieldHighlight {
    std::string text;  // the highlighted text
    float score;
    uint32_t id;       // text id
};

// resolve_token_hits_to_text_id with snippets + get_text_score_id_from_result(false, ..) over the part's final term hits
// (after the per-part bound and boost).  Without `snippet` the reference has no highlighted text to return for a hit (it
// indexes an empty map and panics): InvalidRequest here; likewise for a field that is not tokenized.
inline std::vector<FieldHighlight> highlight_field(const Persistence& p, const HighlightRequest& r, const std::vector<vdev::TermHit>& hits) {
    std::string path = r.part.path;
    if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
    if (!r.snippet) throw vplan::InvalidRequest("highlight needs `snippet: true`: without it no highlighted text exists for a hit");
    if (!p.is_tokenized(path)) throw vplan::InvalidRequest("highlight needs a tokenized field: " + path);
    const KeyValueStore& tokens_to_text_id = p.get_valueid_to_parent(path + ".tokens_to_text_id");
    struct TokenHit {
        uint32_t parent;
        float score;
        uint32_t token;
    };
    std::vector<TokenHit> token_hits;
    std::vector<uint32_t> parents;
    for (const vdev::TermHit& h : hits) {
        parents.clear();
        if (!tokens_to_text_id.get_values(h.id, parents)) continue;
        for (uint32_t parent : parents) token_hits.push_back(TokenHit{parent, h.score, h.id});
    }
    std::stable_sort(token_hits.begin(), token_hits.end(), [](const TokenHit& a, const TokenHit& b) { return a.parent < b.parent; });
    std::vector<FieldHighlight> out;
    for (size_t i = 0; i < token_hits.size();) {
        size_t j = i;
        float max_score = token_hits[i].score;
        std::set<uint32_t> tokens;
        for (; j < token_hits.size() && token_hits[j].parent == token_hits[i].parent; ++j) {
            if (std::fabs(token_hits[j].score) >= std::fabs(max_score)) max_score = token_hits[j].score;  // max_by_key(|score|): the last of equals
            tokens.insert(token_hits[j].token);
        }
        std::string text;
        if (highlight_by_token_ids(p, path, token_hits[i].parent, tokens, r.info, text)) out.push_back(FieldHighlight{std::move(text), max_score, token_hits[i].parent});
        i = j;
    }
    // get_text_score_id_from_result(false, ..): by score, then the part's own skip / top (search.rs:230-239)
    std::stable_sort(out.begin(), out.end(), [](const FieldHighlight& a, const FieldHighlight& b) { return a.score > b.score; });
    if (r.part.skip) out.erase(out.begin(), out.begin() + (long)std::min<uint64_t>(*r.part.skip, out.size()));
    if (r.part.top && out.size() > *r.part.top) out.resize((size_t)*r.part.top);
    return out;
}

}  // namespace vhost
