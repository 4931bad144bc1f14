// query_generator::search_query / suggest_query (src/query_generator.rs:175-257, 297-322): the convenience layer that
// turns a user's query string plus a few settings into the `search::Request` the batch planner takes, so a server can
// feed raw query strings to vgpu_batch_prepare.  The output is the request as JSON text, with the keys and the omission
// rules of the reference's serde derive (src/search/request/mod.rs:14-87, search_request.rs:126-176), i.e. what
// `serde_json::to_string(&request)` gives there and what `serde_json::from_str::<Request>` accepts.
//
// Where the reference iterates a hash map the order here is fixed and stated: fields and boost_terms in key order, phrase
// pairs in order of first appearance.  The hit sets do not depend on it.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>
#include <optional>
#include <string>
#include <vector>

#include "../vjson.hpp"
#include "persistence.hpp"
#include "query_parser.hpp"

namespace vquery {

struct GeneratorError : std::runtime_error {  // VelociError::FieldNotFound / AllFieldsFiltered (src/error.rs:13-17)
    using std::runtime_error::runtime_error;
};
struct ParamsError : std::runtime_error {  // SearchQueryGeneratorParameters that do not deserialize
    using std::runtime_error::runtime_error;
};

// SearchQueryGeneratorParameters (src/query_generator.rs:44-83); `operator`, `select`, `stopword_lists` and `stopwords` are
// read and have no effect on the request, as in the reference (the filtered tree of query_parser_to_veloci_request.rs:12
// is dropped).
struct GeneratorParams {
    std::string search_term;
    ParserOptions parser_options, filter_parser_options;
    std::optional<uint64_t> top, skip, levenshtein, levenshtein_auto_limit, facetlimit;
    std::optional<bool> ignore_case, why_found, text_locality, phrase_pairs, explain;
    std::optional<std::vector<vjson::Value>> boost_queries;
    std::optional<std::vector<std::string>> facets, fields;
    std::optional<std::map<std::string, float>> boost_fields, boost_terms;
    std::optional<std::string> filter;
};

namespace detail {

inline const vjson::Value* field(const vjson::Value& o, const char* key) {
    const vjson::Value* v = o.get(key);
    return v && !v->is_null() ? v : nullptr;
}
inline std::optional<uint64_t> opt_usize(const vjson::Value& o, const char* key) {
    const vjson::Value* v = field(o, key);
    if (!v) return std::nullopt;
    if (!v->is_number() || !v->num_is_u64) throw ParamsError(std::string("invalid type for `") + key + "`: expected an unsigned integer");
    return v->u64;
}
inline std::optional<bool> opt_bool(const vjson::Value& o, const char* key) {
    const vjson::Value* v = field(o, key);
    if (!v) return std::nullopt;
    if (!v->is_bool()) throw ParamsError(std::string("invalid type for `") + key + "`: expected a boolean");
    return v->b;
}
inline std::optional<std::vector<std::string>> opt_strings(const vjson::Value& o, const char* key) {
    const vjson::Value* v = field(o, key);
    if (!v) return std::nullopt;
    if (!v->is_array()) throw ParamsError(std::string("invalid type for `") + key + "`: expected a sequence");
    std::vector<std::string> out;
    for (auto& e : v->arr) {
        if (!e.is_string()) throw ParamsError(std::string("invalid type in `") + key + "`: expected a string");
        out.push_back(e.str);
    }
    return out;
}
inline std::optional<std::map<std::string, float>> opt_f32_map(const vjson::Value& o, const char* key) {
    const vjson::Value* v = field(o, key);
    if (!v) return std::nullopt;
    if (!v->is_object()) throw ParamsError(std::string("invalid type for `") + key + "`: expected a map");
    std::map<std::string, float> out;
    for (auto& kv : v->obj) {
        if (!kv.second.is_number()) throw ParamsError(std::string("invalid type in `") + key + "`: expected a number");
        out[kv.first] = (float)kv.second.num;
    }
    return out;
}
inline ParserOptions parser_options(const vjson::Value& o, const char* key) {
    ParserOptions p;
    const vjson::Value* v = field(o, key);
    if (!v) return p;
    if (!v->is_object()) throw ParamsError(std::string("invalid type for `") + key + "`: expected a map");
    p.no_attributes = opt_bool(*v, "no_attributes").value_or(false);
    p.no_parentheses = opt_bool(*v, "no_parentheses").value_or(false);
    p.no_levensthein = opt_bool(*v, "no_levensthein").value_or(false);
    return p;
}

// an f32 the way serde_json writes it (ryu's `format32`): the shortest digits that read back as the same f32, plain
// notation while the decimal point stays within (-6, 13] places of the first digit, "d.ddde±x" beyond
inline void write_f32(std::string& out, float f) {
    if (!std::isfinite(f)) return void(out += "null");
    if (f == 0) return void(out += std::signbit(f) ? "-0.0" : "0.0");
    char buf[32];
    for (int prec = 0; prec <= 8; ++prec) {
        snprintf(buf, sizeof buf, "%.*e", prec, (double)f);
        if (strtof(buf, nullptr) == f) break;
    }
    std::string digits;
    const char* p = buf;
    if (*p == '-') out += '-', ++p;
    for (; *p && *p != 'e'; ++p)
        if (*p != '.') digits += *p;
    const int exp10 = atoi(p + 1);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int len = (int)digits.size(), kk = exp10 + 1;  // value = 0.DIGITS * 10^kk
    if (kk >= len && kk <= 13) {
        out += digits, out.append((size_t)(kk - len), '0'), out += ".0";
    } else if (kk > 0 && kk <= 13) {
        out.append(digits, 0, (size_t)kk), out += '.', out.append(digits, (size_t)kk, std::string::npos);
    } else if (kk > -6 && kk <= 0) {
        out += "0.", out.append((size_t)(-kk), '0'), out += digits;
    } else {
        out += digits[0];
        if (len > 1) out += '.', out.append(digits, 1, std::string::npos);
        out += 'e', out += std::to_string(kk - 1);
    }
}

inline void write_usize_or_null(std::string& out, const std::optional<uint64_t>& v) { out += v ? std::to_string(*v) : std::string("null"); }

inline size_t count_scalars(const std::string& s) {
    size_t n = 0;
    for (unsigned char c : s) n += (c & 0xC0) != 0x80;
    return n;
}

// regex::escape: a backslash before every meta character of the regex crate's syntax
inline std::string regex_escape(const std::string& s) {
    std::string out;
    for (char c : s) {
        if (strchr("\\.+*?()|[]{}^$#&-~", c) && c) out += '\\';
        out += c;
    }
    return out;
}

}  // namespace detail

inline GeneratorParams params_from_json(const vjson::Value& o) {
    using namespace detail;
    if (!o.is_object()) throw ParamsError("SearchQueryGeneratorParameters: expected a map");
    GeneratorParams p;
    if (const vjson::Value* v = o.get("search_term")) {
        if (!v->is_string()) throw ParamsError("invalid type for `search_term`: expected a string");
        p.search_term = v->str;
    }
    p.parser_options = parser_options(o, "parser_options");
    p.filter_parser_options = parser_options(o, "filter_parser_options");
    p.top = opt_usize(o, "top"), p.skip = opt_usize(o, "skip");
    p.levenshtein = opt_usize(o, "levenshtein"), p.levenshtein_auto_limit = opt_usize(o, "levenshtein_auto_limit");
    p.facetlimit = opt_usize(o, "facetlimit");
    p.ignore_case = opt_bool(o, "ignore_case"), p.why_found = opt_bool(o, "why_found"), p.text_locality = opt_bool(o, "text_locality");
    p.phrase_pairs = opt_bool(o, "phrase_pairs"), p.explain = opt_bool(o, "explain");
    if (const vjson::Value* v = field(o, "boost_queries")) {
        if (!v->is_array()) throw ParamsError("invalid type for `boost_queries`: expected a sequence");
        p.boost_queries = v->arr;
    }
    p.facets = opt_strings(o, "facets"), p.fields = opt_strings(o, "fields");
    (void)opt_strings(o, "stopword_lists"), (void)opt_strings(o, "stopwords");
    p.boost_fields = opt_f32_map(o, "boost_fields"), p.boost_terms = opt_f32_map(o, "boost_terms");
    if (const vjson::Value* v = field(o, "filter")) {
        if (!v->is_string()) throw ParamsError("invalid type for `filter`: expected a string");
        p.filter = v->str;
    }
    return p;
}

// What the generator needs to know of an index (Persistence::metadata.get_all_fields, has_token_to_anchor:
// src/metadata.rs:28-30, src/persistence.rs:329-332).
struct FieldCatalog {
    std::vector<std::string> all_fields;     // every column, in key order
    std::vector<std::string> search_fields;  // the columns with a `<field>.textindex.to_anchor_id_score` index

    static FieldCatalog of(const vhost::Persistence& p) {
        FieldCatalog c;
        for (auto& kv : p.metadata.columns) {
            c.all_fields.push_back(kv.first);
            if (p.token_to_anchor_score.count(kv.first + ".textindex.to_anchor_id_score")) c.search_fields.push_back(kv.first);
        }
        return c;
    }
};

class RequestGenerator {
   public:
    explicit RequestGenerator(const FieldCatalog& cat) : cat_(cat) {}

    // search_query (src/query_generator.rs:175-257)
    std::string search_query(const GeneratorParams& opt) const {
        const std::vector<std::string> fields = search_field_names(opt.fields);
        Ast ast = parse(opt.search_term, opt.parser_options);
        std::string out = "{\"search_req\":";
        write_tree(out, lower(ast, fields, opt));
        if (opt.boost_queries) {
            out += ",\"boost\":[";
            for (size_t i = 0; i < opt.boost_queries->size(); ++i) {
                if (i) out += ',';
                write_boost_part(out, (*opt.boost_queries)[i]);
            }
            out += ']';
        }
        if (opt.boost_terms) {
            out += ",\"boost_term\":[";
            bool first = true;
            for (auto& kv : *opt.boost_terms) boost_term_parts(out, kv.first, kv.second, first);
            out += ']';
        }
        if (opt.facets) {
            out += ",\"facets\":[";
            for (size_t i = 0; i < opt.facets->size(); ++i) {
                check_field((*opt.facets)[i], cat_.all_fields);
                if (i) out += ',';
                out += "{\"field\":";
                vjson::write_string(out, (*opt.facets)[i]);
                out += ",\"top\":" + std::to_string(opt.facetlimit.value_or(5)) + "}";
            }
            out += ']';
        }
        const auto pairs = ast.phrase_pairs();
        if (opt.phrase_pairs.value_or(false) && !pairs.empty()) {
            out += ",\"phrase_boosts\":[";
            bool first = true;
            for (auto& pr : pairs)
                for (auto& f : fields) {
                    if (!first) out += ',';
                    first = false;
                    out += "{\"search1\":";
                    write_phrase_part(out, f, pr.first, opt);
                    out += ",\"search2\":";
                    write_phrase_part(out, f, pr.second, opt);
                    out += '}';
                }
            out += ']';
        }
        out += ",\"select\":null";  // (the generator never fills `select`; the key has no skip rule in the reference)
        if (opt.filter) {
            GeneratorParams fopt;
            fopt.levenshtein = 0;
            Ast fast = parse(*opt.filter, opt.filter_parser_options);
            out += ",\"filter\":";
            write_tree(out, lower(fast, cat_.all_fields, fopt));
        }
        if (opt.top) out += ",\"top\":" + std::to_string(*opt.top);
        if (opt.skip) out += ",\"skip\":" + std::to_string(*opt.skip);
        if (opt.why_found.value_or(false)) out += ",\"why_found\":true";
        if (opt.text_locality.value_or(false)) out += ",\"text_locality\":true";
        if (opt.explain.value_or(false)) out += ",\"explain\":true";
        return out + "}";
    }

    // suggest_query (src/query_generator.rs:297-322): one starts_with part per search field
    std::string suggest_query(const std::string& text, std::optional<uint64_t> top, std::optional<uint64_t> skip, std::optional<uint64_t> levenshtein,
                              const std::optional<std::vector<std::string>>& fields, std::optional<uint64_t> levenshtein_auto_limit) const {
        if (!top) top = 10;
        const uint64_t lev = levenshtein ? *levenshtein : default_levenshtein(text, levenshtein_auto_limit.value_or(1), true);
        std::string out = "{\"suggest\":[";
        bool first = true;
        for (auto& f : search_field_names(fields)) {
            if (!first) out += ',';
            first = false;
            out += "{\"path\":";
            vjson::write_string(out, f);
            out += ",\"terms\":[";
            vjson::write_string(out, text);
            out += "],\"levenshtein_distance\":" + std::to_string((uint32_t)lev) + ",\"starts_with\":true,\"top\":" + std::to_string(*top);
            if (skip) out += ",\"skip\":" + std::to_string(*skip);
            out += '}';
        }
        out += "],\"select\":null,\"top\":" + std::to_string(*top);
        if (skip) out += ",\"skip\":" + std::to_string(*skip);
        return out + "}";
    }

    // get_default_levenshtein / get_levenshteinn (src/query_generator.rs:85-99,129-132)
    static uint64_t default_levenshtein(const std::string& term, uint64_t auto_limit, bool wildcard) {
        const size_t n = detail::count_scalars(term);
        const size_t none_up_to = wildcard ? 3 : 2;
        if (n <= none_up_to) return 0;
        return std::min<uint64_t>(n <= 5 ? 1 : 2, auto_limit);
    }
    static uint32_t levenshtein_for(const std::string& term, const std::optional<uint64_t>& levenshtein, const std::optional<uint64_t>& auto_limit, bool wildcard) {
        const uint64_t lev = levenshtein ? *levenshtein : default_levenshtein(term, auto_limit.value_or(1), wildcard);
        // "at most the length minus one"; for an empty term the reference's `0usize - 1` wraps in a release build (and
        // panics in a debug build): the release value is kept
        const uint64_t cap = (uint64_t)detail::count_scalars(term) - 1;
        return (uint32_t)std::min(lev, cap);
    }

   private:
    // The search tree (src/search/request/search_request.rs:6-24) before it is written: operator nodes with children,
    // leaves as finished JSON text of a RequestSearchPart.
    struct Tree {
        enum Kind { Search, Or, And } kind = Search;
        std::string part;
        std::vector<Tree> queries;
    };

    // get_all_search_field_names (src/query_generator.rs:101-127)
    std::vector<std::string> search_field_names(const std::optional<std::vector<std::string>>& whitelist) const {
        std::vector<std::string> out;
        if (whitelist) {
            for (auto& f : cat_.all_fields)
                if (std::find(whitelist->begin(), whitelist->end(), f) != whitelist->end()) out.push_back(f);
        } else {
            out = cat_.search_fields;
        }
        if (out.empty()) {
            std::string msg = "All fields filtered all_fields: " + debug_list(cat_.all_fields) + " filter: ";
            msg += whitelist ? "Some(" + debug_list(*whitelist) + ")" : std::string("None");
            throw GeneratorError(msg);
        }
        return out;
    }
    static std::string debug_list(const std::vector<std::string>& v) {
        std::string out = "[";
        for (size_t i = 0; i < v.size(); ++i) {
            if (i) out += ", ";
            vjson::write_string(out, v[i]);
        }
        return out + "]";
    }
    static void check_field(const std::string& f, const std::vector<std::string>& all) {  // src/query_generator.rs:134-144
        if (std::find(all.begin(), all.end(), f) == all.end()) throw GeneratorError("Field " + f + " not found in " + debug_list(all));
    }

    // ast_to_search_request + SearchRequest::simplify (query_parser_to_veloci_request.rs:11-15, search_request.rs:27-76)
    Tree lower(const Ast& ast, const std::vector<std::string>& fields, const GeneratorParams& opt) const {
        if (fields.empty()) throw GeneratorError("All fields filtered all_fields: [] filter: None");  // (an index without columns: a panic in the reference)
        Tree t = lower_node(ast, ast.root, fields, opt, nullptr);
        simplify(t);
        return t;
    }

    // expand_fields_in_query_ast and query_ast_to_request in one walk (query_parser_to_veloci_request.rs:23-114): a term
    // outside any attribute becomes an OR over the fields (last field outermost), a term inside one a single part
    Tree lower_node(const Ast& ast, int32_t i, const std::vector<std::string>& fields, const GeneratorParams& opt, const std::string* attr) const {
        const Node& n = ast.nodes[i];
        Tree t;
        if (n.kind == Node::Binary) {
            t.kind = n.op == Operator::And ? Tree::And : Tree::Or;
            t.queries.push_back(lower_node(ast, n.left, fields, opt, attr));
            t.queries.push_back(lower_node(ast, n.right, fields, opt, attr));
        } else if (n.kind == Node::Attributed) {
            if (!attr) check_field(n.text, fields);  // only the outermost attribute is checked, as the reference
            return lower_node(ast, n.left, fields, opt, &n.text);
        } else if (attr) {
            t.part = search_part(n, *attr, opt);
        } else {
            t.part = search_part(n, fields[0], opt);
            for (size_t f = 1; f < fields.size(); ++f) {
                Tree wrap;
                wrap.kind = Tree::Or;
                Tree leaf;
                leaf.part = search_part(n, fields[f], opt);
                wrap.queries.push_back(std::move(leaf));
                wrap.queries.push_back(std::move(t));
                t = std::move(wrap);
            }
        }
        return t;
    }

    // the Leaf arm of query_ast_to_request (query_parser_to_veloci_request.rs:40-80): "term*" = prefix search that may
    // still be fuzzy; any other '*' = a regex with ".*" for each star and no edit distance
    std::string search_part(const Node& leaf, const std::string& field, const GeneratorParams& opt) const {
        std::string term = leaf.text;
        const bool starts_with = !term.empty() && term.back() == '*' && std::count(term.begin(), term.end(), '*') == 1;
        if (starts_with) term.pop_back();
        const bool is_regex = term.find('*') != std::string::npos;
        std::optional<uint32_t> lev;
        if (is_regex) {
            std::string pattern;
            size_t from = 0;
            while (true) {
                const size_t star = term.find('*', from);
                pattern += detail::regex_escape(term.substr(from, star == std::string::npos ? std::string::npos : star - from));
                if (star == std::string::npos) break;
                pattern += ".*";
                from = star + 1;
            }
            term = pattern;
        } else {
            lev = leaf.levenshtein >= 0 ? (uint32_t)leaf.levenshtein : levenshtein_for(term, opt.levenshtein, opt.levenshtein_auto_limit, starts_with);
        }
        std::string out = "{\"path\":";
        vjson::write_string(out, field);
        out += ",\"terms\":[";
        vjson::write_string(out, term);
        out += ']';
        if (lev) out += ",\"levenshtein_distance\":" + std::to_string(*lev);
        if (starts_with) out += ",\"starts_with\":true";
        if (is_regex) out += ",\"is_regex\":true";
        write_field_boost(out, field, opt);
        if (opt.ignore_case) out += *opt.ignore_case ? ",\"ignore_case\":true" : ",\"ignore_case\":false";
        return out + "}";
    }
    static void write_field_boost(std::string& out, const std::string& field, const GeneratorParams& opt) {
        if (!opt.boost_fields) return;
        auto it = opt.boost_fields->find(field);
        if (it == opt.boost_fields->end()) return;
        out += ",\"boost\":";
        detail::write_f32(out, it->second);
    }

    // generate_phrase_queries_for_searchterm (src/query_generator.rs:268-295)
    void write_phrase_part(std::string& out, const std::string& field, const std::string& term, const GeneratorParams& opt) const {
        out += "{\"path\":";
        vjson::write_string(out, field);
        out += ",\"terms\":[";
        vjson::write_string(out, term);
        out += "],\"levenshtein_distance\":" + std::to_string(levenshtein_for(term, opt.levenshtein, opt.levenshtein_auto_limit, false));
        write_field_boost(out, field, opt);
        out += '}';
    }

    // handle_boost_term_query (src/query_generator.rs:146-168): "field:term" restricts the boost to that field; the text
    // between the first and the second ':' is the term, every other piece a field name
    void boost_term_parts(std::string& out, const std::string& spec, float boost, bool& first) const {
        std::string term = spec;
        std::optional<std::vector<std::string>> only;
        if (spec.find(':') != std::string::npos) {
            std::vector<std::string> pieces;
            size_t from = 0;
            while (true) {
                const size_t c = spec.find(':', from);
                pieces.push_back(spec.substr(from, c == std::string::npos ? std::string::npos : c - from));
                if (c == std::string::npos) break;
                from = c + 1;
            }
            term = pieces[1];
            pieces.erase(pieces.begin() + 1);
            only = std::move(pieces);
        }
        for (auto& f : search_field_names(only)) {
            if (!first) out += ',';
            first = false;
            out += "{\"path\":";
            vjson::write_string(out, f);
            out += ",\"terms\":[";
            vjson::write_string(out, term);
            out += "],\"boost\":";
            detail::write_f32(out, boost);
            out += '}';
        }
    }

    // RequestBoostPart (src/search/request/boost_request.rs:3-21): every key written, null when absent
    static void write_boost_part(std::string& out, const vjson::Value& b) {
        if (!b.is_object()) throw ParamsError("invalid type in `boost_queries`: expected a map");
        const vjson::Value* path = b.get("path");
        if (!path || !path->is_string()) throw ParamsError("`boost_queries`: missing field `path`");
        out += "{\"path\":";
        vjson::write_string(out, path->str);
        out += ",\"boost_fun\":";
        if (const vjson::Value* f = detail::field(b, "boost_fun")) {
            static const char* funs[] = {"Log2", "Log10", "Multiply", "Add", "Replace"};
            const bool known = f->is_string() && std::any_of(std::begin(funs), std::end(funs), [&](const char* n) { return f->str == n; });
            if (!known) throw ParamsError("`boost_queries`: unknown variant of `boost_fun`");
            vjson::write_string(out, f->str);
        } else {
            out += "null";
        }
        out += ",\"param\":";
        if (const vjson::Value* p = detail::field(b, "param")) {
            if (!p->is_number()) throw ParamsError("`boost_queries`: invalid type for `param`");
            detail::write_f32(out, (float)p->num);
        } else {
            out += "null";
        }
        out += ",\"skip_when_score\":";
        if (const vjson::Value* s = detail::field(b, "skip_when_score")) {
            if (!s->is_array()) throw ParamsError("`boost_queries`: invalid type for `skip_when_score`");
            out += '[';
            for (size_t i = 0; i < s->arr.size(); ++i) {
                if (!s->arr[i].is_number()) throw ParamsError("`boost_queries`: invalid type in `skip_when_score`");
                if (i) out += ',';
                detail::write_f32(out, (float)s->arr[i].num);
            }
            out += ']';
        } else {
            out += "null";
        }
        out += ",\"expression\":";
        if (const vjson::Value* e = detail::field(b, "expression")) {
            if (!e->is_string()) throw ParamsError("`boost_queries`: invalid type for `expression`");
            vjson::write_string(out, e->str);
        } else {
            out += "null";
        }
        out += '}';
    }

    // SearchRequest::simplify: children first; then the operator's own children of the same kind are taken out back to
    // front and their children appended after the ones that stay (so a right-leaning chain a,(b,(c,d)) comes out a,b,c,d)
    static void simplify(Tree& t) {
        if (t.kind == Tree::Search) return;
        for (auto& q : t.queries) simplify(q);
        std::vector<Tree> lifted;
        for (size_t i = t.queries.size(); i-- > 0;) {
            if (t.queries[i].kind != t.kind) continue;
            Tree sub = std::move(t.queries[i]);
            t.queries.erase(t.queries.begin() + (long)i);
            for (auto& q : sub.queries) lifted.push_back(std::move(q));
        }
        for (auto& q : lifted) t.queries.push_back(std::move(q));
    }

    static void write_tree(std::string& out, const Tree& t) {
        if (t.kind == Tree::Search) {
            out += "{\"search\":" + t.part + "}";
            return;
        }
        out += t.kind == Tree::Or ? "{\"or\":{\"queries\":[" : "{\"and\":{\"queries\":[";
        for (size_t i = 0; i < t.queries.size(); ++i) {
            if (i) out += ',';
            write_tree(out, t.queries[i]);
        }
        out += "]}}";
    }

    const FieldCatalog& cat_;
};

// JSON in, JSON out: the two entry points as the C ABI exposes them.
inline std::string search_query_json(const FieldCatalog& cat, const char* params_json, size_t len) {
    return RequestGenerator(cat).search_query(params_from_json(vjson::parse(params_json, len)));
}
// `params_json`: {"request": "...", "top", "skip", "levenshtein", "fields", "levenshtein_auto_limit"} = suggest_query's arguments
inline std::string suggest_query_json(const FieldCatalog& cat, const char* params_json, size_t len) {
    const vjson::Value o = vjson::parse(params_json, len);
    if (!o.is_object()) throw ParamsError("suggest_query: expected a map");
    const vjson::Value* text = o.get("request");
    if (!text || !text->is_string()) throw ParamsError("suggest_query: missing field `request`");
    using namespace detail;
    return RequestGenerator(cat).suggest_query(text->str, opt_usize(o, "top"), opt_usize(o, "skip"), opt_usize(o, "levenshtein"), opt_strings(o, "fields"),
                                               opt_usize(o, "levenshtein_auto_limit"));
}

}  // namespace vquery
