// The `search::Request` JSON surface, kept verbatim.
//
// Mirrors src/search/request/mod.rs:14-87 (Request, RequestPhraseBoost),
// search_request.rs:6-201 (SearchRequest or/and/search tree, SearchRequestOptions,
// RequestSearchPart, simplify), boost_request.rs:3-33 (RequestBoostPart,
// BoostFunction) and facet_request.rs:1-11 (FacetRequest, default top 10).
// serde semantics kept: unknown keys are ignored, `top` defaults to Some(10)
// when the key is absent and to None when it is null.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../vjson.hpp"

namespace vhost {

struct RequestError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

enum class BoostFun : int { None = 0, Log2 = 1, Log10 = 2, Multiply = 3, Add = 4, Replace = 5 };

struct BoostPart {
    std::string path;
    BoostFun boost_fun = BoostFun::None;
    std::optional<float> param;
    std::optional<std::vector<float>> skip_when_score;
    std::optional<std::string> expression;
    std::string key() const {
        std::string k = path + "\x1f" + std::to_string((int)boost_fun) + "\x1f";
        if (param) k += std::to_string(*param);
        k += "\x1f";
        if (skip_when_score)
            for (float f : *skip_when_score) k += std::to_string(f) + ",";
        k += "\x1f";
        if (expression) k += *expression;
        return k;
    }
};

struct SearchOptions {
    bool present = false;
    bool explain = false;
    std::optional<uint64_t> top, skip;
    std::optional<std::vector<BoostPart>> boost;
};

struct SearchPart {
    std::string path;
    std::vector<std::string> terms;
    std::optional<uint32_t> levenshtein_distance;
    bool starts_with = false;
    bool is_regex = false;
    std::optional<BoostPart> token_value;
    std::optional<float> boost;
    std::optional<bool> ignore_case;
    std::optional<uint64_t> top, skip;
    SearchOptions options;

    // RequestSearchPart derives Hash/Eq over every field: the FieldRequestCache key
    // (execution_plan.rs:13,108-130).
    std::string key() const {
        std::string k;
        k.reserve(path.size() + 24 + (terms.empty() ? 0 : terms[0].size()));
        k += path;
        for (auto& t : terms) k += '\x1e', k += t;
        k += '\x1f';
        if (levenshtein_distance) k += std::to_string(*levenshtein_distance);
        k += starts_with ? "\x1fS" : "\x1fs";
        k += is_regex ? 'R' : 'r';
        k += '\x1f';
        if (token_value) k += token_value->key();
        k += '\x1f';
        if (boost) k += std::to_string(*boost);
        k += '\x1f';
        if (ignore_case) k += *ignore_case ? '1' : '0';
        k += '\x1f';
        if (top) k += std::to_string(*top);
        k += '\x1f';
        if (skip) k += std::to_string(*skip);
        k += '\x1f';
        if (options.present) {
            k += options.explain ? 'E' : 'e';
            if (options.top) k += std::to_string(*options.top);
            k += ',';
            if (options.skip) k += std::to_string(*options.skip);
            if (options.boost)
                for (auto& b : *options.boost) k += '|', k += b.key();
        }
        return k;
    }
};

struct SearchRequest {
    enum Kind { Or, And, Search } kind = Search;
    std::vector<SearchRequest> queries;  // Or / And
    SearchOptions options;               // Or / And
    SearchPart part;                     // Search

    const std::optional<std::vector<BoostPart>>& get_boost() const { return kind == Search ? part.options.boost : options.boost; }

    // search_request.rs:26-72
    void simplify() {
        if (kind == Search) return;
        for (auto& q : queries) q.simplify();
        std::vector<SearchRequest> pulled;
        for (size_t i = queries.size(); i-- > 0;) {
            if (queries[i].kind == kind && !queries[i].options.present) {
                SearchRequest sub = std::move(queries[i]);
                queries.erase(queries.begin() + (long)i);
                for (auto& q : sub.queries) pulled.push_back(std::move(q));
            }
        }
        for (auto& q : pulled) queries.push_back(std::move(q));
    }
};

struct PhraseBoost {
    SearchPart search1, search2;
};

struct FacetRequest {
    std::string field;
    std::optional<uint64_t> top = 10;
};

struct Request {
    std::optional<SearchRequest> search_req;
    std::optional<std::vector<SearchPart>> suggest;  // or/and/search and suggest are mutually exclusive (request/mod.rs:21-22)
    std::optional<std::vector<BoostPart>> boost;
    std::optional<std::vector<SearchPart>> boost_term;
    std::optional<std::vector<FacetRequest>> facets;
    std::optional<std::vector<PhraseBoost>> phrase_boosts;
    std::optional<std::vector<std::string>> select;
    std::shared_ptr<SearchRequest> filter;
    std::optional<uint64_t> top = 10;
    std::optional<uint64_t> skip;
    bool why_found = false;
    bool text_locality = false;
    bool explain = false;
};

// RequestSearchPart::is_explain (search_request.rs:182-184), and whether a request asks for explanations anywhere: its own
// `explain` (merged into every part, execution_plan.rs:46-90) or a part's own options.
inline bool part_explains(const SearchPart& p) { return p.options.present && p.options.explain; }
inline bool tree_explains(const SearchRequest& r) {
    if (r.kind == SearchRequest::Search) return part_explains(r.part);
    for (auto& q : r.queries)
        if (tree_explains(q)) return true;
    return false;
}
inline bool request_explains(const Request& r) { return r.explain || (r.search_req && tree_explains(*r.search_req)); }

// ------------------------------------------------------------- parsing ------
namespace detail {
inline const vjson::Value* field(const vjson::Value& o, const char* k) {
    const vjson::Value* v = o.get(k);
    return (v && !v->is_null()) ? v : nullptr;
}
inline uint64_t as_u64(const vjson::Value& v, const char* what) {
    if (!v.is_number() || !v.num_is_u64) throw RequestError(std::string("invalid type for ") + what + ": expected unsigned integer");
    return v.u64;
}
inline float as_f32(const vjson::Value& v, const char* what) {
    if (!v.is_number()) throw RequestError(std::string("invalid type for ") + what + ": expected number");
    return (float)v.num;
}
inline bool as_bool(const vjson::Value& v, const char* what) {
    if (!v.is_bool()) throw RequestError(std::string("invalid type for ") + what + ": expected bool");
    return v.b;
}
inline std::string as_str(const vjson::Value& v, const char* what) {
    if (!v.is_string()) throw RequestError(std::string("invalid type for ") + what + ": expected string");
    return v.str;
}
}  // namespace detail

inline BoostPart parse_boost_part(const vjson::Value& v) {
    using namespace detail;
    if (!v.is_object()) throw RequestError("boost part must be an object");
    BoostPart b;
    const vjson::Value* p = v.get("path");
    if (!p) throw RequestError("missing field `path`");
    b.path = as_str(*p, "path");
    if (auto* f = field(v, "boost_fun")) {
        std::string s = as_str(*f, "boost_fun");
        if (s == "Log2") b.boost_fun = BoostFun::Log2;
        else if (s == "Log10") b.boost_fun = BoostFun::Log10;
        else if (s == "Multiply") b.boost_fun = BoostFun::Multiply;
        else if (s == "Add") b.boost_fun = BoostFun::Add;
        else if (s == "Replace") b.boost_fun = BoostFun::Replace;
        else throw RequestError("unknown variant `" + s + "`, expected one of `Log2`, `Log10`, `Multiply`, `Add`, `Replace`");
    }
    if (auto* f = field(v, "param")) b.param = as_f32(*f, "param");
    if (auto* f = field(v, "skip_when_score")) {
        if (!f->is_array()) throw RequestError("skip_when_score must be an array");
        std::vector<float> s;
        for (auto& e : f->arr) s.push_back(as_f32(e, "skip_when_score"));
        b.skip_when_score = std::move(s);
    }
    if (auto* f = field(v, "expression")) b.expression = as_str(*f, "expression");
    return b;
}

inline SearchOptions parse_options(const vjson::Value& v) {
    using namespace detail;
    SearchOptions o;
    if (!v.is_object()) throw RequestError("options must be an object");
    o.present = true;
    if (auto* f = field(v, "explain")) o.explain = as_bool(*f, "explain");
    if (auto* f = field(v, "top")) o.top = as_u64(*f, "top");
    if (auto* f = field(v, "skip")) o.skip = as_u64(*f, "skip");
    if (auto* f = field(v, "boost")) {
        if (!f->is_array()) throw RequestError("options.boost must be an array");
        std::vector<BoostPart> bs;
        for (auto& e : f->arr) bs.push_back(parse_boost_part(e));
        o.boost = std::move(bs);
    }
    return o;
}

inline SearchPart parse_search_part(const vjson::Value& v) {
    using namespace detail;
    if (!v.is_object()) throw RequestError("search part must be an object");
    SearchPart s;
    const vjson::Value* p = v.get("path");
    if (!p) throw RequestError("missing field `path`");
    s.path = as_str(*p, "path");
    const vjson::Value* t = v.get("terms");
    if (!t) throw RequestError("missing field `terms`");
    if (!t->is_array()) throw RequestError("terms must be an array");
    for (auto& e : t->arr) s.terms.push_back(as_str(e, "terms"));
    if (auto* f = field(v, "levenshtein_distance")) s.levenshtein_distance = (uint32_t)as_u64(*f, "levenshtein_distance");
    if (auto* f = field(v, "starts_with")) s.starts_with = as_bool(*f, "starts_with");
    if (auto* f = field(v, "is_regex")) s.is_regex = as_bool(*f, "is_regex");
    if (auto* f = field(v, "token_value")) s.token_value = parse_boost_part(*f);
    if (auto* f = field(v, "boost")) s.boost = as_f32(*f, "boost");
    if (auto* f = field(v, "ignore_case")) s.ignore_case = as_bool(*f, "ignore_case");
    if (auto* f = field(v, "top")) s.top = as_u64(*f, "top");
    if (auto* f = field(v, "skip")) s.skip = as_u64(*f, "skip");
    if (auto* f = field(v, "options")) s.options = parse_options(*f);
    return s;
}

inline SearchRequest parse_search_request(const vjson::Value& v) {
    using namespace detail;
    if (!v.is_object() || v.obj.size() != 1) throw RequestError("search request must be an object with exactly one of `or`, `and`, `search`");
    const std::string& tag = v.obj[0].first;
    const vjson::Value& body = v.obj[0].second;
    SearchRequest r;
    if (tag == "search") {
        r.kind = SearchRequest::Search;
        r.part = parse_search_part(body);
        return r;
    }
    if (tag == "or") r.kind = SearchRequest::Or;
    else if (tag == "and") r.kind = SearchRequest::And;
    else throw RequestError("unknown variant `" + tag + "`, expected one of `or`, `and`, `search`");
    if (!body.is_object()) throw RequestError("search tree must be an object");
    const vjson::Value* q = body.get("queries");
    if (!q || !q->is_array()) throw RequestError("missing field `queries`");
    for (auto& e : q->arr) r.queries.push_back(parse_search_request(e));
    if (auto* f = field(body, "options")) r.options = parse_options(*f);
    return r;
}

inline Request parse_request(const vjson::Value& v) {
    using namespace detail;
    if (!v.is_object()) throw RequestError("request must be a JSON object");
    Request r;
    if (auto* f = field(v, "search_req")) r.search_req = parse_search_request(*f);
    if (auto* f = field(v, "suggest")) {
        if (!f->is_array()) throw RequestError("suggest must be an array");
        std::vector<SearchPart> ps;
        for (auto& e : f->arr) ps.push_back(parse_search_part(e));
        r.suggest = std::move(ps);
    }
    if (auto* f = field(v, "boost")) {
        if (!f->is_array()) throw RequestError("boost must be an array");
        std::vector<BoostPart> bs;
        for (auto& e : f->arr) bs.push_back(parse_boost_part(e));
        r.boost = std::move(bs);
    }
    if (auto* f = field(v, "boost_term")) {
        if (!f->is_array()) throw RequestError("boost_term must be an array");
        std::vector<SearchPart> ps;
        for (auto& e : f->arr) ps.push_back(parse_search_part(e));
        r.boost_term = std::move(ps);
    }
    if (auto* f = field(v, "facets")) {
        if (!f->is_array()) throw RequestError("facets must be an array");
        std::vector<FacetRequest> fs;
        for (auto& e : f->arr) {
            if (!e.is_object()) throw RequestError("facet must be an object");
            FacetRequest fr;
            const vjson::Value* fld = e.get("field");
            if (!fld) throw RequestError("missing field `field`");
            fr.field = as_str(*fld, "field");
            if (const vjson::Value* t = e.get("top")) {
                if (t->is_null()) fr.top.reset();
                else fr.top = as_u64(*t, "top");
            }
            fs.push_back(std::move(fr));
        }
        r.facets = std::move(fs);
    }
    if (auto* f = field(v, "phrase_boosts")) {
        if (!f->is_array()) throw RequestError("phrase_boosts must be an array");
        std::vector<PhraseBoost> ps;
        for (auto& e : f->arr) {
            const vjson::Value* s1 = e.get("search1");
            const vjson::Value* s2 = e.get("search2");
            if (!s1 || !s2) throw RequestError("phrase boost needs `search1` and `search2`");
            ps.push_back(PhraseBoost{parse_search_part(*s1), parse_search_part(*s2)});
        }
        r.phrase_boosts = std::move(ps);
    }
    if (auto* f = field(v, "select")) {
        if (!f->is_array()) throw RequestError("select must be an array");
        std::vector<std::string> s;
        for (auto& e : f->arr) s.push_back(as_str(e, "select"));
        r.select = std::move(s);
    }
    if (auto* f = field(v, "filter")) r.filter = std::make_shared<SearchRequest>(parse_search_request(*f));
    if (const vjson::Value* t = v.get("top")) {
        if (t->is_null()) r.top.reset();
        else r.top = as_u64(*t, "top");
    }
    if (auto* f = field(v, "skip")) r.skip = as_u64(*f, "skip");
    if (auto* f = field(v, "why_found")) r.why_found = as_bool(*f, "why_found");
    if (auto* f = field(v, "text_locality")) r.text_locality = as_bool(*f, "text_locality");
    if (auto* f = field(v, "explain")) r.explain = as_bool(*f, "explain");
    return r;
}

// One pass over the request text straight into the Request structs, without a DOM: the planner's way in (a batch is
// thousands of requests; the DOM's allocations were the larger part of planning a request).  Same surface and serde
// semantics as parse_request: unknown keys are skipped, `null` leaves an Option field absent, a repeated key takes its
// last value, the messages of type errors are the same.  tests/test_request_reader.py holds the two against each other.
class RequestReader : private vjson::Parser {
  public:
    RequestReader(const char* p, size_t n) : Parser(p, n) {}

    Request read() {
        Request r;
        try {
            skip_ws();
            read_request(r);
            skip_ws();
            if (p_ != end_) fail("trailing characters");
        } catch (const vjson::ParseError& e) {
            throw RequestError(e.what());
        }
        return r;
    }

  private:
    std::string key_;
    vjson::Value num_;

    template <size_t N>
    static bool is(const std::string& key, const char (&name)[N]) {
        return key.size() == N - 1 && memcmp(key.data(), name, N - 1) == 0;
    }
    char peek() {
        skip_ws();
        if (p_ == end_) fail("unexpected end");
        return *p_;
    }
    bool null_next() {
        if (peek() != 'n') return false;
        expect("null");
        return true;
    }
    void skip_value() {
        vjson::Value ignored;
        parse_value_into(ignored);
    }
    // Walks the members of the object at the cursor: on_key(key) consumes the member's value.
    template <class F>
    void members(const char* not_object, F&& on_key) {
        if (peek() != '{') throw RequestError(not_object);
        ++p_;
        if (peek() == '}') {
            ++p_;
            return;
        }
        while (true) {
            if (peek() != '"') fail("expected object key");
            parse_string(key_);
            if (peek() != ':') fail("expected :");
            ++p_;
            skip_ws();
            on_key(key_);
            const char c = peek();
            ++p_;
            if (c == ',') continue;
            if (c == '}') return;
            fail("expected , or }");
        }
    }
    // Walks the elements of the array at the cursor: on_element() consumes one element.
    template <class F>
    void elements(const char* not_array, F&& on_element) {
        if (peek() != '[') throw RequestError(not_array);
        ++p_;
        if (peek() == ']') {
            ++p_;
            return;
        }
        while (true) {
            skip_ws();
            on_element();
            const char c = peek();
            ++p_;
            if (c == ',') continue;
            if (c == ']') return;
            fail("expected , or ]");
        }
    }
    [[noreturn]] static void bad_type(const char* what, const char* expected) { throw RequestError(std::string("invalid type for ") + what + ": expected " + expected); }
    void number(const char* what, const char* expected) {
        const char c = peek();
        if (c != '-' && (c < '0' || c > '9')) {
            skip_value();  // a syntax error is reported as such
            bad_type(what, expected);
        }
        num_.num_is_u64 = num_.num_is_i64 = false;
        parse_number(num_);
    }
    uint64_t u64(const char* what) {
        number(what, "unsigned integer");
        if (!num_.num_is_u64) bad_type(what, "unsigned integer");
        return num_.u64;
    }
    float f32(const char* what) {
        number(what, "number");
        return (float)num_.num;
    }
    bool boolean(const char* what) {
        const char c = peek();
        if (c == 't') {
            expect("true");
            return true;
        }
        if (c == 'f') {
            expect("false");
            return false;
        }
        skip_value();
        bad_type(what, "bool");
    }
    void str(const char* what, std::string& out) {
        if (peek() != '"') {
            skip_value();
            bad_type(what, "string");
        }
        parse_string(out);
    }
    template <class T, class F>
    void optional_field(std::optional<T>& field, F&& read) {  // null = None
        if (null_next()) field.reset();
        else field = read();
    }

    // (the read_* functions fill a freshly constructed object)
    void read_boost_part(BoostPart& b) {
        bool have_path = false;
        std::string fun;
        members("boost part must be an object", [&](const std::string& k) {
            if (is(k, "path")) {
                str("path", b.path), have_path = true;
            } else if (is(k, "boost_fun")) {
                if (null_next()) {
                    b.boost_fun = BoostFun::None;
                    return;
                }
                str("boost_fun", fun);
                if (fun == "Log2") b.boost_fun = BoostFun::Log2;
                else if (fun == "Log10") b.boost_fun = BoostFun::Log10;
                else if (fun == "Multiply") b.boost_fun = BoostFun::Multiply;
                else if (fun == "Add") b.boost_fun = BoostFun::Add;
                else if (fun == "Replace") b.boost_fun = BoostFun::Replace;
                else throw RequestError("unknown variant `" + fun + "`, expected one of `Log2`, `Log10`, `Multiply`, `Add`, `Replace`");
            } else if (is(k, "param")) {
                optional_field(b.param, [&] { return f32("param"); });
            } else if (is(k, "skip_when_score")) {
                optional_field(b.skip_when_score, [&] {
                    std::vector<float> s;
                    elements("skip_when_score must be an array", [&] { s.push_back(f32("skip_when_score")); });
                    return s;
                });
            } else if (is(k, "expression")) {
                optional_field(b.expression, [&] {
                    std::string e;
                    str("expression", e);
                    return e;
                });
            } else {
                skip_value();
            }
        });
        if (!have_path) throw RequestError("missing field `path`");
    }
    std::vector<BoostPart> read_boost_list(const char* not_array) {
        std::vector<BoostPart> bs;
        elements(not_array, [&] {
            bs.emplace_back();
            read_boost_part(bs.back());
        });
        return bs;
    }
    void read_options(SearchOptions& o) {
        o.present = true;
        members("options must be an object", [&](const std::string& k) {
            if (is(k, "explain")) {
                o.explain = null_next() ? false : boolean("explain");
            } else if (is(k, "top")) {
                optional_field(o.top, [&] { return u64("top"); });
            } else if (is(k, "skip")) {
                optional_field(o.skip, [&] { return u64("skip"); });
            } else if (is(k, "boost")) {
                optional_field(o.boost, [&] { return read_boost_list("options.boost must be an array"); });
            } else {
                skip_value();
            }
        });
    }
    void read_search_part(SearchPart& s) {
        bool have_path = false, have_terms = false;
        members("search part must be an object", [&](const std::string& k) {
            if (is(k, "path")) {
                str("path", s.path), have_path = true;
            } else if (is(k, "terms")) {
                s.terms.clear();
                elements("terms must be an array", [&] {
                    s.terms.emplace_back();
                    str("terms", s.terms.back());
                });
                have_terms = true;
            } else if (is(k, "levenshtein_distance")) {
                optional_field(s.levenshtein_distance, [&] { return (uint32_t)u64("levenshtein_distance"); });
            } else if (is(k, "starts_with")) {
                s.starts_with = null_next() ? false : boolean("starts_with");
            } else if (is(k, "is_regex")) {
                s.is_regex = null_next() ? false : boolean("is_regex");
            } else if (is(k, "token_value")) {
                optional_field(s.token_value, [&] {
                    BoostPart b;
                    read_boost_part(b);
                    return b;
                });
            } else if (is(k, "boost")) {
                optional_field(s.boost, [&] { return f32("boost"); });
            } else if (is(k, "ignore_case")) {
                optional_field(s.ignore_case, [&] { return boolean("ignore_case"); });
            } else if (is(k, "top")) {
                optional_field(s.top, [&] { return u64("top"); });
            } else if (is(k, "skip")) {
                optional_field(s.skip, [&] { return u64("skip"); });
            } else if (is(k, "options")) {
                s.options = SearchOptions();
                if (!null_next()) read_options(s.options);
            } else {
                skip_value();
            }
        });
        if (!have_path) throw RequestError("missing field `path`");
        if (!have_terms) throw RequestError("missing field `terms`");
    }
    std::vector<SearchPart> read_part_list(const char* not_array) {
        std::vector<SearchPart> ps;
        elements(not_array, [&] {
            ps.emplace_back();
            read_search_part(ps.back());
        });
        return ps;
    }
    void read_search_request(SearchRequest& r) {
        static const char* const shape = "search request must be an object with exactly one of `or`, `and`, `search`";
        uint32_t n_keys = 0;
        members(shape, [&](const std::string& tag) {
            if (++n_keys > 1) throw RequestError(shape);
            if (is(tag, "search")) {
                r.kind = SearchRequest::Search;
                read_search_part(r.part);
                return;
            }
            if (is(tag, "or")) r.kind = SearchRequest::Or;
            else if (is(tag, "and")) r.kind = SearchRequest::And;
            else throw RequestError("unknown variant `" + tag + "`, expected one of `or`, `and`, `search`");
            bool have_queries = false;
            members("search tree must be an object", [&](const std::string& k) {
                if (is(k, "queries")) {
                    r.queries.clear();
                    r.queries.reserve(4);
                    elements("missing field `queries`", [&] {
                        r.queries.emplace_back();
                        read_search_request(r.queries.back());
                    });
                    have_queries = true;
                } else if (is(k, "options")) {
                    r.options = SearchOptions();
                    if (!null_next()) read_options(r.options);
                } else {
                    skip_value();
                }
            });
            if (!have_queries) throw RequestError("missing field `queries`");
        });
        if (n_keys != 1) throw RequestError(shape);
    }
    void read_request(Request& r) {
        members("request must be a JSON object", [&](const std::string& k) {
            if (is(k, "search_req")) {
                optional_field(r.search_req, [&] {
                    SearchRequest s;
                    read_search_request(s);
                    return s;
                });
            } else if (is(k, "boost")) {
                optional_field(r.boost, [&] { return read_boost_list("boost must be an array"); });
            } else if (is(k, "suggest")) {
                optional_field(r.suggest, [&] { return read_part_list("suggest must be an array"); });
            } else if (is(k, "boost_term")) {
                optional_field(r.boost_term, [&] { return read_part_list("boost_term must be an array"); });
            } else if (is(k, "facets")) {
                optional_field(r.facets, [&] {
                    std::vector<FacetRequest> fs;
                    elements("facets must be an array", [&] {
                        FacetRequest fr;
                        bool have_field = false;
                        members("facet must be an object", [&](const std::string& fk) {
                            if (is(fk, "field")) str("field", fr.field), have_field = true;
                            else if (is(fk, "top")) optional_field(fr.top, [&] { return u64("top"); });
                            else skip_value();
                        });
                        if (!have_field) throw RequestError("missing field `field`");
                        fs.push_back(std::move(fr));
                    });
                    return fs;
                });
            } else if (is(k, "phrase_boosts")) {
                optional_field(r.phrase_boosts, [&] {
                    std::vector<PhraseBoost> ps;
                    elements("phrase_boosts must be an array", [&] {
                        PhraseBoost pb;
                        bool have1 = false, have2 = false;
                        members("phrase boost needs `search1` and `search2`", [&](const std::string& pk) {
                            if (is(pk, "search1")) pb.search1 = SearchPart(), read_search_part(pb.search1), have1 = true;
                            else if (is(pk, "search2")) pb.search2 = SearchPart(), read_search_part(pb.search2), have2 = true;
                            else skip_value();
                        });
                        if (!have1 || !have2) throw RequestError("phrase boost needs `search1` and `search2`");
                        ps.push_back(std::move(pb));
                    });
                    return ps;
                });
            } else if (is(k, "select")) {
                optional_field(r.select, [&] {
                    std::vector<std::string> s;
                    elements("select must be an array", [&] {
                        s.emplace_back();
                        str("select", s.back());
                    });
                    return s;
                });
            } else if (is(k, "filter")) {
                if (null_next()) {
                    r.filter.reset();
                } else {
                    auto f = std::make_shared<SearchRequest>();
                    read_search_request(*f);
                    r.filter = std::move(f);
                }
            } else if (is(k, "top")) {
                optional_field(r.top, [&] { return u64("top"); });  // absent: Some(10); null: None
            } else if (is(k, "skip")) {
                optional_field(r.skip, [&] { return u64("skip"); });
            } else if (is(k, "why_found")) {
                r.why_found = null_next() ? false : boolean("why_found");
            } else if (is(k, "text_locality")) {
                r.text_locality = null_next() ? false : boolean("text_locality");
            } else if (is(k, "explain")) {
                r.explain = null_next() ? false : boolean("explain");
            } else {
                skip_value();
            }
        });
    }
};

inline Request read_request_json(const char* json, size_t len) { return RequestReader(json, len).read(); }

// Canonical text of a parsed request: every field, in declaration order (the tests compare the two parsers with it).
inline void describe(const BoostPart& b, std::string& out) { out += "B(" + b.key() + ")"; }
inline void describe(const SearchOptions& o, std::string& out) {
    if (!o.present) return;
    out += "O(";
    out += o.explain ? "E" : "e";
    out += o.top ? std::to_string(*o.top) : "-";
    out += ",";
    out += o.skip ? std::to_string(*o.skip) : "-";
    if (o.boost) {
        out += "[";
        for (auto& b : *o.boost) describe(b, out);
        out += "]";
    }
    out += ")";
}
inline void describe(const SearchPart& p, std::string& out) { out += "P(" + p.key() + ")"; }
inline void describe(const SearchRequest& r, std::string& out) {
    if (r.kind == SearchRequest::Search) {
        describe(r.part, out);
        return;
    }
    out += r.kind == SearchRequest::Or ? "or{" : "and{";
    for (auto& q : r.queries) describe(q, out), out += ";";
    describe(r.options, out);
    out += "}";
}
inline std::string describe(const Request& r) {
    std::string out;
    auto u = [&](const char* name, const std::optional<uint64_t>& v) { out += name, out += v ? "=" + std::to_string(*v) : "=None", out += " "; };
    if (r.search_req) out += "search_req=", describe(*r.search_req, out), out += " ";
    if (r.suggest) {
        out += "suggest=[";
        for (auto& p : *r.suggest) describe(p, out);
        out += "] ";
    }
    if (r.boost) {
        out += "boost=[";
        for (auto& b : *r.boost) describe(b, out);
        out += "] ";
    }
    if (r.boost_term) {
        out += "boost_term=[";
        for (auto& p : *r.boost_term) describe(p, out);
        out += "] ";
    }
    if (r.facets) {
        out += "facets=[";
        for (auto& f : *r.facets) out += f.field + ":" + (f.top ? std::to_string(*f.top) : "None") + ",";
        out += "] ";
    }
    if (r.phrase_boosts) {
        out += "phrase_boosts=[";
        for (auto& p : *r.phrase_boosts) describe(p.search1, out), out += "+", describe(p.search2, out), out += ",";
        out += "] ";
    }
    if (r.select) {
        out += "select=[";
        for (auto& f : *r.select) out += f + ",";
        out += "] ";
    }
    if (r.filter) out += "filter=", describe(*r.filter, out), out += " ";
    u("top", r.top), u("skip", r.skip);
    out += r.why_found ? "W" : "w";
    out += r.text_locality ? "T" : "t";
    out += r.explain ? "E" : "e";
    return out;
}

inline Request parse_request_json(const char* json, size_t len) {
    vjson::Value v;
    try {
        v = vjson::parse(json, len);
    } catch (const vjson::ParseError& e) {
        throw RequestError(e.what());
    }
    return parse_request(v);
}

}  // namespace vhost
