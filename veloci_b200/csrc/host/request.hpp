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
        std::string k = path;
        for (auto& t : terms) k += "\x1e" + t;
        k += "\x1f";
        if (levenshtein_distance) k += std::to_string(*levenshtein_distance);
        k += starts_with ? "\x1fS" : "\x1fs";
        k += is_regex ? "R" : "r";
        k += "\x1f";
        if (token_value) k += token_value->key();
        k += "\x1f";
        if (boost) k += std::to_string(*boost);
        k += "\x1f";
        if (ignore_case) k += *ignore_case ? "1" : "0";
        k += "\x1f";
        if (top) k += std::to_string(*top);
        k += "\x1f";
        if (skip) k += std::to_string(*skip);
        k += "\x1f";
        if (options.present) {
            k += options.explain ? "E" : "e";
            if (options.top) k += std::to_string(*options.top);
            k += ",";
            if (options.skip) k += std::to_string(*options.skip);
            if (options.boost)
                for (auto& b : *options.boost) k += "|" + b.key();
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
