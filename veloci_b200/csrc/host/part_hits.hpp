// Host arithmetic on the (term id, score) hits of one search part, after the device has matched and scored it: the per-part
// top / skip bound and the token_value boost of get_term_ids_in_field (search_field.rs:292-294,322-331,359-376,391-395), and the
// `x op y` boost expression both this and the planner read (src/expression.rs:25-100).  No device code: the index helper
// library exposes these to the CPU tests (vidx_bound_part_hits).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../cuda/device_types.cuh"
#include "persistence.hpp"
#include "request.hpp"

namespace vplan {

using vdev::BoostStep;

struct Unsupported : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct InvalidRequest : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// `x op y` boost expression (src/expression.rs:25-100)
inline void parse_expression(const std::string& expression, BoostStep& b) {
    enum Kind { Division, Mul, Add, Sub, Score, Float };
    std::vector<std::pair<Kind, float>> ops;
    std::string current;
    auto try_float = [&](const std::string& s) {
        if (s.empty()) return;
        char* end = nullptr;
        float v = strtof(s.c_str(), &end);
        if (end && *end == 0 && end != s.c_str()) ops.emplace_back(Float, v);
    };
    for (char c : expression) {
        if (c == ' ') {
            try_float(current);
            current.clear();
        } else {
            current.push_back(c);
        }
        if (current == "+") ops.emplace_back(Add, 0.f), current.clear();
        else if (current == "-") ops.emplace_back(Sub, 0.f), current.clear();
        else if (current == "/") ops.emplace_back(Division, 0.f), current.clear();
        else if (current == "*") ops.emplace_back(Mul, 0.f), current.clear();
        else if (current == "$SCORE") ops.emplace_back(Score, 0.f), current.clear();
    }
    try_float(current);
    if (ops.size() < 3) throw InvalidRequest("boost expression must be `x op y`");
    auto operand = [&](const std::pair<Kind, float>& o, uint32_t& is_score, float& val) {
        if (o.first == Score) is_score = 1, val = 0.f;
        else if (o.first == Float) is_score = 0, val = o.second;
        else throw InvalidRequest("boost expression operand must be a float or $SCORE");
    };
    operand(ops[0], b.expr_left_is_score, b.expr_left);
    operand(ops[2], b.expr_right_is_score, b.expr_right);
    switch (ops[1].first) {
        case Division: b.expr_op = vdev::kExprDiv; break;
        case Mul: b.expr_op = vdev::kExprMul; break;
        case Add: b.expr_op = vdev::kExprAdd; break;
        case Sub: b.expr_op = vdev::kExprSub; break;
        default: throw InvalidRequest("boost expression operator must be one of * + - /");
    }
}

}  // namespace vplan

namespace vdev {

struct TermHit {
    uint32_t id;
    float score;
};

// The per-part `top` bound of get_term_ids_in_field over hits in FST order (= ascending term id): once top + skip + 200
// hits are held, the best top + skip by (score, id) stay and a hit scoring below the worst of them is dropped from then
// on (search_field.rs:322-331, sort.rs:25-34); the part's boost comes after (:359-364), then the best top + skip by score
// stay (:366-369; the reference's unstable sort leaves the choice among equal scores open, a stable one is used here).
inline void bound_part_hits(const vhost::SearchPart& part, std::vector<TermHit>& hits) {
    const size_t top_n = part.top ? (size_t)(*part.top + part.skip.value_or(0)) : 0;
    if (part.top) {
        float worst = -3.40282347e+38f;
        std::vector<TermHit> kept;
        for (const TermHit& h : hits) {
            if (h.score < worst) continue;
            if (!kept.empty() && kept.size() == top_n + 200) {
                std::sort(kept.begin(), kept.end(), [](const TermHit& a, const TermHit& b) { return a.score != b.score ? a.score > b.score : a.id > b.id; });
                kept.resize(top_n);
                if (!kept.empty()) worst = kept.back().score;
            }
            kept.push_back(h);
        }
        hits.swap(kept);
    }
    if (part.boost)
        for (TermHit& h : hits) h.score *= *part.boost;
    if (part.top) {
        std::stable_sort(hits.begin(), hits.end(), [](const TermHit& a, const TermHit& b) { return a.score > b.score; });
        if (hits.size() > top_n) hits.resize(top_n);
    }
}

// `token_value` of a search part (search_field.rs:391-395): after the bound, add_boost runs over the part's term hits with the
// values of `<path>.textindex.token_values.boost_valid_to_value`, a 1:1 store keyed by term id (written by
// create/token_values_to_tokens.rs:26-82).  skip_when_score, boost function and `x op y` expression as boost.rs:470-504 /
// :283-377; a few hundred hits of host arithmetic on the output of the device match.
// apply_boost (boost.rs:283-377) on one score with the boost value `v`: boost function, then the `x op y` expression.  `log`
// receives what the reference's explain records on the way: the logarithm factor (Log10 only, :297-300) and the new score (:371-374).
inline void apply_boost_value(const vhost::BoostPart& boost, float v, float& score, std::vector<float>* log = nullptr) {
    const float param = boost.param.value_or(0.0f);
    if (log && boost.boost_fun == vhost::BoostFun::Log10) log->push_back(log10f(v + param));
    switch (boost.boost_fun) {
        case vhost::BoostFun::Log10: score *= log10f(v + param); break;
        case vhost::BoostFun::Log2: score *= log2f(v + param); break;
        case vhost::BoostFun::Multiply: score *= v + param; break;
        case vhost::BoostFun::Add: score += v + param; break;
        case vhost::BoostFun::Replace: score = v + param; break;
        case vhost::BoostFun::None: break;
    }
    if (boost.expression) {
        BoostStep expr;
        memset(&expr, 0, sizeof expr);
        vplan::parse_expression(*boost.expression, expr);
        const float l = expr.expr_left_is_score ? v : expr.expr_left, r = expr.expr_right_is_score ? v : expr.expr_right;
        score += expr.expr_op == vdev::kExprDiv ? l / r : expr.expr_op == vdev::kExprMul ? l * r : expr.expr_op == vdev::kExprAdd ? l + r : l - r;
    }
    if (log) log->push_back(score);
}

inline void apply_token_value(const vhost::Persistence& host, const vhost::SearchPart& part, std::vector<TermHit>& hits, std::map<uint32_t, std::vector<float>>* explain_log = nullptr) {
    if (!part.token_value) return;
    const vhost::BoostPart& tb = *part.token_value;
    const vhost::KeyValueStore& store = host.get_boost(tb.path + ".textindex.token_values.boost_valid_to_value");
    if (tb.expression) {  // a malformed expression fails the part whether or not a hit carries a value
        BoostStep probe;
        memset(&probe, 0, sizeof probe);
        vplan::parse_expression(*tb.expression, probe);
    }
    for (TermHit& h : hits) {
        bool skip = false;
        if (tb.skip_when_score)
            for (float x : *tb.skip_when_score) skip = skip || fabsf(x - h.score) < 0.00001f;
        uint32_t bits = 0;
        if (skip || !store.get_value(h.id, bits)) continue;
        float v;
        memcpy(&v, &bits, 4);
        apply_boost_value(tb, v, h.score, explain_log ? &(*explain_log)[h.id] : nullptr);
    }
}

}  // namespace vdev
