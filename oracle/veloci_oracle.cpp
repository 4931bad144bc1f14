// ORACLE -- TEST INFRASTRUCTURE ONLY.  A single-threaded CPU restatement of
// veloci's query-time hit pipeline, function by function, used by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker for
// the CUDA path.  Nothing under veloci_b200/ may call into this file.
//
// Parity pinning: the Rust reference cannot be compiled in this image (no
// cargo/rustc), so this restatement is pinned against the reference's own
// unit-test vectors and integration fixtures (tests/test_oracle_*.py cite
// them) -- hit sets, order, counts and facet vectors.  What stays UNPINNED
// (no vector exists in the reference): absolute f32 score values, the match
// set of veloci_levenshtein_automata beyond the lev/ignore_case/starts_with
// fixtures, and the third-party byte formats (fst 0.4, vint32 most-common
// encoding).  See DESIGN.md "Oracle".
//
// Each function cites the reference file:line it follows.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../veloci_b200/csrc/format/codecs.hpp"
#include "../veloci_b200/csrc/format/unicode.hpp"
#include "../veloci_b200/csrc/host/persistence.hpp"
#include "../veloci_b200/csrc/host/request.hpp"
#include "../veloci_b200/csrc/index/indexer.hpp"
#include "../veloci_b200/csrc/vjson.hpp"
#include "regex_sim.hpp"
#include "rust_lower.hpp"

using vhost::BoostFun;
using vhost::BoostPart;
using vhost::Persistence;
using vhost::Request;
using vhost::SearchPart;
using vhost::SearchRequest;

namespace oracle {

struct InvalidRequest : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Hit {  // search.rs:53-57
    uint32_t id;
    float score;
};

struct Explain {  // result/explain.rs:1-21
    enum Kind { Boost, MaxTokenToTextId, TermToAnchor, LevenshteinScore, OrSumOverDistinctTerms } kind;
    float a = 0.f, b = 0.f, c = 0.f;  // Boost / MaxTokenToTextId / OrSum: a.  TermToAnchor: term_score, anchor_score, final_score.  LevenshteinScore: score
    uint32_t term_id = 0;
    std::string text;  // LevenshteinScore::text_or_token_id
    explicit Explain(Kind k, float value = 0.f) : kind(k), a(value) {}
};
typedef std::map<uint32_t, std::vector<Explain>> ExplainMap;

static bool is_explain(const SearchPart& part) {  // search_request.rs:182-184
    return part.options.present && part.options.explain;
}

struct SearchFieldResult {  // result/field_result.rs:6-30 (hit-bearing fields only)
    ExplainMap explain;
    std::vector<Hit> hits_scores;
    std::vector<uint32_t> hits_ids;
    std::vector<Hit> boost_ids;
    SearchPart request;
    std::optional<vhost::PhraseBoost> phrase_boost;
    std::map<std::string, std::map<std::string, std::vector<uint32_t>>> term_id_hits_in_field;
    std::map<uint32_t, std::string> terms;  // term id -> text, when the plan asks for it (suggest)
    static SearchFieldResult new_from(const SearchFieldResult& o) {  // :42-52
        SearchFieldResult r;
        r.request = o.request;
        r.phrase_boost = o.phrase_boost;
        r.term_id_hits_in_field = o.term_id_hits_in_field;
        return r;
    }
};

struct FilterResult {  // result/filter_result.rs:4-22
    bool is_set = false;
    std::vector<uint32_t> vec;
    std::unordered_set<uint32_t> set;
    static FilterResult from_result(const std::vector<uint32_t>& res) {
        FilterResult f;
        if (res.size() > 100000) {
            f.vec = res;
        } else {
            f.is_set = true;
            f.set.insert(res.begin(), res.end());
        }
        return f;
    }
};

struct PlanRequestSearchPart {  // execution_plan.rs:16-44
    SearchPart request;
    bool get_scores = false;
    bool get_ids = false;
    bool store_term_id_hits = false;
    bool return_term = false, return_term_lowercase = false;
};

// search.rs:123-130
static bool sort_by_score_and_id_less(const Hit& a, const Hit& b) {
    if (a.score != b.score) return a.score > b.score;
    return a.id > b.id;
}

// search_field.rs:27-33
static float get_default_score_for_distance(uint8_t distance, bool prefix_matches) {
    if (prefix_matches) return 2.0f / (log2f((float)distance + 1.0f) + 0.2f);
    return 2.0f / ((float)distance + 0.2f);
}

// search_field.rs:705-732 (u8 cells, wrapping like release-mode Rust)
static uint8_t distance(const std::string& s1, const std::string& s2) {
    if (s1.size() >= 255 || s2.size() >= 255) return 255;
    std::vector<uint32_t> c1, c2;
    vfmt::utf8_decode(s1, c1);
    vfmt::utf8_decode(s2, c2);
    size_t len_s1 = c1.size();
    uint8_t column[255] = {0};
    for (size_t i = 0; i < len_s1 + 1 && i < 255; ++i) column[i] = (uint8_t)i;
    for (size_t x = 0; x < c2.size(); ++x) {
        column[0] = (uint8_t)(x + 1);
        uint8_t lastdiag = (uint8_t)x;
        for (size_t y = 0; y < c1.size(); ++y) {
            if (c1[y] != c2[x]) lastdiag = (uint8_t)(lastdiag + 1);
            uint8_t olddiag = column[y + 1];
            column[y + 1] = std::min<uint8_t>((uint8_t)(column[y + 1] + 1), std::min<uint8_t>((uint8_t)(column[y] + 1), lastdiag));
            lastdiag = olddiag;
        }
    }
    return column[len_s1];
}

// Edit distance over Unicode scalars with optional adjacent-transposition at
// cost one (the `transposition_cost_one` flag of LevenshteinAutomatonBuilder::new).
static uint32_t edit_distance(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b, bool transposition) {
    size_t n = a.size(), m = b.size();
    std::vector<uint32_t> pp(m + 1), p(m + 1), c(m + 1);
    for (size_t j = 0; j <= m; ++j) p[j] = (uint32_t)j;
    for (size_t i = 1; i <= n; ++i) {
        c[0] = (uint32_t)i;
        for (size_t j = 1; j <= m; ++j) {
            uint32_t v = std::min(std::min(p[j] + 1, c[j - 1] + 1), p[j - 1] + (a[i - 1] != b[j - 1] ? 1u : 0u));
            if (transposition && i >= 2 && j >= 2 && a[i - 1] == b[j - 2] && a[i - 2] == b[j - 1]) v = std::min(v, pp[j - 2] + 1);
            c[j] = v;
        }
        pp.swap(p);
        p.swap(c);
    }
    return p[m];
}

// search_field.rs:691-702: scoring DFA built with (d, transposition=true), case-sensitive,
// run over the lower-cased hit; Exact(k) for k <= d, otherwise the plain DP above.
static uint8_t distance_dfa(const std::string& lower_hit, const std::string& lower_term, uint32_t d) {
    std::vector<uint32_t> a, b;
    vfmt::utf8_decode(lower_hit, a);
    vfmt::utf8_decode(lower_term, b);
    uint32_t k = edit_distance(a, b, true);
    if (k <= (d & 0xFF)) return (uint8_t)k;
    return distance(lower_hit, lower_term);
}

// The Levenshtein automaton intersected with the FST (search_field.rs:54-99):
// walks the byte-sorted dictionary with one DP row per scalar of the current
// key, reusing the rows shared with the previous key and jumping over whole
// prefix ranges once every cell of a row exceeds d.  Visits matches in
// ascending key (= ascending id) order, exactly like the fst stream.
static void levenshtein_search(const vhost::TermDict& dict, const std::string& query, uint32_t d, bool transposition, bool case_insensitive, bool starts_with,
                               const std::function<void(size_t slot)>& on_match) {
    std::vector<uint32_t> q;
    vfmt::utf8_decode(query, q);
    if (case_insensitive)
        for (auto& c : q) c = olow::lower_one(c);  // (the matching automaton folds scalar by scalar: no context, no expansion)
    const size_t m = q.size();
    std::vector<std::vector<uint32_t>> rows(1, std::vector<uint32_t>(m + 1));
    for (size_t j = 0; j <= m; ++j) rows[0][j] = (uint32_t)j;
    std::vector<uint32_t> chars;      // folded scalars of the prefix the rows belong to
    std::vector<uint32_t> byte_end;   // byte offset in the key after each scalar
    std::vector<uint8_t> cur_prefix;  // bytes of that prefix
    std::vector<char> acc(1, (starts_with && rows[0][m] <= d) ? 1 : 0);  // prefix mode: some prefix so far accepted
    const size_t n = dict.size();
    size_t i = 0;
    while (i < n) {
        const uint8_t* key = &dict.bytes[dict.offsets[i]];
        const size_t klen = dict.offsets[i + 1] - dict.offsets[i];
        size_t common_bytes = 0;
        const size_t lim = std::min(klen, cur_prefix.size());
        while (common_bytes < lim && key[common_bytes] == cur_prefix[common_bytes]) ++common_bytes;
        size_t keep = 0;
        while (keep < chars.size() && byte_end[keep] <= common_bytes) ++keep;
        chars.resize(keep);
        byte_end.resize(keep);
        rows.resize(keep + 1);
        acc.resize(keep + 1);
        cur_prefix.resize(keep ? byte_end[keep - 1] : 0);
        size_t pos = cur_prefix.size();
        bool accepted = acc[keep] != 0;
        bool pruned = false;
        while (pos < klen && !accepted) {
            size_t next = pos;
            uint32_t cp = vfmt::utf8_next(key, klen, next);
            if (case_insensitive) cp = olow::lower_one(cp);
            const size_t k = chars.size();
            std::vector<uint32_t> row(m + 1);
            row[0] = (uint32_t)k + 1;
            uint32_t rmin = row[0];
            for (size_t j = 1; j <= m; ++j) {
                uint32_t v = std::min(std::min(rows[k][j] + 1, row[j - 1] + 1), rows[k][j - 1] + (q[j - 1] != cp ? 1u : 0u));
                if (transposition && k >= 1 && j >= 2 && cp == q[j - 2] && chars[k - 1] == q[j - 1]) v = std::min(v, rows[k - 1][j - 2] + 1);
                row[j] = v;
                rmin = std::min(rmin, v);
            }
            const bool a = starts_with && row[m] <= d;
            chars.push_back(cp);
            byte_end.push_back((uint32_t)next);
            rows.push_back(std::move(row));
            acc.push_back(a ? 1 : 0);
            cur_prefix.insert(cur_prefix.end(), key + pos, key + next);
            pos = next;
            if (a) accepted = true;
            else if (rmin > d) {
                pruned = true;
                break;
            }
        }
        if (pruned) {  // no key below this prefix can match: jump behind the prefix range
            size_t lo = i + 1, hi = n;
            const size_t plen = cur_prefix.size();
            while (lo < hi) {
                size_t mid = (lo + hi) / 2;
                const uint8_t* mk = &dict.bytes[dict.offsets[mid]];
                size_t ml = dict.offsets[mid + 1] - dict.offsets[mid];
                if (ml >= plen && memcmp(mk, cur_prefix.data(), plen) == 0) lo = mid + 1;
                else hi = mid;
            }
            i = lo;
            continue;
        }
        const bool match = starts_with ? accepted : rows.back()[m] <= d;
        if (match) on_match(i);
        ++i;
    }
}

// search_field.rs:277-398 get_term_ids_in_field (token_value boost, per-part top/skip
// pruning and the explain map included)
static void add_boost(const Persistence& p, const BoostPart& boost, SearchFieldResult& hits);

static SearchFieldResult get_term_ids_in_field(const Persistence& p, PlanRequestSearchPart& options) {
    SearchPart& req = options.request;
    if (!vfmt::ends_with(req.path, ".textindex")) req.path += ".textindex";
    SearchFieldResult result;
    result.request = req;
    if (req.terms.empty()) throw InvalidRequest("search part without terms");
    std::string lower_term = olow::to_lowercase(req.terms[0]);
    if (req.levenshtein_distance) {
        uint32_t chars = (uint32_t)vfmt::utf8_count(lower_term);
        req.levenshtein_distance = std::min(*req.levenshtein_distance, chars - 1u);  // wraps for the empty term like release Rust
    }
    result.request = req;
    const bool limit_result = req.top.has_value();
    float worst_score = -3.40282347e+38f;
    const uint32_t top_n_search = (uint32_t)(req.top.value_or(10) + req.skip.value_or(0));
    const uint32_t d_score = req.levenshtein_distance.value_or(0);
    const bool should_check_prefix_match = req.starts_with || d_score != 0;

    const vhost::TermDict& dict = p.get_dict(req.path);  // FstNotFound
    const uint32_t d_match = std::min<uint32_t>(req.levenshtein_distance.value_or(0), 4);
    const bool transposition = req.ignore_case.value_or(false);   // search_field.rs:87 (sic)
    const bool case_insensitive = req.ignore_case.value_or(true);  // :88
    const std::function<void(size_t)> on_match = [&](size_t slot) {
        uint32_t token_text_id = dict.ids[slot];
        if (options.get_ids) result.hits_ids.push_back(token_text_id);
        if (options.get_scores) {
            std::string line_lower = olow::to_lowercase(dict.term(slot));
            bool prefix_matches = should_check_prefix_match && line_lower.compare(0, lower_term.size(), lower_term) == 0 && line_lower.size() >= lower_term.size();
            float score = get_default_score_for_distance(distance_dfa(line_lower, lower_term, d_score), prefix_matches);
            if (limit_result) {
                if (score < worst_score) return;
                if (!result.hits_scores.empty() && result.hits_scores.size() == (size_t)top_n_search + 200) {  // sort.rs:25-34
                    std::sort(result.hits_scores.begin(), result.hits_scores.end(), sort_by_score_and_id_less);
                    result.hits_scores.resize(top_n_search);
                    if (!result.hits_scores.empty()) worst_score = result.hits_scores.back().score;
                }
            }
            result.hits_scores.push_back(Hit{token_text_id, score});
            if (is_explain(req)) {  // :334-344
                Explain e(Explain::LevenshteinScore);
                e.a = score, e.term_id = token_text_id, e.text = dict.term(slot);
                result.explain[token_text_id] = {e};
            }
        }
        if (options.return_term) result.terms[token_text_id] = options.return_term_lowercase ? olow::to_lowercase(dict.term(slot)) : dict.term(slot);  // :331-337
    };
    if (req.is_regex) {  // search_field.rs:72-83: the pattern's DFA instead of the Levenshtein automaton, same stream order
        std::unique_ptr<oracle_regex::Program> program;
        try {
            program.reset(new oracle_regex::Program(req.terms[0], case_insensitive));
        } catch (const oracle_regex::BadPattern& e) {
            throw InvalidRequest(std::string("regex \"") + req.terms[0] + "\": " + e.what());
        } catch (const oracle_regex::OutsideSubset& e) {
            throw InvalidRequest(std::string("regex \"") + req.terms[0] + "\" outside the oracle's subset: " + e.what());
        }
        std::vector<uint32_t> scalars;
        for (size_t slot = 0; slot < dict.size(); ++slot) {
            scalars.clear();
            vfmt::utf8_decode(dict.term(slot), scalars);
            if (program->accepts(scalars, req.starts_with)) on_match(slot);
        }
    } else {
        levenshtein_search(dict, req.terms[0], d_match, transposition, case_insensitive, req.starts_with, on_match);
    }
    if (req.boost)
        for (auto& h : result.hits_scores) h.score *= *req.boost;
    if (limit_result) {
        std::stable_sort(result.hits_scores.begin(), result.hits_scores.end(), [](const Hit& a, const Hit& b) { return a.score > b.score; });
        if (result.hits_scores.size() > top_n_search) result.hits_scores.resize(top_n_search);
    }
    if (options.store_term_id_hits && !result.hits_scores.empty()) {
        std::vector<uint32_t> ids;
        for (auto& h : result.hits_scores) ids.push_back(h.id);
        result.term_id_hits_in_field[req.path][req.terms[0]] = ids;
    }
    if (req.token_value) {
        BoostPart tb = *req.token_value;
        tb.path = tb.path + ".textindex" + ".token_values";
        add_boost(p, tb, result);
    }
    return result;
}

static float f16_roundtrip(float f) {  // half::f16::from_f32(x).to_f32(), round-to-nearest-even
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xFF);
    uint32_t man = x & 0x7FFFFFu;
    uint16_t h;
    if (exp == 255) {
        h = (uint16_t)(sign | 0x7C00u | (man ? 0x200u | (man >> 13) : 0));
    } else {
        int32_t e = exp - 127 + 15;
        if (e >= 31) h = (uint16_t)(sign | 0x7C00u);
        else if (e <= 0) {
            if (e < -10) h = (uint16_t)sign;
            else {
                man |= 0x800000u;
                uint32_t shift = (uint32_t)(14 - e);
                uint32_t hm = man >> shift;
                uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
                if (rem > half || (rem == half && (hm & 1))) hm++;
                h = (uint16_t)(sign | hm);
            }
        } else {
            uint32_t hm = man >> 13, rem = man & 0x1FFFu;
            uint32_t v = ((uint32_t)e << 10) | hm;
            if (rem > 0x1000u || (rem == 0x1000u && (v & 1))) v++;
            h = (uint16_t)(sign | v);
        }
    }
    uint32_t hs = (uint32_t)(h & 0x8000u) << 16, he = (h >> 10) & 0x1F, hm = h & 0x3FFu, out;
    if (he == 0) {
        if (hm == 0) out = hs;
        else {
            int e = -1;
            do {
                hm <<= 1;
                e++;
            } while (!(hm & 0x400u));
            out = hs | ((uint32_t)(127 - 15 - e) << 23) | ((hm & 0x3FFu) << 13);
        }
    } else if (he == 31) out = hs | 0x7F800000u | (hm << 13);
    else out = hs | ((he + 112) << 23) | (hm << 13);
    float r;
    memcpy(&r, &out, 4);
    return r;
}

// search_field.rs:400-504
static SearchFieldResult resolve_token_to_anchor(const Persistence& p, const SearchPart& options_in, const std::optional<std::shared_ptr<FilterResult>>& filter, const SearchFieldResult& result) {
    std::string path = options_in.path;
    if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
    SearchFieldResult res = SearchFieldResult::new_from(result);
    std::vector<Hit> anchor_ids_hits;
    const vfmt::AnchorScoreView& store = p.get_token_to_anchor(path);
    const FilterResult* f = (filter && *filter) ? filter->get() : nullptr;
    for (const Hit& hit : result.hits_scores) {
        store.for_each(hit.id, [&](uint32_t anchor, uint32_t raw) {
            if (f && f->is_set && !f->set.count(anchor)) return;  // should_filter :540-548
            float el_score = f16_roundtrip((float)raw);
            float final_score = hit.score * (el_score / 100.0f);
            if (is_explain(options_in)) {  // :429-441
                std::vector<Explain>& vecco = res.explain[anchor];
                Explain e(Explain::TermToAnchor);
                e.term_id = hit.id, e.a = hit.score, e.b = el_score / 100.0f, e.c = final_score;
                vecco.push_back(e);
                auto exp = result.explain.find(hit.id);
                if (exp != result.explain.end()) vecco.insert(vecco.end(), exp->second.begin(), exp->second.end());
            }
            anchor_ids_hits.push_back(Hit{anchor, final_score});
        });
    }
    std::stable_sort(anchor_ids_hits.begin(), anchor_ids_hits.end(), [](const Hit& a, const Hit& b) { return a.id < b.id; });
    {  // dedup_by keeping the max score
        size_t w = 0;
        for (size_t r = 0; r < anchor_ids_hits.size(); ++r) {
            if (w > 0 && anchor_ids_hits[w - 1].id == anchor_ids_hits[r].id) {
                if (anchor_ids_hits[r].score > anchor_ids_hits[w - 1].score) anchor_ids_hits[w - 1].score = anchor_ids_hits[r].score;
            } else {
                anchor_ids_hits[w++] = anchor_ids_hits[r];
            }
        }
        anchor_ids_hits.resize(w);
    }
    std::vector<uint32_t> fast_field_res_ids;
    if (!result.hits_ids.empty()) {
        if (p.is_anchor_identity_column(path)) {
            fast_field_res_ids = result.hits_ids;
        } else {
            const vhost::KeyValueStore& t2a = p.get_valueid_to_parent(path + ".text_id_to_anchor");
            for (uint32_t id : result.hits_ids) t2a.append_values(id, fast_field_res_ids);
        }
    }
    res.hits_ids = std::move(fast_field_res_ids);
    res.hits_scores = std::move(anchor_ids_hits);
    return res;
}

// search_field.rs:640-689
static void resolve_token_hits_to_text_id_ids_only(const Persistence& p, const SearchPart& options, SearchFieldResult& result) {
    std::string path = options.path;
    if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
    if (!p.is_tokenized(path)) return;
    const vhost::KeyValueStore& kv = p.get_valueid_to_parent(path + ".tokens_to_text_id");
    std::vector<uint32_t> token_hits, tmp;
    for (const Hit& hit : result.hits_scores) {
        if (kv.get_values(hit.id, tmp)) token_hits.insert(token_hits.end(), tmp.begin(), tmp.end());
        else token_hits.push_back(hit.id);
    }
    std::sort(token_hits.begin(), token_hits.end());
    token_hits.erase(std::unique(token_hits.begin(), token_hits.end()), token_hits.end());
    result.hits_ids = std::move(token_hits);
    result.hits_scores.clear();
}

// search.rs:281-315
static SearchFieldResult join_to_parent_ids(const Persistence& p, const SearchFieldResult& input, const std::string& path) {
    const vhost::KeyValueStore& kv = p.get_valueid_to_parent(path);
    std::vector<uint32_t> hits, tmp;
    for (uint32_t id : input.hits_ids)
        if (kv.get_values(id, tmp)) hits.insert(hits.end(), tmp.begin(), tmp.end());
    std::sort(hits.begin(), hits.end());
    hits.erase(std::unique(hits.begin(), hits.end()), hits.end());
    SearchFieldResult res = SearchFieldResult::new_from(input);
    res.hits_ids = std::move(hits);
    return res;
}

// expression.rs:25-100
struct ScoreExpression {
    enum Op { Division, Mul, Add, Sub, Score, Float };
    std::vector<std::pair<Op, float>> ops;
    explicit ScoreExpression(const std::string& expression) {
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
            }
            if (c != ' ') current.push_back(c);
            if (current == "+") ops.emplace_back(Add, 0.f), current.clear();
            else if (current == "-") ops.emplace_back(Sub, 0.f), current.clear();
            else if (current == "/") ops.emplace_back(Division, 0.f), current.clear();
            else if (current == "*") ops.emplace_back(Mul, 0.f), current.clear();
            else if (current == "$SCORE") ops.emplace_back(Score, 0.f), current.clear();
        }
        try_float(current);
    }
    float get_score(float rank) const {
        if (ops.size() < 3) throw InvalidRequest("boost expression must be `x op y`");
        auto val = [&](const std::pair<Op, float>& o) -> float {
            if (o.first == Score) return rank;
            if (o.first == Float) return o.second;
            throw InvalidRequest("boost expression operand must be a float or $SCORE");
        };
        float left = val(ops[0]), right = val(ops[2]);
        switch (ops[1].first) {
            case Division: return left / right;
            case Mul: return left * right;
            case Add: return left + right;
            case Sub: return left - right;
            default: throw InvalidRequest("boost expression operator must be one of * + - /");
        }
    }
};

// boost.rs:283-377
static void apply_boost(Hit& hit, float boost_value, float boost_param, BoostFun fun, const std::optional<ScoreExpression>& expre, ExplainMap* explain = nullptr) {
    if (explain && fun == BoostFun::Log10) (*explain)[hit.id].push_back(Explain(Explain::Boost, log10f(boost_value + boost_param)));  // :297-300 (Log10 only)
    switch (fun) {
        case BoostFun::Log10: hit.score *= log10f(boost_value + boost_param); break;
        case BoostFun::Log2: hit.score *= log2f(boost_value + boost_param); break;
        case BoostFun::Multiply: hit.score *= boost_value + boost_param; break;
        case BoostFun::Add: hit.score += boost_value + boost_param; break;
        case BoostFun::Replace: hit.score = boost_value + boost_param; break;
        case BoostFun::None: break;
    }
    if (expre) hit.score += expre->get_score(boost_value);
    if (explain) (*explain)[hit.id].push_back(Explain(Explain::Boost, hit.score));  // :371-374
}

// boost.rs:470-504
static void add_boost(const Persistence& p, const BoostPart& boost, SearchFieldResult& hits) {
    const vhost::KeyValueStore& store = p.get_boost(boost.path + ".boost_valid_to_value");
    float boost_param = boost.param.value_or(0.0f);
    std::optional<ScoreExpression> expre;
    if (boost.expression) expre.emplace(*boost.expression);
    std::vector<float> skip = boost.skip_when_score.value_or(std::vector<float>());
    for (Hit& hit : hits.hits_scores) {
        bool skipped = false;
        for (float x : skip)
            if (fabsf(x - hit.score) < 0.00001f) skipped = true;
        if (skipped) continue;
        uint32_t bits;
        if (store.get_value(hit.id, bits)) {
            float v;
            memcpy(&v, &bits, 4);
            apply_boost(hit, v, boost_param, boost.boost_fun, expre, is_explain(hits.request) ? &hits.explain : nullptr);  // :484
        }
    }
}

// boost.rs:255-281
static void apply_boost_values_anchor(SearchFieldResult& results, const BoostPart& boost, const std::vector<Hit>& boosts) {
    float boost_param = boost.param.value_or(0.0f);
    std::optional<ScoreExpression> expre;
    if (boost.expression) expre.emplace(*boost.expression);
    size_t bi = 0;
    if (bi >= boosts.size()) return;
    ExplainMap* explain = is_explain(results.request) ? &results.explain : nullptr;  // :258
    Hit hit_curr = boosts[bi++];
    for (Hit& hit : results.hits_scores) {
        if (hit_curr.id < hit.id) {
            while (bi < boosts.size()) {
                Hit b_hit = boosts[bi++];
                if (b_hit.id > hit.id) {
                    hit_curr = b_hit;
                    break;
                } else if (b_hit.id == hit.id) {
                    hit_curr = b_hit;
                    apply_boost(hit, b_hit.score, boost_param, boost.boost_fun, expre, explain);
                }
            }
        } else if (hit_curr.id == hit.id) {
            apply_boost(hit, hit_curr.score, boost_param, boost.boost_fun, expre, explain);
        }
    }
}

// boost.rs:197-237
static void apply_boost_from_iter(SearchFieldResult& results, const std::vector<Hit>& boost_iter) {
    size_t bi = 0;
    auto move_boost = [&](Hit& hit, Hit& hit_curr) {
        while (bi < boost_iter.size()) {
            Hit b_hit = boost_iter[bi++];
            if (b_hit.id > hit.id) {
                hit_curr = b_hit;
                break;
            } else if (b_hit.id == hit.id) {
                hit_curr = b_hit;
                hit.score *= b_hit.score;
                if (is_explain(results.request)) results.explain[hit.id].push_back(Explain(Explain::Boost, b_hit.score));  // :213-217 (only here, not for the first match below)
            }
        }
    };
    if (bi < boost_iter.size()) {
        Hit hit_curr = boost_iter[bi++];
        for (Hit& hit : results.hits_scores) {
            if (hit_curr.id < hit.id) {
                move_boost(hit, hit_curr);
            } else if (hit_curr.id == hit.id) {
                hit.score *= hit_curr.score;
                move_boost(hit, hit_curr);
            }
        }
    }
}

// k-way merge by id of already-sorted lists (itertools kmerge_by |a,b| a.id < b.id; ties: lower list first)
static std::vector<Hit> kmerge_hits(const std::vector<std::vector<Hit>>& lists) {
    std::vector<Hit> out;
    std::vector<size_t> pos(lists.size(), 0);
    while (true) {
        int best = -1;
        for (size_t l = 0; l < lists.size(); ++l)
            if (pos[l] < lists[l].size() && (best < 0 || lists[l][pos[l]].id < lists[(size_t)best][pos[(size_t)best]].id)) best = (int)l;
        if (best < 0) break;
        out.push_back(lists[(size_t)best][pos[(size_t)best]++]);
    }
    return out;
}

// boost.rs:380-402
static void boost_hits_ids_vec_multi(SearchFieldResult& results, std::vector<SearchFieldResult>& boost) {
    std::stable_sort(results.hits_scores.begin(), results.hits_scores.end(), [](const Hit& a, const Hit& b) { return a.id < b.id; });
    std::vector<std::vector<Hit>> lists;
    for (auto& res : boost) {
        std::sort(res.hits_ids.begin(), res.hits_ids.end());
        float boost_val = res.request.boost.value_or(2.0f);
        std::vector<Hit> l;
        for (uint32_t id : res.hits_ids) l.push_back(Hit{id, boost_val});
        lists.push_back(std::move(l));
    }
    apply_boost_from_iter(results, kmerge_hits(lists));
}

// boost.rs:432-468
static void get_boost_ids_and_resolve_to_anchor(const Persistence& p, const std::string& boost_path, SearchFieldResult& hits) {
    // FieldPath::from_path(path).as_string() round-trips a path without an index suffix
    const vhost::KeyValueStore& boostkv = p.get_boost(boost_path + ".boost_valid_to_value");
    std::sort(hits.hits_ids.begin(), hits.hits_ids.end());
    for (uint32_t value_id : hits.hits_ids) {
        uint32_t bits;
        if (boostkv.get_value(value_id, bits)) {
            float v;
            memcpy(&v, &bits, 4);
            hits.boost_ids.push_back(Hit{value_id, v});
        }
    }
    hits.hits_ids.clear();
    std::vector<Hit> data;
    const vhost::KeyValueStore& kv = p.get_valueid_to_parent(boost_path + ".value_id_to_anchor");
    for (const Hit& bp : hits.boost_ids) {
        uint32_t anchor;
        if (kv.get_value(bp.id, anchor)) data.push_back(Hit{anchor, bp.score});
    }
    hits.boost_ids = std::move(data);
}

// set_op.rs:87-220
static SearchFieldResult union_hits_score(std::vector<SearchFieldResult> or_results) {
    if (or_results.empty()) return SearchFieldResult();
    if (or_results.size() == 1) return std::move(or_results[0]);
    SearchFieldResult res;
    for (auto& el : or_results)  // merge_term_id_hits :29-47
        for (auto& attr : el.term_id_hits_in_field)
            for (auto& th : attr.second) res.term_id_hits_in_field[attr.first][th.first] = th.second;
    for (auto& r : or_results) std::stable_sort(r.hits_scores.begin(), r.hits_scores.end(), [](const Hit& a, const Hit& b) { return a.id < b.id; });
    std::vector<std::string> terms;
    for (auto& r : or_results) terms.push_back(r.request.terms.empty() ? std::string() : r.request.terms[0]);
    std::sort(terms.begin(), terms.end());
    terms.erase(std::unique(terms.begin(), terms.end()), terms.end());
    std::vector<uint8_t> term_id(or_results.size());
    for (size_t i = 0; i < or_results.size(); ++i) {
        const std::string t = or_results[i].request.terms.empty() ? std::string() : or_results[i].request.terms[0];
        term_id[i] = (uint8_t)(std::find(terms.begin(), terms.end(), t) - terms.begin());
    }
    const bool should_explain = is_explain(or_results[0].request);  // :120
    if (should_explain)  // :133-137: a later input's explanations replace an earlier one's for the same anchor
        for (auto& r : or_results)
            for (auto& kv : r.explain) res.explain[kv.first] = kv.second;
    std::vector<size_t> pos(or_results.size(), 0);
    std::vector<float> max_scores_per_term(terms.size(), 0.0f);
    while (true) {
        bool any = false;
        uint32_t id = 0;
        for (size_t l = 0; l < or_results.size(); ++l)
            if (pos[l] < or_results[l].hits_scores.size()) {
                uint32_t v = or_results[l].hits_scores[pos[l]].id;
                if (!any || v < id) id = v;
                any = true;
            }
        if (!any) break;
        for (auto& m : max_scores_per_term) m = 0.0f;
        for (size_t l = 0; l < or_results.size(); ++l)
            while (pos[l] < or_results[l].hits_scores.size() && or_results[l].hits_scores[pos[l]].id == id) {
                float s = or_results[l].hits_scores[pos[l]].score;
                float& m = max_scores_per_term[term_id[l]];
                m = fmaxf(m, s);  // f32::max
                ++pos[l];
            }
        float num_distinct_terms = 0.f;
        for (float m : max_scores_per_term)
            if (m >= 0.00001f) num_distinct_terms += 1.f;
        float sum = 0.0f;
        for (float m : max_scores_per_term) sum += m;
        res.hits_scores.push_back(Hit{id, sum * num_distinct_terms * num_distinct_terms});
        if (should_explain) res.explain[id].push_back(Explain(Explain::OrSumOverDistinctTerms, sum));  // :187-190
    }
    if (should_explain)  // :199-208
        for (const Hit& hit : res.hits_scores)
            for (auto& r : or_results) {
                auto exp = r.explain.find(hit.id);
                if (exp != r.explain.end()) {
                    std::vector<Explain>& e = res.explain[hit.id];
                    e.insert(e.end(), exp->second.begin(), exp->second.end());
                }
            }
    res.request = or_results[0].request;
    return res;
}

// set_op.rs:222-258
static SearchFieldResult union_hits_ids(std::vector<SearchFieldResult> or_results) {
    if (or_results.empty()) return SearchFieldResult();
    if (or_results.size() == 1) return std::move(or_results[0]);
    std::vector<uint32_t> all;
    for (auto& r : or_results) all.insert(all.end(), r.hits_ids.begin(), r.hits_ids.end());
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    SearchFieldResult res;
    res.hits_ids = std::move(all);
    res.request = or_results[0].request;
    return res;
}

// set_op.rs:311-326
static SearchFieldResult intersect_score_hits_with_ids(SearchFieldResult score_results, SearchFieldResult id_hits) {
    std::stable_sort(score_results.hits_scores.begin(), score_results.hits_scores.end(), [](const Hit& a, const Hit& b) { return a.id < b.id; });
    std::sort(id_hits.hits_ids.begin(), id_hits.hits_ids.end());
    if (!id_hits.hits_ids.empty()) {
        size_t it = 0;
        uint32_t current = id_hits.hits_ids[it++];
        std::vector<Hit> kept;
        for (const Hit& hit : score_results.hits_scores) {
            while (current < hit.id) current = it < id_hits.hits_ids.size() ? id_hits.hits_ids[it++] : UINT32_MAX;
            if (hit.id == current) kept.push_back(hit);
        }
        score_results.hits_scores = std::move(kept);
    }
    return score_results;
}

// set_op.rs:368-446
static SearchFieldResult intersect_hits_score(std::vector<SearchFieldResult> and_results) {
    if (and_results.empty()) return SearchFieldResult();
    if (and_results.size() == 1) return std::move(and_results[0]);
    SearchFieldResult res;
    for (auto& el : and_results)
        for (auto& attr : el.term_id_hits_in_field)
            for (auto& th : attr.second) res.term_id_hits_in_field[attr.first][th.first] = th.second;
    const bool should_explain = is_explain(and_results[0].request);  // :384 (before the shortest input is removed)
    size_t index_shortest = 0;
    uint64_t shortest = UINT64_MAX;
    for (size_t i = 0; i < and_results.size(); ++i)
        if ((uint64_t)and_results[i].hits_scores.size() < shortest) {
            shortest = and_results[i].hits_scores.size();
            index_shortest = i;
        }
    for (auto& r : and_results) std::stable_sort(r.hits_scores.begin(), r.hits_scores.end(), [](const Hit& a, const Hit& b) { return a.id < b.id; });
    // swap_remove
    std::vector<Hit> shortest_result = std::move(and_results[index_shortest].hits_scores);
    std::swap(and_results[index_shortest], and_results.back());
    and_results.pop_back();
    struct Cursor {
        const std::vector<Hit>* v;
        size_t next;
        Hit current;
    };
    std::vector<Cursor> its;
    for (auto& r : and_results)
        if (!r.hits_scores.empty()) its.push_back(Cursor{&r.hits_scores, 1, r.hits_scores[0]});
    auto check = [](Cursor& c, uint32_t id) {  // check_score_iter_for_id :347-366
        if (c.current.id == id) return true;
        if (c.current.id > id) return false;
        while (c.next < c.v->size()) {
            Hit el = (*c.v)[c.next++];
            c.current = el;
            if (el.id > id) return false;
            if (el.id == id) return true;
        }
        return false;
    };
    for (const Hit& cur : shortest_result) {
        bool all = true;
        for (auto& c : its)
            if (!check(c, cur.id)) {
                all = false;
                break;
            }
        if (all) {
            float score = 0.0f;
            for (auto& c : its) score += c.current.score;
            score += cur.score;
            res.hits_scores.push_back(Hit{cur.id, score});
        }
    }
    if (should_explain)  // :421-432: the explanations of the remaining inputs (the shortest one has been removed)
        for (const Hit& hit : res.hits_scores)
            for (auto& r : and_results) {
                auto exp = r.explain.find(hit.id);
                if (exp != r.explain.end()) {
                    std::vector<Explain>& e = res.explain[hit.id];
                    e.insert(e.end(), exp->second.begin(), exp->second.end());
                }
            }
    res.request = and_results[0].request;
    return res;
}

// set_op.rs:468-509
static SearchFieldResult intersect_hits_ids(std::vector<SearchFieldResult> and_results) {
    if (and_results.empty()) return SearchFieldResult();
    if (and_results.size() == 1) return std::move(and_results[0]);
    size_t index_shortest = 0;
    uint64_t shortest = UINT64_MAX;
    for (size_t i = 0; i < and_results.size(); ++i)
        if ((uint64_t)and_results[i].hits_ids.size() < shortest) {
            shortest = and_results[i].hits_ids.size();
            index_shortest = i;
        }
    for (auto& r : and_results) std::sort(r.hits_ids.begin(), r.hits_ids.end());
    std::vector<uint32_t> shortest_result = std::move(and_results[index_shortest].hits_ids);
    std::swap(and_results[index_shortest], and_results.back());
    and_results.pop_back();
    SearchFieldResult res;
    std::vector<size_t> pos(and_results.size(), 0);
    for (uint32_t id : shortest_result) {
        bool all = true;
        for (size_t l = 0; l < and_results.size(); ++l) {
            const auto& v = and_results[l].hits_ids;
            if (v.empty()) continue;  // filtered out of iterators_and_current (:489)
            while (pos[l] < v.size() && v[pos[l]] < id) ++pos[l];
            if (pos[l] >= v.size() || v[pos[l]] != id) {
                all = false;
                break;
            }
        }
        if (all) res.hits_ids.push_back(id);
    }
    return res;
}

// search_field.rs:263-275
static SearchFieldResult get_anchor_for_phrases_in_field(const Persistence& p, const std::string& path, const std::vector<uint32_t>& ids1, const std::vector<uint32_t>& ids2) {
    SearchFieldResult result;
    const vfmt::PhrasePairView& store = p.get_phrase_pair_to_anchor(path);
    for (uint32_t t1 : ids1)
        for (uint32_t t2 : ids2) store.get_values(t1, t2, result.hits_ids);
    std::sort(result.hits_ids.begin(), result.hits_ids.end());
    return result;
}

// boost.rs:34-87
static std::vector<Hit> boost_text_locality(const Persistence& p, const std::string& path, const std::map<std::string, std::vector<uint32_t>>& term_to_ids) {
    std::vector<Hit> boost_anchor;
    if (term_to_ids.size() <= 1) return boost_anchor;
    const vhost::KeyValueStore& t2t = p.get_valueid_to_parent(path + ".tokens_to_text_id");
    std::vector<uint32_t> all;
    for (auto& kv : term_to_ids)
        for (uint32_t id : kv.second) t2t.append_values(id, all);  // search.rs:113-120
    std::sort(all.begin(), all.end());
    std::vector<std::pair<uint32_t, size_t>> boost_text_ids;
    for (size_t i = 0; i < all.size();) {
        size_t j = i;
        while (j < all.size() && all[j] == all[i]) ++j;
        if (j - i > 1) boost_text_ids.emplace_back(all[i], j - i);
        i = j;
    }
    if (p.is_anchor_identity_column(path)) {
        for (auto& t : boost_text_ids) boost_anchor.push_back(Hit{t.first, 2.f * (float)t.second * (float)t.second});
    } else {
        const vhost::KeyValueStore& t2a = p.get_valueid_to_parent(path + ".text_id_to_anchor");
        std::vector<uint32_t> anchors;
        for (auto& t : boost_text_ids) {
            anchors.clear();
            t2a.append_values(t.first, anchors);
            for (uint32_t a : anchors) boost_anchor.push_back(Hit{a, 2.f * (float)t.second * (float)t.second});
        }
    }
    std::stable_sort(boost_anchor.begin(), boost_anchor.end(), [](const Hit& a, const Hit& b) { return a.id < b.id; });
    return boost_anchor;
}

// boost.rs:11-32.  The comparator passed to max_by is reversed, so the group's
// MINIMUM wins (last of equal minima -- irrelevant for f32 values).
static std::vector<Hit> boost_text_locality_all(const Persistence& p, const std::map<std::string, std::map<std::string, std::vector<uint32_t>>>& term_id_hits_in_field) {
    std::vector<std::vector<Hit>> boosts;
    for (auto& kv : term_id_hits_in_field) boosts.push_back(boost_text_locality(p, kv.first, kv.second));
    std::vector<Hit> merged = kmerge_hits(boosts), out;
    for (size_t i = 0; i < merged.size();) {
        size_t j = i;
        float best = merged[i].score;
        while (j < merged.size() && merged[j].id == merged[i].id) {
            if (merged[j].score < best) best = merged[j].score;
            ++j;
        }
        out.push_back(Hit{merged[i].id, best});
        i = j;
    }
    return out;
}

// sort.rs:5-22
static std::vector<Hit> top_n_sort(const std::vector<Hit>& data, uint32_t top_n) {
    float worst_score = -3.40282347e+38f;
    std::vector<Hit> new_data;
    for (const Hit& el : data) {
        if (el.score < worst_score) continue;
        if (!new_data.empty() && new_data.size() == (size_t)top_n + 200) {
            std::sort(new_data.begin(), new_data.end(), sort_by_score_and_id_less);
            new_data.resize(top_n);
            if (!new_data.empty()) worst_score = new_data.back().score;
        }
        new_data.push_back(el);
    }
    std::sort(new_data.begin(), new_data.end(), sort_by_score_and_id_less);
    return new_data;
}

struct FacetGroup {
    uint32_t id;
    uint32_t count;
    std::string text;
};

// facet.rs:31-73.  Ties in count are left unordered by the reference
// (sort_unstable over a hash map); here they are broken by ascending value id.
static std::vector<FacetGroup> get_facet(const Persistence& p, const vhost::FacetRequest& req, const std::vector<uint32_t>& ids) {
    std::vector<std::string> steps = vfmt::get_steps_to_anchor(req.field);
    std::map<uint32_t, uint32_t> counts;
    std::vector<uint32_t> tmp;
    if (steps.size() == 1 || p.has_index(steps.back() + ".anchor_to_text_id")) {
        std::string path = steps.size() == 1 ? steps.front() + ".parent_to_value_id" : steps.back() + ".anchor_to_text_id";
        const vhost::KeyValueStore& kv = p.get_valueid_to_parent(path);
        for (uint32_t id : ids)
            if (kv.get_values(id, tmp))
                for (uint32_t v : tmp) counts[v]++;
    } else {
        std::vector<uint32_t> cur = ids, next;  // join_anchor_to_leaf :75-83
        for (auto& step : steps) {
            const vhost::KeyValueStore& kv = p.get_valueid_to_parent(step + ".parent_to_value_id");
            next.clear();
            for (uint32_t id : cur)
                if (kv.get_values(id, tmp)) next.insert(next.end(), tmp.begin(), tmp.end());
            cur.swap(next);
        }
        for (uint32_t v : cur) counts[v]++;
    }
    std::vector<FacetGroup> groups;
    for (auto& kv : counts) groups.push_back(FacetGroup{kv.first, kv.second, std::string()});
    std::stable_sort(groups.begin(), groups.end(), [](const FacetGroup& a, const FacetGroup& b) { return a.count > b.count; });
    if (req.top && groups.size() > *req.top) groups.resize(*req.top);
    for (auto& g : groups) g.text = p.get_text_for_id(steps.back(), g.id);
    return groups;
}

struct SearchResult {  // result/search_result.rs:8-26
    uint64_t num_hits = 0;
    std::vector<Hit> data;
    std::vector<std::pair<std::string, std::vector<FacetGroup>>> facets;
    bool has_facets = false;
    ExplainMap explain;  // of the returned hits (search.rs:174 keeps the plan result's whole map; to_documents :86 reads it per hit)
};

// plan_creator + execute_steps, evaluated as a tree walk with the same dataflow.
class Executor {
  public:
    Executor(const Persistence& p, const Request& header) : p_(p), header_(header) {}

    SearchResult run() {
        Request request = header_;
        request.top = request.top ? request.top : std::optional<uint64_t>(10);  // search.rs:146
        if (!request.search_req) throw InvalidRequest("search_req is None, but is required in search");
        if (request.explain) {  // get_all_field_request_parts_and_propagate_settings (execution_plan.rs:46-90)
            if (request.phrase_boosts)
                for (auto& pb : *request.phrase_boosts) set_explain(pb.search1), set_explain(pb.search2);
            propagate_explain(*request.search_req);
            if (request.filter) propagate_explain(*request.filter);
        }
        // collect_all_field_request_into_cache (execution_plan.rs:91-130)
        if (request.phrase_boosts)
            for (auto& pb : *request.phrase_boosts) {
                add_to_cache(pb.search1, false);
                add_to_cache(pb.search2, false);
            }
        collect(*request.search_req, false);
        if (request.filter) collect(*request.filter, true);
        if (request.phrase_boosts)
            for (auto& pb : *request.phrase_boosts) {  // add_phrase_boost_plan_steps :229
                lookup(pb.search1).req.get_ids = true;
                lookup(pb.search2).req.get_ids = true;
            }
        mark_store_flags(*request.search_req);
        if (request.filter) mark_store_flags(*request.filter);

        std::optional<std::shared_ptr<FilterResult>> filter;
        SearchFieldResult filter_res;
        if (request.filter) {
            filter_res = eval(*request.filter, true, {}, std::nullopt);
            filter = std::make_shared<FilterResult>(FilterResult::from_result(filter_res.hits_ids));
        }
        std::vector<BoostPart> boosts = request.boost.value_or(std::vector<BoostPart>());
        SearchFieldResult res = eval(*request.search_req, false, boosts, filter);
        if (request.filter) res = intersect_score_hits_with_ids(std::move(res), filter_res);
        for (auto& b : boosts)  // execution_plan.rs:175-189
            if (b.path.find("[]") == std::string::npos) add_boost(p_, b, res);
        if (request.phrase_boosts) {  // :202-262 + plan_steps.rs:235-293
            std::vector<SearchFieldResult> phrase_results;
            for (auto& pb : *request.phrase_boosts) {
                const SearchFieldResult& r1 = field_result(pb.search1);
                const SearchFieldResult& r2 = field_result(pb.search2);
                if (pb.search1.path != pb.search2.path) throw InvalidRequest("phrase boost parts must be on the same path");
                std::string path = pb.search1.path;
                if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
                if (!vfmt::ends_with(path, ".phrase_pair_to_anchor")) path += ".phrase_pair_to_anchor";
                SearchFieldResult r = get_anchor_for_phrases_in_field(p_, path, r1.hits_ids, r2.hits_ids);
                r.phrase_boost = pb;
                phrase_results.push_back(std::move(r));
            }
            // sort_and_group_boosts_by_phrase_terms
            std::stable_sort(phrase_results.begin(), phrase_results.end(), [](const SearchFieldResult& a, const SearchFieldResult& b) {
                auto ka = std::make_pair(a.phrase_boost->search1.terms[0], a.phrase_boost->search2.terms[0]);
                auto kb = std::make_pair(b.phrase_boost->search1.terms[0], b.phrase_boost->search2.terms[0]);
                return ka < kb;
            });
            std::vector<SearchFieldResult> grouped;
            for (size_t i = 0; i < phrase_results.size();) {
                size_t j = i;
                std::vector<uint32_t> merged;
                while (j < phrase_results.size() && phrase_results[j].phrase_boost->search1.terms[0] == phrase_results[i].phrase_boost->search1.terms[0] &&
                       phrase_results[j].phrase_boost->search2.terms[0] == phrase_results[i].phrase_boost->search2.terms[0]) {
                    merged.insert(merged.end(), phrase_results[j].hits_ids.begin(), phrase_results[j].hits_ids.end());
                    ++j;
                }
                std::sort(merged.begin(), merged.end());
                merged.erase(std::unique(merged.begin(), merged.end()), merged.end());
                SearchFieldResult g;
                g.hits_ids = std::move(merged);
                g.request.boost = 5.0f;
                grouped.push_back(std::move(g));
                i = j;
            }
            boost_hits_ids_vec_multi(res, grouped);
        }
        ExplainMap plan_explain = res.explain;  // search.rs:174: taken before boost_term and text locality
        // search.rs:176-228
        if (request.boost_term) {
            std::vector<SearchFieldResult> data;
            for (auto& part : *request.boost_term) {
                PlanRequestSearchPart pr;
                pr.request = part;
                pr.get_ids = true;
                SearchFieldResult r = get_term_ids_in_field(p_, pr);
                data.push_back(resolve_token_to_anchor(p_, pr.request, std::nullopt, r));
            }
            boost_hits_ids_vec_multi(res, data);
        }
        if (request.text_locality) {
            std::vector<Hit> boost_anchor = boost_text_locality_all(p_, res.term_id_hits_in_field);
            apply_boost_from_iter(res, boost_anchor);
        }
        SearchResult out;
        if (request.facets) {
            std::vector<uint32_t> hit_ids;
            for (auto& h : res.hits_scores) hit_ids.push_back(h.id);
            std::sort(hit_ids.begin(), hit_ids.end());
            out.has_facets = true;
            for (auto& fr : *request.facets) out.facets.emplace_back(fr.field, get_facet(p_, fr, hit_ids));
        }
        out.num_hits = res.hits_scores.size();
        uint64_t skip = request.skip.value_or(0);
        out.data = top_n_sort(res.hits_scores, (uint32_t)*request.top + (uint32_t)skip);
        // apply_top_skip :230-239
        if (request.skip) out.data.erase(out.data.begin(), out.data.begin() + (long)std::min<uint64_t>(skip, out.data.size()));
        if (out.data.size() > *request.top) out.data.resize(*request.top);
        for (const Hit& h : out.data) {
            auto e = plan_explain.find(h.id);
            if (e != plan_explain.end()) out.explain[h.id] = e->second;
        }
        return out;
    }

  private:
    struct FieldSearch {
        PlanRequestSearchPart req;
        bool done = false;
        SearchFieldResult result;
    };
    const Persistence& p_;
    const Request& header_;
    std::map<std::string, FieldSearch> cache_;

    static void set_explain(SearchPart& part) { part.options.present = true, part.options.explain = true; }
    static void propagate_explain(SearchRequest& r) {
        if (r.kind == SearchRequest::Search) set_explain(r.part);
        else
            for (auto& q : r.queries) propagate_explain(q);
    }
    void add_to_cache(const SearchPart& part, bool ids_only) {
        auto it = cache_.find(part.key());
        if (it != cache_.end()) {
            it->second.req.get_ids |= ids_only;
            it->second.req.get_scores |= !ids_only;
            return;
        }
        FieldSearch fs;
        fs.req.request = part;
        fs.req.get_scores = !ids_only;
        fs.req.get_ids = ids_only;
        cache_.emplace(part.key(), std::move(fs));
    }
    void collect(const SearchRequest& r, bool ids_only) {
        if (r.kind == SearchRequest::Search) add_to_cache(r.part, ids_only);
        else
            for (auto& q : r.queries) collect(q, ids_only);
    }
    FieldSearch& lookup(const SearchPart& part) {
        auto it = cache_.find(part.key());
        if (it == cache_.end()) throw InvalidRequest("PlanCreator: Could not find request in field_search_cache");
        return it->second;
    }
    void mark_store_flags(const SearchRequest& r) {  // execution_plan.rs:401,417
        if (r.kind == SearchRequest::Search) {
            lookup(r.part).req.store_term_id_hits |= header_.why_found || header_.text_locality;
        } else
            for (auto& q : r.queries) mark_store_flags(q);
    }
    const SearchFieldResult& field_result(const SearchPart& part) {
        FieldSearch& fs = lookup(part);
        if (!fs.done) {
            fs.result = get_term_ids_in_field(p_, fs.req);
            fs.done = true;
        }
        return fs.result;
    }

    SearchFieldResult eval(const SearchRequest& r, bool is_filter, std::vector<BoostPart> boosts, const std::optional<std::shared_ptr<FilterResult>>& filter) {
        if (r.kind != SearchRequest::Search) {
            std::vector<SearchFieldResult> inputs;
            for (auto& q : r.queries) {
                std::vector<BoostPart> b = boosts;  // merge_vec :263-270
                if (q.get_boost()) b.insert(b.end(), q.get_boost()->begin(), q.get_boost()->end());
                inputs.push_back(eval(q, is_filter, b, filter));
            }
            if (r.kind == SearchRequest::Or) return is_filter ? union_hits_ids(std::move(inputs)) : union_hits_score(std::move(inputs));
            return is_filter ? intersect_hits_ids(std::move(inputs)) : intersect_hits_score(std::move(inputs));
        }
        // plan_creator_search_part :389-534
        const SearchPart& part = r.part;
        const SearchFieldResult& fr = field_result(part);
        size_t pos = part.path.rfind("[]");
        if (pos != std::string::npos) {
            std::string end_obj = part.path.substr(0, pos);
            std::vector<const BoostPart*> boosto;
            for (auto& b : boosts) {
                size_t bp = b.path.rfind("[]");
                if (bp != std::string::npos && b.path.substr(0, bp) == end_obj) boosto.push_back(&b);
            }
            if (!boosto.empty()) {
                if (boosto.size() != 1) throw InvalidRequest("more than one boost on the same 1:n level");
                SearchFieldResult anchors = resolve_token_to_anchor(p_, part, filter, fr);
                // BoostToAnchor plan_steps.rs:173-196
                SearchFieldResult field_result = fr;
                resolve_token_hits_to_text_id_ids_only(p_, part, field_result);
                field_result = join_to_parent_ids(p_, field_result, part.path + ".textindex" + ".value_id_to_parent");
                get_boost_ids_and_resolve_to_anchor(p_, boosto[0]->path, field_result);
                apply_boost_values_anchor(anchors, *boosto[0], field_result.boost_ids);  // ApplyAnchorBoost
                return anchors;
            }
        }
        return resolve_token_to_anchor(p_, part, filter, fr);
    }
};

// serde's externally tagged form of result/explain.rs:1-21 (floats with 9 significant digits)
static void write_explain(std::string& s, const Explain& e) {
    char buf[160];
    switch (e.kind) {
        case Explain::Boost: snprintf(buf, sizeof buf, "{\"Boost\":%.9g}", (double)e.a), s += buf; break;
        case Explain::MaxTokenToTextId: snprintf(buf, sizeof buf, "{\"MaxTokenToTextId\":%.9g}", (double)e.a), s += buf; break;
        case Explain::OrSumOverDistinctTerms: snprintf(buf, sizeof buf, "{\"OrSumOverDistinctTerms\":%.9g}", (double)e.a), s += buf; break;
        case Explain::TermToAnchor:
            snprintf(buf, sizeof buf, "{\"TermToAnchor\":{\"term_score\":%.9g,\"anchor_score\":%.9g,\"final_score\":%.9g,\"term_id\":%u}}", (double)e.a, (double)e.b, (double)e.c, e.term_id);
            s += buf;
            break;
        case Explain::LevenshteinScore:
            snprintf(buf, sizeof buf, "{\"LevenshteinScore\":{\"score\":%.9g,\"text_or_token_id\":", (double)e.a);
            s += buf;
            vjson::write_string(s, e.text);
            s += ",\"term_id\":" + std::to_string(e.term_id) + "}}";
            break;
    }
}

static std::string result_to_json(const SearchResult& r) {
    std::string s = "{\"num_hits\":" + std::to_string(r.num_hits) + ",\"data\":[";
    char buf[64];
    for (size_t i = 0; i < r.data.size(); ++i) {
        if (i) s += ",";
        uint32_t bits;
        memcpy(&bits, &r.data[i].score, 4);
        snprintf(buf, sizeof buf, "[%u,%.9g,%u]", r.data[i].id, (double)r.data[i].score, bits);
        s += buf;
    }
    s += "]";
    if (r.has_facets) {
        s += ",\"facets\":{";
        for (size_t i = 0; i < r.facets.size(); ++i) {
            if (i) s += ",";
            vjson::write_string(s, r.facets[i].first);
            s += ":[";
            for (size_t j = 0; j < r.facets[i].second.size(); ++j) {
                if (j) s += ",";
                s += "[";
                vjson::write_string(s, r.facets[i].second[j].text);
                s += "," + std::to_string(r.facets[i].second[j].count) + "," + std::to_string(r.facets[i].second[j].id) + "]";
            }
            s += "]";
        }
        s += "}";
    }
    if (!r.explain.empty()) {
        s += ",\"explain\":{";
        bool first = true;
        for (auto& kv : r.explain) {
            s += (first ? "\"" : ",\"") + std::to_string(kv.first) + "\":[";
            first = false;
            for (size_t j = 0; j < kv.second.size(); ++j) {
                if (j) s += ",";
                write_explain(s, kv.second[j]);
            }
            s += "]";
        }
        s += "}";
    }
    s += "}";
    return s;
}

static std::vector<Hit> hits_from_json(const vjson::Value& v) {
    std::vector<Hit> out;
    for (auto& e : v.arr) out.push_back(Hit{(uint32_t)e.arr[0].num, (float)e.arr[1].num});
    return out;
}
static std::string hits_to_json(const std::vector<Hit>& hits) {
    std::string s = "[";
    char buf[64];
    for (size_t i = 0; i < hits.size(); ++i) {
        if (i) s += ",";
        snprintf(buf, sizeof buf, "[%u,%.9g]", hits[i].id, (double)hits[i].score);
        s += buf;
    }
    return s + "]";
}
static std::string ids_to_json(const std::vector<uint32_t>& ids) {
    std::string s = "[";
    for (size_t i = 0; i < ids.size(); ++i) s += (i ? "," : "") + std::to_string(ids[i]);
    return s + "]";
}

// search_field.rs:147-176 get_text_score_id_from_result(suggest_text = true) and :178-217 suggest_multi / suggest.
// The reference sorts with sort_unstable_by (by text, then by score): which of several equal texts keeps its id, and
// the order among equal scores, are whatever its sort leaves; stable sorts here.
struct Suggestion {
    std::string text;
    float score;
    uint32_t id;
};
static std::vector<Suggestion> suggest_multi(const Persistence& p, const Request& req) {
    if (!req.suggest) throw InvalidRequest("only suggest allowed in suggest function");
    std::vector<Suggestion> out;
    for (const SearchPart& part : *req.suggest) {
        PlanRequestSearchPart plan;
        plan.request = part;
        plan.get_scores = true, plan.return_term = true, plan.return_term_lowercase = true;
        SearchFieldResult res = get_term_ids_in_field(p, plan);
        for (const Hit& h : res.hits_scores) out.push_back(Suggestion{res.terms[h.id], h.score, h.id});
    }
    std::stable_sort(out.begin(), out.end(), [](const Suggestion& a, const Suggestion& b) { return b.text < a.text; });
    std::vector<Suggestion> merged;  // dedup_by: the first of a run of equal texts stays and takes the largest score
    for (const Suggestion& sgg : out) {
        if (!merged.empty() && merged.back().text == sgg.text) {
            if (sgg.score > merged.back().score) merged.back().score = sgg.score;
        } else {
            merged.push_back(sgg);
        }
    }
    std::stable_sort(merged.begin(), merged.end(), [](const Suggestion& a, const Suggestion& b) { return a.score > b.score; });
    if (req.skip) merged.erase(merged.begin(), merged.begin() + (long)std::min<uint64_t>(*req.skip, merged.size()));  // search.rs:230-239
    if (req.top && merged.size() > *req.top) merged.resize(*req.top);
    return merged;
}
static std::string suggestions_to_json(const std::vector<Suggestion>& v) {
    std::string s = "[";
    char buf[64];
    for (size_t i = 0; i < v.size(); ++i) {
        if (i) s += ",";
        s += "[";
        vjson::write_string(s, v[i].text);
        snprintf(buf, sizeof buf, ",%.9g,%u]", (double)v[i].score, v[i].id);
        s += buf;
    }
    return s + "]";
}

// Named entry points for the unit-level known-answer tests.
static std::string call(const Persistence* p, const std::string& fn, const vjson::Value& a) {
    auto results_from = [&](const vjson::Value& arr) {
        std::vector<SearchFieldResult> rs;
        for (auto& e : arr.arr) {
            SearchFieldResult r;
            if (auto* h = e.get("hits_scores")) r.hits_scores = hits_from_json(*h);
            if (auto* h = e.get("hits_ids"))
                for (auto& x : h->arr) r.hits_ids.push_back((uint32_t)x.num);
            if (auto* t = e.get("term")) r.request.terms.push_back(t->str);
            if (auto* b = e.get("boost")) r.request.boost = (float)b->num;
            rs.push_back(std::move(r));
        }
        return rs;
    };
    if (fn == "suggest_multi") {  // the whole Request (suggest, top, skip); Request.top is None unless given (`..Default::default()` is not serde's default)
        Request r = vhost::parse_request(*a.get("request"));
        return suggestions_to_json(suggest_multi(*p, r));
    }
    if (fn == "suggest") {  // search_field.rs:219-228: one part, its own top/skip also bound the merged list
        Request r;
        SearchPart part = vhost::parse_search_part(*a.get("part"));
        r.suggest = std::vector<SearchPart>{part};
        r.top = part.top, r.skip = part.skip;
        return suggestions_to_json(suggest_multi(*p, r));
    }
    if (fn == "union_hits_score") return hits_to_json(union_hits_score(results_from(*a.get("inputs"))).hits_scores);
    if (fn == "union_hits_ids") return ids_to_json(union_hits_ids(results_from(*a.get("inputs"))).hits_ids);
    if (fn == "intersect_hits_score") return hits_to_json(intersect_hits_score(results_from(*a.get("inputs"))).hits_scores);
    if (fn == "intersect_hits_ids") return ids_to_json(intersect_hits_ids(results_from(*a.get("inputs"))).hits_ids);
    if (fn == "intersect_score_hits_with_ids") {
        auto rs = results_from(*a.get("inputs"));
        return hits_to_json(intersect_score_hits_with_ids(rs[0], rs[1]).hits_scores);
    }
    if (fn == "get_facet") {  // facet.rs:31-73 over the given hit ids -> [[text, count, value id]]
        vhost::FacetRequest fr;
        fr.field = a.get("field")->str;
        if (auto* t = a.get("top")) {
            if (t->is_null()) fr.top.reset();
            else fr.top = (uint64_t)t->num;
        }
        std::vector<uint32_t> ids;
        for (auto& x : a.get("ids")->arr) ids.push_back((uint32_t)x.num);
        std::string out = "[";
        bool first = true;
        for (auto& g : get_facet(*p, fr, ids)) {
            out += first ? "" : ",";
            first = false;
            out += "[";
            vjson::write_string(out, g.text);
            out += "," + std::to_string(g.count) + "," + std::to_string(g.id) + "]";
        }
        return out + "]";
    }
    if (fn == "apply_boost_values_anchor") {
        SearchFieldResult r;
        r.hits_scores = hits_from_json(*a.get("hits_scores"));
        BoostPart b = vhost::parse_boost_part(*a.get("boost"));
        apply_boost_values_anchor(r, b, hits_from_json(*a.get("boost_ids")));
        return hits_to_json(r.hits_scores);
    }
    if (fn == "boost_text_locality") {  // boost.rs:11-87 for one field: terms = {"term": [token ids]}
        std::string path = a.get("path")->str;
        if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
        std::map<std::string, std::map<std::string, std::vector<uint32_t>>> in;
        for (auto& kv : a.get("terms")->obj)
            for (auto& x : kv.second.arr) in[path][kv.first].push_back((uint32_t)x.num);
        return hits_to_json(boost_text_locality_all(*p, in));
    }
    if (fn == "boost_to_anchor") {  // the BoostToAnchor step, plan_steps.rs:174-196
        SearchFieldResult r;
        SearchPart part = vhost::parse_search_part(*a.get("part"));
        if (auto* h = a.get("hits_scores")) r.hits_scores = hits_from_json(*h);
        if (auto* h = a.get("hits_ids"))
            for (auto& x : h->arr) r.hits_ids.push_back((uint32_t)x.num);
        resolve_token_hits_to_text_id_ids_only(*p, part, r);
        r = join_to_parent_ids(*p, r, part.path + ".textindex" + ".value_id_to_parent");
        get_boost_ids_and_resolve_to_anchor(*p, vhost::parse_boost_part(*a.get("boost")).path, r);
        return hits_to_json(r.boost_ids);
    }
    if (fn == "get_anchor_for_phrases_in_field") {  // search_field.rs:247-275
        std::string path = a.get("path")->str;
        if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
        if (!vfmt::ends_with(path, ".phrase_pair_to_anchor")) path += ".phrase_pair_to_anchor";
        const vfmt::PhrasePairView& store = p->get_phrase_pair_to_anchor(path);
        std::vector<uint32_t> out;
        for (auto& t1 : a.get("ids1")->arr)
            for (auto& t2 : a.get("ids2")->arr) store.get_values((uint32_t)t1.num, (uint32_t)t2.num, out);
        std::sort(out.begin(), out.end());
        return ids_to_json(out);
    }
    if (fn == "boost_anchor_from_phrase_results") {  // plan_steps.rs:230-277; boosts: [{"hits_ids": [...], "phrase": ["a", "b"]}]
        SearchFieldResult r;
        r.hits_scores = hits_from_json(*a.get("hits_scores"));
        std::map<std::pair<std::string, std::string>, std::vector<std::vector<uint32_t>>> by_phrase;  // sort_unstable_by_key + group_by on the term pair
        for (auto& e : a.get("boosts")->arr) {
            std::vector<uint32_t> ids;
            for (auto& x : e.get("hits_ids")->arr) ids.push_back((uint32_t)x.num);
            by_phrase[{e.get("phrase")->arr[0].str, e.get("phrase")->arr[1].str}].push_back(std::move(ids));
        }
        std::vector<SearchFieldResult> boosts;
        for (auto& kv : by_phrase) {
            std::vector<uint32_t> merged;  // kmerge of the (sorted) lists, then dedup of neighbours
            for (auto& l : kv.second) merged.insert(merged.end(), l.begin(), l.end());
            std::stable_sort(merged.begin(), merged.end());
            merged.erase(std::unique(merged.begin(), merged.end()), merged.end());
            SearchFieldResult b;
            b.hits_ids = std::move(merged);
            b.request.boost = 5.0f;
            boosts.push_back(std::move(b));
        }
        boost_hits_ids_vec_multi(r, boosts);
        return hits_to_json(r.hits_scores);
    }
    if (fn == "boost_hits_ids_vec_multi") {
        SearchFieldResult r;
        r.hits_scores = hits_from_json(*a.get("hits_scores"));
        auto bs = results_from(*a.get("boosts"));
        boost_hits_ids_vec_multi(r, bs);
        return hits_to_json(r.hits_scores);
    }
    if (fn == "regex_accepts") {  // regex_sim.hpp: '0' / '1' per term; "bad: .." / "outside: .." when the pattern does not compile
        try {
            oracle_regex::Program prog(a.get("pattern")->str, a.get("case_insensitive")->b);
            std::string bits = "\"";
            std::vector<uint32_t> scalars;
            for (auto& t : a.get("terms")->arr) {
                scalars.clear();
                vfmt::utf8_decode(t.str, scalars);
                bits += prog.accepts(scalars, a.get("starts_with")->b) ? '1' : '0';
            }
            return bits + "\"";
        } catch (const oracle_regex::BadPattern& e) {
            return std::string("\"bad\"");
        } catch (const oracle_regex::OutsideSubset& e) {
            return std::string("\"outside\"");
        }
    }
    if (fn == "to_lowercase") {
        std::string out;
        vjson::write_string(out, olow::to_lowercase(a.get("text")->str));
        return out;
    }
    if (fn == "distance") return std::to_string((int)distance(a.get("a")->str, a.get("b")->str));
    if (fn == "distance_dfa") return std::to_string((int)distance_dfa(a.get("hit")->str, a.get("term")->str, (uint32_t)a.get("d")->num));
    if (fn == "default_score") {
        char buf[64];
        snprintf(buf, sizeof buf, "%.9g", (double)get_default_score_for_distance((uint8_t)a.get("distance")->num, a.get("prefix")->b));
        return buf;
    }
    if (fn == "expression") {
        char buf[64];
        snprintf(buf, sizeof buf, "%.9g", (double)ScoreExpression(a.get("expr")->str).get_score((float)a.get("value")->num));
        return buf;
    }
    if (fn == "top_n_sort") return hits_to_json(top_n_sort(hits_from_json(*a.get("hits")), (uint32_t)a.get("top")->num));
    if (fn == "f16_roundtrip") {
        char buf[64];
        snprintf(buf, sizeof buf, "%.9g", (double)f16_roundtrip((float)a.get("value")->num));
        return buf;
    }
    if (fn == "token_score") return std::to_string(vindex::calculate_token_score_for_entry((uint32_t)a.get("pos")->num, (uint32_t)a.get("nocc")->num, (uint32_t)a.get("ntok")->num, a.get("exact")->b));
    if (fn == "steps_to_anchor") {
        std::string s = "[";
        auto st = vfmt::get_steps_to_anchor(a.get("path")->str);
        for (size_t i = 0; i < st.size(); ++i) {
            if (i) s += ",";
            vjson::write_string(s, st[i]);
        }
        return s + "]";
    }
    if (fn == "tokenize") {
        std::vector<std::pair<std::string, bool>> toks;
        vindex::tokenize(a.get("text")->str, vindex::default_separators(), toks);
        std::string s = "[";
        for (size_t i = 0; i < toks.size(); ++i) {
            if (i) s += ",";
            vjson::write_string(s, toks[i].first);
        }
        return s + "]";
    }
    // ---- codec round trips (writer -> bytes -> reader) for the reference's codec unit tests
    auto u32s = [](const vjson::Value& arr) {
        std::vector<uint32_t> v;
        for (auto& x : arr.arr) v.push_back((uint32_t)x.num);
        return v;
    };
    if (fn == "codec_indirect") {
        vfmt::IndirectWriter w;
        for (auto& e : a.get("adds")->arr) w.add((uint32_t)e.arr[0].num, u32s(e.arr[1]));
        vfmt::IndirectView v;
        v.start_pos = (const uint8_t*)w.ids.data();
        v.n_ids = w.ids.size();
        v.data = w.data.data();
        v.data_len = w.data.size();
        std::string s = "{\"values\":[";
        std::vector<uint32_t> out;
        bool first = true;
        std::map<uint32_t, uint32_t> counts;
        for (auto& q : a.get("queries")->arr) {
            s += first ? "" : ",";
            first = false;
            if (v.get_values((uint64_t)q.num, out)) {
                s += ids_to_json(out);
                for (uint32_t x : out) counts[x]++;
            } else s += "null";
        }
        s += "],\"counts\":{";
        first = true;
        for (auto& kv : counts) {
            s += (first ? "\"" : ",\"") + std::to_string(kv.first) + "\":" + std::to_string(kv.second);
            first = false;
        }
        return s + "},\"data_len\":" + std::to_string(w.data.size()) + "}";
    }
    if (fn == "codec_packed") {
        std::vector<uint32_t> vals = u32s(*a.get("stored"));  // already value+1, like encode_vals' input
        uint32_t mx = 0;
        for (uint32_t x : vals) mx = std::max(mx, x);
        int width = vfmt::packed_bytes_required(mx);
        std::vector<uint8_t> bytes(vals.size() * (size_t)width);
        for (size_t i = 0; i < vals.size(); ++i) memcpy(&bytes[i * width], &vals[i], (size_t)width);
        vfmt::PackedView v;
        v.bytes = bytes.data();
        v.len = bytes.size();
        v.width = width;
        std::string s = "{\"width\":" + std::to_string(width) + ",\"values\":[";
        bool first = true;
        for (auto& q : a.get("queries")->arr) {
            uint32_t out;
            s += first ? "" : ",";
            first = false;
            s += v.get_value((uint64_t)q.num, out) ? std::to_string(out) : "null";
        }
        return s + "]}";
    }
    if (fn == "codec_phrase") {
        vfmt::PhrasePairWriter w;
        for (auto& e : a.get("adds")->arr) {
            std::vector<uint32_t> vals = u32s(e.arr[2]);
            w.add((uint32_t)e.arr[0].num, (uint32_t)e.arr[1].num, vals.data(), vals.size());
        }
        vfmt::PhrasePairView v;
        v.recs = w.recs.data();
        v.n = w.recs.size() / 12;
        v.data = w.data.data();
        v.data_len = w.data.size();
        std::string s = "{\"size\":" + std::to_string(v.n) + ",\"offsets\":[";
        for (size_t i = 0; i < v.n; ++i) s += (i ? "," : "") + std::to_string(vfmt::load_u32(v.recs + i * 12 + 8));
        s += "],\"values\":[";
        bool first = true;
        for (auto& q : a.get("queries")->arr) {
            std::vector<uint32_t> out;
            s += first ? "" : ",";
            first = false;
            s += v.get_values((uint32_t)q.arr[0].num, (uint32_t)q.arr[1].num, out) ? ids_to_json(out) : "null";
        }
        return s + "]}";
    }
    if (fn == "codec_anchor_score") {
        vfmt::AnchorScoreWriter w;
        for (auto& e : a.get("adds")->arr) {
            std::vector<uint32_t> pairs = u32s(e.arr[1]);
            w.set_scores((uint32_t)e.arr[0].num, pairs.data(), pairs.size());
        }
        std::vector<uint8_t> sp = w.encode_start_pos();
        vfmt::AnchorScoreView v;
        if (a.get("wide") && a.get("wide")->b) {  // data_type U64 (token_to_anchor_score_vint.rs:242-248): 8-byte start positions
            std::vector<uint8_t> wide(sp.size() * 2, 0);
            for (size_t i = 0; i < sp.size() / 4; ++i) memcpy(&wide[i * 8], &sp[i * 4], 4);
            sp.swap(wide);
            v.wide = true;
        }
        v.start_pos = sp.data();
        v.start_len = sp.size();
        v.data = w.data.data();
        v.data_len = w.data.size();
        std::string s = "[";
        bool first = true;
        for (auto& q : a.get("queries")->arr) {
            s += first ? "[" : ",[";
            first = false;
            bool f2 = true;
            v.for_each((uint32_t)q.num, [&](uint32_t anchor, uint32_t score) {
                s += (f2 ? "[" : ",[") + std::to_string(anchor) + "," + std::to_string(score) + "]";
                f2 = false;
            });
            s += "]";
        }
        return s + "]";
    }
    if (fn == "fst_roundtrip") {
        vfmt::FstWriter w;
        uint64_t i = 0;
        for (auto& k : a.get("keys")->arr) w.insert(k.str, a.get("values") ? (uint64_t)a.get("values")->arr[i].num : i), ++i;
        std::vector<uint8_t> bytes = w.finish();
        vfmt::FstReader r(bytes.data(), bytes.size());
        std::string s = "{\"len\":" + std::to_string(r.len()) + ",\"bytes\":" + std::to_string(bytes.size()) + ",\"items\":[";
        bool first = true;
        r.for_each([&](const std::string& k, uint64_t v) {
            s += first ? "[" : ",[";
            first = false;
            vjson::write_string(s, k);
            s += "," + std::to_string(v) + "]";
        });
        s += "],\"ord_to_term\":[";
        first = true;
        if (a.get("ords"))
            for (auto& o : a.get("ords")->arr) {
                std::string t;
                s += first ? "" : ",";
                first = false;
                if (r.ord_to_term((uint64_t)o.num, t)) vjson::write_string(s, t);
                else s += "null";
            }
        return s + "]}";
    }
    if (!p) throw InvalidRequest("call needs an index: " + fn);
    if (fn == "field_search") {  // get_term_ids_in_field
        PlanRequestSearchPart pr;
        pr.request = vhost::parse_search_part(*a.get("part"));
        pr.get_scores = true;
        pr.get_ids = a.get("get_ids") && a.get("get_ids")->b;
        SearchFieldResult r = get_term_ids_in_field(*p, pr);
        std::string s = "{\"hits_scores\":" + hits_to_json(r.hits_scores) + ",\"hits_ids\":" + ids_to_json(r.hits_ids) + ",\"terms\":[";
        const vhost::TermDict& dict = p->get_dict(pr.request.path);
        for (size_t i = 0; i < r.hits_scores.size(); ++i) {
            if (i) s += ",";
            size_t slot;
            vjson::write_string(s, dict.find_id(r.hits_scores[i].id, slot) ? dict.term(slot) : std::string());
        }
        return s + "]}";
    }
    if (fn == "resolve_token_to_anchor") {
        PlanRequestSearchPart pr;
        pr.request = vhost::parse_search_part(*a.get("part"));
        pr.get_scores = true;
        pr.get_ids = a.get("get_ids") && a.get("get_ids")->b;
        SearchFieldResult r = get_term_ids_in_field(*p, pr);
        SearchFieldResult res = resolve_token_to_anchor(*p, pr.request, std::nullopt, r);
        return "{\"hits_scores\":" + hits_to_json(res.hits_scores) + ",\"hits_ids\":" + ids_to_json(res.hits_ids) + "}";
    }
    if (fn == "dict") {
        const vhost::TermDict& dict = p->get_dict(a.get("path")->str);
        std::string s = "[";
        for (size_t i = 0; i < dict.size(); ++i) {
            if (i) s += ",";
            s += "[";
            vjson::write_string(s, dict.term(i));
            s += "," + std::to_string(dict.ids[i]) + "]";
        }
        return s + "]";
    }
    if (fn == "get_values") {
        const vhost::KeyValueStore& kv = p->get_valueid_to_parent(a.get("path")->str);
        std::vector<uint32_t> v;
        if (!kv.get_values((uint64_t)a.get("id")->num, v)) return "null";
        return ids_to_json(v);
    }
    if (fn == "phrase_pairs") {
        std::vector<uint32_t> v;
        const vfmt::PhrasePairView& pp = p->get_phrase_pair_to_anchor(a.get("path")->str);
        if (!pp.get_values((uint32_t)a.get("t1")->num, (uint32_t)a.get("t2")->num, v)) return "null";
        return ids_to_json(v);
    }
    if (fn == "postings") {
        std::string s = "[";
        bool first = true;
        p->get_token_to_anchor(a.get("path")->str).for_each((uint32_t)a.get("id")->num, [&](uint32_t anchor, uint32_t score) {
            s += (first ? "[" : ",[") + std::to_string(anchor) + "," + std::to_string(score) + "]";
            first = false;
        });
        return s + "]";
    }
    throw InvalidRequest("unknown oracle function " + fn);
}

}  // namespace oracle

// ----------------------------------------------------------------- C API ----
extern "C" {

struct vo_index {
    std::unique_ptr<Persistence> p;
};

static void set_err(char* err, size_t n, const std::string& msg) {
    if (err && n) {
        snprintf(err, n, "%s", msg.c_str());
    }
}
static char* dup_str(const std::string& s) {
    char* r = (char*)malloc(s.size() + 1);
    memcpy(r, s.c_str(), s.size() + 1);
    return r;
}

vo_index* vo_open(const char* dir, char* err, size_t errlen) {
    try {
        vo_index* h = new vo_index();
        h->p = Persistence::load(dir);
        return h;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}
void vo_close(vo_index* h) { delete h; }
void vo_free(char* s) { free(s); }

// status: 0 ok, 1 InvalidRequest, 2 FstNotFound, 3 path not found, 4 io, 5 json
static int run_guarded(const std::function<std::string()>& f, char** out, char* err, size_t errlen) {
    try {
        *out = dup_str(f());
        return 0;
    } catch (const oracle::InvalidRequest& e) {
        set_err(err, errlen, e.what());
        return 1;
    } catch (const vhost::FstNotFound& e) {
        set_err(err, errlen, e.what());
        return 2;
    } catch (const vhost::PathNotFound& e) {
        set_err(err, errlen, e.what());
        return 3;
    } catch (const vhost::IoError& e) {
        set_err(err, errlen, e.what());
        return 4;
    } catch (const vhost::RequestError& e) {
        set_err(err, errlen, e.what());
        return 5;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return 9;
    }
}

int vo_search(vo_index* h, const char* request_json, char** out, char* err, size_t errlen) {
    return run_guarded(
        [&]() {
            Request r = vhost::parse_request_json(request_json, strlen(request_json));
            oracle::Executor ex(*h->p, r);
            return oracle::result_to_json(ex.run());
        },
        out, err, errlen);
}

int vo_call(vo_index* h, const char* fn, const char* args_json, char** out, char* err, size_t errlen) {
    return run_guarded(
        [&]() {
            vjson::Value a = vjson::parse(args_json, strlen(args_json));
            return oracle::call(h ? h->p.get() : nullptr, fn, a);
        },
        out, err, errlen);
}

// Runs `n` requests on `threads` host threads (one query at a time per thread, like
// the reference); returns wall seconds, fills ids/scores/num_hits for the first
// `k` hits of each query (rows of k, padded with id 0xFFFFFFFF).
double vo_search_batch(vo_index* h, const char* const* requests, uint32_t n, uint32_t threads, uint32_t k, uint32_t* ids, float* scores, uint64_t* num_hits, int32_t* status) {
    if (threads == 0) threads = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            for (uint32_t q = t; q < n; q += threads) {
                int st = 0;
                try {
                    Request r = vhost::parse_request_json(requests[q], strlen(requests[q]));
                    oracle::Executor ex(*h->p, r);
                    oracle::SearchResult res = ex.run();
                    if (num_hits) num_hits[q] = res.num_hits;
                    for (uint32_t i = 0; i < k; ++i) {
                        if (ids) ids[(size_t)q * k + i] = i < res.data.size() ? res.data[i].id : 0xFFFFFFFFu;
                        if (scores) scores[(size_t)q * k + i] = i < res.data.size() ? res.data[i].score : 0.f;
                    }
                } catch (const std::exception&) {
                    st = 9;
                }
                if (status) status[q] = st;
            }
        });
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
