// Deterministic synthetic corpora written straight into
// the veloci on-disk layout, plus the matching request streams.  Test/bench
// infrastructure, not on the accelerated path.
//
// The files are what src/create.rs would write for documents
//   {"body": "<w1> <w2> ... <wn>", "commonness": <f32>, "tags": ["..",".."]}
// under a field config where `body` texts are longer than
// `do_not_store_text_longer_than` (so only tokens enter the FST,
// create_fulltext.rs:60-64,99-104): token ids are ranks in byte order with the
// separator token " " ranked first, token positions count separators
// (create.rs:243-251), posting scores follow calculate_score.rs:34-49, long
// text ids follow get_text_info (create.rs:143-161).  Deliberate omission: the
// per-document "whole text" postings of the long text ids (create.rs:221-225),
// which no FST key can reach.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "indexer.hpp"

namespace vindex {

struct Rng {  // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

struct SynthParams {
    uint64_t num_docs = 100000;
    uint32_t vocab = 20000;
    uint32_t tokens_per_doc = 8;
    double zipf_s = 1.07;
    uint32_t len_min = 3, len_max = 12;
    uint64_t seed = 42;
    bool boost = true;           // `commonness` f32 column, LogUniform[1, 1e6]
    uint32_t tags = 0;           // number of distinct tags (0 = no `tags[]` facet field), 2 per doc
    bool text_locality = false;  // tokens_to_text_id + text_id_to_anchor
    bool phrase = false;         // phrase_pair_to_anchor
    // requests
    uint32_t num_queries = 1000;
    uint64_t query_seed = 43;
    std::string query_kind = "or3";  // or3 | and | single
    uint32_t levenshtein = 1;
    double edit_prob = 0.5;
    uint32_t top = 10;

    static SynthParams from_json(const std::string& json) {
        SynthParams p;
        vjson::Value v = vjson::parse(json);
        auto num = [&](const char* k, double def) {
            const vjson::Value* x = v.get(k);
            return (x && x->is_number()) ? x->num : def;
        };
        auto flag = [&](const char* k, bool def) {
            const vjson::Value* x = v.get(k);
            return (x && x->is_bool()) ? x->b : def;
        };
        p.num_docs = (uint64_t)num("num_docs", (double)p.num_docs);
        p.vocab = (uint32_t)num("vocab", p.vocab);
        p.tokens_per_doc = (uint32_t)num("tokens_per_doc", p.tokens_per_doc);
        p.zipf_s = num("zipf_s", p.zipf_s);
        p.len_min = (uint32_t)num("len_min", p.len_min);
        p.len_max = (uint32_t)num("len_max", p.len_max);
        p.seed = (uint64_t)num("seed", (double)p.seed);
        p.boost = flag("boost", p.boost);
        p.tags = (uint32_t)num("tags", p.tags);
        p.text_locality = flag("text_locality", p.text_locality);
        p.phrase = flag("phrase", p.phrase);
        p.num_queries = (uint32_t)num("num_queries", p.num_queries);
        p.query_seed = (uint64_t)num("query_seed", (double)p.query_seed);
        if (const vjson::Value* k = v.get("query_kind")) p.query_kind = k->str;
        p.levenshtein = (uint32_t)num("levenshtein", p.levenshtein);
        p.edit_prob = num("edit_prob", p.edit_prob);
        p.top = (uint32_t)num("top", p.top);
        return p;
    }
};

// Zipf(s) over ranks 0..n-1 by inverse CDF with a guide table.
struct ZipfSampler {
    std::vector<double> cdf;
    std::vector<uint32_t> guide;
    ZipfSampler(uint32_t n, double s) : cdf(n), guide(65537) {
        double acc = 0;
        for (uint32_t k = 0; k < n; ++k) {
            acc += 1.0 / pow((double)(k + 1), s);
            cdf[k] = acc;
        }
        for (auto& c : cdf) c /= acc;
        cdf[n - 1] = 1.0;
        uint32_t k = 0;
        for (uint32_t g = 0; g <= 65536; ++g) {
            double u = (double)g / 65536.0;
            while (k < n - 1 && cdf[k] < u) ++k;
            guide[g] = k;
        }
    }
    uint32_t sample(Rng& r) const {
        double u = r.uniform();
        uint32_t g = (uint32_t)(u * 65536.0);
        uint32_t lo = guide[g], hi = guide[g + 1];
        while (lo < hi) {
            uint32_t mid = (lo + hi) / 2;
            if (cdf[mid] < u) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    }
};

// Vocabulary in Zipf-rank order (rank 0 = most frequent); same for corpus and requests.
inline std::vector<std::string> make_vocabulary(const SynthParams& p) {
    Rng r(p.seed * 0x51ED2701u + 17);
    std::unordered_set<std::string> seen;
    std::vector<std::string> words;
    words.reserve(p.vocab);
    seen.reserve((size_t)p.vocab * 2);
    while (words.size() < p.vocab) {
        uint32_t len = p.len_min + r.below(p.len_max - p.len_min + 1);
        std::string w(len, 'a');
        for (auto& c : w) c = (char)('a' + r.below(26));
        if (seen.insert(w).second) words.push_back(std::move(w));
    }
    return words;
}

inline void write_vec(const std::string& path, const std::vector<uint8_t>& v) { vhost::write_file(path, v.data(), v.size()); }
inline void write_u32s(const std::string& path, const std::vector<uint32_t>& v) { vhost::write_file(path, v.data(), v.size() * 4); }

inline void write_synthetic_index(const std::string& dir, const SynthParams& p) {
    mkdir(dir.c_str(), 0755);
    const uint64_t A = p.num_docs;
    const uint32_t V = p.vocab, T = p.tokens_per_doc;
    std::vector<std::string> words = make_vocabulary(p);
    // term ids: " " is the smallest key (id 0); words follow in byte order
    std::vector<uint32_t> order(V);
    for (uint32_t i = 0; i < V; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return words[a] < words[b]; });
    std::vector<uint32_t> id_of_rank(V);
    for (uint32_t i = 0; i < V; ++i) id_of_rank[order[i]] = i + 1;
    const uint32_t num_terms = V + 1;

    // documents: T Zipf ranks each; fixed 64 chunks so the corpus does not depend on the thread count
    std::vector<uint32_t> doc_tokens((size_t)A * T);
    ZipfSampler zipf(V, p.zipf_s);
    {
        const uint32_t chunks = 64;
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < hw; ++t)
            pool.emplace_back([&, t]() {
                for (uint32_t c = t; c < chunks; c += hw) {
                    uint64_t lo = A * c / chunks, hi = A * (c + 1) / chunks;
                    Rng r(p.seed * 1000003ull + c);
                    for (uint64_t d = lo; d < hi; ++d)
                        for (uint32_t k = 0; k < T; ++k) doc_tokens[d * T + k] = id_of_rank[zipf.sample(r)];
                }
            });
        for (auto& th : pool) th.join();
    }
    // global occurrence counts (TermInfo::num_occurences)
    std::vector<uint32_t> nocc(num_terms, 0);
    for (uint32_t t : doc_tokens) nocc[t]++;
    nocc[0] = (uint32_t)std::min<uint64_t>(A * (T - 1), UINT32_MAX);
    // postings per term: distinct tokens of a doc, best (first) position
    std::vector<uint64_t> df(num_terms + 1, 0);
    auto for_doc_distinct = [&](uint64_t d, auto&& f) {
        const uint32_t* tk = &doc_tokens[d * T];
        for (uint32_t k = 0; k < T; ++k) {
            bool first = true;
            for (uint32_t j = 0; j < k; ++j)
                if (tk[j] == tk[k]) first = false;
            if (first) f(tk[k], 2 * k);
        }
        if (T > 1) f(0u, 1u);
    };
    for (uint64_t d = 0; d < A; ++d) for_doc_distinct(d, [&](uint32_t t, uint32_t) { df[t + 1]++; });
    for (uint32_t t = 0; t < num_terms; ++t) df[t + 1] += df[t];
    const uint64_t P = df[num_terms];
    std::vector<uint32_t> post_anchor(P);
    std::vector<uint16_t> post_score(P);
    {
        std::vector<uint64_t> cur(df.begin(), df.end() - 1);
        const uint32_t ntok = T > 1 ? 2 * T - 1 : 1;
        // score depends only on (pos, nocc): cache per term the value for pos 0 to skip most log calls
        for (uint64_t d = 0; d < A; ++d)
            for_doc_distinct(d, [&](uint32_t t, uint32_t pos) {
                uint64_t at = cur[t]++;
                post_anchor[at] = (uint32_t)d;
                post_score[at] = (uint16_t)calculate_token_score_for_entry(pos, nocc[t], ntok, false);
            });
    }
    vhost::Metadata meta;
    meta.num_docs = A;
    // ---- body.textindex.to_anchor_id_score + fst
    {
        vfmt::AnchorScoreWriter w;
        w.data.reserve((size_t)P * 3);
        std::vector<uint32_t> pairs;
        for (uint32_t t = 0; t < num_terms; ++t) {
            uint64_t lo = df[t], hi = df[t + 1];
            if (lo == hi) continue;
            pairs.resize((size_t)(hi - lo) * 2);
            for (uint64_t i = lo; i < hi; ++i) {
                pairs[(i - lo) * 2] = post_anchor[i];
                pairs[(i - lo) * 2 + 1] = post_score[i];
            }
            w.set_scores(t, pairs.data(), pairs.size());
        }
        if (w.pos.size() < num_terms) w.pos.resize(num_terms, 0);
        w.finish();
        write_vec(dir + "/body.textindex.to_anchor_id_score.indirect", w.encode_start_pos());
        write_vec(dir + "/body.textindex.to_anchor_id_score.data", w.data);
        vfmt::FstWriter fw;
        if (T > 1) fw.insert(std::string(" "), 0);
        for (uint32_t i = 0; i < V; ++i) fw.insert(words[order[i]], i + 1);
        write_vec(dir + "/body.textindex.fst", fw.finish());
        FieldInfo fi;
        fi.name = "body";
        fi.has_fst = true;
        fi.num_text_ids = num_terms;
        fi.num_long_text_ids = A;
        fi.do_not_store_text_longer_than = 16;
        IndexMetadata im;
        im.path = "body.textindex.to_anchor_id_score";
        im.category = IndexCategory::AnchorScore;
        im.meta = w.meta;
        im.data_type_u64 = w.needs_u64();
        fi.indices.push_back(im);
        meta.columns["body"] = fi;
    }
    // long text ids (create.rs:143-161): terms.len() + 1 + counter, counter continuing after pass 1
    const uint32_t text_id_base = num_terms + 1 + (uint32_t)A + 1;
    if (p.text_locality) {
        // tokens_to_text_id: token -> sorted unique text ids of the documents containing it
        vfmt::IndirectWriter w;
        std::vector<uint32_t> vals;
        uint32_t max_v = 0;
        for (uint32_t t = 0; t < num_terms; ++t) {
            uint64_t lo = df[t], hi = df[t + 1];
            if (lo == hi) continue;
            vals.resize(hi - lo);
            for (uint64_t i = lo; i < hi; ++i) vals[i - lo] = text_id_base + post_anchor[i];
            max_v = std::max(max_v, vals.back());
            w.add(t, vals);
        }
        w.finish();
        w.meta.max_value_id = max_v;
        write_u32s(dir + "/body.textindex.tokens_to_text_id.indirect", w.ids);
        write_vec(dir + "/body.textindex.tokens_to_text_id.data", w.data);
        IndexMetadata im;
        im.path = "body.textindex.tokens_to_text_id";
        im.meta = w.meta;
        meta.columns["body"].indices.push_back(im);
        // text_id_to_anchor: one anchor per long text id (inline)
        std::vector<uint32_t> ids((size_t)text_id_base + A, 0);
        for (uint64_t d = 0; d < A; ++d) ids[text_id_base + d] = (uint32_t)d | vfmt::kHighBit;
        write_u32s(dir + "/body.textindex.text_id_to_anchor.indirect", ids);
        write_vec(dir + "/body.textindex.text_id_to_anchor.data", std::vector<uint8_t>(1, 0));
        IndexMetadata im2;
        im2.path = "body.textindex.text_id_to_anchor";
        im2.meta.num_values = A;
        im2.meta.num_ids = (uint32_t)A;
        im2.meta.max_value_id = (uint32_t)(A - 1);
        im2.meta.avg_join_size = 1.f;
        meta.columns["body"].indices.push_back(im2);
    }
    if (p.phrase) {
        std::vector<std::pair<uint64_t, uint32_t>> pairs;
        pairs.reserve((size_t)A * (T - 1));
        for (uint64_t d = 0; d < A; ++d)
            for (uint32_t k = 0; k + 1 < T; ++k) pairs.emplace_back(((uint64_t)doc_tokens[d * T + k] << 32) | doc_tokens[d * T + k + 1], (uint32_t)d);
        std::sort(pairs.begin(), pairs.end());
        vfmt::PhrasePairWriter w;
        std::vector<uint32_t> group;
        for (size_t i = 0; i < pairs.size();) {
            size_t j = i;
            group.clear();
            while (j < pairs.size() && pairs[j].first == pairs[i].first) group.push_back(pairs[j++].second);
            group.erase(std::unique(group.begin(), group.end()), group.end());
            w.add((uint32_t)(pairs[i].first >> 32), (uint32_t)pairs[i].first, group.data(), group.size());
            i = j;
        }
        w.finish();
        write_vec(dir + "/body.textindex.phrase_pair_to_anchor.indirect", w.recs);
        write_vec(dir + "/body.textindex.phrase_pair_to_anchor.data", w.data);
        IndexMetadata im;
        im.path = "body.textindex.phrase_pair_to_anchor";
        im.category = IndexCategory::Phrase;
        im.is_empty = w.empty();
        im.meta = w.meta;
        meta.columns["body"].indices.push_back(im);
    }
    if (p.boost) {
        Rng r(p.seed * 7919ull + 5);
        std::vector<uint32_t> ids(A);
        for (uint64_t d = 0; d < A; ++d) {
            float v = (float)pow(10.0, r.uniform() * 6.0);
            uint32_t bits;
            memcpy(&bits, &v, 4);
            ids[d] = bits | vfmt::kHighBit;  // single value inlined in the .indirect slot (create_indirect.rs:67-70)
        }
        write_u32s(dir + "/commonness.boost_valid_to_value.indirect", ids);
        write_vec(dir + "/commonness.boost_valid_to_value.data", std::vector<uint8_t>(1, 0));
        FieldInfo fi;
        fi.name = "commonness";
        fi.has_fst = false;
        IndexMetadata im;
        im.path = "commonness.boost_valid_to_value";
        im.category = IndexCategory::Boost;
        im.meta.num_values = A;
        im.meta.num_ids = (uint32_t)A;
        im.meta.avg_join_size = 1.f;
        fi.indices.push_back(im);
        meta.columns["commonness"] = fi;
    }
    if (p.tags) {
        Rng r(p.seed * 104729ull + 9);
        ZipfSampler tz(p.tags, 1.0);
        char buf[32];
        std::vector<std::string> tag_names(p.tags);
        for (uint32_t i = 0; i < p.tags; ++i) {
            snprintf(buf, sizeof buf, "tag%05u", i);  // byte order == numeric order -> term id == i
            tag_names[i] = buf;
        }
        vfmt::FstWriter fw;
        for (uint32_t i = 0; i < p.tags; ++i) fw.insert(tag_names[i], i);
        write_vec(dir + "/tags[].textindex.fst", fw.finish());
        vfmt::IndirectWriter w;
        std::vector<uint32_t> vals(2);
        for (uint64_t d = 0; d < A; ++d) {
            vals[0] = tz.sample(r);
            vals[1] = tz.sample(r);
            w.add((uint32_t)d, vals);  // anchor_to_text_id keeps insertion order and duplicates (no sort_and_dedup)
        }
        w.finish();
        w.meta.max_value_id = p.tags - 1;
        write_u32s(dir + "/tags[].textindex.anchor_to_text_id.indirect", w.ids);
        write_vec(dir + "/tags[].textindex.anchor_to_text_id.data", w.data);
        FieldInfo fi;
        fi.name = "tags[]";
        fi.has_fst = true;
        fi.tokenize = true;
        fi.num_text_ids = p.tags;
        IndexMetadata im;
        im.path = "tags[].textindex.anchor_to_text_id";
        im.meta = w.meta;
        fi.indices.push_back(im);
        meta.columns["tags[]"] = fi;
    }
    std::string mj = vjson::to_string(vhost::metadata_to_json(meta), 2);
    vhost::write_file(dir + "/metaData.json", mj.data(), mj.size());
}

inline std::string edit_word(const std::string& w, Rng& r) {
    std::string s = w;
    uint32_t kind = r.below(3);
    char c = (char)('a' + r.below(26));
    if (kind == 0 && !s.empty()) s[r.below((uint32_t)s.size())] = c;               // substitute
    else if (kind == 1) s.insert(s.begin() + r.below((uint32_t)s.size() + 1), c);  // insert
    else if (s.size() > 2) s.erase(s.begin() + r.below((uint32_t)s.size()));       // delete
    return s;
}

inline void write_synthetic_requests(const std::string& out_path, const SynthParams& p) {
    std::vector<std::string> words = make_vocabulary(p);
    ZipfSampler zipf(p.vocab, p.zipf_s);
    Rng r(p.query_seed * 2654435761ull + 3);
    std::string out;
    auto part = [&](const std::string& w, uint32_t lev) {
        std::string s = "{\"terms\":[\"" + w + "\"],\"path\":\"body\"";
        if (lev) s += ",\"levenshtein_distance\":" + std::to_string(lev);
        return s + "}";
    };
    for (uint32_t q = 0; q < p.num_queries; ++q) {
        if (p.query_kind == "single") {
            std::string w = words[zipf.sample(r)];
            if (r.uniform() < p.edit_prob) w = edit_word(w, r);
            out += "{\"search_req\":{\"search\":" + part(w, p.levenshtein) + "},\"top\":" + std::to_string(p.top) + "}\n";
        } else if (p.query_kind == "and") {
            uint32_t n = 2 + r.below(2);
            std::vector<std::string> ws;
            std::vector<uint32_t> levs;
            for (uint32_t i = 0; i < n; ++i) {
                ws.push_back(words[zipf.sample(r)]);
                levs.push_back(r.below(2) ? p.levenshtein : 0);
            }
            std::string s = "{\"search_req\":{\"and\":{\"queries\":[";
            for (uint32_t i = 0; i < n; ++i) s += (i ? "," : "") + ("{\"search\":" + part(ws[i], levs[i]) + "}");
            s += "]}}";
            if (p.phrase) {
                s += ",\"phrase_boosts\":[";
                for (uint32_t i = 0; i + 1 < n; ++i) s += (i ? "," : "") + ("{\"search1\":" + part(ws[i], levs[i]) + ",\"search2\":" + part(ws[i + 1], levs[i + 1]) + "}");
                s += "]";
            }
            if (p.text_locality) s += ",\"text_locality\":true";
            if (p.tags) s += ",\"facets\":[{\"field\":\"tags[]\"}]";
            if (p.boost) s += ",\"boost\":[{\"path\":\"commonness\",\"boost_fun\":\"Log10\",\"param\":1}]";
            out += s + ",\"top\":" + std::to_string(p.top) + "}\n";
        } else {  // or3
            std::string s = "{\"search_req\":{\"or\":{\"queries\":[";
            for (uint32_t i = 0; i < 3; ++i) {
                std::string w = words[zipf.sample(r)];
                if (r.uniform() < p.edit_prob) w = edit_word(w, r);
                s += (i ? "," : "") + ("{\"search\":" + part(w, p.levenshtein) + "}");
            }
            s += "]}}";
            if (p.boost) s += ",\"boost\":[{\"path\":\"commonness\",\"boost_fun\":\"Log10\",\"param\":1}]";
            out += s + ",\"top\":" + std::to_string(p.top) + "}\n";
        }
    }
    vhost::write_file(out_path, out.data(), out.size());
}

}  // namespace vindex
